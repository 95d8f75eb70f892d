"""The device-side synthetic transmitter / channel through the C ABI against the host modulator
(gr_lora_b200/tx.py) and through the receiver itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


@pytest.mark.parametrize("sf", [7, 10, 12])
def test_tx_symbols_equal_host_modulator(torch, sf):
    """No noise, no CFO, the host's chirp table: bit-identical to tx.modulate_shifts (a cyclic shift is a table look-up).
    With the decoder's own table: the reference's ideal up-chirp, shifted."""
    import gr_lora_b200 as G
    from gr_lora_b200 import tx
    nb, sps = 1 << sf, 8 << sf
    vals = np.concatenate([[0, 1, nb - 1, nb // 2], np.random.default_rng(sf).integers(0, nb, 29)]).astype(np.uint32)
    dec = G.decoder(1e6, 125000, sf, False, 4, True, demod="fft", quiet=True)
    up = tx.base_upchirp(sf).astype(np.complex64)
    out = torch.empty((len(vals), sps), dtype=torch.complex64, device="cuda")
    v = torch.from_numpy(vals.astype(np.int32)).cuda()
    dec.tx_symbols(v, out, len(vals), up_table_dev=torch.from_numpy(up).cuda())
    torch.cuda.synchronize()
    want = tx.modulate_shifts(vals, sf).astype(np.complex64).reshape(len(vals), sps)
    assert np.array_equal(out.cpu().numpy(), want)
    dec.tx_symbols(v, out, len(vals))                      # the decoder's own ideal up-chirp
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.abs(np.abs(got) - np.sqrt(2.0)).max() < 1e-3      # the reference's table is (1 + 1j) e^{j phase} (lib/decoder_impl.cc:159-160)
    bins = torch.empty(len(vals), dtype=torch.int32, device="cuda")
    dec.demod_fft(out, len(vals), bins)
    torch.cuda.synchronize()
    assert np.array_equal(bins.cpu().numpy().astype(np.uint32), vals)
    dec.close()


def test_tx_noise_and_cfo(torch):
    """AWGN: zero mean, variance sigma^2 per component, reproducible for a seed, different across seeds and across
    symbols; CFO: a rotation by 2 pi f n / fs."""
    import gr_lora_b200 as G
    from gr_lora_b200 import tx
    sf, n = 8, 512
    sps = 8 << sf
    dec = G.decoder(1e6, 125000, sf, False, 4, True, quiet=True)
    v = torch.zeros(n, dtype=torch.int32, device="cuda")
    a = torch.empty((n, sps), dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    clean = torch.empty_like(a)
    up = torch.from_numpy(tx.base_upchirp(sf).astype(np.complex64)).cuda()
    dec.tx_symbols(v, clean, n, up_table_dev=up)
    dec.tx_symbols(v, a, n, noise_sigma=0.25, seed=7, up_table_dev=up)
    dec.tx_symbols(v, b, n, noise_sigma=0.25, seed=7, up_table_dev=up)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    z = torch.view_as_real(a - clean).cpu().numpy().astype(np.float64)
    assert abs(z.mean()) < 2e-3 and abs(z.std() - 0.25) < 2e-3
    assert abs(np.mean(z[..., 0] * z[..., 1])) < 2e-3                      # I and Q uncorrelated
    assert abs(np.mean(z[0] * z[1])) < 5e-3                                # symbols get different noise
    k4 = np.mean(z ** 4) / np.mean(z ** 2) ** 2
    assert abs(k4 - 3.0) < 0.05                                            # Gaussian kurtosis
    dec.tx_symbols(v, b, n, noise_sigma=0.25, seed=8, up_table_dev=up)
    torch.cuda.synchronize()
    assert not torch.equal(a, b)
    cfo = torch.full((n,), 1234.5, dtype=torch.float32, device="cuda")
    dec.tx_symbols(v, b, n, cfo_hz_dev=cfo, up_table_dev=up)
    torch.cuda.synchronize()
    want = clean.cpu().numpy()[0] * np.exp(2j * np.pi * 1234.5 * np.arange(sps) / 1e6)
    assert np.abs(b.cpu().numpy()[3] - want).max() < 2e-4
    dec.close()


def test_tx_expand_feeds_the_receiver(torch):
    """tx_expand: K host-built captures -> 24 streams with their own noise on the device -> every frame decodes."""
    import gr_lora_b200 as G
    from conftest import make_capture
    caps, pays = [], []
    for k in range(3):
        p = bytes([k + 1, 2, 3, 4, 5, 6, 7])
        caps.append(make_capture(p, 7, 4, False, seed=100 + k, n_frames=1, lead=2.5, snr_db=60.0))
        pays.append(p)
    n = max(c.size for c in caps) // 2 * 2
    base = np.zeros((3, n), np.complex64)
    for k, c in enumerate(caps):
        base[k, : min(n, c.size)] = c[:n]
    ns = 24
    dec = G.decoder(1e6, 125000, 7, False, 4, False, n_streams=ns, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    out = torch.empty((ns, n), dtype=torch.complex64, device="cuda")
    dec.tx_expand(torch.from_numpy(base).cuda(), 3, n, ns, out, noise_sigma=float(np.sqrt(10 ** (-3.5) / 2)), seed=5)
    torch.cuda.synchronize()
    x = out.cpu().numpy()
    assert not np.array_equal(x[0], x[3]) and np.abs(x[0] - x[3]).max() < 0.2     # same capture, different noise
    dec.work_batch(out, n_items=n, stride_items=n, host=0, callbacks=False)
    fr = dec.frames_last()
    assert len(fr) == ns
    for r in fr:
        assert bytes(r["bytes"][18: int(r["len"])])[: len(pays[0])] == pays[int(r["stream"]) % 3]
    dec.close()
