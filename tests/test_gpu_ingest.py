"""GPU: the host-buffer entry points of the drop-in call.

lora_b200_work_batch with host memory stages the streams in groups (copy of group g + 1 under the state machine of
group g); lora_b200_work_batch_sc16 / lora_b200_demod_fft_host_sc16 take SDR-native int16 I/Q and convert on the
device.  Both must publish exactly what the gr_complex device path publishes, which in turn equals the oracle."""
import numpy as np
import pytest

from conftest import make_capture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _streams(sf, n_streams, seed0):
    caps, pays = [], []
    rng = np.random.default_rng(seed0)
    for k in range(n_streams):
        p = bytes(rng.integers(0, 256, 6 + k % 5, dtype=np.uint8))
        caps.append(make_capture(p, sf, 1 + k % 4, bool(k & 1), seed=seed0 + k, n_frames=1 + k % 2, lead=2.0 + 0.37 * (k % 6), snr_db=30.0))
        pays.append(p)
    n = max(c.size for c in caps)
    x = np.zeros((n_streams, n), np.complex64)
    for k, c in enumerate(caps):
        x[k, :c.size] = c
    return x, pays


@pytest.mark.parametrize("sf,n_streams", [(7, 21), (9, 5)])
def test_host_groups_equal_device_path_and_oracle(torch, oracle, sf, n_streams):
    """21 streams -> 8 staging groups of 3 (ragged last group).  Per stream: frames and consumed items equal the
    device-pointer path and the oracle (every stream has its own coding rate in the header; the decoder is created with
    cr 4 like the reference's apps do for explicit-header captures)."""
    import gr_lora_b200 as G
    x, pays = _streams(sf, n_streams, 900 + sf)
    n = x.shape[1]
    want = []
    for k in range(n_streams):
        od = oracle.Decoder(sf=sf, cr=4, crc=True)
        oc, _ = od.run(x[k])
        want.append((oc, od.frames()))
    res = {}
    for mode in ("host", "device"):
        dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=n_streams, quiet=True, max_items_per_call=n, max_frames_per_call=4)
        if mode == "host":
            consumed = dec.work_batch(x)
        else:
            consumed = dec.work_batch(torch.from_numpy(x).cuda(), n_items=n, stride_items=n, host=0)
        per = [[] for _ in range(n_streams)]
        for s, f in dec.frames:
            per[s].append(f)
        res[mode] = (list(map(int, consumed)), per)
        dec.close()
    assert res["host"] == res["device"]
    for k in range(n_streams):
        assert res["host"][0][k] == want[k][0], k
        assert res["host"][1][k] == want[k][1], k
    assert sum(len(w[1]) for w in want) >= n_streams


def _to_sc16(x, scale):
    """What an SDR front end delivers: int16 I/Q; and the host-side conversion back to gr_complex (x * scale in fp32)."""
    q = np.empty(x.shape + (2,), np.int16)
    q[..., 0] = np.clip(np.rint(x.real / scale), -32768, 32767).astype(np.int16)
    q[..., 1] = np.clip(np.rint(x.imag / scale), -32768, 32767).astype(np.int16)
    back = (q[..., 0].astype(np.float32) * np.float32(scale) + 1j * (q[..., 1].astype(np.float32) * np.float32(scale))).astype(np.complex64)
    return q, back


@pytest.mark.parametrize("sf", [7, 10])
def test_work_batch_sc16_equals_host_converted(torch, oracle, sf):
    import gr_lora_b200 as G
    scale = 1.0 / 4096.0
    x, pays = _streams(sf, 9, 1300 + sf)
    q, back = _to_sc16(x, scale)
    n = x.shape[1]
    dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=9, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    consumed = dec.work_batch(q, sc16_scale=scale)
    per = [[] for _ in range(9)]
    for s, f in dec.frames:
        per[s].append(f)
    dec.close()
    for k in range(9):
        od = oracle.Decoder(sf=sf, cr=4, crc=True)
        oc, _ = od.run(back[k])
        assert int(consumed[k]) == oc and per[k] == od.frames(), k
    assert sum(len(p) for p in per) >= 9
    # misaligned device pointer / odd item counts take the scalar tail of the converter
    dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=1, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    c1 = dec.work_batch(q[3:4, : n - 3], sc16_scale=scale)
    od = oracle.Decoder(sf=sf, cr=4, crc=True)
    oc, _ = od.run(back[3, : n - 3])
    assert int(c1[0]) == oc and [f for _, f in dec.frames] == od.frames()
    dec.close()


@pytest.mark.parametrize("sf,n", [(7, 9001), (12, 300)])
def test_demod_fft_host_sc16(torch, oracle, sf, n):
    """K1 from int16 host memory, several 64 MiB chunks: bins and magnitudes equal those of the gr_complex entry point on
    the host-converted buffer bit for bit (same kernel, same inputs) and the oracle's bins."""
    import gr_lora_b200 as G
    from gr_lora_b200 import tx
    nb = 1 << sf
    rng = np.random.default_rng(77 + sf)
    vals = rng.integers(0, nb, n)
    x = np.empty(n * (8 << sf), np.complex64)
    step = max(1, (64 << 20) // (16 * (8 << sf)))
    for s in range(0, n, step):
        e = min(n, s + step)
        x[s * (8 << sf): e * (8 << sf)] = tx.synth_symbols(vals[s:e], sf, snr_db=0.0, seed=s + 1)
    scale = 1.0 / 2048.0
    q, back = _to_sc16(x, scale)
    dec = G.decoder(1e6, 125000, sf, False, 4, True, demod="fft", quiet=True)
    b16, m16 = dec.demod_fft_host_sc16(q, scale)
    b32, m32 = dec.demod_fft_host(back)
    assert np.array_equal(b16, b32) and np.array_equal(m16, m32)
    ob, om = oracle.Decoder(sf=sf).demod_fft_batch(back)
    assert np.array_equal(b16, ob)
    np.testing.assert_allclose(m16, om, rtol=1e-4)
    assert np.mean(b16 == vals) > 0.999
    dec.close()


def test_frames_last_equals_callbacks(torch, oracle):
    """lora_b200_frames_last: the bulk view of what the last call published = the callback sequence (stream, bytes),
    also when no callback is registered at all."""
    import gr_lora_b200 as G
    x, pays = _streams(7, 12, 2100)
    n = x.shape[1]
    dec = G.decoder(1e6, 125000, 7, False, 4, True, n_streams=12, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    dec.work_batch(x)
    fr = dec.frames_last()
    assert len(fr) == len(dec.frames) > 0
    for rec, (s, f) in zip(fr, dec.frames):
        assert int(rec["stream"]) == s and bytes(rec["bytes"][: int(rec["len"])]) == f
    dec.close()
    dec = G.decoder(1e6, 125000, 7, False, 4, True, n_streams=12, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    dec.work_batch(x, callbacks=False)
    fr2 = dec.frames_last()
    assert dec.frames == [] and fr2.tobytes() == fr.tobytes()
    dec.close()


def test_reset_replays_identically(torch, oracle):
    """lora_b200_reset: after it the same capture decodes to the same frames and consume amounts as on a new decoder, also
    when the first pass stopped in the middle of a frame."""
    import gr_lora_b200 as G
    x, pays = _streams(7, 12, 2100)
    n = x.shape[1]
    dec = G.decoder(1e6, 125000, 7, False, 4, True, n_streams=12, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    def published():
        return [(int(r["stream"]), bytes(r["bytes"][: int(r["len"])])) for r in dec.frames_last()]

    c0 = dec.work_batch(x, callbacks=False).copy()
    fr0 = published()
    dec.reset()
    cut = (n * 2 // 3) // 1024 * 1024                    # stop inside a frame, then restart from scratch
    dec.work_batch(np.ascontiguousarray(x[:, :cut]), callbacks=False)
    dec.reset()
    assert all(dec.state(s) == 0 for s in range(12))
    c1 = dec.work_batch(x, callbacks=False).copy()
    assert np.array_equal(c0, c1) and published() == fr0 and len(fr0) > 0
    dec.close()


def test_work_batch_sc8_equals_host_converted(torch, oracle):
    """int8 I/Q: frames and consume amounts equal the oracle's on the host-converted buffer (x * scale in fp32), vector path
    and the scalar tail of the converter."""
    import gr_lora_b200 as G
    sf, scale = 7, 1.0 / 64.0
    x, pays = _streams(sf, 9, 4100)
    q = np.stack([np.clip(np.round(x.real / scale), -127, 127), np.clip(np.round(x.imag / scale), -127, 127)], axis=-1).astype(np.int8)
    back = (q[..., 0].astype(np.float32) * np.float32(scale) + 1j * (q[..., 1].astype(np.float32) * np.float32(scale))).astype(np.complex64)
    n = x.shape[1]
    dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=9, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    consumed = dec.work_batch(q, sc8_scale=scale)
    per = [[] for _ in range(9)]
    for s, f in dec.frames:
        per[s].append(f)
    dec.close()
    for k in range(9):
        od = oracle.Decoder(sf=sf, cr=4, crc=True)
        oc, _ = od.run(back[k])
        assert int(consumed[k]) == oc and per[k] == od.frames(), k
    assert sum(len(p) for p in per) >= 9
    dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=1, quiet=True, max_items_per_call=n, max_frames_per_call=4)
    c1 = dec.work_batch(q[3:4, : n - 5], sc8_scale=scale)
    od = oracle.Decoder(sf=sf, cr=4, crc=True)
    oc, _ = od.run(back[3, : n - 5])
    assert int(c1[0]) == oc and [f for _, f in dec.frames] == od.frames()
    dec.close()
