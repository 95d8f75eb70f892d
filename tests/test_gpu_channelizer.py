"""GPU: the channelizer against a float64 restatement of GNU Radio's firdes::low_pass +
freq_xlating_fir_filter_ccf (gr-filter is not in the reference tree: parity is "unpinned"),
and end to end in front of the decoder like python/lora_receiver.py:51-68 wires it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def xlating_fir_reference(x, taps, f_off, fs, decim, n0=0, hist=None):
    """y[n] = e^{-j w D (n0+n)} sum_k taps[k] e^{j w k} x[nD - k], zero (or given) history, float64."""
    w = 2 * np.pi * f_off / fs
    k = np.arange(taps.size)
    ct = taps.astype(np.float64) * np.exp(1j * w * k)
    h = np.zeros(taps.size - 1, np.complex128) if hist is None else hist
    xe = np.concatenate([h, x.astype(np.complex128)])
    full = np.convolve(xe, ct)[taps.size - 1: taps.size - 1 + x.size]
    y = full[::decim]
    return y * np.exp(-1j * w * decim * (n0 + np.arange(y.size)))


def test_taps_match_firdes_formula(torch):
    import gr_lora_b200 as G
    from gr_lora_b200.channelizer import firdes_low_pass_reference
    for fs in (1e6, 4e6, 10e6):
        ch = G.channelizer(fs, 868e6, [868.1e6], 125000, 1)
        ref = firdes_low_pass_reference(fs, 125000 // 2 + 15000, 10000)
        assert ch.ntaps == ref.size and ch.ntaps % 2 == 1
        np.testing.assert_allclose(ch.taps(), ref, rtol=2e-6, atol=1e-9)
        assert abs(ch.taps().sum() - 1.0) < 1e-5
        ch.close()
    assert G.channelizer(1e6, 0, [0], 125000, 1).ntaps == 241


@pytest.mark.parametrize("fs,decim,offsets", [(1e6, 1, [100e3]), (4e6, 4, [1.1e6, -700e3, 0.0, 300e3, -1.5e6]), (10e6, 10, [2.4e6, -3.1e6])])
def test_fir_bank_matches_float64_reference_and_streams(torch, fs, decim, offsets):
    import gr_lora_b200 as G
    rng = np.random.default_rng(int(fs) % 97 + decim)
    n_in = 20000 * decim
    x = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
    center = 868e6
    ch = G.channelizer(fs, center, [center + f for f in offsets], 125000, decim)
    taps = ch.taps()
    d_in = torch.from_numpy(x).cuda()
    n_out = n_in // decim
    d_out = torch.zeros((len(offsets), n_out), dtype=torch.complex64, device="cuda")
    # two calls (history and rotator phase must carry over) == one call == reference
    cut = (n_out // 3) * decim
    got1 = ch.work_dev(d_in[:cut], cut, d_out, n_out)
    got2 = ch.work_dev(d_in[cut:], n_in - cut, d_out[:, got1:], n_out)
    torch.cuda.synchronize()
    assert got1 + got2 == n_out
    y = d_out.cpu().numpy()
    for c, f in enumerate(offsets):
        # float32(center + f) - float32(center) is what the library sees (channel_list is float, like the reference's)
        f_eff = float(np.float32(center + f)) - float(np.float32(center))
        ref = xlating_fir_reference(x, taps, f_eff, fs, decim)
        err = np.max(np.abs(y[c] - ref)) / np.max(np.abs(ref))
        assert err < 2e-5, (c, f, err)
    ch.close()


def _wideband_capture(payload, sf, fs_in, decim, f_off, snr_db, seed):
    from scipy.signal import resample_poly
    from gr_lora_b200 import tx
    frame = tx.modulate_frame(tx.encode_frame(payload, sf, 4), sf)
    base = tx.channel([frame] * 2, sf=sf, snr_db=None, seed=seed).astype(np.complex128)
    up = resample_poly(base, decim, 1) if decim > 1 else base
    n = np.arange(up.size)
    x = up * np.exp(2j * np.pi * f_off * n / fs_in)
    x = x + tx.awgn(x.size, snr_db, np.random.default_rng(seed))
    return x.astype(np.complex64)


@pytest.mark.parametrize("fs_in,decim,f_off", [(1e6, 1, 100e3), (4e6, 4, 1.1e6)])
def test_receiver_with_channelizer_decodes_offset_channel(torch, fs_in, decim, f_off):
    """lora_receiver(samp_rate, center, [channel], ...) exactly like apps/lora_receive_file_nogui.py:31:
    a frame 100 kHz / 1.1 MHz off centre at +15 dB wide-band SNR (too noisy for the decoder without the
    channel filter) is decoded to the transmitted bytes; the IQ stays on the GPU between the blocks."""
    import gr_lora_b200 as G
    payload = bytes.fromhex("deadbeef700d")
    x = _wideband_capture(payload, 7, fs_in, decim, f_off, 15.0, 5)
    center = 868.0e6
    rx = G.lora_receiver(fs_in, center, [center + f_off], 125000, 7, False, 4, True, decimation=decim, quiet=True)
    rx.run(x)
    assert [f[15:].hex() for _, f in rx.frames] == ["049040deadbeef700d"] * 2
    # without channelization the same capture (shifted back by hand) does not even synchronise
    if decim == 1:
        raw = G.lora_receiver(fs_in, center, [center], 125000, 7, False, 4, True, disable_channelization=True, quiet=True)
        raw.run((x * np.exp(-2j * np.pi * f_off * np.arange(x.size) / fs_in)).astype(np.complex64))
        assert len(raw.frames) == 0


def test_apply_cfo_retunes(torch):
    import gr_lora_b200 as G
    fs, center = 1e6, 868e6
    ch = G.channelizer(fs, center, [center + 50e3], 125000, 1)
    n = 8000
    tone = np.exp(2j * np.pi * 60e3 * np.arange(n) / fs).astype(np.complex64)       # 10 kHz above the channel centre
    d_in = torch.from_numpy(tone).cuda()
    d_out = torch.zeros((1, n), dtype=torch.complex64, device="cuda")
    ch.work_dev(d_in, n, d_out, n)
    torch.cuda.synchronize()
    y = d_out.cpu().numpy()[0, 1000:]
    f_before = np.angle(np.mean(y[1:] * np.conj(y[:-1]))) * fs / (2 * np.pi)
    ch.apply_cfo(10e3)                                                               # control message ("cfo" . 10e3)
    ch.work_dev(d_in, n, d_out, n)
    torch.cuda.synchronize()
    y = d_out.cpu().numpy()[0, 1000:]
    f_after = np.angle(np.mean(y[1:] * np.conj(y[:-1]))) * fs / (2 * np.pi)
    assert abs(f_before - 10e3) < 50 and abs(f_after) < 50
    ch.close()


def test_receiver_conj_after_channelizer(torch):
    """lora_receiver(conj=True): channelizer -> conjugate_cc -> decoder (python/lora_receiver.py:62-63,70-75).  A spectrally
    inverted capture (conjugated IQ, mirrored offset) decodes only with conj=True; the conjugation happens in the
    channelizer's output stage on the device."""
    import gr_lora_b200 as G
    payload = bytes.fromhex("deadbeef700d")
    f_off = 100e3
    x = _wideband_capture(payload, 7, 1e6, 1, f_off, 15.0, 5)
    xi = np.conj(x)                                           # inverted spectrum: the channel now sits at -f_off
    center = 868.0e6
    rx = G.lora_receiver(1e6, center, [center - f_off], 125000, 7, False, 4, True, conj=True, quiet=True)
    rx.run(xi)
    assert [f[15:].hex() for _, f in rx.frames] == ["049040deadbeef700d"] * 2
    plain = G.lora_receiver(1e6, center, [center - f_off], 125000, 7, False, 4, True, conj=False, quiet=True)
    plain.run(xi)
    assert [f[15:].hex() for _, f in plain.frames] != ["049040deadbeef700d"] * 2
    # conjugate flag on the raw channelizer: output == conj(output without it)
    ch = G.channelizer(1e6, center, [center + 50e3, center - 75e3], 125000, 1)
    n = 4096
    d_in = torch.from_numpy(x[:n].copy()).cuda()
    a = torch.zeros((2, n), dtype=torch.complex64, device="cuda")
    ch.work_dev(d_in, n, a, n)
    ch2 = G.channelizer(1e6, center, [center + 50e3, center - 75e3], 125000, 1)
    ch2.set_conjugate(True)
    b = torch.zeros((2, n), dtype=torch.complex64, device="cuda")
    ch2.work_dev(d_in, n, b, n)
    torch.cuda.synchronize()
    assert torch.equal(b, torch.conj(a).resolve_conj())
    ch.close(); ch2.close()


def test_receiver_rejects_multi_stream_decoder(torch):
    import gr_lora_b200 as G
    with pytest.raises(ValueError, match="one decoder stream"):
        G.lora_receiver(1e6, 868e6, [868.1e6], 125000, 7, False, 4, True, quiet=True, n_streams=4)


def test_cpp_channelizer_shim(torch, tmp_path):
    """The compiled C++ drop-in (gr_lora_b200/host/channelizer_impl.cc: lora::channelizer::make with the reference's
    signature, include/lora/channelizer.h:49) driven with GNU-Radio-sized buffers: every channel's output equals the float64
    restatement, history and rotator phase carry across work() calls, apply_cfo retunes channel 0."""
    import os
    import subprocess
    from gr_lora_b200 import build as B
    import gr_lora_b200 as G
    exe = B.build_chan_shim()
    fs, decim, center = 4e6, 4, 868e6
    offsets = [1.1e6, -700e3, 300e3]
    rng = np.random.default_rng(3)
    n_in = 60000
    x = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
    f_in = tmp_path / "in.cf32"
    x.tofile(f_in)
    cmd = [str(exe), str(f_in), str(tmp_path / "out"), str(fs), str(center), "125000", str(decim), "4096"] + [repr(center + f) for f in offsets]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert f"PRODUCED {n_in // decim}" in p.stderr
    taps = G.channelizer(fs, center, [center], 125000, decim).taps()
    for c, f in enumerate(offsets):
        y = np.fromfile(tmp_path / f"out.{c}.cf32", np.complex64)
        f_eff = float(np.float32(center + f)) - float(np.float32(center))
        ref = xlating_fir_reference(x, taps, f_eff, fs, decim)
        assert y.size == ref.size
        assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 2e-5, c
    # apply_cfo after the first buffer: channel 0 of the rest is tuned 10 kHz higher
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env={**os.environ, "CHAN_SHIM_CFO": "10000"})
    assert p.returncode == 0, p.stderr
    y = np.fromfile(tmp_path / "out.0.cf32", np.complex64)
    f_eff = float(np.float32(center + offsets[0])) - float(np.float32(center))
    ref0 = xlating_fir_reference(x, taps, f_eff, fs, decim)
    assert np.max(np.abs(y[:4096] - ref0[:4096])) / np.max(np.abs(ref0)) < 2e-5
    assert np.max(np.abs(y[8192:] - ref0[8192:])) / np.max(np.abs(ref0)) > 0.1


def test_cfo_feedback_retunes_the_channelizer(torch):
    """decoder 'control' -> channelizer 'control' (python/lora_receiver.py:64; lib/decoder_impl.cc:774-776 commented out in the
    reference): with cfo_feedback the estimate of one run() is applied with apply_cfo, the next run sees the residual."""
    import gr_lora_b200 as G
    payload = bytes.fromhex("deadbeef700d")
    f_off, cfo = 100e3, 3000.0
    x = _wideband_capture(payload, 7, 1e6, 1, f_off + cfo, 25.0, 9)
    center = 868.0e6
    rx = G.lora_receiver(1e6, center, [center + f_off], 125000, 7, False, 4, True, quiet=True, cfo_feedback=True)
    rx.run(x)
    first, n1 = rx.decoder.last_cfo()
    assert n1 >= 1 and 0.3 * cfo < first < 1.25 * cfo          # part of the offset is absorbed as timing by SYNC (chirp ambiguity)
    rx.run(x)                                                 # same capture again: the channelizer is now tuned `first` Hz higher
    second, n2 = rx.decoder.last_cfo()
    assert n2 > n1 and abs(second) < 0.75 * abs(first)        # the loop converges: the residual shrinks
    off = G.lora_receiver(1e6, center, [center + f_off], 125000, 7, False, 4, True, quiet=True)
    off.run(x)
    assert off.decoder.last_cfo() == (0.0, 0)
