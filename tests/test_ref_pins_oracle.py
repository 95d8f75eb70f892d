"""CPU: the restatement (oracle/lora_oracle.c) against the REFERENCE'S OWN code.

oracle/_ref/liblora_ref.so is the reference's unmodified lib/decoder_impl.cc compiled against stand-in headers for the
absent third-party libraries (oracle/ref_wrap.cc, oracle/ref_standins/README.md).  Everything the reference's own
source decides must be reproduced by the restatement exactly: derived parameters and banner (A1), chirp tables (A2),
instantaneous frequency (A3), get_shift_fft (A4), max_frequency_gradient_idx (A5), fine_sync (A6), the three detectors
and the energy (A8-A11), the integer chain (B1-B4), and the whole work() state machine (A7, A12, B5-B7): per-step state,
consume amount, demodulated bin, fine-sync correction, published frames and stdout.

Float comparisons are bit-exact wherever both sides add in the same order (both use in-order scalar loops; the stand-in
VOLK is VOLK's generic protokernel order).  The FFT is the one place with a tolerance (two different radix-2
factorisations in fp32): bins equal, magnitudes within rtol 2e-5."""
import numpy as np
import pytest

from conftest import FRAME_CASES, case_decoder_args, make_capture, make_case_iq
from gr_lora_b200 import tx

SFS = range(7, 13)


def _noisy_symbols(sf, n, snr_db, seed):
    rng = np.random.default_rng(seed)
    vals = rng.integers(0, 1 << sf, n)
    return vals, tx.synth_symbols(vals, sf, snr_db=snr_db, seed=seed + 1)


@pytest.mark.parametrize("sf", SFS)
def test_parameters_banner_tables(oracle, ref, sf):
    """A1 lib/decoder_impl.cc:69-103, A2 :141-175."""
    for cr, implicit, ddc in ((4, False, False), (1, True, False), (3, False, True)):
        o = oracle.Decoder(sf=sf, cr=cr, implicit=implicit, disable_drift_correction=ddc)
        r = ref.RefDecoder(sf=sf, cr=cr, implicit=implicit, disable_drift_correction=ddc)
        assert (o.sps, o.n_bins, o.decim) == (r.sps, r.n_bins, r.decim)
        assert r.output_multiple == 2 * r.sps and r.delay_after_sync == r.sps // 4 and r.n_bins_hdr == r.n_bins // 4
        assert o.stdout == r.stdout
    for name in ("downchirp", "upchirp", "downchirp_ifreq", "upchirp_ifreq", "upchirp_ifreq_v"):
        a, b = getattr(o, name), getattr(r, name)
        assert a.tobytes() == b.tobytes(), name


def test_sf_range(ref):
    for sf in (5, 14):
        with pytest.raises(ValueError):
            ref.RefDecoder(sf=sf)
    ref.RefDecoder(sf=6)
    ref.RefDecoder(sf=13)


@pytest.mark.parametrize("sf", [7, 10, 12])
def test_instantaneous_frequency(oracle, ref, sf):
    """A3 :224-244: noise, a chirp, and samples with zero real or imaginary part (atan2 edge cases)."""
    o, r = oracle.Decoder(sf=sf), ref.RefDecoder(sf=sf)
    rng = np.random.default_rng(sf)
    n = 2 * o.sps
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    x[::97] = 0
    x[5::131] = x[5::131].real
    x[7::113] = 1j * x[7::113].imag
    for v in (x, tx.synth_symbols([3, (1 << sf) - 1], sf, snr_db=5.0, seed=1), x[:2], x[:3]):
        assert o.ifreq(v).tobytes() == r.ifreq(v).tobytes()


@pytest.mark.parametrize("sf", SFS)
def test_get_shift_fft(oracle, ref, sf):
    """A4 :430-464, the north-star K1: bins equal on clean, +0 dB and -12 dB symbols; magnitudes within fp32 FFT rounding."""
    o, r = oracle.Decoder(sf=sf), ref.RefDecoder(sf=sf)
    n = 24 if sf < 11 else 6
    for snr in (None, 0.0, -12.0):
        vals, x = _noisy_symbols(sf, n, snr, 100 * sf)
        vals[:3] = [0, (1 << sf) // 2, (1 << sf) - 1]
        x = tx.synth_symbols(vals, sf, snr_db=snr, seed=7)
        ob, om = o.demod_fft_batch(x)
        rb, rm = r.demod_fft_batch(x)
        assert np.array_equal(ob, rb)
        np.testing.assert_allclose(om, rm, rtol=2e-5)
        if snr is None or snr >= 0:
            assert np.array_equal(rb, vals)
    # the kept N bins of one symbol (bins 0..N/2-1 | sps-N/2..sps-1, plus the tmp[N/2] += F[N/2] quirk)
    spec = r.spectrum(x[: r.sps])
    mult = x[: r.sps].astype(np.complex128) * r.downchirp.astype(np.complex128)
    F = np.fft.fft(mult)
    N = r.n_bins
    want = np.concatenate([F[: N // 2], F[r.sps - N // 2:]])
    want[N // 2] += F[N // 2]
    np.testing.assert_allclose(spec, want, rtol=0, atol=2e-4 * np.abs(want).max())


@pytest.mark.parametrize("sf", SFS)
def test_gradient_demod_fine_sync_detectors(oracle, ref, sf):
    """A5 :466-491, A6 :300-338, A8-A11 :340-425 on the same inputs, bit-exact."""
    o, r = oracle.Decoder(sf=sf), ref.RefDecoder(sf=sf)
    sps, N = o.sps, o.n_bins
    vals, x = _noisy_symbols(sf, 8 if sf < 11 else 4, 25.0, 31 * sf)
    assert np.array_equal(o.demod_grad_batch(x), r.demod_grad_batch(x))
    for k, v in enumerate(vals):
        sym = x[k * sps:(k + 1) * sps]
        b = int((int(v) - 1) % N)
        for off in (0, 1, -1):                         # a late / early window makes the lag non-zero
            w = np.roll(sym, off)
            for bin_idx, space in ((b, 2), (b, max(o.decim // 4, 2)), (-1, 4 * o.decim)):
                if bin_idx == N - 1:
                    continue                            # reads past d_upchirp_ifreq_v in the reference (D1)
                assert o.fine_sync(w, bin_idx, space) == r.fine_sync(w, bin_idx, space)
    # detectors on a preamble: 2 up-chirps | up-chirp + down-chirp, plus noise-only windows
    up = tx.synth_symbols([0, 0, 0], sf, snr_db=30.0, seed=5)
    down = np.conj(tx.synth_symbols([0], sf)).astype(np.complex64)
    rng = np.random.default_rng(9)
    noise = (rng.standard_normal(2 * sps) + 1j * rng.standard_normal(2 * sps)).astype(np.complex64) * 0.05
    for w in (up[: 2 * sps], np.roll(up, 37)[: 2 * sps], noise, np.concatenate([up[:sps], down]) + noise):
        assert o.autocorr(w) == r.autocorr(w) or (np.isnan(o.autocorr(w)) and np.isnan(r.autocorr(w)))
        oc, oi = o.detect_upchirp(w)
        rc, ri = r.detect_upchirp(w)
        assert (oc, oi) == (rc, ri)
        assert o.detect_downchirp(w) == r.detect_downchirp(w)
        assert o.detect_downchirp(w[sps:]) == r.detect_downchirp(w[sps:])
        assert o.energy(w) == r.energy(w)


def test_integer_chain(oracle, ref):
    """B1 deinterleave :535-565, B2-B4 decode() :567-586 through the reference's member functions."""
    rng = np.random.default_rng(2024)
    r = ref.RefDecoder(sf=12)
    for sf in SFS:
        for ppm in (sf, sf - 2):
            for nw in (5, 6, 7, 8):
                words = rng.integers(0, 1 << ppm, nw, dtype=np.uint32)
                assert np.array_equal(oracle.deinterleave(words, ppm), r.deinterleave(words, ppm))
    for v in range(64):
        for c in range(0, 9):
            for size in (5, 8, 12):
                assert ref.rotl(v, c, size) == int(oracle.lib().lo_rotl(v, c, size))
    book = [ref.hamming_encode_soft(v) for v in range(16)]
    assert bytes(book).hex() == "00d25587994bcc1ee133b46678aa2dff"
    for v in range(256):
        assert ref.hamming_decode_soft_byte(v) == int(oracle.lib().lo_hamming_decode_soft_byte(v))
    # code-word vectors: clean and single-bit-error Hamming(8,4) words (pinned), all four coding rates, header and payload
    for trial in range(300):
        cr = 1 + trial % 4
        n = int(rng.integers(1, 60))
        nib = rng.integers(0, 16, n)
        cw = np.array([book[v] for v in nib], np.uint8)
        flip = rng.integers(0, 9, n)                    # 8 = no error
        cw = np.where(flip < 8, cw ^ (1 << np.minimum(flip, 7)).astype(np.uint8), cw).astype(np.uint8)
        for is_header in (False, True):
            if is_header and n < 5:
                continue
            assert oracle.decode_codewords(cw, is_header, cr) == r.decode_codewords(cw, is_header, cr), (trial, cr, is_header)
    # arbitrary bytes (>= 2 bit errors): depends on the Hamming table of the absent liquid-dsp; both sides take the
    # nearest code word, lowest symbol on ties -- agreement here is between two stand-ins, recorded, not a pin
    cw = rng.integers(0, 256, 64, dtype=np.uint8)
    for cr in (1, 2, 3, 4):
        assert oracle.decode_codewords(cw, False, cr) == r.decode_codewords(cw, False, cr)


def _assert_same_run(o, r, x, cr):
    oc, os_ = o.run(x)
    rc, rs = r.run(x)
    assert oc == rc and len(os_) == len(rs)
    for f in ("state", "consumed", "bin", "fine_sync"):
        assert np.array_equal(os_[f], rs[f]), f
    m = ~(np.isnan(os_["metric"]) & np.isnan(rs["metric"]))
    assert np.array_equal(os_["metric"][m], rs["metric"][m])
    of, rf = o.frames(), r.frames()
    assert of == rf
    so, sr = o.stdout, r.stdout
    if cr == 3:
        # header print with cr = 3: fec_decode produces ceil(6*4/7) = 4 bytes from 8 code words of which 6 exist
        # (lib/decoder_impl.cc:658-661); the reference decodes whatever the vector's spare capacity holds (stale words of
        # the previous payload), the restatement reads zeros (D3).  The 4th printed byte is excluded.
        def strip(s):
            out = []
            for ln in s.splitlines():
                if ln.startswith(" ") and len(ln) > 12:
                    ln = ln[:9] + " xx" + ln[12:]
                out.append(ln)
            return out
        assert strip(so) == strip(sr)
    else:
        assert so == sr
    return of


@pytest.mark.parametrize("case", FRAME_CASES, ids=[c[0] for c in FRAME_CASES])
def test_work_state_machine_on_golden_cases(oracle, ref, case):
    """A12 :740-903 + A7 + B5-B7: the 13 golden frame cases (explicit / implicit / reduced rate, CR1-4, SF7-12)."""
    x, fs, payload = make_case_iq(case)
    args = case_decoder_args(case)
    frames = _assert_same_run(oracle.Decoder(**args), ref.RefDecoder(**args), x, case[2])
    assert len(frames) == 2


def test_readme_golden_through_the_reference(ref):
    """README.md:77-85 through the reference's own code: banner and ' 04 90 40 de ad be ef 70 0d' x5."""
    r = ref.RefDecoder(sf=7, cr=4, crc=True)
    banner = "Bits (nominal) per symbol: \t3.5\nBins per symbol: \t128\nSamples per symbol: \t1024\nDecimation: \t\t8\n"
    assert r.stdout == banner
    fsy = tx.encode_frame(bytes.fromhex("deadbeef700d"), 7, 4)
    x = tx.channel([tx.modulate_frame(fsy, 7)] * 5, sf=7, snr_db=40.0, seed=0x4C6F5201, gap_symbols=97.66)
    r.run(x)
    lines = r.stdout[len(banner):].splitlines()
    assert len(lines) == 5 and all(ln.startswith(" 04 90 40 de ad be ef 70 0d") for ln in lines)
    assert [f[15:].hex() for f in r.frames()] == ["049040deadbeef700d"] * 5


SHORT = [("deadbeef", True), ("88", False), ("ffff", True)]


@pytest.mark.parametrize("sf", SFS)
def test_suite_short_matrix(oracle, ref, sf):
    """The reference's `short` suite shape (apps/generate_test_suites.py:199-201) on synthetic captures: SF x CR x payload."""
    for cr in ((1, 2, 3, 4) if sf <= 10 else (1, 4)):
        for k, (hexs, crc) in enumerate(SHORT):
            payload = bytes.fromhex(hexs) + (b"\x12\x34" if crc else b"")
            x = make_capture(payload, sf, cr, crc, seed=1000 * sf + 10 * cr + k)
            o = oracle.Decoder(sf=sf, cr=cr, crc=crc, reduced_rate=sf > 10)
            r = ref.RefDecoder(sf=sf, cr=cr, crc=crc, reduced_rate=sf > 10)
            assert len(_assert_same_run(o, r, x, cr)) == 1


@pytest.mark.parametrize("sf,ppm", [(7, 200.0), (7, -200.0), (9, 100.0), (11, -20.0)])
def test_clock_drift(oracle, ref, sf, ppm):
    x = make_capture(bytes(range(40)), sf, 4, False, seed=5 + sf, sfo_ppm=ppm)
    o, r = oracle.Decoder(sf=sf, cr=4, crc=False, reduced_rate=sf > 10), ref.RefDecoder(sf=sf, cr=4, crc=False, reduced_rate=sf > 10)
    _assert_same_run(o, r, x, 4)


@pytest.mark.parametrize("lead", [2.0, 2.13, 2.5, 2.999, 3.37])
def test_frame_offsets_cfo_noise(oracle, ref, lead):
    for cfo, snr in ((0.0, 38.0), (-600.0, 20.0), (900.0, 8.0)):
        x = make_capture(bytes.fromhex("0123456789abcdef"), 8, 2, True, seed=int(lead * 1000), lead=lead, cfo_hz=cfo, snr_db=snr)
        _assert_same_run(oracle.Decoder(sf=8, cr=2, crc=True), ref.RefDecoder(sf=8, cr=2, crc=True), x, 2)


def test_noise_only_and_silence(oracle, ref):
    rng = np.random.default_rng(1)
    noise = (rng.standard_normal(40 * 1024) + 1j * rng.standard_normal(40 * 1024)).astype(np.complex64)
    for x in (noise, np.zeros(8 * 1024, np.complex64), np.zeros(100, np.complex64)):
        _assert_same_run(oracle.Decoder(sf=7), ref.RefDecoder(sf=7), x, 4)
