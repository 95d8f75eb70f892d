"""ctypes binding of the CPU oracle (oracle/liblora_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liblora_oracle.so"


def build(force: bool = False) -> Path:
    src_m = max((HERE / f).stat().st_mtime for f in ("lora_oracle.c", "lora_oracle.h", "lo_tables.h"))
    if force or not LIB.exists() or LIB.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return LIB


class Step(C.Structure):
    _fields_ = [("state", C.c_int32), ("consumed", C.c_int32), ("bin", C.c_int32),
                ("fine_sync", C.c_int32), ("metric", C.c_float)]


STEP_DTYPE = np.dtype([("state", "<i4"), ("consumed", "<i4"), ("bin", "<i4"), ("fine_sync", "<i4"), ("metric", "<f4")])
STATES = ["DETECT", "SYNC", "FIND_SFD", "PAUSE", "DECODE_HEADER", "DECODE_PAYLOAD", "STOP"]
DEMOD_GRADIENT, DEMOD_FFT = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        vp, u32, i32, f32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_float, C.c_size_t
        L.lo_create.restype = vp
        L.lo_create.argtypes = [f32, u32, C.c_uint8, C.c_int, C.c_uint8, C.c_int, C.c_int, C.c_int]
        L.lo_destroy.argtypes = [vp]
        L.lo_set_demod.argtypes = [vp, C.c_int]
        for n in ("lo_sps", "lo_bins", "lo_decim"):
            getattr(L, n).restype = u32
            getattr(L, n).argtypes = [vp]
        L.lo_bits_per_symbol.restype = C.c_double
        L.lo_bits_per_symbol.argtypes = [vp]
        for n in ("lo_downchirp", "lo_upchirp", "lo_downchirp_ifreq", "lo_upchirp_ifreq", "lo_upchirp_ifreq_v"):
            getattr(L, n).restype = vp
            getattr(L, n).argtypes = [vp]
        L.lo_instantaneous_frequency.argtypes = [vp, vp, u32]
        L.lo_get_shift_fft.restype = u32
        L.lo_get_shift_fft.argtypes = [vp, vp, vp]
        L.lo_max_frequency_gradient_idx.restype = u32
        L.lo_max_frequency_gradient_idx.argtypes = [vp, vp]
        L.lo_fine_sync.restype = i32
        L.lo_fine_sync.argtypes = [vp, vp, i32, i32]
        L.lo_detect_preamble_autocorr.restype = f32
        L.lo_detect_preamble_autocorr.argtypes = [vp, vp]
        L.lo_detect_upchirp.restype = f32
        L.lo_detect_upchirp.argtypes = [vp, vp, vp]
        L.lo_detect_downchirp.restype = f32
        L.lo_detect_downchirp.argtypes = [vp, vp]
        L.lo_determine_energy.restype = f32
        L.lo_determine_energy.argtypes = [vp, vp]
        L.lo_demod_fft_batch.argtypes = [vp, vp, sz, vp, vp]
        L.lo_demod_grad_batch.argtypes = [vp, vp, sz, vp]
        L.lo_work.restype = C.c_int
        L.lo_work.argtypes = [vp, vp, vp]
        L.lo_run.restype = sz
        L.lo_run.argtypes = [vp, vp, sz, vp, sz, vp]
        L.lo_state.restype = C.c_int
        L.lo_state.argtypes = [vp]
        L.lo_frame_count.restype = sz
        L.lo_frame_count.argtypes = [vp]
        L.lo_frame_len.restype = sz
        L.lo_frame_len.argtypes = [vp, sz]
        L.lo_frame_data.restype = vp
        L.lo_frame_data.argtypes = [vp, sz]
        L.lo_frames_clear.argtypes = [vp]
        L.lo_stdout.restype = C.c_char_p
        L.lo_stdout.argtypes = [vp]
        L.lo_rotl.restype = u32
        L.lo_rotl.argtypes = [u32, u32, u32]
        L.lo_gray.restype = u32
        L.lo_gray.argtypes = [u32]
        L.lo_reduce_bin.restype = u32
        L.lo_reduce_bin.argtypes = [u32, u32]
        L.lo_deinterleave_words.argtypes = [vp, u32, u32, vp]
        for n in ("lo_deshuffle_byte", "lo_hamming84_encode", "lo_hamming84_decode", "lo_hamming_decode_soft_byte"):
            getattr(L, n).restype = C.c_uint8
            getattr(L, n).argtypes = [C.c_uint8]
        L.lo_decode_codewords.restype = sz
        L.lo_decode_codewords.argtypes = [vp, sz, C.c_int, C.c_uint8, vp, sz, vp]
        L.lo_payload_symbols.restype = i32
        L.lo_payload_symbols.argtypes = [u32, C.c_uint8, C.c_uint8, C.c_int]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Decoder:
    """Oracle decoder instance: same constructor arguments as lora::decoder::make
    (include/lora/decoder.h:705)."""

    def __init__(self, samp_rate=1e6, bandwidth=125000, sf=7, implicit=False, cr=4, crc=True,
                 reduced_rate=False, disable_drift_correction=False, demod=DEMOD_GRADIENT):
        self.L = lib()
        self.h = self.L.lo_create(samp_rate, bandwidth, sf, int(implicit), cr, int(crc), int(reduced_rate),
                                  int(disable_drift_correction))
        if not self.h:
            raise ValueError("spreading factor should be between 6 and 12")
        self.L.lo_set_demod(self.h, demod)
        self.sps = self.L.lo_sps(self.h)
        self.n_bins = self.L.lo_bins(self.h)
        self.decim = self.L.lo_decim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.lo_destroy(self.h)
            self.h = None

    def _table(self, name, n, dtype):
        p = getattr(self.L, name)(self.h)
        nbytes = n * np.dtype(dtype).itemsize
        return np.frombuffer(C.string_at(p, nbytes), dtype=dtype).copy()

    @property
    def downchirp(self):
        return self._table("lo_downchirp", self.sps, np.complex64)

    @property
    def upchirp(self):
        return self._table("lo_upchirp", self.sps, np.complex64)

    @property
    def downchirp_ifreq(self):
        return self._table("lo_downchirp_ifreq", self.sps, np.float32)

    @property
    def upchirp_ifreq(self):
        return self._table("lo_upchirp_ifreq", self.sps, np.float32)

    @property
    def upchirp_ifreq_v(self):
        return self._table("lo_upchirp_ifreq_v", 3 * self.sps, np.float32)

    @staticmethod
    def _iq(x):
        return np.ascontiguousarray(x, dtype=np.complex64)

    def ifreq(self, x):
        x = self._iq(x)
        out = np.empty(x.size, np.float32)
        self.L.lo_instantaneous_frequency(_ptr(x), _ptr(out), x.size)
        return out

    def get_shift_fft(self, x):
        x = self._iq(x)
        assert x.size >= self.sps
        mag = C.c_float()
        b = self.L.lo_get_shift_fft(self.h, _ptr(x), C.addressof(mag))
        return int(b), float(mag.value)

    def grad_idx(self, x):
        x = self._iq(x)
        assert x.size >= self.sps
        return int(self.L.lo_max_frequency_gradient_idx(self.h, _ptr(x)))

    def fine_sync(self, x, bin_idx, search_space):
        x = self._iq(x)
        assert x.size >= self.sps
        return int(self.L.lo_fine_sync(self.h, _ptr(x), bin_idx, search_space))

    def autocorr(self, x):
        x = self._iq(x)
        assert x.size >= 2 * self.sps
        return float(self.L.lo_detect_preamble_autocorr(self.h, _ptr(x)))

    def detect_upchirp(self, x):
        x = self._iq(x)
        assert x.size >= 2 * self.sps
        idx = C.c_int32(0)
        c = self.L.lo_detect_upchirp(self.h, _ptr(x), C.addressof(idx))
        return float(c), int(idx.value)

    def detect_downchirp(self, x):
        x = self._iq(x)
        assert x.size >= self.sps
        return float(self.L.lo_detect_downchirp(self.h, _ptr(x)))

    def energy(self, x):
        x = self._iq(x)
        return float(self.L.lo_determine_energy(self.h, _ptr(x)))

    def demod_fft_batch(self, iq):
        iq = self._iq(iq)
        n = iq.size // self.sps
        bins = np.empty(n, np.uint32)
        mags = np.empty(n, np.float32)
        self.L.lo_demod_fft_batch(self.h, _ptr(iq), n, _ptr(bins), _ptr(mags))
        return bins, mags

    def demod_grad_batch(self, iq):
        iq = self._iq(iq)
        n = iq.size // self.sps
        bins = np.empty(n, np.uint32)
        self.L.lo_demod_grad_batch(self.h, _ptr(iq), n, _ptr(bins))
        return bins

    def run(self, iq, max_steps=1 << 18):
        """Fake scheduler over a whole capture; returns (consumed, steps ndarray)."""
        iq = self._iq(iq)
        steps = np.zeros(max_steps, STEP_DTYPE)
        n = C.c_size_t(0)
        consumed = self.L.lo_run(self.h, _ptr(iq), iq.size, _ptr(steps), max_steps, C.addressof(n))
        return int(consumed), steps[: min(n.value, max_steps)]

    def work(self, iq):
        iq = self._iq(iq)
        assert iq.size >= 2 * self.sps
        st = Step()
        c = self.L.lo_work(self.h, _ptr(iq), C.addressof(st))
        return int(c), st

    @property
    def state(self):
        return int(self.L.lo_state(self.h))

    def frames(self, clear=True):
        out = []
        for i in range(self.L.lo_frame_count(self.h)):
            n = self.L.lo_frame_len(self.h, i)
            p = self.L.lo_frame_data(self.h, i)
            out.append(bytes(C.string_at(p, n)))
        if clear:
            self.L.lo_frames_clear(self.h)
        return out

    @property
    def stdout(self):
        return self.L.lo_stdout(self.h).decode()


def decode_codewords(codewords, is_header, cr):
    L = lib()
    cw = np.ascontiguousarray(codewords, dtype=np.uint8)
    out = np.zeros(1024, np.uint8)
    consumed = C.c_size_t(0)
    n = L.lo_decode_codewords(_ptr(cw), cw.size, int(is_header), cr, _ptr(out), out.size, C.addressof(consumed))
    return bytes(out[:n]), int(consumed.value)


def deinterleave(words, ppm):
    L = lib()
    w = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.zeros(ppm, np.uint8)
    L.lo_deinterleave_words(_ptr(w), w.size, ppm, _ptr(out))
    return out
