"""Host-side mirror of the reference's decoder block over the C ABI.

``decoder`` keeps the name, constructor arguments, ``set_sf``/``set_samp_rate`` behaviour,
stdout side effects and message-port semantics of ``gr::lora::decoder``
(include/lora/decoder.h:693-709, lib/decoder_impl.cc:41-122, 740-915) so that tests read like
the reference's own; the GNU Radio scheduler is replaced by ``run()`` (a fake scheduler
honouring ``set_output_multiple(2*sps)`` and the consume protocol).  All arithmetic happens in
liblora_b200.so on the GPU: there is no Python/NumPy compute path here.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _native as N
from .loraphy import parse_frame

LORATAP_LEN = 15   # sizeof(loratap_header_t), include/lora/loratap.h:48-55
LORAPHY_LEN = 3    # sizeof(loraphy_header_t), include/lora/loraphy.h:25-32


def _dev_ptr(x) -> int:
    """Device pointer of a torch tensor / anything with data_ptr(), or a raw int."""
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


class decoder:
    """lora.decoder(samp_rate, bandwidth, sf, implicit, cr, crc, reduced_rate, disable_drift_correction)

    Extra keyword arguments select what the reference cannot express: ``n_streams`` independent
    (channel, SF) streams sharing the configuration, the demodulator (``"gradient"`` = the
    reference's live path, ``"fft"`` = the north-star dechirp+FFT path) and the CUDA device.
    """

    def __init__(self, samp_rate, bandwidth, sf, implicit, cr, crc, reduced_rate=False,
                 disable_drift_correction=False, *, n_streams=1, demod="gradient", device=-1,
                 max_items_per_call=0, max_frames_per_call=0, trace_capacity=0, quiet=False):
        self._L = N.lib()
        self._h = None
        demod_id = {"gradient": N.DEMOD_GRADIENT, "fft": N.DEMOD_FFT}[demod] if isinstance(demod, str) else int(demod)
        cfg = N.Config(samp_rate=float(samp_rate), bandwidth=int(bandwidth), sf=int(sf), implicit=int(bool(implicit)),
                       cr=int(cr), crc=int(bool(crc)), reduced_rate=int(bool(reduced_rate)),
                       disable_drift_correction=int(bool(disable_drift_correction)), demod=demod_id,
                       n_streams=int(n_streams), device=int(device), max_items_per_call=int(max_items_per_call),
                       max_frames_per_call=int(max_frames_per_call), trace_capacity=int(trace_capacity))
        self.cfg = cfg
        h = self._L.lora_b200_create(C.byref(cfg))
        if not h:
            msg = self._L.lora_b200_last_error().decode(errors="replace")
            if sf < 6 or sf > 13:
                # the reference prints this to std::cerr and exit(1)s (lib/decoder_impl.cc:57-61)
                print(msg, file=sys.stderr)
                raise SystemExit(1)
            raise RuntimeError("lora_b200_create failed: " + msg)
        self._h = h
        self.n_streams = int(n_streams)
        self.sps = self._L.lora_b200_samples_per_symbol(h)
        self.n_bins = self._L.lora_b200_bins(h)
        self.decim = self._L.lora_b200_decimation(h)
        self.quiet = quiet
        self.frames = []            # what was published on message port "frames": (stream, bytes)
        self.implicit = bool(implicit)
        self.header_checks = {"ok": 0, "bad": 0}   # explicit-header checksum of the published frames (loraphy.py; the
                                                   # reference never checks it, include/lora/utilities.h:396-404)
        self._handlers = []
        self._cb = N.FRAME_CB(self._on_frame)
        buf = C.create_string_buffer(512)
        self._L.lora_b200_banner(h, buf, len(buf))
        self.banner = buf.value.decode()
        if not quiet:
            sys.stdout.write(self.banner)     # lib/decoder_impl.cc:93-103

    # -- gr::lora::decoder::make -------------------------------------------------------------
    @classmethod
    def make(cls, samp_rate, bandwidth, sf, implicit, cr, crc, reduced_rate, disable_drift_correction, **kw):
        return cls(samp_rate, bandwidth, sf, implicit, cr, crc, reduced_rate, disable_drift_correction, **kw)

    def close(self):
        if self._h:
            self._L.lora_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- unsupported setters, lib/decoder_impl.cc:905-915 --------------------------------------
    def set_sf(self, sf):
        self._L.lora_b200_set_sf(self._h, int(sf))
        print(self._L.lora_b200_last_error().decode(), file=sys.stderr)

    def set_samp_rate(self, samp_rate):
        self._L.lora_b200_set_samp_rate(self._h, float(samp_rate))
        print(self._L.lora_b200_last_error().decode(), file=sys.stderr)

    # -- message port "frames" -----------------------------------------------------------------
    def message_port_subscribe(self, handler):
        """handler(stream, frame_bytes) is called for every published frame."""
        self._handlers.append(handler)

    def _on_frame(self, _user, stream, data, length):
        blob = bytes(C.string_at(data, length))
        self.frames.append((int(stream), blob))
        if not self.implicit and len(blob) >= 18:     # side channel only: the published bytes are never touched
            self.header_checks["ok" if parse_frame(blob).header_ok else "bad"] += 1
        for h in self._handlers:
            h(int(stream), blob)

    def output_multiple(self):
        return 2 * self.sps          # set_output_multiple, lib/decoder_impl.cc:91

    # -- work(): one scheduler call -------------------------------------------------------------
    def work(self, input_items, stream=0):
        """Hand a host buffer of gr_complex to the GPU state machine; returns items consumed.
        Frames completed inside the call are published before it returns."""
        x = np.ascontiguousarray(input_items, dtype=np.complex64)
        consumed = C.c_size_t(0)
        N.check(self._L.lora_b200_work(self._h, int(stream), x.ctypes.data, x.size, C.byref(consumed), self._cb, None),
                "lora_b200_work")
        self._emit_stdout(stream)
        return int(consumed.value)

    FRAME_DTYPE = np.dtype([("stream", "<u4"), ("seq", "<u4"), ("len", "<u4"), ("n_hdr_print", "u1"), ("hdr_print", "u1", (4,)),
                            ("pad", "u1", (3,)), ("bytes", "u1", (564,))])

    def frames_last(self) -> np.ndarray:
        """Structured array (FRAME_DTYPE) of the frames the last work call published, in delivery order: the bulk
        alternative to one Python callback per frame (lora_b200_frames_last)."""
        ptr = C.c_void_p(0)
        n = int(self._L.lora_b200_frames_last(self._h, C.byref(ptr)))
        if n == 0:
            return np.zeros(0, self.FRAME_DTYPE)
        buf = C.string_at(ptr.value, n * self.FRAME_DTYPE.itemsize)
        return np.frombuffer(buf, dtype=self.FRAME_DTYPE)

    def work_batch(self, iq, n_items=None, stride_items=None, host=None, sc16_scale=None, callbacks=True, sc8_scale=None):
        """All streams at once. ``iq``: host ndarray [n_streams, n_items] (complex64, or int16 / int8 [n_streams, n_items, 2]
        together with ``sc16_scale`` / ``sc8_scale``) or a device tensor/pointer.  ``sc16_scale`` / ``sc8_scale`` select the
        int16 / int8 I/Q entry points: the device computes x * scale, PCIe moves 4 / 2 bytes per sample."""
        if isinstance(iq, np.ndarray) and (sc16_scale is not None or sc8_scale is not None):
            x = np.ascontiguousarray(iq, dtype=np.int16 if sc16_scale is not None else np.int8)
            assert x.ndim == 3 and x.shape[0] == self.n_streams and x.shape[2] == 2
            ptr, n_items, stride_items, host = x.ctypes.data, x.shape[1], x.shape[1], 1
        elif isinstance(iq, np.ndarray):
            x = np.ascontiguousarray(iq, dtype=np.complex64)
            assert x.ndim == 2 and x.shape[0] == self.n_streams
            ptr, n_items, stride_items, host = x.ctypes.data, x.shape[1], x.shape[1], 1
        else:
            ptr = _dev_ptr(iq)                                    # device tensor, or a raw host / device address with host=1 / 0
            host = 0 if host is None else int(host)
            if stride_items is None:
                stride_items = n_items
        consumed = np.zeros(self.n_streams, dtype=np.uint64)      # size_t[n_streams]
        cptr = consumed.ctypes.data_as(C.POINTER(C.c_size_t))
        cb = self._cb if callbacks else N.FRAME_CB()              # callbacks=False: drain with frames_last() instead
        if sc16_scale is not None:
            N.check(self._L.lora_b200_work_batch_sc16(self._h, ptr, float(sc16_scale), int(n_items), int(stride_items), host,
                                                      cptr, cb, None), "lora_b200_work_batch_sc16")
        elif sc8_scale is not None:
            N.check(self._L.lora_b200_work_batch_sc8(self._h, ptr, float(sc8_scale), int(n_items), int(stride_items), host,
                                                     cptr, cb, None), "lora_b200_work_batch_sc8")
        else:
            N.check(self._L.lora_b200_work_batch(self._h, ptr, int(n_items), int(stride_items), host, cptr, cb, None),
                    "lora_b200_work_batch")
        if not self.quiet:
            for s in range(self.n_streams):
                self._emit_stdout(s)
        return consumed.astype(np.int64)

    def _emit_stdout(self, stream):
        if self.quiet:
            return
        buf = C.create_string_buffer(1 << 16)
        n = self._L.lora_b200_stdout_last(self._h, int(stream), buf, len(buf))
        if n > 0:
            sys.stdout.write(buf.value.decode(errors="replace"))   # lib/decoder_impl.cc:832,872

    def run(self, samples, stream=0, chunk_items=None):
        """Fake GNU Radio scheduler over a whole capture (host array): repeatedly calls work()
        with at least 2*sps items and drops what work() consumed.  Returns total consumed."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        limit = chunk_items or (self.cfg.max_items_per_call or (1 << 20))
        pos, need = 0, 2 * self.sps
        while x.size - pos >= need:
            c = self.work(x[pos:pos + limit], stream)
            if c == 0:
                break
            pos += c
        return pos

    def set_cfo_estimate(self, enable=True):
        """Also run the reference's experimental_determine_cfo at every SYNC (lib/decoder_impl.cc:730-738,774); off by default."""
        N.check(self._L.lora_b200_set_cfo_estimate(self._h, int(bool(enable))), "lora_b200_set_cfo_estimate")

    def last_cfo(self, stream=0):
        """(latest CFO estimate in Hz, number of estimates so far) of a stream."""
        cfo, n = C.c_float(0.0), C.c_uint32(0)
        N.check(self._L.lora_b200_last_cfo(self._h, int(stream), C.byref(cfo), C.byref(n)), "lora_b200_last_cfo")
        return float(cfo.value), int(n.value)

    def reset(self):
        """Every stream back to a freshly made block's state (a flowgraph restart); buffers and tables are kept."""
        N.check(self._L.lora_b200_reset(self._h), "lora_b200_reset")
        self.frames = []

    def state(self, stream=0):
        return N.check(self._L.lora_b200_stream_state(self._h, int(stream)), "lora_b200_stream_state")

    def trace(self, stream=0):
        cap = max(int(self.cfg.trace_capacity), 1)
        steps = (N.Step * cap)()
        n = C.c_size_t(0)
        rc = self._L.lora_b200_trace_read(self._h, int(stream), steps, cap, C.byref(n))
        if rc not in (N.OK, N.EOVERFLOW):
            N.check(rc, "lora_b200_trace_read")
        m = min(int(n.value), cap)
        return [(steps[i].state, steps[i].consumed, steps[i].bin, steps[i].fine_sync, steps[i].metric) for i in range(m)]

    # -- batch kernels (device-resident) ---------------------------------------------------------
    def demod_fft(self, iq_dev, n_symbols, bins_dev, mags_dev=None, cuda_stream=0):
        """K1 on device memory: dechirp + FFT + argmax of n_symbols aligned windows."""
        N.check(self._L.lora_b200_demod_fft_dev(self._h, _dev_ptr(iq_dev), int(n_symbols), _dev_ptr(bins_dev),
                                               _dev_ptr(mags_dev), int(cuda_stream)), "lora_b200_demod_fft_dev")

    def demod_fft_host(self, iq_host, bins_out=None, mags_out=None):
        """K1 end to end from host memory (copies inside). iq_host: complex64 ndarray or (ptr, n_symbols)."""
        if isinstance(iq_host, tuple):
            ptr, n = int(iq_host[0]), int(iq_host[1])
        else:
            x = np.ascontiguousarray(iq_host, dtype=np.complex64)
            ptr, n = x.ctypes.data, x.size // self.sps
        bins = np.empty(n, np.uint32) if bins_out is None else bins_out
        mags = np.empty(n, np.float32) if mags_out is None else mags_out
        bp = bins.ctypes.data if isinstance(bins, np.ndarray) else int(bins)
        mp = mags.ctypes.data if isinstance(mags, np.ndarray) else int(mags)
        N.check(self._L.lora_b200_demod_fft_host(self._h, ptr, n, bp, mp), "lora_b200_demod_fft_host")
        return bins, mags

    def demod_fft_host_sc16(self, iq_sc16, scale, bins_out=None, mags_out=None):
        """K1 end to end from int16 I/Q host memory. iq_sc16: int16 ndarray [..., 2] or (ptr, n_symbols)."""
        if isinstance(iq_sc16, tuple):
            ptr, n = int(iq_sc16[0]), int(iq_sc16[1])
        else:
            x = np.ascontiguousarray(iq_sc16, dtype=np.int16)
            ptr, n = x.ctypes.data, x.size // (2 * self.sps)
        bins = np.empty(n, np.uint32) if bins_out is None else bins_out
        mags = np.empty(n, np.float32) if mags_out is None else mags_out
        bp = bins.ctypes.data if isinstance(bins, np.ndarray) else int(bins)
        mp = mags.ctypes.data if isinstance(mags, np.ndarray) else int(mags)
        N.check(self._L.lora_b200_demod_fft_host_sc16(self._h, ptr, float(scale), n, bp, mp), "lora_b200_demod_fft_host_sc16")
        return bins, mags

    def tx_symbols(self, values_dev, out_dev, n_symbols, noise_sigma=0.0, seed=0, cfo_hz_dev=None, up_table_dev=None, cuda_stream=0):
        """Synthetic aligned data symbols on the device (chirp shift = value, optional per-symbol CFO and AWGN)."""
        N.check(self._L.lora_b200_tx_symbols_dev(self._h, _dev_ptr(up_table_dev), _dev_ptr(values_dev), _dev_ptr(cfo_hz_dev), float(noise_sigma), int(seed),
                                                int(n_symbols), _dev_ptr(out_dev), int(cuda_stream)), "lora_b200_tx_symbols_dev")

    def tx_expand(self, base_dev, k, n_items, n_streams, out_dev, noise_sigma=0.0, seed=0, cuda_stream=0):
        """n_streams channels from k base captures plus every stream's own AWGN, on the device."""
        N.check(self._L.lora_b200_tx_expand_dev(self._h, _dev_ptr(base_dev), int(k), int(n_items), float(noise_sigma), int(seed),
                                               int(n_streams), _dev_ptr(out_dev), int(cuda_stream)), "lora_b200_tx_expand_dev")

    def ifreq(self, iq_dev, n_windows, window, out_dev, cuda_stream=0):
        """A3 instantaneous_frequency of n_windows windows of `window` samples (device tensors)."""
        N.check(self._L.lora_b200_ifreq_dev(self._h, _dev_ptr(iq_dev), int(n_windows), int(window), _dev_ptr(out_dev),
                                           int(cuda_stream)), "lora_b200_ifreq_dev")

    def demod_gradient(self, iq_dev, n_symbols, bins_dev, cuda_stream=0):
        N.check(self._L.lora_b200_demod_gradient_dev(self._h, _dev_ptr(iq_dev), int(n_symbols), _dev_ptr(bins_dev),
                                                    int(cuda_stream)), "lora_b200_demod_gradient_dev")

    def decode_codewords(self, cw_dev, lengths_dev, stride, cr_dev, is_header_dev, n_vec, out_dev, out_stride,
                         out_len_dev, cuda_stream=0):
        N.check(self._L.lora_b200_decode_codewords_dev(self._h, _dev_ptr(cw_dev), _dev_ptr(lengths_dev), int(stride),
                                                      _dev_ptr(cr_dev), _dev_ptr(is_header_dev), int(n_vec),
                                                      _dev_ptr(out_dev), int(out_stride), _dev_ptr(out_len_dev),
                                                      int(cuda_stream)), "lora_b200_decode_codewords_dev")

    def deinterleave(self, words_dev, n_words, ppm, n_blocks, cw_dev, cuda_stream=0):
        N.check(self._L.lora_b200_deinterleave_dev(self._h, _dev_ptr(words_dev), int(n_words), int(ppm), int(n_blocks),
                                                  _dev_ptr(cw_dev), int(cuda_stream)), "lora_b200_deinterleave_dev")

    # -- tables ------------------------------------------------------------------------------------
    def tables_bytes(self):
        return int(self._L.lora_b200_tables_bytes(self._h))

    def tables_export(self) -> np.ndarray:
        out = np.empty(self.tables_bytes(), np.uint8)
        N.check(self._L.lora_b200_tables_export(self._h, out.ctypes.data, out.size), "lora_b200_tables_export")
        return out

    def tables_import(self, blob):
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        N.check(self._L.lora_b200_tables_import(self._h, b.ctypes.data, b.size), "lora_b200_tables_import")

    def tables_device_view(self):
        """Zero-copy view of the device table blob as an object with __cuda_array_interface__
        (wrap with torch.as_tensor(view, device="cuda") for the init-time NCCL broadcast)."""
        ptr, n = int(self._L.lora_b200_tables_device_ptr(self._h)), self.tables_bytes()

        class _View:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return _View()

    def tables_commit(self):
        N.check(self._L.lora_b200_tables_commit(self._h), "lora_b200_tables_commit")

    def launch_count(self):
        return int(self._L.lora_b200_launch_count(self._h))


def tables_build_host(samp_rate=1e6, bandwidth=125000, sf=7) -> np.ndarray:
    """The chirp/ifreq/twiddle blob built on the host only (no GPU needed)."""
    L = N.lib()
    cfg = N.Config(samp_rate=float(samp_rate), bandwidth=int(bandwidth), sf=int(sf), n_streams=1)
    n = L.lora_b200_tables_build_host(C.byref(cfg), None, 0)
    if not n:
        raise RuntimeError(L.lora_b200_last_error().decode())
    out = np.empty(n, np.uint8)
    L.lora_b200_tables_build_host(C.byref(cfg), out.ctypes.data, out.size)
    return out


def split_tables(blob: np.ndarray, sps: int) -> dict:
    """Views into the blob: layout documented in include/lora_b200.h."""
    o, out = 0, {}
    for name, dt, n in (("downchirp", np.complex64, sps), ("upchirp", np.complex64, sps),
                        ("downchirp_ifreq", np.float32, sps), ("upchirp_ifreq", np.float32, sps),
                        ("upchirp_ifreq_v", np.float32, 3 * sps), ("twiddles", np.complex64, sps)):
        nb = np.dtype(dt).itemsize * n
        out[name] = blob[o:o + nb].view(dt)
        o += nb
    return out


def dissect_frame(blob: bytes):
    """(loratap, phy header, payload) of a published frame (lib/decoder_impl.cc:588-601)."""
    return blob[:LORATAP_LEN], blob[LORATAP_LEN:LORATAP_LEN + LORAPHY_LEN], blob[LORATAP_LEN + LORAPHY_LEN:]
