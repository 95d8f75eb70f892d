"""Host-side mirror of the reference's ``lora_receiver`` hier block (python/lora_receiver.py:26-89):
same constructor arguments, same wiring (optional conjugate -> channelizer -> decoder) and the
same 'frames' message port, without GNU Radio.  Decoder and channelizer both run
on the GPU; between them the IQ never leaves device memory."""
from __future__ import annotations

import numpy as np

from .decoder import decoder


class lora_receiver:
    def __init__(self, samp_rate, center_freq, channel_list, bandwidth, sf, implicit, cr, crc, reduced_rate=False,
                 conj=False, decimation=1, disable_channelization=False, disable_drift_correction=False, cfo_feedback=False,
                 **decoder_kw):
        self.samp_rate, self.center_freq, self.channel_list = samp_rate, center_freq, list(channel_list)
        self.bandwidth, self.sf, self.implicit, self.cr, self.crc = bandwidth, sf, implicit, cr, crc
        self.decimation, self.conj = decimation, conj
        self.disable_channelization = disable_channelization
        self.disable_drift_correction = disable_drift_correction
        self.channelizer = None
        if not disable_channelization:
            # python/lora_receiver.py:52: lora.channelizer(samp_rate, center_freq, channel_list, bandwidth, decimation)
            from .channelizer import channelizer
            self.channelizer = channelizer(samp_rate, center_freq, self.channel_list, bandwidth, decimation,
                                           device=decoder_kw.get("device", -1))
            if conj:                                 # channelizer -> conjugate_cc -> decoder (:62-63,70-75), conjugated on the device
                self.channelizer.set_conjugate(True)
        if disable_channelization and decimation != 1:
            raise NotImplementedError("fractional_resampler_cc path (python/lora_receiver.py:58-61) is host plumbing, not built")
        # python/lora_receiver.py:53
        self.decoder = decoder(samp_rate / decimation, bandwidth, sf, implicit, cr, crc, reduced_rate,
                               disable_drift_correction, **decoder_kw)
        if self.decoder.n_streams != 1:
            raise ValueError("lora_receiver feeds one channel (channel_list[0]) to one decoder stream, like the reference; "
                             "use decoder(..., n_streams=N).work_batch for many channels")
        self.frames = self.decoder.frames            # hier-block message port 'frames' (:56,:68)
        # decoder 'control' -> channelizer 'control' (:64): the ("cfo" . x) message the reference's decoder would publish at
        # SYNC (lib/decoder_impl.cc:774-776, commented out there) and lib/controller_impl.cc:52-57 turns into apply_cfo(x).
        # Opt-in here as well: with cfo_feedback the estimates made during a run() retune the channelizer afterwards.
        self.cfo_feedback = bool(cfo_feedback) and self.channelizer is not None
        self._cfo_seen = 0
        if self.cfo_feedback:
            self.decoder.set_cfo_estimate(True)

    def message_port_subscribe(self, handler):
        self.decoder.message_port_subscribe(handler)

    def _front(self, samples):
        x = np.asarray(samples, dtype=np.complex64)
        return np.conj(x) if self.conj else x        # blocks.conjugate_cc, :50,:70-75

    def work(self, samples, stream=0):
        if self.channelizer is None:
            return self.decoder.work(self._front(samples), stream)
        return self.run(samples, stream)

    def run(self, samples, stream=0):
        """Whole capture through (channelizer ->) [conj ->] decoder.  With the channelizer the filtered IQ
        stays in device memory: the decoder consumes the channelizer's output buffer directly.  Only
        channel_list[0] reaches the decoder, as in the reference (lib/channelizer_impl.cc:47,56-57)."""
        if self.channelizer is None:
            return self.decoder.run(self._front(samples), stream)
        x = np.asarray(samples, dtype=np.complex64)
        x = x[: (x.size // self.decimation) * self.decimation]
        limit = int(self.decoder.cfg.max_items_per_call or (1 << 20))
        pos_out, need = 0, 2 * self.decoder.sps
        # the channelizer is stateful (history, rotator), so the capture is filtered once into one device
        # buffer and the decoder walks over it call by call
        n_out = self.channelizer.work(x)
        ptr, stride = self.channelizer.output_ptr(0)
        while n_out - pos_out >= need:
            n = min(limit, n_out - pos_out)
            c = int(self.decoder.work_batch(ptr + 8 * pos_out, n_items=n, stride_items=n, host=0)[0])
            if c == 0:
                break
            pos_out += c
        if self.cfo_feedback:
            cfo, n = self.decoder.last_cfo(0)
            if n > self._cfo_seen:
                self._cfo_seen = n
                self.channelizer.apply_cfo(cfo)      # channelizer_impl::apply_cfo, lib/channelizer_impl.cc:68-71
        return pos_out * self.decimation

    def get_sf(self):
        return self.sf

    def set_sf(self, sf):                            # :80-82 (decoder warns: unsupported at run time)
        self.sf = sf
        self.decoder.set_sf(sf)

    def get_center_freq(self):
        return self.center_freq
