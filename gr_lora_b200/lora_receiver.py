"""Host-side mirror of the reference's ``lora_receiver`` hier block (python/lora_receiver.py:26-89):
same constructor arguments, same wiring (optional conjugate -> channelizer -> decoder) and the
same 'frames' message port, without GNU Radio.  The decoder is the GPU path; the channelizer
(GNU Radio's freq_xlating_fir_filter_ccf inside lib/channelizer_impl.cc:40-60) is SURVEY.md 8f
row N1 ("next") and is not built yet, so only already-channelised input is accepted."""
from __future__ import annotations

import numpy as np

from .decoder import decoder


class lora_receiver:
    def __init__(self, samp_rate, center_freq, channel_list, bandwidth, sf, implicit, cr, crc, reduced_rate=False,
                 conj=False, decimation=1, disable_channelization=False, disable_drift_correction=False, **decoder_kw):
        self.samp_rate, self.center_freq, self.channel_list = samp_rate, center_freq, list(channel_list)
        self.bandwidth, self.sf, self.implicit, self.cr, self.crc = bandwidth, sf, implicit, cr, crc
        self.decimation, self.conj = decimation, conj
        self.disable_channelization = disable_channelization
        self.disable_drift_correction = disable_drift_correction
        needs_channelizer = not disable_channelization and (decimation != 1 or float(self.channel_list[0]) != float(center_freq))
        if needs_channelizer:
            raise NotImplementedError("channelizer (SURVEY.md 8f N1) is not built yet: feed channelised IQ "
                                      "(channel_list[0] == center_freq, decimation == 1) or disable_channelization=True")
        if disable_channelization and decimation != 1:
            raise NotImplementedError("fractional_resampler_cc path (python/lora_receiver.py:58-61) is host plumbing, not built")
        # python/lora_receiver.py:53
        self.decoder = decoder(samp_rate / decimation, bandwidth, sf, implicit, cr, crc, reduced_rate,
                               disable_drift_correction, **decoder_kw)
        self.frames = self.decoder.frames            # hier-block message port 'frames' (:56,:68)

    def message_port_subscribe(self, handler):
        self.decoder.message_port_subscribe(handler)

    def _front(self, samples):
        x = np.asarray(samples, dtype=np.complex64)
        return np.conj(x) if self.conj else x        # blocks.conjugate_cc, :50,:70-75

    def work(self, samples, stream=0):
        return self.decoder.work(self._front(samples), stream)

    def run(self, samples, stream=0):
        return self.decoder.run(self._front(samples), stream)

    def get_sf(self):
        return self.sf

    def set_sf(self, sf):                            # :80-82 (decoder warns: unsupported at run time)
        self.sf = sf
        self.decoder.set_sf(sf)

    def get_center_freq(self):
        return self.center_freq
