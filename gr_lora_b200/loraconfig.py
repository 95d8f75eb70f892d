"""LoRa link description used by the apps and the test harness (same fields and string
forms as the reference's python/loraconfig.py:1-30: cr is given as "4/x", cr_num = x - 4)."""


class LoRaConfig:
    def __init__(self, freq, sf, cr, bw=125e3, prlen=8, crc=True, implicit=False):
        self.freq, self.sf, self.cr, self.bw = freq, sf, cr, bw
        self.prlen, self.crc, self.implicit = prlen, crc, implicit
        self.cr_num = int(str(cr).rpartition("/")[2]) - 4

    def file_repr(self):
        s = "{:n}-sf{:n}-cr{:n}-bw{:n}".format(self.freq / 1e6, self.sf, self.cr_num, self.bw / 1e3)
        return s + ("-crc" if self.crc else "") + ("-imp" if self.implicit else "")

    def string_repr(self):
        return "{:n} MHz, SF {:n}, CR {:s}, BW {:n} kHz, prlen {:n}, crc {:s}, implicit {:s}".format(
            self.freq / 1e6, self.sf, self.cr, self.bw / 1e3, self.prlen,
            "on" if self.crc else "off", "on" if self.implicit else "off")
