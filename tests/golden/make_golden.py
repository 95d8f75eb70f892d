#!/usr/bin/env python3
"""Regenerate tests/golden/golden.json.

The reference ships no IQ fixtures, so the golden vectors are: (1) the README
console golden (README.md:62-71) and (2) outputs recorded in THIS container on deterministic
synthetic captures by the CPU oracle (oracle/lora_oracle.c, the restatement of
lib/decoder_impl.cc) AND, where /root/reference is present, by the reference's own
decoder_impl.cc compiled against stand-in headers (oracle/_ref, oracle/ref.py): generation
fails unless both produce the same states, consume amounts, bins, frames and K1 bins, and the
fixture then carries "pinned_by_reference": true.  Tests on any machine regenerate the same
captures from the seeds and must reproduce these outputs with the oracle and with the GPU.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from conftest import FRAME_CASES, make_case_iq, case_decoder_args  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402
from gr_lora_b200 import tx  # noqa: E402


def k1_case(sf, n_sym, snr_db, seed):
    rng = np.random.default_rng(seed)
    vals = rng.integers(0, 1 << sf, n_sym)
    vals[:5] = [0, (1 << sf) // 2, (1 << sf) - 1, (1 << sf) // 2 - 1, (1 << sf) // 2 + 1]
    x = tx.synth_symbols(vals, sf, snr_db=snr_db, seed=seed + 1)
    return vals, x


def main():
    out = {"readme": {"banner": "Bits (nominal) per symbol: \t3.5\nBins per symbol: \t128\nSamples per symbol: \t1024\nDecimation: \t\t8\n",
                      "line": " 04 90 40 de ad be ef 70 0d", "source": "README.md:77-85 of the reference"},
           "frames": {}, "k1": {}}
    have_ref = R.available()
    out["pinned_by_reference"] = bool(have_ref)
    out["reference_sources_sha256"] = (R.HERE / "_ref" / "SOURCES.sha256").read_text().split("\n") if have_ref and R.build() else []
    for case in FRAME_CASES:
        name = case[0]
        x, fs, payload = make_case_iq(case)
        d = O.Decoder(**case_decoder_args(case))
        consumed, steps = d.run(x)
        frames = d.frames()
        if have_ref:      # the reference's own work() on the same capture (gradient demodulator, its live path)
            r = R.RefDecoder(**case_decoder_args(case))
            rc, rsteps = r.run(x)
            assert rc == consumed and r.frames() == frames, name
            for f in ("state", "consumed", "bin", "fine_sync"):
                assert np.array_equal(steps[f], rsteps[f]), (name, f)
        out["frames"][name] = {
            "iq_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
            "n_items": int(x.size),
            "shifts": [int(s) for s in fs.shifts],
            "consumed": consumed,
            "n_steps": int(len(steps)),
            "states": "".join(str(int(s)) for s in steps["state"]),
            "consumes": [int(c) for c in steps["consumed"]],
            "bins": [int(b) for b in steps["bin"] if b >= 0],
            "frames": [f.hex() for f in frames],
            "stdout": d.stdout,
        }
        print(name, len(frames), [f[15:].hex() for f in frames])
    for sf in range(7, 13):
        n = 24 if sf < 11 else 8
        vals, x = k1_case(sf, n, 0.0, 1000 + sf)
        d = O.Decoder(sf=sf)
        fb, fm = d.demod_fft_batch(x)
        if have_ref:      # the reference's get_shift_fft (lib/decoder_impl.cc:430-464) on the same symbols
            rb, rm = R.RefDecoder(sf=sf).demod_fft_batch(x)
            assert np.array_equal(rb, fb) and np.allclose(rm, fm, rtol=2e-5), sf
        gb = d.demod_grad_batch(tx.synth_symbols(vals, sf))          # gradient demod needs a clean input
        out["k1"][str(sf)] = {"n": n, "snr_db": 0.0, "seed": 1000 + sf, "values": [int(v) for v in vals],
                              "fft_bins": [int(b) for b in fb], "fft_mags": [float(m) for m in fm],
                              "grad_bins_clean": [int(b) for b in gb]}
    (HERE / "golden.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
