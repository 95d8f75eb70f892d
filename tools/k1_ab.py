"""A/B harness for the K1 kernel of one SF (default, or LORA_B200_K1=generic / LORA_B200_K1_ROWS=0, or another build of the
library through LORA_B200_LIB): parity against the oracle on true symbols of that SF (ragged count, several grid passes,
edge bins, -3 dB), then device timing on a large resident batch.  Prints one JSON line.

    LORA_B200_K1_ROWS=0 python tools/k1_ab.py --sf 12
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def guard(torch, seconds, what):
    """Wait for the device; a launch that does not finish kills the process (a hung kernel must not burn GPU minutes)."""
    import time
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    t0 = time.time()
    while not ev.query():
        if time.time() - t0 > seconds:
            print(json.dumps({"hang": what}), flush=True)
            os._exit(3)
        time.sleep(0.01)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, required=True)
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--parity-symbols", type=int, default=0)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    import torch
    import gr_lora_b200 as G
    from gr_lora_b200 import tx
    from oracle import oracle as O

    sf = args.sf
    sps, nb = 8 << sf, 1 << sf
    dev = torch.device("cuda", 0)
    dec = G.decoder(1e6, 125000, sf, False, 4, True, demod="fft", quiet=True)
    out = {"sf": sf, "env": {k: v for k, v in os.environ.items() if k.startswith("LORA_B200_")}}
    stream = torch.cuda.current_stream()
    if not args.no_parity:
        n = args.parity_symbols or {7: 7111, 8: 3559, 9: 1783, 10: 907, 11: 607, 12: 461}.get(sf, 300)
        rng = np.random.default_rng(sf)
        vals = rng.integers(0, nb, n)
        vals[:6] = [0, 1, nb // 2 - 1, nb // 2, nb // 2 + 1, nb - 1]
        x = tx.synth_symbols(vals, sf, snr_db=-3.0, seed=11 + sf)
        iq = torch.from_numpy(x).to(dev)
        bins = torch.full((n,), -1, dtype=torch.int32, device=dev)
        mags = torch.zeros(n, dtype=torch.float32, device=dev)
        for _ in range(2):                                   # twice: the second launch reuses every buffer
            dec.demod_fft(iq, n, bins, mags, stream.cuda_stream)
            guard(torch, 10.0, "parity launch")
        ob, om = O.Decoder(sf=sf).demod_fft_batch(x)
        gb = bins.cpu().numpy().astype(np.uint32)
        out["parity"] = {"n": n, "bins_equal": bool(np.array_equal(gb, ob)), "n_diff": int(np.sum(gb != ob)),
                         "mags_close": bool(np.allclose(mags.cpu().numpy(), om, rtol=1e-4)),
                         "vs_tx": float(np.mean(gb == vals))}
    n2 = int(args.gib * (1 << 30)) // (8 * sps)
    iq2 = torch.randn((n2, sps, 2), dtype=torch.float32, device=dev)
    b2 = torch.empty(n2, dtype=torch.int32, device=dev)
    for _ in range(3):
        dec.demod_fft(iq2, n2, b2, None, stream.cuda_stream)
        guard(torch, 10.0, "timing warm-up")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.reps):
        dec.demod_fft(iq2, n2, b2, None, stream.cuda_stream)
    e1.record(stream)
    guard(torch, 20.0, "timed launches")
    ms = e0.elapsed_time(e1) / args.reps
    gbs = n2 * (64 * nb + 8) / (ms * 1e-3) / 1e9
    peak = 6570.0
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    out.update({"symbols": n2, "ms": ms, "symbols_per_s": n2 / (ms * 1e-3), "hbm_gbs": gbs, "frac": gbs / peak})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
