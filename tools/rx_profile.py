#!/usr/bin/env python3
"""Device-resident run of the stream state machine on frame-bearing streams (what bench.py's e2e feeds through host
buffers), for ncu / timing: n_streams streams x `windows` symbol times of SF `sf`, FFT demodulator.  One JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=int, default=7)
    ap.add_argument("--streams", type=int, default=2048)
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--demod", default="fft")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    import gr_lora_b200 as G
    dev = torch.device("cuda", 0)
    sps = 8 << args.sf
    n_items = args.windows * sps
    caps, pays = zip(*[bench.frame_stream(args.sf, n_items, 0x4C6F5201 + k, payload_len=12) for k in range(16)])
    x = bench.expand_streams(torch, list(caps), args.streams, 35.0, dev, 1)
    out = {"sf": args.sf, "streams": args.streams, "windows": args.windows, "demod": args.demod}
    times = []
    for rep in range(args.reps):
        dec = G.decoder(1e6, 125000, args.sf, False, 4, False, args.sf > 10, False, n_streams=args.streams, demod=args.demod, quiet=True,
                        max_items_per_call=n_items, max_frames_per_call=max(len(p) for p in pays) + 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        consumed = dec.work_batch(x, n_items=n_items, stride_items=n_items, host=0, callbacks=False)
        fr = dec.frames_last()
        times.append(time.perf_counter() - t0)
        exp, ok = bench.check_frames(fr, list(pays), 16, args.streams)
        dec.close()
    t = min(times[1:]) if len(times) > 1 else times[0]
    out.update({"s": t, "windows_per_s": float(consumed.sum()) / sps / t, "samples_per_s": float(consumed.sum()) / t,
                "hbm_gbs_algorithmic": float(consumed.sum()) * 8 / t / 1e9, "frames_expected": exp, "frames_ok": ok})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
