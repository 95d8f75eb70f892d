#!/usr/bin/env python3
"""Where the time of bench.py's e2e goes: the H2D copy of the batch alone, the state machine on a device-resident batch,
and lora_b200_work_batch on the pinned host batch (copy and state machine overlapped in stream groups).  One JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    import bench
    import gr_lora_b200 as G
    dev = torch.device("cuda", 0)
    sf, sps = 7, 1024
    n_items = args.windows * sps
    caps, pays = zip(*[bench.frame_stream(sf, n_items, 0x4C6F5201 + k, payload_len=12) for k in range(16)])
    x = bench.expand_streams(torch, list(caps), args.streams, 35.0, dev, 1)
    xh = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    xh.copy_(x)
    torch.cuda.synchronize()
    y = torch.empty_like(x)

    def best(fn):
        ts = []
        for _ in range(args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return min(ts[1:])

    out = {"streams": args.streams, "windows": args.windows, "bytes": x.numel() * x.element_size()}
    out["h2d_s"] = best(lambda: y.copy_(xh, non_blocking=True))
    out["h2d_gbs"] = out["bytes"] / out["h2d_s"] / 1e9
    dec = G.decoder(1e6, 125000, sf, False, 4, False, False, False, n_streams=args.streams, demod="fft", quiet=True,
                    max_items_per_call=n_items, max_frames_per_call=max(len(p) for p in pays) + 2)

    def run(buf, host):
        dec.reset()
        dec.work_batch(buf, n_items=n_items, stride_items=n_items, host=host, callbacks=False)

    out["device_resident_s"] = best(lambda: run(x, 0))
    out["host_s"] = best(lambda: run(xh, 1))
    fr = dec.frames_last()
    exp, ok = bench.check_frames(fr, list(pays), 16, args.streams)
    out.update({"frames_expected": exp, "frames_ok": ok, "host_minus_copy_s": out["host_s"] - out["h2d_s"]})
    dec.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
