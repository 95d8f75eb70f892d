#!/usr/bin/env python3
"""Throughput of the whole GPU receive path (state machine kernel + K8) on a config-4-like load:
64 channels x SF7..SF12, every stream 1 MS/s post-channelizer IQ with LoRa frames in it, all six
decoders running concurrently (one host thread + one CUDA stream per SF).

Algorithmic bytes: every IQ sample is needed at least once = 8 B/sample.  Prints one JSON object.
    python tools/bench_rx_path.py [--channels 64] [--seconds 1.0] [--demod gradient|fft]
"""
import argparse
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def capture(sf, n_items, seed, tx):
    """One stream: back-to-back frames with 16-byte payloads, CR4/8, gaps of 6 symbols."""
    rng = np.random.default_rng(seed)
    frames, total, sps = [], 0, 8 << sf
    while True:
        payload = bytes(rng.integers(0, 256, 16, dtype=np.uint8))
        f = tx.modulate_frame(tx.encode_frame(payload, sf, 4, has_crc=False, reduced_rate=sf > 10), sf,
                              sync_word=0x78 if sf >= 11 else 0x12)
        if total + f.size + 6 * sps > n_items - 8 * sps:
            break
        frames.append(f)
        total += f.size + 6 * sps
    x = tx.channel(frames, sf=sf, snr_db=38.0, seed=seed, gap_symbols=6.0, lead_symbols=3.0, tail_symbols=3.0)
    out = np.zeros(n_items, np.complex64)
    out[: min(n_items, x.size)] = x[:n_items]
    return out, len(frames)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--demod", default="gradient")
    ap.add_argument("--sfs", default="7,8,9,10,11,12")
    args = ap.parse_args()
    import torch
    import gr_lora_b200 as G
    from gr_lora_b200 import tx
    dev = torch.device("cuda", 0)
    n_items = int(args.seconds * 1e6)
    sfs = [int(s) for s in args.sfs.split(",")]
    decs, bufs, expected = {}, {}, {}
    for sf in sfs:
        x, nf = capture(sf, n_items, 0x4C6F5204 + sf, tx)
        t = torch.from_numpy(x).to(dev)
        bufs[sf] = t.unsqueeze(0).repeat(args.channels, 1).contiguous()
        expected[sf] = nf
        decs[sf] = G.decoder(1e6, 125000, sf, False, 4, False, sf > 10, False, n_streams=args.channels, demod=args.demod,
                             quiet=True, max_items_per_call=n_items, max_frames_per_call=max(8, nf + 2))
    torch.cuda.synchronize()
    res = {}

    def run(sf):
        t0 = time.perf_counter()
        consumed = decs[sf].work_batch(bufs[sf], n_items=n_items, stride_items=n_items, host=0)
        res[sf] = (time.perf_counter() - t0, int(consumed.min()), len(decs[sf].frames))

    for sf in sfs:                     # per-SF timings, one decoder at a time
        run(sf)
    solo = dict(res)
    for sf in sfs:                     # reset streams for the concurrent run
        decs[sf].close()
        decs[sf] = G.decoder(1e6, 125000, sf, False, 4, False, sf > 10, False, n_streams=args.channels, demod=args.demod,
                             quiet=True, max_items_per_call=n_items, max_frames_per_call=max(8, expected[sf] + 2))
    torch.cuda.synchronize()
    ths = [threading.Thread(target=run, args=(sf,)) for sf in sfs]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    wall = time.perf_counter() - t0
    n_streams = len(sfs) * args.channels
    samples = n_streams * n_items
    out = {
        "workload": f"{args.channels} channels x SF{sfs[0]}..SF{sfs[-1]} = {n_streams} streams x {n_items} samples, demod={args.demod}",
        "concurrent": {"wall_s": wall, "msamples_per_s": samples / wall / 1e6, "hbm_gbs_algorithmic": samples * 8 / wall / 1e9,
                       "realtime_streams_supported": samples / wall / 1e6},
        "per_sf_solo": {str(sf): {"s": solo[sf][0], "msamples_per_s": args.channels * n_items / solo[sf][0] / 1e6,
                                   "symbols_per_s": args.channels * n_items / (8 << sf) / solo[sf][0],
                                   "frames_per_stream": solo[sf][2] / args.channels, "frames_expected": expected[sf],
                                   "consumed_min": solo[sf][1]} for sf in sfs},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
