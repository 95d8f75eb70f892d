#!/usr/bin/env python3
"""Per-source-line totals of an ncu --set full --import-source on capture: instructions executed and stall samples.
usage: python tools/ncu_lines.py x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    fname = "?"
    hdr = None
    rows = []
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].rsplit("/", 1)[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0]:
            d = dict(zip(hdr, r))
            rows.append((fname, int(r[0]), r[1].strip(), float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0)))
    ti = sum(x[3] for x in rows) or 1
    ts = sum(x[4] for x in rows) or 1
    print(f"total warp instructions {ti:.4g}, samples {ts:.0f}")
    print("-- by file")
    agg = {}
    for f, _, _, i, s in rows:
        a = agg.setdefault(f, [0, 0]); a[0] += i; a[1] += s
    for f, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{f:24s} inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%")
    print("-- top lines by samples")
    for f, ln, src, i, s in sorted(rows, key=lambda x: -x[4])[:top]:
        print(f"{f}:{ln:<5d} inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%  {src[:110]}")


if __name__ == "__main__":
    main()
