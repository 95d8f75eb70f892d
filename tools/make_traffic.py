#!/usr/bin/env python3
"""profiles/k1_traffic.json and profiles/r2_k1_sf<N>.txt from the ncu captures of tools/capture_profiles.sh
(gpurun_out/r2_k1_sf<N>.ncu-rep): DRAM bytes per launch of the shipped K1 kernel of every SF against the algorithmic
bytes of that launch (64 * 2^SF + 8 per symbol + the chirp table once)."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    out = {"source": "ncu --set full --clock-control none, tools/k1_ab.py --sf N --gib 8 (tools/capture_profiles.sh); dram__bytes_read.sum + "
                     "dram__bytes_write.sum of one launch of the shipped kernel"}
    for sf in range(7, 13):
        rep = ROOT / "gpurun_out" / f"r2_k1_sf{sf}.ncu-rep"
        if not rep.exists():
            continue
        n = (8 << 30) // (64 << sf)
        bps = 64 * (1 << sf) + 8
        txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep), str(n), str(bps)], capture_output=True, text=True).stdout
        (ROOT / "profiles" / f"r2_k1_sf{sf}.txt").write_text(txt)
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, r = rows[0], rows[1], rows[2]
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        tr = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
        alg = n * bps + 64 * (1 << sf)
        out[f"sf{sf}"] = {"kernel": d.get("Kernel Name"), "symbols_per_launch": n, "dram_bytes_per_launch": tr, "algorithmic_bytes_per_launch": alg,
                          "ratio": tr / alg}
    (ROOT / "profiles" / "k1_traffic.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({k: (v["ratio"] if isinstance(v, dict) else v) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
