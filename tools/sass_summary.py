#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of liblora_b200.so, how many instructions of the families that show a
Blackwell-native kernel (TMA: UBLKCP / UTMALDG, mbarrier: SYNCS, tensor memory: LDTM / STTM, packed fp32: FFMA2 / FADD2 /
FMUL2) and that there is no tensor-core MMA (this path has no matrix product).  Run here (no GPU needed):
    python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
FAMILIES = ["UBLKCP", "UTMALDG", "SYNCS", "STAS", "LDTM", "STTM", "CREDUX", "FFMA2", "FADD2", "FMUL2", "FFMA", "LDS", "STS", "LDG", "SHFL", "BAR", "MEMBAR", "UTCALLOC", "UTCMMA", "HMMA"]
DEFAULT = ["k1_sf7_warp_kernelILi12ELi2E", "k1_group_kernelILi8ELi6ELi2E", "k1_group_kernelILi9ELi3ELi2E", "k1_sf10_kernelILi3E",
           "k1_rows_kernelILi11E", "k1_rows_kernelILi12E", "rx_warp_kernelILb1E", "rx_warp_kernelILb0E", "rx_stream_kernelILi8ELb1E", "k8_frames_kernel",
           "chan_fir_kernel", "sc16_to_cf32_kernel", "sc8_to_cf32_kernel", "k1_finalize_kernel"]


def main():
    lib = ROOT / "gr_lora_b200" / "liblora_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", txt))
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", ln)
        if m and cur:
            toks = m.group(1).split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0]
            kernels[cur]["_total"] += 1
            for f in FAMILIES:
                if op == f or (f == "HMMA" and op.startswith(f)) or (f == "UTCALLOC" and op in ("UTCALLOC", "UTCDEALLOC", "UTCATOMSWS")) \
                        or (f == "UTCMMA" and re.match(r"UTC\w*MMA", op)):
                    kernels[cur][f] += 1
    print("# SASS summary of gr_lora_b200/liblora_b200.so (cuobjdump -sass), architectures: " + ", ".join(sorted(arch)))
    print()
    print(f"{len(kernels)} kernels in the library; the ones on the default paths:")
    print()
    cols = ["_total"] + FAMILIES
    print("| kernel | " + " | ".join(c.strip("_") for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for name, c in kernels.items():
        if any(d in name for d in DEFAULT):
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0].replace("void lb::", "").replace("void ", "") or name
            print(f"| `{short}` | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print()
    print("Whole library: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in FAMILIES) + f"; instructions {tot['_total']}.")
    print("UTCMMA / HMMA = 0: no tensor-core matrix product (the north star asks for none); UBLKCP / UTMALDG = TMA bulk and "
          "tensor-map copies, SYNCS = mbarrier, STAS = st.async into the peer CTA (SF12 exchange), CREDUX = the warp argmax (redux.sync), LDTM / STTM = tcgen05.ld / tcgen05.st and UTCALLOC = tcgen05.alloc / dealloc (the dechirp table and twiddles of "
          "k1_rows live in tensor memory), FFMA2 / FADD2 / FMUL2 = packed fp32.")


if __name__ == "__main__":
    sys.exit(main())
