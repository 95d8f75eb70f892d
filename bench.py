#!/usr/bin/env python3
"""bench.py -- LoRa symbols/s through dechirp + FFT + argmax (K1), BASELINE.json's metric.

Workload (BASELINE.json configs[1]): batched synthetic SF7 BW125k at
1 MS/s, 4096 concurrent channels x 256 aligned symbols per channel = 1 048 576 symbols = 8 GiB of
cf32 per GPU per step, symbol values ~ U[0,128), AWGN +10 dB, generated on the device.

A "step" = one pass of K1 over the whole batch.  The input (8 GiB) is far larger than the
126 MB L2, so no L2 flush is needed between timed iterations.

  value      whole-job symbols/s with the batch resident in HBM (CUDA events on the launch stream)
  e2e        the same through the C ABI with HOST buffers: pinned H2D of the batch + D2H of the
             bins inside the timed region (lora_b200_demod_fft_host)
  roofline   algorithmic bytes (64*2^SF + 8 per symbol) / K1 launch time vs the
             measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle's get_shift_fft restatement ("port": the reference itself cannot be
             built here) on the host cores, bounded sample
  --impl reference   times only that CPU path (all host threads) and prints the same JSON shape

Multi-GPU (torchrun): streams are independent, so each rank owns its own 4096-channel batch
(weak scaling, no per-symbol collective); the chirp/twiddle tables are broadcast once from rank 0
with NCCL at init.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0x4C6F5202


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=int, default=7)
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--symbols-per-channel", type=int, default=256)
    ap.add_argument("--snr-db", type=float, default=10.0)
    ap.add_argument("--all-sf", action="store_true", default=True,
                    help="also report K1 for SF8..SF12 under per_sf (default on: the metric is quoted per SF)")
    ap.add_argument("--no-all-sf", dest="all_sf", action="store_false")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def algorithmic_bytes_per_symbol(sf: int) -> int:
    return 64 * (1 << sf) + 8          # read 8*2^SF cf32 once, write u32 bin + f32 magnitude


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle's get_shift_fft on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_fft_rate(sf: int, seconds: float, threads: int):
    from oracle import oracle as O
    from gr_lora_b200 import tx
    O.lib()
    n_bins = 1 << sf
    per_thread = max(8, min(4096, int(8e6 // (8 << sf))))     # symbols in each thread's private buffer (<= 64 MB)
    rng = np.random.default_rng(SEED)
    vals = rng.integers(0, n_bins, per_thread)
    x = tx.synth_symbols(vals, sf, snr_db=10.0, seed=SEED)
    decs = [O.Decoder(sf=sf) for _ in range(threads)]
    counts = [0] * threads
    ok = [True] * threads
    stop = time.perf_counter() + seconds

    def worker(i):
        d = decs[i]
        while time.perf_counter() < stop:
            bins, _ = d.demod_fft_batch(x)          # ctypes releases the GIL
            ok[i] = ok[i] and bool(np.mean(bins == vals) > 0.99)
            counts[i] += per_thread

    t0 = time.perf_counter()
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    total = sum(counts)
    return total / dt, total, dt, all(ok)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc = None
        self.index = index
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, ln in self.lines:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [v.strip() for v in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch(torch, sf, channels, n_sym, snr_db, device, seed):
    """[channels * n_sym, sps] cf32 on the device: chirp shift = value, unit amplitude, AWGN."""
    from gr_lora_b200 import tx
    n_bins, sps = 1 << sf, 8 << sf
    up = torch.from_numpy(tx.base_upchirp(sf).astype(np.complex64)).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    total = channels * n_sym
    vals = torch.randint(0, n_bins, (total,), generator=gen, device=device, dtype=torch.int64)
    iq = torch.empty((total, sps), dtype=torch.complex64, device=device)
    sigma = float(np.sqrt(10.0 ** (-snr_db / 10.0) / 2.0))
    ar = torch.arange(sps, device=device, dtype=torch.int64)
    chunk = max(1, (256 << 20) // (8 * sps))
    iqr = torch.view_as_real(iq)
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        idx = (ar[None, :] + vals[s:e, None] * 8) % sps
        iq[s:e] = up[idx]
        iqr[s:e].add_(torch.randn((e - s, sps, 2), generator=gen, device=device, dtype=torch.float32), alpha=sigma)
    return iq, vals


def workload_name(args, sf):
    """One string for both arms (the driver compares metric + config of the two JSON lines)."""
    return (f"batched synthetic SF{sf} BW125k, 1 MS/s IQ, {args.channels} concurrent channels x "
            f"{args.symbols_per_channel} symbols per GPU (BASELINE.json configs[1])")


K1_KERNEL = {7: "k1_sf7_warp_kernel<12,2>", 8: "k1_group_kernel<8,6,2>", 9: "k1_group_kernel<9,3,2>", 10: "k1_sf10_kernel<2>",
             11: "k1_cluster_kernel<11>", 12: "k1_xchg_kernel<12,256>"}


def run_reference(args):
    """--impl reference: the reference's CPU get_shift_fft path (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    rates = []
    total_syms = 0
    per_step = max(1.0, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r, n, dt, ok = cpu_fft_rate(args.sf, per_step, threads)
        if i >= args.warmup:
            rates.append(r)
            total_syms += n
    value = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": "LoRa symbols/s (dechirp+FFT+argmax)", "value": value, "unit": "symbols/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.sf), "sf": args.sf, "channels_per_gpu": args.channels,
                   "symbols_per_channel": args.symbols_per_channel, "snr_db": args.snr_db,
                   "parallelism": f"host threads x{threads} (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": "symbols/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step:.1f} s of get_shift_fft per step on {threads} threads "
                                   f"({total_syms} symbols timed), CPU {cpu_model()}; the reference cannot be built "
                                   f"here (GNU Radio/VOLK/liquid-dsp absent), so this is the oracle restatement"},
        "e2e": {"value": value, "unit": "symbols/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import gr_lora_b200 as G

    sf = args.sf
    n_sym_total = args.channels * args.symbols_per_channel
    sps = 8 << sf
    dec = G.decoder(1e6, 125000, sf, False, 4, True, n_streams=1, demod="fft", device=local, quiet=True)

    # ---- init-time table broadcast (the only collective on this path) -------------------------
    if world > 1:
        from gr_lora_b200 import sharding
        sharding.broadcast_tables(dec, dist, device=device, src=0)

    iq, vals = synth_batch(torch, sf, args.channels, args.symbols_per_channel, args.snr_db, device, SEED + rank)
    bins = torch.empty(n_sym_total, dtype=torch.int32, device=device)
    mags = torch.empty(n_sym_total, dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream()

    def step():
        dec.demod_fft(iq, n_sym_total, bins, mags, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    acc = float((bins.to(torch.int64) == vals).float().mean().item())
    bins_ref = bins.clone()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = dec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    launches = dec.launch_count() - l0
    ms = e0.elapsed_time(e1)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = ms_max / args.steps
    value = world * n_sym_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (K1 is the only kernel in the step at SF<=10) --------
    peak, peak_src = measured_peak_gbs()
    abytes = n_sym_total * algorithmic_bytes_per_symbol(sf) + 8 * sps      # + the chirp table once per launch
    k1_ms = ms / args.steps                                                # this rank's own launches
    achieved = abytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    tr = ROOT / "profiles" / "k1_traffic.json"
    if tr.exists():
        try:
            traffic = json.loads(tr.read_text()).get(f"sf{sf}")
        except Exception:
            traffic = None

    # ---- per-SF table (extra keys): same bytes per batch, fewer symbols ------------------------
    per_sf = {}
    if args.all_sf and rank == 0:
        for s2 in range(7, 13):
            if s2 == sf:
                per_sf[str(s2)] = {"symbols_per_s": n_sym_total / (k1_ms * 1e-3), "hbm_gbs": achieved, "frac": achieved / peak}
                continue
            d2 = G.decoder(1e6, 125000, s2, False, 4, True, demod="fft", device=local, quiet=True)
            n2 = n_sym_total >> (s2 - 7) if s2 >= 7 else n_sym_total
            n2 = max(1, n2)
            iq2 = iq.view(-1)[: n2 * (8 << s2)]
            b2 = bins[:n2]
            for _ in range(3):
                d2.demod_fft(iq2, n2, b2, None, stream.cuda_stream)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            reps = 5
            for _ in range(reps):
                d2.demod_fft(iq2, n2, b2, None, stream.cuda_stream)
            a1.record(stream)
            torch.cuda.synchronize()
            m2 = a0.elapsed_time(a1) / reps
            gb = n2 * algorithmic_bytes_per_symbol(s2) / (m2 * 1e-3) / 1e9
            per_sf[str(s2)] = {"symbols_per_s": n2 / (m2 * 1e-3), "hbm_gbs": gb, "frac": gb / peak,
                               "note": "timing only (buffer holds SF7 symbols)"}
            d2.close()

    # ---- e2e through the C ABI with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_iq = h_bins = h_mags = None
        err = ""
        try:
            h_iq = torch.empty((n_sym_total, sps), dtype=torch.complex64, pin_memory=True)
            h_iq.copy_(iq)
            h_bins = torch.empty(n_sym_total, dtype=torch.int32, pin_memory=True)
            h_mags = torch.empty(n_sym_total, dtype=torch.float32, pin_memory=True)
            torch.cuda.synchronize()
        except Exception as exc:     # e.g. not enough pinnable host memory on the box
            err = str(exc)[:200]
            h_iq = None
        ok_t = torch.tensor([1 if h_iq is not None else 0], dtype=torch.int32, device=device)
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)      # all ranks take the same branch (no barrier mismatch)
        if int(ok_t.item()) == 1:
            e_steps = max(3, min(args.steps, 8))
            call = lambda: dec.demod_fft_host((h_iq.data_ptr(), n_sym_total), h_bins.numpy().view(np.uint32), h_mags.numpy())
            for _ in range(2):
                call()
            if world > 1:
                dist.barrier()
            ta = time.perf_counter()
            for _ in range(e_steps):
                call()
            torch.cuda.synchronize()
            tb = time.perf_counter()
            dt_t = torch.tensor([tb - ta], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
            e2e = {"value": world * n_sym_total * e_steps / float(dt_t.item()), "unit": "symbols/s",
                   "h2d_bytes_per_step": int(n_sym_total * sps * 8), "d2h_bytes_per_step": int(n_sym_total * 8),
                   "steps": e_steps, "timer": "host wall clock around lora_b200_demod_fft_host (pinned host buffers), max over ranks",
                   "bins_match_device_path": bool(torch.equal(h_bins.to(device), bins_ref))}
        else:
            e2e = {"value": None, "unit": "symbols/s", "error": err or "pinned host allocation failed on another rank"}
        del h_iq

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        r1, n1, d1, ok1 = cpu_fft_rate(sf, args.cpu_seconds / 2, 1)
        thr = host_threads()
        rN, nN, dN, okN = cpu_fft_rate(sf, args.cpu_seconds / 2, thr)
        cpu = {"value": rN, "unit": "symbols/s", "cores": thr, "kind": "port", "single_thread": r1,
               "sample": f"oracle get_shift_fft restatement (reference not buildable here): {n1} symbols on 1 thread in "
                         f"{d1:.1f} s, {nN} symbols on {thr} threads in {dN:.1f} s; SF{sf}, +10 dB; CPU {cpu_model()}",
               "bins_correct": bool(ok1 and okN)}

    if rank == 0:
        out = {
            "metric": "LoRa symbols/s (dechirp+FFT+argmax)", "value": value, "unit": "symbols/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, sf),
                       "sf": sf, "channels_per_gpu": args.channels, "symbols_per_channel": args.symbols_per_channel,
                       "snr_db": args.snr_db, "batch_bytes_per_gpu": int(n_sym_total * sps * 8),
                       "l2": "inputs (8 GiB) larger than L2, no flush needed", "parallelism": f"streams sharded x{world}",
                       "demod_accuracy_vs_tx": acc},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": K1_KERNEL.get(sf, "?") if os.environ.get("LORA_B200_K1", "w12x2") == "w12x2" else f"k1_fft_kernel<{sf}>",
                         "algorithmic_bytes_per_launch": int(abytes)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": cpu,
        }
        if per_sf:
            out["per_sf"] = per_sf
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
