#!/usr/bin/env python3
"""bench.py -- throughput of the B200 re-alignment hot path on synthetic 30x long-read noisy regions.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mbp M] [--tech hifi|ont]

A *step* is one pass of the hot path over one batch: all noisy regions of `--mbp` megabases of
reference per GPU (default 50 Mb: BASELINE.json configs[1], "synthetic HiFi 30x, 50 Mb ref, 1 GPU").
The metric is reference megabases called per second.  `value` is the device-resident number (inputs in
HBM, CUDA events on the library stream); `e2e` goes through the host-buffer C-ABI batch calls with
H2D/D2H inside the timed region.  One JSON line on stdout (rank 0).

`--impl reference` times the UNMODIFIED reference libraries (oracle/_ref/libref_shim.so: WFA2-lib /
edlib / abPOA compiled from /root/reference) on the same workload with all host threads, each step a
bounded sample of the batch.  That is the only place this file touches oracle/ (as the baseline, never
as the product path).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ref_Mbp_per_s_called"
UNIT = "Mbp/s"
WFA_BYTES_PER_CELL = 48          # 5 components written + 7 read, int32


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ workload
def build_workload(mbp, tech, seed):
    from longcalld_b200 import synth
    from longcalld_b200.capi import pack_pairs
    regions = synth.make_regions(mbp, tech, seed=seed, with_reads=False)
    pairs = synth.wfa_problems(regions)
    seqs, po, pl, to, tl = pack_pairs(pairs)
    return {"n_regions": len(regions), "wfa": (seqs, po, pl, to, tl), "n_wfa": len(pairs)}


# ------------------------------------------------------------------------------------------ reference arm
def ref_shim():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            return None
    return C.CDLL(path)


def reference_step(lib, wl, sample_idx, n_threads):
    """The reference's CPU implementation of the same stages on the problems in sample_idx."""
    from longcalld_b200.capi import WFA_PARAMS_DTYPE, WFA_RESULT_DTYPE, wfa_params
    seqs, po, pl, to, tl = wl["wfa"]
    idx = np.asarray(sample_idx, dtype=np.int64)
    n = len(idx)
    spo, spl, sto, stl = (np.ascontiguousarray(a[idx]) for a in (po, pl, to, tl))
    par = np.zeros(n, dtype=WFA_PARAMS_DTYPE)
    par[:] = wfa_params()
    cap = 2 * (spl.astype(np.int64) + stl) + 8
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cap, out=off[1:])
    ops = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
    res = np.zeros(n, dtype=WFA_RESULT_DTYPE)
    t0 = time.perf_counter()
    lib.ref_wfa_batch(C.c_int(n), seqs.ctypes.data_as(C.c_void_p), spo.ctypes.data_as(C.c_void_p),
                      spl.ctypes.data_as(C.c_void_p), sto.ctypes.data_as(C.c_void_p), stl.ctypes.data_as(C.c_void_p),
                      par.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                      res.ctypes.data_as(C.c_void_p), C.c_int(n_threads))
    return time.perf_counter() - t0


def region_sample(wl, frac, seed=1):
    """Whole regions (both haplotype problems), uniformly sampled: a bounded slice of the batch."""
    rng = np.random.default_rng(seed)
    n_reg = wl["n_regions"]
    k = max(1, int(round(n_reg * frac)))
    regs = np.sort(rng.choice(n_reg, size=k, replace=False))
    return regs, np.stack([2 * regs, 2 * regs + 1], axis=1).ravel()


def calibrate_sample(lib, wl, n_threads, target_s):
    """Pick the sample fraction so one reference step costs about target_s seconds."""
    frac = min(1.0, 200.0 / wl["n_regions"])
    regs, idx = region_sample(wl, frac)
    dt = reference_step(lib, wl, idx, n_threads)
    per_region = dt / len(regs)
    return float(min(1.0, max(frac, target_s / max(per_region, 1e-9) / wl["n_regions"])))


def run_reference(args, rank):
    if rank != 0:
        return
    lib = ref_shim()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_shim.so missing and /root/reference absent"}))
        return
    n_threads = os.cpu_count() or 1
    wl = build_workload(args.mbp, args.tech, args.seed)
    frac = calibrate_sample(lib, wl, n_threads, target_s=max(2.0, 60.0 / max(1, args.steps + args.warmup)))
    regs, idx = region_sample(wl, frac)
    mbp_sample = args.mbp * len(regs) / wl["n_regions"]
    for _ in range(args.warmup):
        reference_step(lib, wl, idx, n_threads)
    t = [reference_step(lib, wl, idx, n_threads) for _ in range(args.steps)]
    total = sum(t)
    value = mbp_sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, wl),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "reference",
                             "sample": f"{len(regs)} of {wl['n_regions']} regions ({mbp_sample:.3f} Mb) per step"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, wl):
    return {"workload": f"synthetic {args.tech.upper()} 30x noisy-region re-alignment, {args.mbp:g} Mb ref per GPU "
                        f"(BASELINE configs[1] shape), {wl['n_regions']} regions",
            "stages": ["K6 WFA gap-affine-2p ref-vs-consensus (align.c:565)"],
            "stages_not_yet_on_gpu": ["K5 abPOA consensus", "K7 edlib", "K1-K4 pileup/phasing"],
            "n_wfa": wl["n_wfa"], "l2": "flushed between timed steps (256 MiB write)", "seed": args.seed}


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    import longcalld_b200 as lcd
    from longcalld_b200.capi import WFA_PARAMS_DTYPE
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lcd.init(local, 0)
    stream = torch.cuda.ExternalStream(lcd.stream(), device=local)
    wl = build_workload(args.mbp, args.tech, args.seed + rank)      # weak scaling: one 50 Mb shard per GPU
    seqs, po, pl, to, tl = wl["wfa"]
    par = lcd.wfa_params()
    plan = lcd.WfaPlan(seqs, po, pl, to, tl, par)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            plan.run()
            e1.record(stream)
        return e0, e1

    for _ in range(args.warmup):
        device_step()
    barrier()
    launches0 = lcd.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [device_step() for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()
    launches = lcd.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    cells = plan.work_units()
    res, _, _ = plan.fetch(want_ops=False)
    assert (res["status"] == 0).all()

    # ---- e2e: host buffers -> lcd_wfa_batch -> host results, copies inside the timed region
    par_arr = np.zeros(len(pl), dtype=WFA_PARAMS_DTYPE)
    par_arr[:] = par
    cap = 2 * (pl.astype(np.int64) + tl) + 8
    off = np.zeros(len(pl) + 1, dtype=np.int64)
    np.cumsum(cap, out=off[1:])
    ops = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
    out = np.zeros(len(pl), dtype=lcd.capi.WFA_RESULT_DTYPE)
    L = lcd.lib()

    def e2e_step():
        rc = L.lcd_wfa_batch(C.c_int(len(pl)), seqs.ctypes.data_as(C.c_void_p), C.c_size_t(seqs.size),
                             po.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p),
                             to.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p),
                             par_arr.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p),
                             off.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(L.lcd_gpu_last_error().decode())

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(seqs.size + 64 * len(pl))
    d2h = int(out.nbytes + (pl.astype(np.int64) + tl + 8).sum())

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = t.tolist()
    total_mbp = args.mbp * world
    value = total_mbp * args.steps / (dev_ms_max / 1e3)
    e2e_value = total_mbp * args.steps / (e2e_ms_max / 1e3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lib = ref_shim()
        if lib is not None:
            nt = os.cpu_count() or 1
            frac = calibrate_sample(lib, wl, nt, target_s=15.0)
            regs, idx = region_sample(wl, frac)
            dt = reference_step(lib, wl, idx, nt)
            mbp_sample = args.mbp * len(regs) / wl["n_regions"]
            cpu_baseline = {"value": mbp_sample / dt, "unit": UNIT, "cores": nt, "kind": "reference",
                            "sample": f"{len(regs)} of {wl['n_regions']} regions ({mbp_sample:.3f} Mb), WFA2-lib via oracle/_ref"}
    if rank == 0:
        peak, which = load_peaks()
        kernel_ms = dev_ms / args.steps
        achieved = cells * WFA_BYTES_PER_CELL / (kernel_ms / 1e3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": workload_config(args, wl),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": "wfa_kernel<32>+wfa_kernel<256> (concurrent)",
                             "algorithmic": f"{cells} wavefront cells x {WFA_BYTES_PER_CELL} B", "peak_source": which},
                "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mbp", type=float, default=50.0, help="reference megabases per GPU per step")
    ap.add_argument("--tech", default="hifi", choices=["hifi", "ont"])
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
