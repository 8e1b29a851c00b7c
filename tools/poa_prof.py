#!/usr/bin/env python3
"""Profiling driver for the POA plan: python tools/poa_prof.py MBP [MIN_LEN MAX_LEN [REPS]] -- runs the plan REPS times
over the problems of an MBP-megabase HiFi workload whose longest read is in (MIN_LEN, MAX_LEN]."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import longcalld_b200 as lcd
from longcalld_b200 import synth
from longcalld_b200.capi import pack_poa

mbp = float(sys.argv[1]); lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0; hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lcd.init(0, int(float(os.environ.get("POOL_GB", "0")) * (1 << 30)))
problems = []
for r in synth.make_regions(mbp, "hifi", seed=11, with_reads=True):
    for hap in (1, 2):
        reads = [s for s, h in zip(r.reads, r.read_hap) if h == hap and len(s) > 0]
        if reads and lo < max(len(s) for s in reads) <= hi:
            problems.append(reads)
plan = lcd.PoaPlan(*pack_poa(problems), lcd.poa_params())
import torch
st = torch.cuda.ExternalStream(lcd.stream())
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); plan.run(); e1.record(st); plan.sync(); torch.cuda.synchronize()
    print(f"{len(problems)} problems, {sum(len(p) for p in problems)} reads: {e0.elapsed_time(e1):.2f} ms")
print("cells", plan.work_units())
