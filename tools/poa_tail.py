#!/usr/bin/env python3
"""Tail experiment for the POA plan: python tools/poa_tail.py MBP K1 K2 ... -- times the plan with the K largest problems routed to
the CTA-per-problem kernel (LCD_POA_CTA_TOP) and checks that consensus and statuses do not change."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import longcalld_b200 as lcd
from bench import Workload
import torch

lcd.init(0, 0)
wl = Workload(float(sys.argv[1]), "hifi", 11)
plan = lcd.PoaPlan(wl.seqs, wl.first, wl.n_reads, wl.read_off, wl.read_len, lcd.poa_params())
st = torch.cuda.ExternalStream(lcd.stream())
base = None
for k in [0] + [int(x) for x in sys.argv[2:]] + [0]:
    os.environ["LCD_POA_CTA_TOP"] = str(k)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); plan.run(); e1.record(st); plan.sync(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = plan.fetch(want_msa=False)
    sig = (res[0]["status"].tobytes(), res[0]["cons_len"].tobytes(), res[1].tobytes())
    if base is None: base = sig
    print(f"CTA_TOP={k}: {' '.join(f'{t:.1f}' for t in ts)} ms  same={sig == base}", flush=True)
