import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import longcalld_b200 as lcd
from longcalld_b200 import synth
lcd.init(0, 0)
regions = synth.make_regions(50, "hifi", seed=11, with_reads=False) if False else synth.make_regions(50, "hifi", seed=11, with_reads=True)
pairs = synth.wfa_problems(regions)
from longcalld_b200.capi import pack_pairs
plan = lcd.WfaPlan(*pack_pairs(pairs), lcd.wfa_params())
st = torch.cuda.ExternalStream(lcd.stream())
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); plan.run(); e1.record(st); plan.sync(); torch.cuda.synchronize()
    print(f"{len(pairs)} WFA problems: {e0.elapsed_time(e1):.2f} ms", flush=True)
