import sys, numpy as np
sys.path.insert(0,'.')
import longcalld_b200 as lcd
from bench import Workload
lcd.init(0,0)
wl = Workload(float(sys.argv[1]), 'hifi', 11)
plan = lcd.PoaPlan(wl.seqs, wl.first, wl.n_reads, wl.read_off, wl.read_len, lcd.poa_params())
for _ in range(3):
    plan.run(); plan.sync()
print('done', plan.work_units())
