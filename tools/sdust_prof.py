"""K0 timing: 100 windows of 500 kb, (a) low-complexity stretches planted every ~2 kb (a genome-like density) and (b) every ~150 b (stress: long segments).
Measurement tooling, not product: the sequences come from the test library and the oracle is the checker of the first window (and is timed beside it);
SDUST_QUICK=1 runs (a) only; LCD_SDUST_SPT sets the segment slots per replay thread."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import longcalld_b200 as lcd
import lcd_testlib as T
lcd.init(0, 0)
orc = T.oracle_lib()
import os
for lc_every in ((2000,) if os.environ.get("SDUST_QUICK") else (2000, 150)):
    rng = np.random.default_rng(95)
    tmpl = [T.sdust_sequence(rng, 500000, lc_every=lc_every) for _ in range(10)]
    seqs = [tmpl[i % 10] for i in range(100)]
    plan = lcd.SdustPlan(seqs, 5, 20)
    st = torch.cuda.ExternalStream(lcd.stream())
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); plan.run(); e1.record(st); plan.sync(); torch.cuda.synchronize()
        print(f"lc_every {lc_every}: sdust of 100 x 500 kb: {e0.elapsed_time(e1):.2f} ms", flush=True)
    out = plan.fetch()
    print(sum(len(x) for x in out), "intervals")
    t0 = time.perf_counter(); ref = T.sdust(orc, "lcd_oracle_sdust", tmpl[0]); dt = time.perf_counter() - t0
    assert np.array_equal(np.asarray(ref), np.asarray(out[0])), "differs from the oracle"
    print(f"oracle (one thread) on one 500 kb window: {1e3 * dt:.1f} ms")
