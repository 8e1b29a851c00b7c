"""How much do partially covering reads cost the POA kernel?  The same problems with and without their sub-graph anchors (scratch probe)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import longcalld_b200 as lcd
import lcd_testlib as T
lcd.init(0, 0)
rng = np.random.default_rng(5)
cases = list(T.partial_cover_problems(float(sys.argv[1]) if len(sys.argv) > 1 else 6.0, "hifi", 77, rng, max_len=6000))
problems = [c[0] for c in cases]
sub = [(c[1], c[2]) for c in cases]
nosub = [(np.where(c[1] < 0, -1, 0).astype(np.int32), np.where(c[1] < 0, -1, 0).astype(np.int32)) for c in cases]
print(len(problems), "problems,", sum(len(p) for p in problems), "reads,", sum(int((c[1] > 0).sum()) for c in cases), "partial; max len", max(max(len(s) for s in p) for p in problems))
for name, sb in (("whole-graph", nosub), ("sub-graph", sub), ("whole-graph", nosub), ("sub-graph", sub)):
    t0 = time.perf_counter(); lcd.poa_batch(problems, lcd.poa_params(1, 10), want_msa=False, sub=sb); t1 = time.perf_counter()
    print(f"{name}: {1e3 * (t1 - t0):.1f} ms (host wall, incl. copies)")
