#!/usr/bin/env python3
"""bench.py at several --reserve-sms values (CTA slots the persistent DP grids leave to the short kernels): python tools/reserve_sweep.py 12 24 32"""
import json, subprocess, sys
for r in sys.argv[1:]:
    out = subprocess.run([sys.executable, "bench.py", "--reserve-sms", r, "--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--no-whole-program"], capture_output=True, text=True).stdout
    for l in out.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            print(f"reserve {r}: value {d['value']:.1f} Mbp/s ({d['ms_per_step']:.1f} ms, POA overlapped {d['poa_ms_overlapped']:.1f}) e2e {d['e2e']['value']:.1f} pileup thread {d['e2e_pileup_ms']}", flush=True)
