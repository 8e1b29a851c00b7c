#!/usr/bin/env python3
"""Which pileup / phasing kernels run alongside the persistent POA grid?  python tools/aux_probe.py MBP RESERVE_SMS -- starts the POA plan
(library stream, own host thread: its run() waits for the launch), then 50 ms later times one pool-free plan on the auxiliary stream."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import longcalld_b200 as lcd
from longcalld_b200 import synth
from bench import Workload, PileupStage
import torch

lcd.init(0, 64 << 30); lcd.split_pool(32 << 30)
lcd.reserve_sms(int(sys.argv[2]))
mbp = float(sys.argv[1])
wl = Workload(mbp, "hifi", 11)
gpu_sites = lambda bare, outs, regs: [synth.site_list_from_sites(o, st) for o, st in zip(outs, lcd.sites_batch(bare, regs))]
ps = PileupStage(mbp, "hifi", 11, lcd.digar_batch, gpu_sites, lcd.pileup_batch, pin=True)
digar = lcd.DigarPlan(ps.chunks); digar.run(); digar.sync()
sites = lcd.SitesPlan(None, ps.regs, min_sv_len=[50] * ps.n_chunks, digar_plan=digar); sites.run(); sites.sync()
k2 = lcd.PileupOnSitesPlan(digar, sites); k3 = lcd.ProfileOnDigarPlan(digar, ps.var_sites, ps.n_reads); phase = lcd.PhasePlan(wl.phase)
poa = lcd.PoaPlan(wl.seqs, wl.first, wl.n_reads, wl.read_off, wl.read_len, lcd.poa_params())
poa.run(); poa.sync()
aux_h = lcd.aux_stream(); aux = torch.cuda.ExternalStream(aux_h); st = torch.cuda.ExternalStream(lcd.stream())
for name, plan in [("K1", digar), ("K1b", sites), ("K2", k2), ("K3", k3), ("K4", phase), ("K1", digar)]:
    for during in (False, True):
        t = None
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if during:
            def run_poa():
                p0.record(st); poa.run(); p1.record(st)
            t = threading.Thread(target=run_poa); t.start(); time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(aux); plan.run(aux_h); e1.record(aux)
        if t: t.join()
        torch.cuda.synchronize()
        print(f"{name} {'during POA' if during else 'alone     '}: {e0.elapsed_time(e1):8.2f} ms" + (f"   (POA {p0.elapsed_time(p1):.1f} ms)" if during else ""), flush=True)

# host-side pieces of the e2e pileup chain, each timed with the wall clock while the POA launch runs
def during_poa(name, fn):
    t = threading.Thread(target=lambda: poa.run()); t0 = time.perf_counter(); t.start(); time.sleep(0.05)
    lcd.set_thread_stream(aux_h)
    t1 = time.perf_counter(); r = fn(); t2 = time.perf_counter()
    lcd.set_thread_stream(0)
    t.join(); t3 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name:34s}: {1e3 * (t2 - t1):8.2f} ms   (POA call {1e3 * (t3 - t0):.1f} ms)", flush=True)
    return r

if len(sys.argv) > 3:
    for rep in range(3):
        during_poa("K1 re-run + sync", lambda: (digar.run(), digar.sync()))
        during_poa("K1b re-run + sync", lambda: (sites.run(), sites.sync()))
        during_poa("K2 re-run + sync", lambda: (k2.run(), k2.sync()))
        during_poa("K4 re-run + sync", lambda: (phase.run(), phase.sync()))
    sys.exit(0)
new_sites = during_poa("SitesPlan create (views K1: D2H)", lambda: lcd.SitesPlan(None, ps.regs, min_sv_len=[50] * ps.n_chunks, digar_plan=digar))
during_poa("SitesPlan first run (2 syncs)", lambda: (new_sites.run(), new_sites.sync()))
during_poa("SitesPlan re-run + sync", lambda: (new_sites.run(), new_sites.sync()))
during_poa("SitesPlan fetch (D2H, pageable)", lambda: new_sites.fetch())
nk2 = during_poa("PileupOnSitesPlan create", lambda: lcd.PileupOnSitesPlan(digar, new_sites))
during_poa("K2 run + fetch", lambda: (nk2.run(), nk2.fetch()))
nk3 = during_poa("ProfileOnDigarPlan create (H2D)", lambda: lcd.ProfileOnDigarPlan(digar, ps.var_sites, ps.n_reads))
during_poa("K3 run + fetch", lambda: (nk3.run(), nk3.fetch()))
during_poa("phase_batch", lambda: lcd.phase_batch(wl.phase))
during_poa("plan destroy x3", lambda: [x.destroy() for x in (nk3, nk2, new_sites)])
