"""GPU: liblcd_gpu.so's candidate-site list (K1b: counting sort on position bins + exact per-bin pass, through the C-ABI) against the
oracle and the golden fixtures, bit-exact: every site (position, type, lengths, alt bases) in collect_all_cand_var_sites' order; then
the chain K1 -> K1b -> K2 on the lists left in HBM against the oracle chain."""
import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth
from test_oracle_sites import sites_cases

pytestmark = pytest.mark.gpu


def view(d, o):
    return T.sites_view(d, o["site_pos"], o["site_type"], o["site_ref_len"], o["site_alt_len"], o["site_src"], o["n_sites"])


def test_gpu_vs_reference_fixtures(gpu):
    cases = [T.sites_case_from_json(c) for c in T.load_golden("sites_lcd")["cases"]]
    res = gpu.sites_batch([d for d, _, _ in cases], [reg for _, reg, _ in cases])
    for (d, _, want), o in zip(cases, res):
        assert view(d, o) == want


def test_gpu_vs_oracle_random(gpu, oracle):
    cases = list(sites_cases(47, 160))
    res = gpu.sites_batch([d for d, _ in cases], [reg for _, reg in cases])          # one batch of 160 chunks
    tot = fuzzy = 0
    for i, ((d, (b, e)), o) in enumerate(zip(cases, res)):
        got = view(d, o)
        assert got == T.collect_sites(oracle, "lcd_oracle_collect_sites", d, b, e), (i, d["n_reads"])
        # the record standing for a site is the lowest-indexed identical one: the result does not depend on the atomics' order
        src = o["site_src"]
        assert all(not d["digar_low_qual"][k] and d["digar_type"][k] in (1, 2, 8) for k in src.tolist())
        tot += len(got); fuzzy += sum(1 for s in got if s[1] == 1 and s[3] >= d["min_sv_len"])
    assert tot > 20000 and fuzzy > 100, (tot, fuzzy)
    assert gpu.sites_batch([], []) == []


def test_gpu_edge_cases(gpu, oracle):
    rng = np.random.default_rng(48)
    none = synth.make_pileup_chunk(rng, ref_len=3000, n_reads=8); none["is_skipped"][:] = 1
    empty = synth.make_pileup_chunk(rng, ref_len=3000, n_reads=3); empty["n_reads"] = 0
    one = synth.make_pileup_chunk(rng, ref_len=600, n_reads=1, read_len=(200, 500), err_every=20)
    dense = synth.make_pileup_chunk(rng, ref_len=2000, n_reads=300, read_len=(1500, 2000), var_every=40, err_every=15)     # hundreds of candidates per bin
    cases = [none, empty, one, dense, one]
    lo = int(dense["read_beg"].min())
    regs = [(-1, -1), (-1, -1), (-1, -1), (lo + 500, lo + 1500), (10**9, 10**9 + 5)]          # the last window holds no record
    res = gpu.sites_batch(cases, regs)
    for i, (d, reg, o) in enumerate(zip(cases, regs, res)):
        assert view(d, o) == T.collect_sites(oracle, "lcd_oracle_collect_sites", d, reg[0], reg[1]), i
    assert res[0]["n_sites"] == res[1]["n_sites"] == res[4]["n_sites"] == 0 and res[3]["n_sites"] > 500


def test_gpu_chain_digar_sites_pileup(gpu, oracle):
    """K1 -> K1b -> K2 with everything left in HBM (resident plans, re-run twice) against the oracle chain
    collect_digar_eqx -> collect_sites -> collect_cand_vars."""
    rng = np.random.default_rng(49)
    cases = [synth.make_digar_chunk(rng, n_reads=240, read_len=(10000, 20000), err_every=300, ref_len=120000) for _ in range(3)]
    cases.append(synth.make_digar_chunk(rng, n_reads=5)); cases[-1]["is_skipped"][:] = 1
    regs = [(int(d["reg_beg"]), int(d["reg_end"])) for d in cases]; regs[1] = (-1, -1)
    k1 = gpu.DigarPlan(cases)
    k1.run(); k1.sync()
    k1b = gpu.SitesPlan(None, regs, min_sv_len=[50] * len(cases), digar_plan=k1)
    for _ in range(2):
        k1b.run(); k1b.sync()
    assert k1b.work_units() > 0
    k2 = gpu.PileupOnSitesPlan(k1, k1b)
    for _ in range(2):
        k2.run(); k2.sync()
    recs, sites, counts = k1.fetch(), k1b.fetch(), k2.fetch()
    n_tot = 0
    for i, (d, o, st, cnt) in enumerate(zip(cases, recs, sites, counts)):
        nr = d["n_reads"]
        pile = dict(n_reads=nr, n_sites=0, min_bq=d["min_bq"], min_sv_len=50, ordered_read_ids=d["ordered_read_ids"],
                    is_skipped=np.maximum(d["is_skipped"], o["skip"][:nr]), read_beg=o["read_beg"], read_end=o["read_end"], read_is_rev=d["read_is_rev"],
                    digar_first=o["digar_first"], n_digar=o["n_digar"], qual_off=d["qual_off"], qual=d["qual"], digar_pos=o["digar_pos"], digar_type=o["digar_type"],
                    digar_len=o["digar_len"], digar_qi=o["digar_qi"], digar_low_qual=o["digar_low_qual"], digar_alt_off=o["digar_alt_off"], digar_alt=o["digar_alt"],
                    **{k: np.zeros(1, t) for k, t in T.PILEUP_IN_FIELDS if k.startswith("site_")})
        want = T.collect_sites(oracle, "lcd_oracle_collect_sites", pile, regs[i][0], regs[i][1])
        assert view(pile, st) == want, i
        ns = st["n_sites"]; n_tot += ns
        if ns == 0: continue
        src = st["site_src"]
        alt, aoff = [], []
        for k in src.tolist():
            aoff.append(len(alt))
            if o["digar_type"][k] != 2: alt.extend(o["digar_alt"][int(o["digar_alt_off"][k]):int(o["digar_alt_off"][k]) + int(o["digar_len"][k])].tolist())
        pile.update(n_sites=ns, site_pos=np.append(st["site_pos"], 0), site_type=np.append(st["site_type"], 0).astype(np.int32),
                    site_ref_len=np.append(st["site_ref_len"], 0).astype(np.int32), site_alt_len=np.append(st["site_alt_len"], 0).astype(np.int32),
                    site_alt_off=np.array(aoff + [0], np.int64), site_alt=np.array(alt + [0], np.uint8))
        assert np.array_equal(cnt, T.pileup(oracle, "lcd_oracle_collect_cand_vars", pile)), i
    assert n_tot > 1000
