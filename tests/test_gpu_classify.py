"""GPU: liblcd_gpu.so's per-site category kernel (K2b: one thread per site, through the C-ABI) against the oracle and the golden fixtures,
bit-exact: the LONGCALLD_* category of every candidate site (depth / allele-fraction thresholds, homopolymer and repeat context of small indels)."""
import collections

import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth
from test_oracle_classify import classify_cases

pytestmark = pytest.mark.gpu


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("classify_lcd")
    chunks = [T.classify_case_from_json(c["in"]) for c in g["cases"]]
    for c, got in zip(g["cases"], gpu.classify_batch(chunks)):
        assert got.tolist() == c["cate"]


def test_gpu_vs_oracle_random(gpu, oracle):
    cases = list(classify_cases(65, 200))
    seen = collections.Counter()
    for i, (d, got) in enumerate(zip(cases, gpu.classify_batch(cases))):          # one batch of 200 chunks
        assert np.array_equal(got, T.classify(oracle, "lcd_oracle_classify_sites", d)), (i, d["n_sites"])
        seen.update(got.tolist())
    assert all(seen[c] > 100 for c in (0x001, 0x400, 0x080, 0x010, 0x004, 0x008)), seen
    assert gpu.classify_batch([]) == []


def test_gpu_vs_oracle_ont(gpu, oracle):
    """ONT chunks (BASELINE configs[2]): the strand-bias Fisher test (var_is_strand_bias, src/collect_var.c:270) decides exactly where the oracle
    -- pinned on the live reference -- does, deep sites beyond the lgamma cache included; HiFi and ONT chunks mixed in one batch."""
    cases = list(classify_cases(71, 120, is_ont=1))
    for n, d in enumerate(cases):
        if n % 10 == 0:
            d["site_counts"][:, :] *= 9
    cases += list(classify_cases(73, 20))
    seen = collections.Counter()
    for i, (d, got) in enumerate(zip(cases, gpu.classify_batch(cases))):
        assert np.array_equal(got, T.classify(oracle, "lcd_oracle_classify_sites", d)), (i, d["n_sites"], d["is_ont"])
        seen.update(got.tolist())
    assert seen[0x002] > 500, seen


def test_gpu_chunk_shaped_plan_and_rejections(gpu, oracle):
    """Chunks shaped like 500 kb (8 700 sites on a 600 kb window), resident plan re-run; sites at the window's edge are rejected loudly."""
    rng = np.random.default_rng(67)
    cases = [synth.make_classify_chunk(rng, ref_len=600000, n_sites=8700) for _ in range(4)]
    plan = gpu.ClassifyPlan(cases)
    for _ in range(2):
        plan.run(); plan.sync()
    assert plan.work_units() == 4 * 8700
    for d, got in zip(cases, plan.fetch()):
        assert np.array_equal(got, T.classify(oracle, "lcd_oracle_classify_sites", d))
    edge = synth.make_classify_chunk(rng, ref_len=600, n_sites=20)
    edge["site_type"] = edge["site_type"].copy(); edge["site_type"][0] = 2; edge["site_ref_len"] = edge["site_ref_len"].copy(); edge["site_ref_len"][0] = 2
    edge["site_pos"] = edge["site_pos"].copy(); edge["site_pos"][0] = edge["ref_beg"] + 3
    with pytest.raises(gpu.LcdGpuError, match="reference window"):
        gpu.classify_batch([edge])


def test_gpu_chain_in_place(gpu, oracle):
    """K1 -> K1b -> K2 -> K2b with everything left in HBM (only the reference windows are uploaded for K2b) against the oracle run on the
    fetched sites and counters; a window that ends too close to a small indel is reported, not read past."""
    rng = np.random.default_rng(69)
    cases = [synth.make_digar_chunk(rng, n_reads=240, read_len=(10000, 20000), err_every=300, ref_len=120000) for _ in range(3)]
    regs = [(int(d["reg_beg"]), int(d["reg_end"])) for d in cases]
    k1 = gpu.DigarPlan(cases); k1.run(); k1.sync()
    k1b = gpu.SitesPlan(None, regs, min_sv_len=[50] * len(cases), digar_plan=k1); k1b.run(); k1b.sync()
    k2 = gpu.PileupOnSitesPlan(k1, k1b); k2.run(); k2.sync()
    recs, sites, counts = k1.fetch(), k1b.fetch(), k2.fetch()
    lists = [synth.site_list_from_sites(o, st) for o, st in zip(recs, sites)]
    cls = [synth.classify_input_from_sites(d, sl, c, 1000 + i) for i, (d, sl, c) in enumerate(zip(cases, lists, counts))]
    k2b = gpu.ClassifyOnPileupPlan(k2, cls, [c["n_sites"] for c in cls])
    for _ in range(2):
        k2b.run(); k2b.sync()
    wants = []
    for d, got in zip(cls, k2b.fetch()):
        want = T.classify(oracle, "lcd_oracle_classify_sites", d)
        assert np.array_equal(got, want)
        wants.append(want)
    assert sum(c["n_sites"] for c in cls) > 1000
    ctx = np.nonzero((wants[0] == 0x010) | (wants[0] == 0x008))[0]             # indels whose category needed the reference context
    if len(ctx):
        short = [dict(c) for c in cls]
        short[0]["ref_end"] = int(cls[0]["site_pos"][ctx[-1]]) + 5; short[0]["ref_seq"] = cls[0]["ref_seq"][:short[0]["ref_end"] - short[0]["ref_beg"] + 1]
        bad = gpu.ClassifyOnPileupPlan(k2, short, [c["n_sites"] for c in cls]); bad.run()
        with pytest.raises(gpu.LcdGpuError, match="reference window"):
            bad.fetch()
