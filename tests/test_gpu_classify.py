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


def test_gpu_chunk_shaped_plan_and_rejections(gpu, oracle):
    """Chunks shaped like 500 kb (8 700 sites on a 600 kb window), resident plan re-run; ONT chunks and sites at the window's edge are rejected loudly."""
    rng = np.random.default_rng(67)
    cases = [synth.make_classify_chunk(rng, ref_len=600000, n_sites=8700) for _ in range(4)]
    plan = gpu.ClassifyPlan(cases)
    for _ in range(2):
        plan.run(); plan.sync()
    assert plan.work_units() == 4 * 8700
    for d, got in zip(cases, plan.fetch()):
        assert np.array_equal(got, T.classify(oracle, "lcd_oracle_classify_sites", d))
    ont = dict(cases[0], is_ont=1)
    with pytest.raises(gpu.LcdGpuError, match="ONT"):
        gpu.classify_batch([ont])
    edge = synth.make_classify_chunk(rng, ref_len=600, n_sites=20)
    edge["site_type"] = edge["site_type"].copy(); edge["site_type"][0] = 2; edge["site_ref_len"] = edge["site_ref_len"].copy(); edge["site_ref_len"][0] = 2
    edge["site_pos"] = edge["site_pos"].copy(); edge["site_pos"][0] = edge["ref_beg"] + 3
    with pytest.raises(gpu.LcdGpuError, match="reference window"):
        gpu.classify_batch([edge])
