"""GPU: liblcd_gpu.so's pileup kernel (one thread per read, atomics per site; through the C-ABI) against the oracle and
the golden fixtures, bit-exact: total / low-quality / allele / strand-by-allele coverage of every candidate site."""
import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth
from test_oracle_pileup import pileup_cases

pytestmark = pytest.mark.gpu


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("pileup_lcd")
    chunks = [{k: (np.array(v, dtype=dict(T.PILEUP_IN_FIELDS)[k]) if k in dict(T.PILEUP_IN_FIELDS) else v) for k, v in c["in"].items()} for c in g["cases"]]
    for c, got in zip(g["cases"], gpu.pileup_batch(chunks)):
        assert got.tolist() == c["counts"]


def test_gpu_vs_oracle_random(gpu, oracle):
    cases = list(pileup_cases(27, 200))
    for i, (d, got) in enumerate(zip(cases, gpu.pileup_batch(cases))):          # one batch of 200 chunks
        assert np.array_equal(got, T.pileup(oracle, "lcd_oracle_collect_cand_vars", d)), (i, d["n_reads"], d["n_sites"])
    assert gpu.pileup_batch([]) == []


def test_gpu_chunk_shaped_plan(gpu, oracle):
    """Chunks shaped like 500 kb at 30x (scaled: 100 kb windows, 15 kb reads), resident plan re-run twice."""
    rng = np.random.default_rng(29)
    cases = [synth.make_pileup_chunk(rng, ref_len=100000, n_reads=220, read_len=(10000, 20000), var_every=160, err_every=500) for _ in range(6)]
    plan = gpu.PileupPlan(cases)
    for _ in range(2):
        plan.run(); plan.sync()
    assert plan.work_units() > 0
    for d, got in zip(cases, plan.fetch()):
        assert np.array_equal(got, T.pileup(oracle, "lcd_oracle_collect_cand_vars", d))
