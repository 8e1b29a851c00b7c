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


def _rows(d, o):
    rows = []
    for r in range(d["n_reads"]):
        ps, pe = int(o["prof_start"][r]), int(o["prof_end"][r])
        n = max(0, pe - ps + 1) if ps >= 0 else 0
        a = int(o["allele_off"][r])
        rows.append((ps, pe, tuple(o["alleles"][a:a + n].tolist()), tuple(o["alt_qi"][a:a + n].tolist())))
    return rows


def test_gpu_profile_vs_oracle_and_into_phasing(gpu, oracle):
    """K3 rows bit-exact vs the oracle, then handed to K4 unchanged: the profile arrays ARE lcd_phase_input_t's."""
    from test_oracle_pileup import profile_cases
    cases = list(profile_cases(31, 120))
    res = gpu.profile_batch(cases)
    for i, (d, o) in enumerate(zip(cases, res)):
        assert _rows(d, o) == T.read_var_profile(oracle, "lcd_oracle_read_var_profile", d), i
    # chain: coverage (K2) + profile (K3) -> phasing (K4) on the device path vs the oracle chain
    d, o = cases[-1], res[-1]
    cov = gpu.pileup_batch([d])[0]
    nv = d["n_sites"]
    ph = dict(n_reads=d["n_reads"], n_vars=nv, ordered_read_ids=d["ordered_read_ids"], is_skipped=d["is_skipped"], prof_start=o["prof_start"],
              prof_end=o["prof_end"], allele_off=o["allele_off"], alleles=o["alleles"], var_cate=d["var_cate"][:nv + 1], var_type=d["site_type"][:nv + 1],
              is_hp_indel=np.zeros(nv + 1, np.int32), n_uniq_alles=np.full(nv + 1, 2, np.int32),
              alle_covs=np.concatenate([cov[:, 2:4], np.zeros((nv, 2), np.int32)], axis=1), total_cov=cov[:, 0], pos=d["site_pos"][:nv + 1])
    got = gpu.phase_batch([(ph, T.CATE_CLEAN, 0)])[0]
    want = T.phase(oracle, "lcd_oracle_assign_hap", ph, T.CATE_CLEAN, 0)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
