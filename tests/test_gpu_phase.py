"""GPU: liblcd_gpu.so's phasing kernel (one CTA per chunk, through the C-ABI) against the oracle and the golden
fixtures, bit-exact: read haplotypes and phase sets, per-variant consensus alleles, per-haplotype allele counts, phase
sets, and the clean-SNP agree / conflict counters."""
import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_phase import phase_cases, CMP

pytestmark = pytest.mark.gpu


def _cmp(got, want, nr, nv, tag):
    sizes = {"haps": nr, "phase_sets": nr, "hap_to_cons_alle": 3 * nv, "hap_to_alle_profile": 12 * nv, "var_phase_set": nv,
             "n_clean_agree_snps": nr, "n_clean_conflict_snps": nr}
    for k in CMP:
        assert np.array_equal(got[k][:sizes[k]], want[k][:sizes[k]]), (tag, k)


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("phase_lcd")
    chunks = []
    for c in g["cases"]:
        d = {k: (np.array(v, dtype=dict(T.PHASE_IN_FIELDS)[k]) if k in dict(T.PHASE_IN_FIELDS) else v) for k, v in c["in"].items()}
        d["alle_covs"] = d["alle_covs"].reshape(-1, 4)
        chunks.append((d, c["target"], c["is_ont"]))
    res = gpu.phase_batch(chunks)
    for i, (c, r) in enumerate(zip(g["cases"], res)):
        d = chunks[i][0]
        _cmp(r, {k: np.array(c["out"][k]) for k in CMP}, d["n_reads"], d["n_vars"], i)


def test_gpu_vs_oracle_random(gpu, oracle):
    cases = list(phase_cases(91, 600))
    res = gpu.phase_batch(cases)                                   # one batch: 600 chunks, one CTA each
    for i, ((d, target, is_ont), r) in enumerate(zip(cases, res)):
        want = T.phase(oracle, "lcd_oracle_assign_hap", d, target, is_ont)
        _cmp(r, want, d["n_reads"], d["n_vars"], (i, d["n_reads"], d["n_vars"], target))


def test_gpu_chunk_shaped_batch(gpu, oracle):
    """Chunks shaped like a 500 kb HiFi region chunk (~1000 reads, a few thousand candidate variants), both category masks
    (the first call of collect_var_main uses the clean mask, later calls the germline mask)."""
    rng = np.random.default_rng(93)
    cases = []
    for i in range(12):
        d = T.make_phase_chunk(rng, n_vars=int(rng.integers(1500, 4000)), n_reads=int(rng.integers(800, 1300)), err=0.01,
                               tech="hifi" if i % 3 else "ont")
        cases.append((d, T.CATE_CLEAN if i % 2 else T.CATE_GERMLINE, int(i % 3 == 0)))
    plan = gpu.PhasePlan(cases)
    plan.run(); plan.sync()
    res = plan.fetch()
    assert plan.work_units() > 0
    for i, ((d, target, is_ont), r) in enumerate(zip(cases, res)):
        want = T.phase(oracle, "lcd_oracle_assign_hap", d, target, is_ont)
        _cmp(r, want, d["n_reads"], d["n_vars"], i)
    assert gpu.phase_batch([]) == []
