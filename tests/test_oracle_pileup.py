"""CPU: pins oracle/pileup.c (per-site coverage of the pileup scan: collect_cand_vars / update_cand_vars_from_digar,
reference src/collect_var.c:238, src/bam_utils.c:287) against the unmodified reference (oracle/_ref/libref_shim.so:
ref_collect_cand_vars builds digar_t / var_site_t records around the same flat arrays) and committed reference outputs."""
import numpy as np
import pytest

import lcd_testlib as T
from longcalld_b200 import synth


def pileup_cases(seed, n):
    rng = np.random.default_rng(seed)
    for it in range(n):
        yield synth.make_pileup_chunk(rng, ref_len=int(rng.choice([600, 3000, 20000])), n_reads=int(rng.choice([1, 8, 60, 250])),
                                      read_len=(200, 500) if it % 4 == 0 else (1500, 6000), var_every=int(rng.choice([40, 150, 600])),
                                      err_every=int(rng.choice([60, 800])))


def test_oracle_vs_live_reference(oracle, ref):
    n = tot = 0
    for d in pileup_cases(7, 150):
        a = T.pileup(oracle, "lcd_oracle_collect_cand_vars", d)
        b = T.pileup(ref, "ref_collect_cand_vars", d)
        assert np.array_equal(a, b), (n, d["n_reads"], d["n_sites"], np.nonzero((a != b).any(axis=1))[0][:5])
        n += 1; tot += int(a[:, 3].sum())
    assert tot > 10000          # alt-allele observations were actually matched


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("pileup_lcd")
    assert len(g["cases"]) >= 20
    for c in g["cases"]:
        d = {k: (np.array(v, dtype=dict(T.PILEUP_IN_FIELDS)[k]) if k in dict(T.PILEUP_IN_FIELDS) else v) for k, v in c["in"].items()}
        assert T.pileup(oracle, "lcd_oracle_collect_cand_vars", d).tolist() == c["counts"]


def profile_cases(seed, n):
    rng = np.random.default_rng(seed)
    for d in pileup_cases(seed + 1, n):
        yield synth.add_profile_inputs(rng, d)


def test_profile_oracle_vs_live_reference(oracle, ref):
    n = tot = 0
    for d in profile_cases(9, 150):
        a = T.read_var_profile(oracle, "lcd_oracle_read_var_profile", d)
        b = T.read_var_profile(ref, "ref_read_var_profile", d)
        bad = [r for r in range(len(a)) if a[r] != b[r]]
        assert not bad, (n, d["n_reads"], d["n_sites"], bad[:3], a[bad[0]][:2], b[bad[0]][:2])
        n += 1; tot += sum(sum(1 for x in row[2] if x == 1) for row in a)
    assert tot > 10000
