"""CPU: pins oracle/phase.c (read -> haplotype assignment and phasing, reference src/assign_hap.c:16-547) against the
unmodified reference (oracle/_ref/libref_shim.so: ref_assign_hap builds a bam_chunk_t around the same flat arrays and
calls assign_hap_based_on_germline_het_vars_kmeans) and against committed reference outputs."""
import numpy as np
import pytest

import lcd_testlib as T

CMP = ("haps", "phase_sets", "hap_to_cons_alle", "hap_to_alle_profile", "var_phase_set", "n_clean_agree_snps", "n_clean_conflict_snps")


def phase_cases(seed, n):
    rng = np.random.default_rng(seed)
    for it in range(n):
        nv = int(rng.choice([0, 1, 2, 5, 20, 60, 150, 400]))
        nr = int(rng.choice([1, 3, 30, 70, 200, 600]))
        tech = "ont" if it % 3 == 2 else "hifi"
        d = T.make_phase_chunk(rng, nv, nr, err=float(rng.choice([0.0, 0.01, 0.05, 0.2])), tech=tech, shuffle_order=(it % 5 == 4))
        yield d, (T.CATE_CLEAN if it % 2 == 0 else T.CATE_GERMLINE), int(tech == "ont")


def same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in CMP)


def test_cr_order_matches_reference(oracle, ref):
    """The order reads come back from cgranges (in-place MSD radix sort: not stable above 64 intervals)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 64, 65, 200, 1000, 5000):
        for spread in (3, 50, 3000):
            start = np.sort(rng.integers(0, spread, n)).astype(np.int32) if n % 2 else rng.integers(0, spread, n).astype(np.int32)
            label = np.arange(n, dtype=np.int32)
            a, b = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
            oracle.lcd_oracle_cr_order(n, start.ctypes.data_as(C.c_void_p), label.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p))
            ref.ref_cr_order(n, start.ctypes.data_as(C.c_void_p), label.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
            assert np.array_equal(a, b), (n, spread)


def test_oracle_vs_live_reference(oracle, ref):
    n = 0
    for d, target, is_ont in phase_cases(5, 400):
        a = T.phase(oracle, "lcd_oracle_assign_hap", d, target, is_ont)
        b = T.phase(ref, "ref_assign_hap", d, target, is_ont)
        assert same(a, b), (n, d["n_reads"], d["n_vars"], target, [k for k in CMP if not np.array_equal(a[k], b[k])])
        n += 1
    assert n == 400


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("phase_lcd")
    assert len(g["cases"]) >= 40
    for c in g["cases"]:
        d = {k: (np.array(v, dtype=dict(T.PHASE_IN_FIELDS)[k]) if k in dict(T.PHASE_IN_FIELDS) else v) for k, v in c["in"].items()}
        d["alle_covs"] = d["alle_covs"].reshape(-1, 4)
        got = T.phase(oracle, "lcd_oracle_assign_hap", d, c["target"], c["is_ont"])
        for k in CMP:
            assert got[k].tolist() == c["out"][k], k
