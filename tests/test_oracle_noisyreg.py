"""CPU: pins oracle/noisyreg.c (the chunk's noisy-region set and the candidate sites that stay in the clean regions: pre_process_noisy_regs +
classify_cand_vars after classify_var_cate, with the reference's own interval algebra) against the UNMODIFIED reference functions
(src/collect_var.c:557,902 on src/cgranges.c), called through oracle/_ref/libref_shim.so on chunks assembled from the same flat arrays."""
import numpy as np
import pytest
import lcd_testlib as T


def noisyreg_cases(oracle):
    from longcalld_b200 import synth
    rng = np.random.default_rng(81)
    for i in range(14):          # chunks with many dense clusters / clips (noisy intervals), tight sites
        d = synth.make_digar_chunk(rng, n_reads=int(rng.integers(60, 260)), read_len=(2000, 9000), err_every=int(rng.choice([120, 300, 700])), ref_len=int(rng.choice([20000, 60000])),
                                   tech="ont" if i % 4 == 3 else "hifi")
        yield T.noisyreg_case(oracle, d, 900 + i, is_ont=int(i % 4 == 3), low_every=int(rng.choice([150, 400, 100000])))
    for i, d in enumerate(synth.digar_chunks_30x(6, seed=82, chunk_len=40000, read_mean=6000) + synth.digar_chunks_30x(3, tech="ont", seed=83, chunk_len=30000, read_mean=6000)):
        yield T.noisyreg_case(oracle, d, 950 + i, is_ont=int(i >= 6), low_every=300)


def test_oracle_noisy_regs_vs_live_reference(oracle, ref):
    n = n_regs = n_kept = n_sites = 0
    cates = set()
    for case, ci in noisyreg_cases(oracle):
        kept, regs, cate = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", case)
        rkept, rregs = T.ref_noisy_regs(ref, ci, case)
        assert regs == rregs, (n, regs[:5], rregs[:5])
        assert kept == rkept, (n, len(kept), len(rkept))
        n += 1; n_regs += len(regs); n_kept += len(kept); n_sites += case["n_sites"]; cates |= set(c for _, _, _, c in kept)
    assert n >= 20 and n_regs > 100 and n_kept > 200 and n_sites > 5 * n_kept and len(cates) >= 3, (n, n_regs, n_kept, n_sites, cates)


def test_oracle_noisy_regs_edge_cases(oracle, ref):
    from longcalld_b200 import synth
    rng = np.random.default_rng(84)
    d = synth.make_digar_chunk(rng, n_reads=80, read_len=(2000, 6000), err_every=300, ref_len=20000)
    case, ci = T.noisyreg_case(oracle, d, 990)
    variants = []
    c = dict(case); c.update(n_low=0); variants.append((c, ci))                                       # no low-complexity intervals
    c = dict(case); c.update(n_cnreg=0); variants.append((c, ci))                                     # no noisy intervals from the reads
    c = dict(case); c.update(min_alt_dp=1, noisy_reg_flank_len=0); variants.append((c, dict(ci, min_alt_dp=1)))
    c = dict(case); c.update(n_sites=0); variants.append((c, dict(ci, n_sites=0)))                   # no candidate sites: classify_cand_vars is not called
    for c, k in variants:
        if c["n_sites"] and k.get("min_alt_dp", 2) != 2:
            c = dict(c, var_cate=np.append(T.classify(oracle, "lcd_oracle_classify_sites", k), 0))
        kept, regs, _ = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", c)
        rkept, rregs = T.ref_noisy_regs(ref, k, c)
        assert regs == rregs and kept == rkept


def noisyreg_fixture_cases():
    for c in T.load_golden("noisyreg_lcd")["cases"]:
        yield T.noisyreg_case_from_json(c["in"]), [tuple(k) for k in c["kept"]], [tuple(r) for r in c["regs"]]


def test_oracle_noisy_regs_vs_reference_fixtures(oracle):
    """committed outputs of the unmodified pre_process_noisy_regs + classify_cand_vars (tests/golden/make_golden.py noisyreg_lcd)"""
    n = 0
    for case, kept, regs in noisyreg_fixture_cases():
        got = T.noisy_regs(oracle, "lcd_oracle_noisy_regs", case)
        assert got[0] == kept and got[1] == regs, n
        n += 1
    assert n == 8
