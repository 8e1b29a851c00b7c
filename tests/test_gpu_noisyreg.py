"""GPU: liblcd_gpu.so's K2c (noisyreg_kernel: the chunk's noisy-region set and the sites that stay clean-region candidates, one CTA per chunk)
through the C-ABI against the oracle: working categories, kept sites and the final region list, bit-exact."""
import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_noisyreg import noisyreg_cases, noisyreg_fixture_cases

pytestmark = pytest.mark.gpu


def _same(got, want):
    kept, regs, cate = want
    g_regs = list(zip(got["reg_beg"].tolist(), got["reg_end"].tolist(), got["reg_label"].tolist()))
    return g_regs == regs and np.array_equal(got["var_cate"], cate) and np.nonzero(got["keep"])[0].size == len(kept)


def test_gpu_noisy_regs_vs_oracle(gpu, oracle):
    cases = [c for c, _ in noisyreg_cases(oracle)]
    batch = []
    for c in cases:
        batch += [c, dict(c, n_low=0), dict(c, n_cnreg=0), dict(c, n_sites=0), dict(c, min_alt_dp=1, noisy_reg_flank_len=0)]
    got = gpu.noisyreg_batch(batch)
    bad = [i for i, (g, c) in enumerate(zip(got, batch)) if not _same(g, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", c))]
    assert not bad and len(batch) >= 100, bad[:10]
    assert sum(g["n_regs"] for g in got) > 500 and sum(int(g["keep"].sum()) for g in got) > 1000
    # a plan is re-runnable and gives the same answer (scratch is rebuilt by every run)
    plan = gpu.NoisyRegPlan(batch[:20])
    for _ in range(2):
        plan.run(); plan.sync()
    for g, c in zip(plan.fetch(), batch[:20]):
        assert _same(g, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", c))


def test_gpu_noisy_regs_fabricated(gpu, oracle):
    """packed sites in list order (same / neighbouring anchors, long deletions), arbitrary categories: one batch of 300 chunks"""
    from test_emu_noisyreg import fabricated_case
    rng = np.random.default_rng(87)
    batch = [fabricated_case(rng, n_sites=int(rng.integers(2, 220)), span=int(rng.choice([60, 300, 900]))) for _ in range(300)]
    got = gpu.noisyreg_batch(batch)
    bad = [i for i, (g, c) in enumerate(zip(got, batch)) if not _same(g, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", c))]
    assert not bad, bad[:10]


def test_gpu_noisy_regs_chunk_shaped(gpu, oracle):
    """A 500 kb chunk at 30x (the bench's shape: ~1 000 reads, thousands of candidate sites)"""
    from longcalld_b200 import synth
    d = synth.digar_chunks_30x(1, seed=85, chunk_len=500000)[0]
    case, _ = T.noisyreg_case(oracle, d, 991, low_every=250)
    got = gpu.noisyreg_batch([case])[0]
    assert _same(got, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", case)) and case["n_sites"] > 3000


def test_gpu_noisy_regs_rejections(gpu, oracle):
    case = next(noisyreg_cases(oracle))[0]
    bad = dict(case, site_pos=np.asarray(case["site_pos"])[::-1].copy())
    with pytest.raises(gpu.LcdGpuError, match="ascend by anchor"):
        gpu.noisyreg_batch([bad])
    if case["n_low"] > 1:
        bad = dict(case, low_beg=np.asarray(case["low_beg"])[::-1].copy())
        with pytest.raises(gpu.LcdGpuError, match="ascend by start"):
            gpu.noisyreg_batch([bad])


def test_gpu_chain_in_place(gpu, oracle):
    """K1 -> K1b -> K2 -> K2b -> K2c with everything left in HBM (only the reference windows, the options and the low-complexity intervals are uploaded):
    K2c's answer against the oracle run on the fetched K1 / K1b / K2b results; the plan is re-runnable."""
    from longcalld_b200 import synth
    rng = np.random.default_rng(88)
    cases = [synth.make_digar_chunk(rng, n_reads=200, read_len=(4000, 12000), err_every=300, ref_len=80000) for _ in range(2)] + synth.digar_chunks_30x(2, seed=89, chunk_len=60000, read_mean=8000)
    regs = [(int(d["reg_beg"]), int(d["reg_end"])) for d in cases]
    k1 = gpu.DigarPlan(cases); k1.run(); k1.sync()
    k1b = gpu.SitesPlan(None, regs, min_sv_len=[50] * len(cases), digar_plan=k1); k1b.run(); k1b.sync()
    k2 = gpu.PileupOnSitesPlan(k1, k1b); k2.run(); k2.sync()
    recs, sites, counts = k1.fetch(), k1b.fetch(), k2.fetch()
    lists = [synth.site_list_from_sites(o, st) for o, st in zip(recs, sites)]
    cls = [synth.classify_input_from_sites(d, sl, c, 2000 + i) for i, (d, sl, c) in enumerate(zip(cases, lists, counts))]
    k2b = gpu.ClassifyOnPileupPlan(k2, cls, [c["n_sites"] for c in cls]); k2b.run(); k2b.sync()
    cates = k2b.fetch()
    inputs = [synth.noisyreg_input_from(d, o, sl, ct, 3000 + i) for i, (d, o, sl, ct) in enumerate(zip(cases, recs, lists, cates))]
    k2c = gpu.NoisyRegOnClassifyPlan(k1, k2b, inputs, [c["n_sites"] for c in cls], [s[2] for s in k1.sizes()])
    for _ in range(2):
        k2c.run(); k2c.sync()
    got = k2c.fetch()
    for g, x in zip(got, inputs):
        assert _same(g, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", x))
    assert sum(g["n_regs"] for g in got) > 20 and sum(int(g["keep"].sum()) for g in got) > 100
    # the host-buffer form on the same inputs gives the same answer
    for g, h in zip(got, gpu.noisyreg_batch(inputs)):
        assert np.array_equal(g["var_cate"], h["var_cate"]) and np.array_equal(g["keep"], h["keep"]) and np.array_equal(g["reg_beg"], h["reg_beg"]) and np.array_equal(g["reg_end"], h["reg_end"])

    # K0 -> K2c: the low-complexity intervals read where an sdust plan leaves them (their number too), against the oracle on K0's fetched intervals
    wins = [synth.sdust_window(rng, b - a + 2000, lc_every=400) for a, b in regs]
    k0 = gpu.SdustPlan(wins, 5, 20, base=[a - 1000 for a, _ in regs]); k0.run()
    k2c0 = gpu.NoisyRegOnClassifyPlan(k1, k2b, inputs, [c["n_sites"] for c in cls], [s[2] for s in k1.sizes()], sdust_plan=k0)
    for _ in range(2):
        k2c0.run(); k2c0.sync()
    low = k0.fetch()
    assert min(len(x) for x in low) > 200
    for g, x, iv in zip(k2c0.fetch(), inputs, low):
        x0 = dict(x, n_low=len(iv), low_beg=np.asarray([p[0] for p in iv], np.int64), low_end=np.asarray([p[1] for p in iv], np.int64))
        assert _same(g, T.noisy_regs(oracle, "lcd_oracle_noisy_regs", x0))
    # intervals from the host AND an sdust plan: rejected
    import ctypes as C
    from longcalld_b200 import capi
    capi.lib().lcd_noisyreg_plan_create_on_sdust.restype = C.c_void_p
    assert not capi.lib().lcd_noisyreg_plan_create_on_sdust(k1.h, k2b.h, k0.h, C.c_int(len(inputs)), k2c.par)
    assert b"of its own" in capi.lib().lcd_gpu_last_error()

def test_gpu_noisy_regs_vs_reference_fixtures(gpu):
    """committed outputs of the unmodified reference functions: no oracle, no /root/reference in this comparison"""
    cases = list(noisyreg_fixture_cases())
    for g, (case, kept, regs) in zip(gpu.noisyreg_batch([c for c, _, _ in cases]), cases):
        idx = np.nonzero(g["keep"])[0]
        got_kept = [(int(case["site_pos"][i]), int(case["site_type"][i]), int(case["site_ref_len"][i]), int(g["var_cate"][i])) for i in idx]
        assert got_kept == kept and list(zip(g["reg_beg"].tolist(), g["reg_end"].tolist(), g["reg_label"].tolist())) == regs
