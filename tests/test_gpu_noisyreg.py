"""GPU: liblcd_gpu.so's K2c (noisyreg_kernel: the chunk's noisy-region set and the sites that stay clean-region candidates, one CTA per chunk)
through the C-ABI against the oracle: working categories, kept sites and the final region list, bit-exact."""
import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_noisyreg import noisyreg_cases

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
