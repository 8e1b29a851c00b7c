"""GPU: liblcd_gpu.so's K0 (sdust_kernel: symmetric DUST of the chunks' reference windows, one CTA per chunk, independent segments replayed by one thread each)
through the C-ABI against the oracle, bit-exact; a 600 kb window at the reference's parameters."""
import numpy as np
import pytest

import lcd_testlib as T

pytestmark = pytest.mark.gpu


def test_gpu_sdust_vs_oracle(gpu, oracle):
    rng = np.random.default_rng(93)
    seqs = [T.sdust_sequence(rng, int(rng.choice([1, 2, 3, 19, 20, 21, 45, 200, 3000, 40000])), lc_every=int(rng.choice([30, 120, 400])), n_frac=float(rng.choice([0.0, 0.002, 0.05])))
            for _ in range(120)]
    seqs += [np.frombuffer(s, np.uint8) for s in (b"A", b"ACG", b"N" * 50, b"A" * 300, b"AC" * 200, b"ACGT" * 100 + b"N" + b"T" * 40, b"acgtnACGTN" * 30, b"ANCNGN" * 50 + b"A" * 30, b"AAAN" * 60)]
    n_iv = 0
    for Tt, W in ((5, 20), (8, 16), (4, 24)):
        got = gpu.sdust_batch(seqs, Tt, W)
        bad = [i for i, (g, s) in enumerate(zip(got, seqs)) if g != T.sdust(oracle, "lcd_oracle_sdust", s, Tt, W)]
        assert not bad, (Tt, W, bad[:10])
        n_iv += sum(len(g) for g in got)
    assert n_iv > 10000
    with pytest.raises(gpu.LcdGpuError, match="windows of"):
        gpu.sdust_batch(seqs[:2], 20, 64)


def test_gpu_sdust_chunk_shaped(gpu, oracle):
    """a 600 kb window (a 500 kb chunk with its flanks) at T = 5, W = 20, positions offset as the loader adds them; the plan is re-runnable"""
    rng = np.random.default_rng(94)
    seqs = [T.sdust_sequence(rng, 600000, lc_every=150) for _ in range(3)]
    plan = gpu.SdustPlan(seqs, 5, 20, base=[1000000 - 1, 1500000 - 1, 7])
    for _ in range(2):
        plan.run(); plan.sync()
    for g, s, b in zip(plan.fetch(), seqs, (999999, 1499999, 7)):
        want = [(x + b, y + b) for x, y in T.sdust(oracle, "lcd_oracle_sdust", s)]
        assert g == want and len(g) > 2000


def test_gpu_sdust_vs_reference_fixtures(gpu):
    """the committed outputs of the unmodified sdust() (tests/golden/sdust_lcd.json.gz), one batch per (T, W)"""
    cases = T.sdust_fixture_cases()
    for Tt, W in sorted({(c[1], c[2]) for c in cases}):
        sub = [c for c in cases if (c[1], c[2]) == (Tt, W)]
        got = gpu.sdust_batch([c[0] for c in sub], Tt, W)
        assert [list(map(tuple, g)) for g in got] == [c[3] for c in sub], (Tt, W)
