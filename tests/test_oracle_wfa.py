"""CPU: pins oracle/wfa.c (the scalar restatement) against
  (1) WFA2-lib's own golden vectors (WFA2-lib/tests/wfa.utest.check/*.alg),
  (2) committed outputs of the unmodified reference at longcallD's parameter points,
  (3) the unmodified reference itself (oracle/_ref/libref_shim.so) on seeded random inputs."""
import numpy as np
import pytest

import lcd_testlib as T


def _codes(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def _rle_to_ops(c):
    return T.unrle(c)


@pytest.mark.parametrize("name", ["affine", "affine2p", "affine.p0", "affine.p1", "affine.p2",
                                  "affine.wfapt0", "affine.wfapt1"])
def test_oracle_vs_wfa2_utest_golden(oracle, name):
    g = T.load_golden("wfa_utest")
    par = T.WfaParams(*g["params"][name])
    step = 1 if name in ("affine2p", "affine") else 3        # the full set for the two main metrics
    n = 0
    for (p, t), (score, cigar) in list(zip(g["pairs"], g["golden"][name]))[::step]:
        st, sc, ops, ev, eh = T.wfa_align(oracle, "lcd_oracle_wfa_align", _codes(p), _codes(t), par)
        assert st == 0
        assert sc == score, (name, n)
        assert ops == _rle_to_ops(cigar), (name, n)
        assert (ev, eh) == (len(p), len(t))
        n += 1
    assert n > 100


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("wfa_lcd")
    assert len(g["cases"]) > 300
    for c in g["cases"]:
        p = np.array([int(x) for x in c["p"]], dtype=np.uint8)
        t = np.array([int(x) for x in c["t"]], dtype=np.uint8)
        got = T.wfa_align(oracle, "lcd_oracle_wfa_align", p, t, T.WfaParams(*c["par"]))
        assert got == (c["status"], c["score"], T.unrle(c["ops"]), c["end_v"], c["end_h"]), c["par"]


def test_oracle_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(7)
    n = 0
    for it in range(400):
        L = int(rng.choice([3, 20, 70, 150, 400, 1200]))
        a = rng.integers(0, 4, L).astype(np.uint8)
        kind = it % 5
        if kind == 0:
            b = T.mutate(rng, a, sub=0.03, ins=0.02, dele=0.02, max_indel=4)
        elif kind == 1:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 2, "ins", max(1, L // 3)))
        elif kind == 2:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 4, "del", max(1, L // 3)))
        elif kind == 3:
            b = np.concatenate([a[:L // 2], rng.integers(0, 4, L // 2).astype(np.uint8)])
        else:
            b = T.mutate(rng, a, sub=0.15, ins=0.05, dele=0.05, max_indel=2)
        for heur, two in ((T.HEUR_NONE, 1), (T.HEUR_ADAPTIVE, 0), (T.HEUR_ZDROP, 1), (T.HEUR_NONE, 0)):
            par = T.wfa_params(heur, two, len(a), len(b))
            assert T.wfa_align(oracle, "lcd_oracle_wfa_align", a, b, par) == \
                T.wfa_align(ref, "ref_wfa_align", a, b, par), (it, heur, two)
            n += 1
    assert n == 1600
