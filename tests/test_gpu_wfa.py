"""GPU: liblcd_gpu.so's WFA kernels (through the C-ABI) against the oracle / golden vectors, bit-exact."""
import numpy as np
import pytest

import lcd_testlib as T

pytestmark = pytest.mark.gpu


def _codes(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def _oracle_all(oracle, pairs, params):
    return [T.wfa_align(oracle, "lcd_oracle_wfa_align", p, t, T.WfaParams(*par)) for (p, t), par in zip(pairs, params)]


@pytest.mark.parametrize("name", ["affine", "affine2p", "affine.p0", "affine.p1", "affine.p2",
                                  "affine.wfapt0", "affine.wfapt1"])
def test_gpu_vs_wfa2_utest_golden(gpu, name):
    g = T.load_golden("wfa_utest")
    pairs = [(_codes(p), _codes(t)) for p, t in g["pairs"]]
    got = gpu.wfa_batch(pairs, tuple(g["params"][name]))
    for i, ((p, t), (score, cigar)) in enumerate(zip(pairs, g["golden"][name])):
        assert got[i] == (0, score, T.unrle(cigar), len(p), len(t)), (name, i)


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("wfa_lcd")
    pairs, params, want = [], [], []
    for c in g["cases"]:
        pairs.append((np.array([int(x) for x in c["p"]], dtype=np.uint8), np.array([int(x) for x in c["t"]], dtype=np.uint8)))
        params.append(tuple(c["par"]))
        want.append((c["status"], c["score"], T.unrle(c["ops"]), c["end_v"], c["end_h"]))
    got = gpu.wfa_batch(pairs, params)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (bad[:10], got[bad[0]][:2], want[bad[0]][:2])


def _random_workload(seed, n, lens):
    rng = np.random.default_rng(seed)
    pairs, params = [], []
    for it in range(n):
        L = int(rng.choice(lens))
        a = rng.integers(0, 4, L).astype(np.uint8)
        kind = it % 6
        if kind == 0:
            b = T.mutate(rng, a, sub=0.03, ins=0.02, dele=0.02, max_indel=4)
        elif kind == 1:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 2, "ins", max(1, L // 3)))
        elif kind == 2:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 4, "del", max(1, L // 3)))
        elif kind == 3:
            b = np.concatenate([a[:L // 2], rng.integers(0, 4, L // 2).astype(np.uint8)])
        elif kind == 4:
            b = T.mutate(rng, a, sub=0.15, ins=0.05, dele=0.05, max_indel=2)
        else:
            b = a.copy()
        heur, two = [(T.HEUR_NONE, 1), (T.HEUR_ADAPTIVE, 0), (T.HEUR_ZDROP, 1), (T.HEUR_NONE, 0)][(it // 6) % 4]
        pairs.append((a, b))
        params.append(T.wfa_params_tuple(heur, two, len(a), len(b)))
    return pairs, params


def test_gpu_vs_oracle_random_small(gpu, oracle):
    pairs, params = _random_workload(11, 3000, [0, 1, 3, 20, 70, 150, 400, 900])
    got = gpu.wfa_batch(pairs, params)
    want = _oracle_all(oracle, pairs, params)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (len(bad), bad[:10])


def test_gpu_vs_oracle_random_large(gpu, oracle):
    """kilobase problems (CTA-per-problem kernel, HBM-resident sequences, arena overflow)."""
    pairs, params = _random_workload(12, 96, [1500, 3000, 6000])
    got = gpu.wfa_batch(pairs, params)
    want = _oracle_all(oracle, pairs, params)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (len(bad), bad[:10])


def test_gpu_empty_batch_and_plan_rerun(gpu, oracle):
    assert gpu.wfa_batch([], gpu.wfa_params()) == []
    pairs, params = _random_workload(13, 500, [50, 200, 2000])
    from longcalld_b200.capi import pack_pairs
    seqs, po, pl, to, tl = pack_pairs(pairs)
    plan = gpu.WfaPlan(seqs, po, pl, to, tl, params)
    first = None
    for _ in range(3):                       # re-running a resident plan is idempotent
        plan.run(); plan.sync()
        res, ops, off = plan.fetch()
        cur = [(int(r["status"]), int(r["score"]), ops[off[i]:off[i] + r["n_ops"]].tobytes()) for i, r in enumerate(res)]
        if first is None:
            first = cur
        assert cur == first
    want = _oracle_all(oracle, pairs, params)
    assert [w[:3] for w in want] == first
    assert plan.work_units() > 0


def test_gpu_full_size_properties(gpu):
    """BASELINE-sized batch (too big for the scalar oracle): size-independent properties --
    every CIGAR consumes exactly both sequences, its gap-affine-2p penalty equals -score, and
    identical pairs give all-M."""
    rng = np.random.default_rng(5)
    pairs = []
    for i in range(20000):
        L = int(rng.integers(20, 600))
        a = rng.integers(0, 4, L).astype(np.uint8)
        pairs.append((a, a.copy() if i % 50 == 0 else T.mutate(rng, a, sub=0.01, ins=0.005, dele=0.005, max_indel=3)))
    got = gpu.wfa_batch(pairs, gpu.wfa_params())
    for i, ((a, b), (st, score, ops, ev, eh)) in enumerate(zip(pairs, got)):
        assert st == 0 and (ev, eh) == (len(a), len(b))
        o = np.frombuffer(ops, dtype=np.uint8)
        nm, nx, ni, nd = [(o == ord(c)).sum() for c in "MXID"]
        assert nm + nx + nd == len(a) and nm + nx + ni == len(b)
        pen, j = 0, 0
        s = ops.decode()
        import re
        for run in re.finditer(r"M+|X+|I+|D+", s):
            ch, ln = run.group()[0], len(run.group())
            if ch == "X":
                pen += 6 * ln
            elif ch in "ID":
                pen += min(6 + 2 * ln, 24 + ln)
        assert pen == -score, i
        if i % 50 == 0:
            assert nm == len(a) and score == 0
