"""GPU: liblcd_gpu.so's edlib kernel (through the C-ABI) against the golden fixtures and the oracle, bit-exact
(edit distance, start / end locations and every byte of the path)."""
import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_edlib import edlib_cases

pytestmark = pytest.mark.gpu


def _run_by_mode(gpu, cases, want_path=1):
    """cases: [(q, t, mode)] -> results in the same order (one batch per mode value, as the reference's call sites are)."""
    out = [None] * len(cases)
    for mode in (0, 2):
        idx = [i for i, c in enumerate(cases) if c[2] == mode]
        if idx:
            res = gpu.edlib_batch([(cases[i][0], cases[i][1]) for i in idx], mode, want_path)
            for i, r in zip(idx, res):
                out[i] = r
    return out


def test_gpu_vs_reference_fixtures(gpu):
    g = T.load_golden("edlib_lcd")
    cases, want = [], []
    for c in g["cases"]:
        cases.append((np.array([int(x) for x in c["q"]], dtype=np.uint8), np.array([int(x) for x in c["t"]], dtype=np.uint8), c["mode"]))
        want.append((0, c["ed"], c["start"], c["end"], bytes(int(x) for x in c["aln"])))
    got = _run_by_mode(gpu, cases)
    bad = [i for i in range(len(want)) if got[i] != want[i]]
    assert not bad, (bad[:10], got[bad[0]][:4], want[bad[0]][:4])


def test_gpu_vs_oracle_random(gpu, oracle):
    rng = np.random.default_rng(41)
    cases = edlib_cases(rng, 3000) + edlib_cases(rng, 24, big=True)
    for want_path in (1, 0):
        got = _run_by_mode(gpu, cases, want_path)
        for i, (q, t, mode) in enumerate(cases):
            assert got[i] == T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, want_path), (i, len(q), len(t), mode, want_path)


def test_gpu_edge_cases(gpu, oracle):
    e = np.zeros(0, dtype=np.uint8)
    a = np.array([0, 1, 2, 3, 4, 1], dtype=np.uint8)
    cases = [(q, t, m) for q, t in ((e, a), (a, e), (e, e), (a[:1], a[:1]), (a[:1], a[1:2]), (a, a[:1]), (a[:1], a)) for m in (0, 2)]
    got = _run_by_mode(gpu, cases)
    for (q, t, m), g in zip(cases, got):
        assert g == T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, m, 1), (len(q), len(t), m)
    assert gpu.edlib_batch([], 0, 1) == []
    with pytest.raises(gpu.LcdGpuError):                       # a byte that is not a base code is rejected loudly
        gpu.edlib_batch([(np.array([0, 9, 1], dtype=np.uint8), a)], 0, 1)


def test_gpu_xgaps_read_vs_consensus_shape(gpu, oracle):
    """The reduction longcallD applies (edlib_xgaps, src/align.c:222): mismatches + gap openings of the NW path, on
    a batch shaped like the partial-read filter (read window vs consensus, ~200 bp median, kilobase tail)."""
    rng = np.random.default_rng(43)
    pairs = []
    for _ in range(4000):
        L = int(np.clip(rng.lognormal(np.log(200), 0.8), 20, 3000))
        cons = rng.integers(0, 4, L).astype(np.uint8)
        pairs.append((T.mutate(rng, cons, sub=0.003, ins=0.004, dele=0.004, max_indel=3), cons))
    plan = gpu.EdlibPlan(*gpu.capi.pack_pairs(pairs), gpu.MODE_NW, 1)
    plan.run(); plan.sync()
    res, aln, off = plan.fetch()
    assert plan.work_units() > 0
    for i in range(0, len(pairs), 7):
        st, ed, s, e2, path = T.edlib_align(oracle, "lcd_oracle_edlib_align", pairs[i][0], pairs[i][1], 0, 1)
        got = aln[off[i]:off[i] + res[i]["aln_len"]].tobytes()
        assert (res[i]["edit_distance"], got) == (ed, path), i
        assert gpu.xgaps(got) == sum(1 for j, c in enumerate(path) if c == 3 or (c in (1, 2) and (j == 0 or path[j - 1] != c)))
