"""CPU: pins oracle/edlib.c (the plain-C restatement of edlib's block bit-vector NW / HW with path) against
  (1) committed outputs of the unmodified reference edlib (tests/golden/edlib_lcd.json.gz),
  (2) the unmodified reference itself (oracle/_ref/libref_shim.so) on seeded random inputs, including
      the Hirschberg regime (> 1 MiB of column data) and the k-doubling regime (distance > 64)."""
import numpy as np
import pytest

import lcd_testlib as T


def edlib_cases(rng, n, big=False):
    """(query, target, mode) triples shaped like longcallD's calls: read-vs-consensus NW paths
    (edlib_xgaps, src/align.c:222), VNTR unit-vs-flank HW infix (edlib_infix_aln, src/align.c:256)."""
    out = []
    for it in range(n):
        L = int(rng.choice([1, 5, 30, 64, 65, 128, 200, 400, 941] if not big else [1500, 2500, 4000]))
        a = rng.integers(0, 4, L).astype(np.uint8)
        kind = it % 6
        if kind == 0:
            b = T.mutate(rng, a, sub=0.01, ins=0.01, dele=0.01, max_indel=3)
        elif kind == 1:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 2, "ins", max(1, L // 3)))
        elif kind == 2:
            b = T.mutate(rng, a, sub=0.002, ins=0.002, dele=0.002, sv=(L // 4, "del", max(1, L // 3)))
        elif kind == 3:
            b = rng.integers(0, 4, max(1, int(L * rng.uniform(0.5, 1.5)))).astype(np.uint8)      # unrelated: distance >> 64
        elif kind == 4:
            unit = rng.integers(0, 4, int(rng.integers(1, 7))).astype(np.uint8)                  # low-complexity: many co-optimal paths
            a = np.resize(unit, L)
            b = T.mutate(rng, a, sub=0.02, ins=0.03, dele=0.03, max_indel=6)
        else:
            b = T.mutate(rng, a, sub=0.10, ins=0.05, dele=0.05, max_indel=2)
        if len(b) == 0:
            b = a[:1].copy()
        if it % 4 == 3:                                   # HW infix: the query inside a longer target
            pad1 = rng.integers(0, 4, int(rng.integers(0, 80))).astype(np.uint8)
            pad2 = rng.integers(0, 4, int(rng.integers(0, 80))).astype(np.uint8)
            out.append((b, np.concatenate([pad1, a, pad2]), 2))
        else:
            out.append((a, b, 0))
        if it % 16 == 5:                                  # both modes on the same shapes
            out[-1] = (out[-1][0], out[-1][1], (it // 16) % 2 * 2)
    return out


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("edlib_lcd")
    assert len(g["cases"]) >= 300
    for c in g["cases"]:
        q = np.array([int(x) for x in c["q"]], dtype=np.uint8)
        t = np.array([int(x) for x in c["t"]], dtype=np.uint8)
        got = T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, c["mode"], 1)
        assert got == (0, c["ed"], c["start"], c["end"], bytes(int(x) for x in c["aln"])), (len(q), len(t), c["mode"])


def test_oracle_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(12)
    n = 0
    for q, t, mode in edlib_cases(rng, 1500):
        for want_path in (1, 0):
            assert T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, want_path) == \
                T.edlib_align(ref, "ref_edlib_align", q, t, mode, want_path), (len(q), len(t), mode)
            n += 1
    assert n == 3000


def test_oracle_vs_live_reference_hirschberg(oracle, ref):
    rng = np.random.default_rng(13)
    for q, t, mode in edlib_cases(rng, 36, big=True):
        assert T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, 1) == \
            T.edlib_align(ref, "ref_edlib_align", q, t, mode, 1), (len(q), len(t), mode)


def test_oracle_empty_and_tiny(oracle, ref):
    e = np.zeros(0, dtype=np.uint8)
    a = np.array([0, 1, 2, 3, 0, 1], dtype=np.uint8)
    for q, t in ((e, a), (a, e), (a[:1], a[:1]), (a[:1], a[1:2]), (a, a[:1]), (a[:1], a)):
        for mode in (0, 2):
            assert T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, 1) == \
                T.edlib_align(ref, "ref_edlib_align", q, t, mode, 1), (len(q), len(t), mode)
