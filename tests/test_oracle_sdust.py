"""CPU: pins oracle/sdust.c (symmetric DUST as the reference's loader runs it over a chunk's reference window: chunk->low_comp_cr) against the UNMODIFIED
sdust() of the reference (src/sdust.c, through oracle/_ref/libref_shim.so)."""
import numpy as np
import pytest
import lcd_testlib as T


def test_oracle_sdust_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(91)
    n_iv = 0
    for it in range(60):
        n = int(rng.choice([1, 2, 3, 19, 20, 21, 200, 5000, 60000]))
        seq = T.sdust_sequence(rng, n, lc_every=int(rng.choice([40, 120, 400])))
        for Tt, W in ((5, 20), (20, 64), (10, 50)):
            a, b = T.sdust(oracle, "lcd_oracle_sdust", seq, Tt, W), T.sdust(ref, "ref_sdust", seq, Tt, W)
            assert a == b, (it, n, Tt, W, a[:3], b[:3])
            n_iv += len(b)
    assert n_iv > 5000


def test_oracle_sdust_edge_cases(oracle, ref):
    for seq in (b"", b"A", b"AC", b"ACG", b"N" * 50, b"A" * 300, b"AC" * 200, b"ACGT" * 100 + b"N" + b"T" * 40, b"acgtnACGTN" * 30, bytes([0, 1, 2, 3] * 40)):
        assert T.sdust(oracle, "lcd_oracle_sdust", seq) == T.sdust(ref, "ref_sdust", seq), seq[:20]


def test_oracle_sdust_every_byte(oracle, ref):
    """every byte value as a 'base': the reference's table makes A / C / G / T in either case and the codes 0 .. 3 bases, the rest breaks a word"""
    rng = np.random.default_rng(94)
    for it in range(60):
        n = int(rng.integers(50, 4000))
        seq = np.frombuffer(b"ACGTacgt\x00\x01\x02\x03", np.uint8)[rng.integers(0, 12, n)].copy()
        hit = rng.random(n) < float(rng.choice([0.0, 0.02, 0.2])); seq[hit] = rng.integers(0, 256, int(hit.sum()))
        if it % 3 == 0: seq[n // 3:n // 3 + 40] = seq[n // 3]
        assert T.sdust(oracle, "lcd_oracle_sdust", seq) == T.sdust(ref, "ref_sdust", seq), it


def test_oracle_sdust_vs_fixtures(oracle):
    """the committed outputs of the unmodified sdust() (tests/golden/sdust_lcd.json.gz): what pins the oracle where /root/reference is absent"""
    cases = T.sdust_fixture_cases()
    assert len(cases) >= 27 and sum(len(c[3]) for c in cases) > 1500
    for k, (seq, Tt, W, iv) in enumerate(cases):
        assert T.sdust(oracle, "lcd_oracle_sdust", seq, Tt, W) == iv, k
