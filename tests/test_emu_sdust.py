"""CPU: the product's K0 device logic (longcalld_b200/csrc/sdust_device.cuh: per-position window statistics, independent segments replayed by one thread each)
compiled for the host with a one-thread CTA (tests/emu) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module", params=["libsdust_emu.so", "libsdust_simt_emu.so"])
def emu(request):
    """one thread per chunk / 32 cooperating fibers per chunk (tests/emu/simt_emu.h)"""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, request.param])
    return C.CDLL(os.path.join(EMU_DIR, request.param))


def test_emu_sdust_vs_oracle(emu, oracle):
    rng = np.random.default_rng(92)
    n_iv = 0
    for it in range(80):
        n = int(rng.choice([1, 2, 3, 19, 20, 21, 45, 200, 3000, 40000]))
        seq = T.sdust_sequence(rng, n, lc_every=int(rng.choice([30, 120, 400])), n_frac=float(rng.choice([0.0, 0.002, 0.05])))
        for Tt, W in ((5, 20), (8, 16), (4, 24)):
            a, b = T.sdust(emu, "emu_sdust", seq, Tt, W), T.sdust(oracle, "lcd_oracle_sdust", seq, Tt, W)
            assert a == b, (it, n, Tt, W, a[:3], b[:3])
            n_iv += len(b)
    assert n_iv > 5000
    for seq in (b"", b"A", b"AC", b"ACG", b"N" * 50, b"A" * 300, b"AC" * 200, b"ACGT" * 100 + b"N" + b"T" * 40, b"acgtnACGTN" * 30, b"ANCNGN" * 50 + b"A" * 30, b"AAAN" * 60):
        assert T.sdust(emu, "emu_sdust", seq) == T.sdust(oracle, "lcd_oracle_sdust", seq), seq[:20]


def test_emu_sdust_long_repeats(emu, oracle):
    """long tandem repeats / homopolymers with mutations sprinkled in: hundreds of perfect intervals alive at once (the incremental scan of that list)"""
    rng = np.random.default_rng(93)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    n_iv = 0
    for it in range(40):
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            u = int(rng.integers(1, 9)); reps = int(rng.integers(20, 400))
            unit = rng.integers(0, 4, u)
            s = np.tile(unit, reps)
            hit = rng.random(len(s)) < float(rng.choice([0.0, 0.01, 0.05, 0.15])); s[hit] = rng.integers(0, 4, int(hit.sum()))
            parts.append(acgt[s])
            parts.append(acgt[rng.integers(0, 4, int(rng.integers(0, 80)))])
        seq = np.concatenate(parts)
        for Tt, W in ((5, 20), (8, 16), (4, 24)):
            a, b = T.sdust(emu, "emu_sdust", seq, Tt, W), T.sdust(oracle, "lcd_oracle_sdust", seq, Tt, W)
            assert a == b, (it, len(seq), Tt, W, a[:3], b[:3])
            n_iv += len(b)
    assert n_iv > 100


def test_emu_sdust_every_byte(emu, oracle):
    """every byte value as a 'base' (the reference's table: A / C / G / T in either case and the codes 0 .. 3 are bases, the rest breaks a word)"""
    rng = np.random.default_rng(94)
    for it in range(30):
        n = int(rng.integers(50, 4000))
        seq = np.frombuffer(b"ACGTacgt\x00\x01\x02\x03", np.uint8)[rng.integers(0, 12, n)].copy()
        hit = rng.random(n) < float(rng.choice([0.0, 0.02, 0.2])); seq[hit] = rng.integers(0, 256, int(hit.sum()))
        if it % 3 == 0: seq[n // 3:n // 3 + 40] = seq[n // 3]            # a run, so that there is something to find
        assert T.sdust(emu, "emu_sdust", seq) == T.sdust(oracle, "lcd_oracle_sdust", seq), it


def test_emu_sdust_vs_fixtures(emu):
    """the device logic against the committed outputs of the unmodified sdust()"""
    for k, (seq, Tt, W, iv) in enumerate(T.sdust_fixture_cases()):
        assert T.sdust(emu, "emu_sdust", seq, Tt, W) == iv, k
