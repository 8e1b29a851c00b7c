"""CPU: the product's K7 device logic (longcalld_b200/csrc/edlib_device.cuh, one thread per problem) compiled for
the host (tests/emu) against the golden fixtures and the oracle, inside exactly the workspace the host plan
hands out (poisoned, guard zone checked)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_edlib import edlib_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return C.CDLL(os.path.join(EMU_DIR, "libedlib_emu.so"))


def test_emu_vs_golden(emu):
    g = T.load_golden("edlib_lcd")
    for c in g["cases"]:
        q = np.array([int(x) for x in c["q"]], dtype=np.uint8)
        t = np.array([int(x) for x in c["t"]], dtype=np.uint8)
        got = T.edlib_align(emu, "emu_edlib_align", q, t, c["mode"], 1)
        assert got == (0, c["ed"], c["start"], c["end"], bytes(int(x) for x in c["aln"])), (len(q), len(t), c["mode"])


def test_emu_vs_oracle(emu, oracle):
    rng = np.random.default_rng(31)
    for q, t, mode in edlib_cases(rng, 700) + edlib_cases(rng, 12, big=True):
        for want_path in (1, 0):
            assert T.edlib_align(emu, "emu_edlib_align", q, t, mode, want_path) == \
                T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, want_path), (len(q), len(t), mode, want_path)


def test_emu_edge_cases(emu, oracle):
    e = np.zeros(0, dtype=np.uint8)
    a = np.array([0, 1, 2, 3, 4, 1], dtype=np.uint8)
    for q, t in ((e, a), (a, e), (e, e), (a[:1], a[:1]), (a[:1], a[1:2]), (a, a[:1]), (a[:1], a)):
        for mode in (0, 2):
            assert T.edlib_align(emu, "emu_edlib_align", q, t, mode, 1) == \
                T.edlib_align(oracle, "lcd_oracle_edlib_align", q, t, mode, 1), (len(q), len(t), mode)
    bad = np.array([0, 9, 2], dtype=np.uint8)                      # not a base code: rejected loudly
    assert T.edlib_align(emu, "emu_edlib_align", bad, a, 0, 1)[0] == -1
