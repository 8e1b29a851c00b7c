"""CPU: the product's K6 device logic (longcalld_b200/csrc/wfa_device.cuh) compiled for the host as a
single-lane group (tests/emu) against the golden vectors and the oracle.  Checks the algorithm the
kernel executes -- including a poisoned arena and a tiny private arena that forces the overflow
path -- without a GPU; lane-parallel behaviour is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return C.CDLL(os.path.join(EMU_DIR, "libwfa_emu.so"))


def emu_align(emu, p, t, par, arena_kib=64):
    p, pp = T._u8(p)
    t, tp = T._u8(t)
    ops = C.create_string_buffer(2 * (len(p) + len(t)) + 16)
    res = T.WfaResult()
    emu.emu_wfa_align(pp, len(p), tp, len(t), C.byref(par), ops, C.byref(res), arena_kib)
    return res.status, res.score, ops.raw[:res.n_ops], res.end_v, res.end_h


def test_emu_vs_reference_fixtures(emu):
    g = T.load_golden("wfa_lcd")
    for i, c in enumerate(g["cases"]):
        p = np.array([int(x) for x in c["p"]], dtype=np.uint8)
        t = np.array([int(x) for x in c["t"]], dtype=np.uint8)
        got = emu_align(emu, p, t, T.WfaParams(*c["par"]), arena_kib=1 if i % 3 == 0 else 64)
        assert got == (c["status"], c["score"], T.unrle(c["ops"]), c["end_v"], c["end_h"]), (i, c["par"])


@pytest.mark.parametrize("name", ["affine2p", "affine.wfapt0", "affine.wfapt1", "affine.p1"])
def test_emu_vs_wfa2_utest_golden(emu, name):
    g = T.load_golden("wfa_utest")
    par = T.WfaParams(*g["params"][name])
    n = 0
    for (p, t), (score, cigar) in list(zip(g["pairs"], g["golden"][name]))[::2]:
        if len(p) > 3000:
            continue
        got = emu_align(emu, np.frombuffer(p.encode(), np.uint8), np.frombuffer(t.encode(), np.uint8), par)
        assert got == (0, score, T.unrle(cigar), len(p), len(t)), (name, n)
        n += 1
    assert n > 80
