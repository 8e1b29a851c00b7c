"""CPU: the product's K4 device logic (longcalld_b200/csrc/phase_device.cuh) compiled for the host as a one-thread CTA
(tests/emu) against the oracle and the golden fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_phase import phase_cases, same, CMP

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(oracle):
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libphase_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libphase_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    n = 0
    for d, target, is_ont in phase_cases(77, 250):
        a = T.phase(emu, "emu_assign_hap", d, target, is_ont)
        b = T.phase(oracle, "lcd_oracle_assign_hap", d, target, is_ont)
        assert same(a, b), (n, d["n_reads"], d["n_vars"], target, [k for k in CMP if not np.array_equal(a[k], b[k])])
        n += 1


def test_emu_vs_fixtures(emu):
    g = T.load_golden("phase_lcd")
    for c in g["cases"]:
        d = {k: (np.array(v, dtype=dict(T.PHASE_IN_FIELDS)[k]) if k in dict(T.PHASE_IN_FIELDS) else v) for k, v in c["in"].items()}
        d["alle_covs"] = d["alle_covs"].reshape(-1, 4)
        got = T.phase(emu, "emu_assign_hap", d, c["target"], c["is_ont"])
        for k in CMP:
            assert got[k].tolist() == c["out"][k], k
