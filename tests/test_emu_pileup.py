"""CPU: the product's K2 device logic (longcalld_b200/csrc/pileup_device.cuh, one thread per read) compiled for the host
(tests/emu) against the oracle and the golden fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_pileup import pileup_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libpileup_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libpileup_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    for i, d in enumerate(pileup_cases(17, 80)):
        assert np.array_equal(T.pileup(emu, "emu_collect_cand_vars", d), T.pileup(oracle, "lcd_oracle_collect_cand_vars", d)), i


def test_emu_vs_fixtures(emu):
    for c in T.load_golden("pileup_lcd")["cases"]:
        d = {k: (np.array(v, dtype=dict(T.PILEUP_IN_FIELDS)[k]) if k in dict(T.PILEUP_IN_FIELDS) else v) for k, v in c["in"].items()}
        assert T.pileup(emu, "emu_collect_cand_vars", d).tolist() == c["counts"]


def test_emu_profile_vs_oracle(emu, oracle):
    from test_oracle_pileup import profile_cases
    for i, d in enumerate(profile_cases(19, 80)):
        assert T.read_var_profile(emu, "emu_read_var_profile", d) == T.read_var_profile(oracle, "lcd_oracle_read_var_profile", d), i
