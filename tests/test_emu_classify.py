"""CPU: the product's K2b device logic (longcalld_b200/csrc/classify_device.cuh, one thread per site) compiled for the host (tests/emu)
against the oracle and the golden fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_classify import classify_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libclassify_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libclassify_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    for n, d in enumerate(classify_cases(63, 100)):
        assert np.array_equal(T.classify(emu, "emu_classify_sites", d), T.classify(oracle, "lcd_oracle_classify_sites", d)), n


def test_emu_vs_oracle_ont(emu, oracle):
    for n, d in enumerate(classify_cases(69, 60, is_ont=1)):
        if n % 10 == 0:
            d["site_counts"][:, :] *= 9          # beyond the lgamma cache
        assert np.array_equal(T.classify(emu, "emu_classify_sites", d), T.classify(oracle, "lcd_oracle_classify_sites", d)), n


def test_emu_vs_fixtures(emu):
    for c in T.load_golden("classify_lcd")["cases"]:
        assert T.classify(emu, "emu_classify_sites", T.classify_case_from_json(c["in"])).tolist() == c["cate"]
