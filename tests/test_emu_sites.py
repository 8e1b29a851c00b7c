"""CPU: the product's candidate-site device logic (longcalld_b200/csrc/sites_device.cuh: counting sort on position bins + exact
per-bin pass) compiled for the host (tests/emu) against the oracle and the golden fixtures."""
import ctypes as C
import os
import subprocess

import pytest

import lcd_testlib as T
from test_oracle_sites import sites_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libsites_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libsites_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    tot = 0
    for n, (d, (b, e)) in enumerate(sites_cases(43, 100)):
        got = T.collect_sites(emu, "emu_collect_sites", d, b, e)
        assert got == T.collect_sites(oracle, "lcd_oracle_collect_sites", d, b, e), n
        tot += len(got)
    assert tot > 10000


def test_emu_vs_fixtures(emu):
    for c in T.load_golden("sites_lcd")["cases"]:
        d, reg, want = T.sites_case_from_json(c)
        assert T.collect_sites(emu, "emu_collect_sites", d, reg[0], reg[1]) == want
