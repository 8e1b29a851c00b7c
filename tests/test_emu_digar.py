"""CPU: the product's K1 device logic (longcalld_b200/csrc/digar_device.cuh: count / fill one thread per read, the quality
histogram as a single lane) compiled for the host (tests/emu) against the oracle and the golden fixtures."""
import ctypes as C
import os
import subprocess

import pytest

import lcd_testlib as T
from test_oracle_digar import digar_cases

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libdigar_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libdigar_emu.so"))


def test_emu_vs_oracle(emu, oracle):
    for i, d in enumerate(digar_cases(23, 100)):
        a = T.collect_digar(emu, "emu_collect_digar_eqx", d)
        b = T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d)
        for r in b["reads"]:
            assert a["reads"][r] == b["reads"][r], (i, r, [x for x, y in zip(a["reads"][r], b["reads"][r]) if x != y][:1])
        assert a["qual_counts"] == b["qual_counts"] and a["chunk_noisy"] == b["chunk_noisy"] and a["totals"] == b["totals"], i


def test_emu_vs_fixtures(emu):
    for c in T.load_golden("digar_lcd")["cases"]:
        a = T.collect_digar(emu, "emu_collect_digar_eqx", T.digar_case_from_json(c["in"]))
        assert T.digar_digest(a) == c["digest"] and a["chunk_noisy"] == [tuple(x) for x in c["chunk_noisy"]]
