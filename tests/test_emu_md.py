"""CPU: the product's MD-tag walk (longcalld_b200/csrc/md_device.cuh: count + fill, thread per read) compiled for the host (tests/emu)
against the oracle's restatement of the reference's walk (oracle/md.c), on random reads and on the corners."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_digar import digar_cases
from test_oracle_md import to_md

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libmd_emu.so"])
    lib = C.CDLL(os.path.join(EMU_DIR, "libmd_emu.so"))
    lib.emu_md_to_eqx.restype = C.c_int64
    return lib


def _conv(lib, fn, ops, md):
    ops = np.ascontiguousarray(ops, np.uint32); cap = int((ops >> 4).sum()) + len(ops) + 8
    out = np.zeros(cap, np.uint32)
    f = getattr(lib, fn); f.restype = C.c_int64
    n = f(C.c_int(len(ops)), ops.ctypes.data_as(C.c_void_p), C.c_char_p(md), out.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    return int(n), out[:max(int(n), 0)].tolist()


def test_emu_vs_oracle(emu, oracle):
    rng = np.random.default_rng(75)
    n_reads = 0
    for d in digar_cases(77, 40):
        e, md_off, md = to_md(d, rng)
        for r in range(e["n_reads"]):
            ops = e["cigar"][int(e["cigar_off"][r]):int(e["cigar_off"][r]) + int(e["n_cigar"][r])]
            tag = bytes(md[int(md_off[r]):]).split(b"\0")[0]
            assert _conv(emu, "emu_md_to_eqx", ops, tag) == _conv(oracle, "lcd_oracle_md_to_eqx", ops, tag), (r, tag[:60])
            n_reads += 1
    assert n_reads > 1500


def test_emu_corners(emu, oracle):
    enc = lambda cigar: [(ln << 4) | {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4, "H": 5, "=": 7, "X": 8}[op] for ln, op in cigar]
    for cigar, tag in (([(5, "M"), (2, "I"), (5, "M")], b"10"), ([(17, "M")], b"10A0C5"), ([(3, "M"), (2, "D"), (4, "M")], b"3^AC0T3"), ([(2, "M")], b"0A0C0"),
                       ([(4, "S"), (6, "M"), (3, "N"), (6, "M"), (2, "H")], b"12"), ([(5, "M")], b"2#2"), ([(5, "=")], b"5"), ([(6, "M")], b"3a2"), ([(4, "M")], b"4")):
        assert _conv(emu, "emu_md_to_eqx", enc(cigar), tag) == _conv(oracle, "lcd_oracle_md_to_eqx", enc(cigar), tag), (cigar, tag)
