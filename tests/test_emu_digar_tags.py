"""CPU: the product's tag front end (longcalld_b200/csrc/md_device.cuh: the cs-tag walk and the no-tag walk against the reference window, next
to the MD walk) chained with the difference-list pass (digar_device.cuh, incl. its pseudo-ops for skipped bases and cs clips), compiled for the
host (tests/emu), against the oracle's restatements of collect_digar_from_cs_tag / collect_digar_from_ref_seq / collect_digar_from_MD_tag."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T
from test_oracle_digar import digar_cases
from test_oracle_digar_cs import to_cs
from test_oracle_digar_refseq import to_refseq, refseq_args

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")
KIND_MD, KIND_CS, KIND_REFSEQ = 0, 1, 2


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libdigar_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libdigar_emu.so"))


def tag_args(n_reads, kind, off=None, text=None, ref=None, ref_beg=0, ref_end=-1):
    k = np.full(n_reads + 1, kind, np.int8)
    off = np.zeros(n_reads + 1, np.int64) if off is None else off
    keep = (k, off, text, ref)
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    return (vp(k), vp(off), vp(text), vp(ref), C.c_int64(ref_beg), C.c_int64(ref_end)), keep


def same(a, b, what):
    for r in b["reads"]:
        assert a["reads"][r] == b["reads"][r], (what, r, [x for x, y in zip(a["reads"][r], b["reads"][r]) if x != y][:1])
    assert a["qual_counts"] == b["qual_counts"] and a["chunk_noisy"] == b["chunk_noisy"] and a["totals"] == b["totals"], what


def test_cs_front_end(emu, oracle):
    rng = np.random.default_rng(101)
    for n, d in enumerate(digar_cases(103, 60)):
        e, off, cs = to_cs(d, rng)
        want = T.collect_digar(oracle, "lcd_oracle_collect_digar_cs", e, mid_args=(off.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p)), cap_like=d, slack=64)
        args, keep = tag_args(e["n_reads"], KIND_CS, off, cs)
        same(T.collect_digar(emu, "emu_collect_digar_tags", e, mid_args=args, cap_like=d, slack=64), want, ("cs", n))


def test_refseq_front_end(emu, oracle):
    rng = np.random.default_rng(105)
    for n, d in enumerate(digar_cases(107, 60)):
        e, ref, rb, re_ = to_refseq(d, rng, trim=n % 3 != 0)
        want = T.collect_digar(oracle, "lcd_oracle_collect_digar_refseq", e, mid_args=refseq_args(ref, rb, re_), cap_like=d, slack=2000)
        args, keep = tag_args(e["n_reads"], KIND_REFSEQ, None, None, ref, rb, re_)
        same(T.collect_digar(emu, "emu_collect_digar_tags", e, mid_args=args, cap_like=d, slack=2000), want, ("refseq", n))


def test_cs_letters_that_differ_from_seq_are_rejected(emu):
    """The reference takes alt bases from the tag; the kernels take them from SEQ -- a tag that spells other bases is refused, loudly."""
    rng = np.random.default_rng(109)
    d = next(iter(digar_cases(111, 1)))
    e, off, cs = to_cs(d, rng)
    cs = cs.copy(); star = [i for i in range(len(cs) - 2) if cs[i] == ord("*")]
    assert star
    i = star[0]; cs[i + 2] = ord("a") if cs[i + 2] != ord("a") else ord("c")
    args, keep = tag_args(e["n_reads"], KIND_CS, off, cs)
    inp, k2 = T.digar_input(e)
    with pytest.raises(AssertionError):
        T.collect_digar(emu, "emu_collect_digar_tags", e, mid_args=args, cap_like=d, slack=64)
