"""CPU: pins oracle/md.c -- the walk of collect_digar_from_MD_tag (reference src/bam_utils.c:1003-1174) over (CIGAR with M, MD tag) restated as
a conversion to the equivalent =/X CIGAR -- against the unmodified reference: collect_digar_from_MD_tag on reads with plain-M CIGARs and MD
tags (oracle/_ref/libref_shim.so: ref_collect_digar_md) must give exactly what the =/X oracle gives on the converted CIGARs."""
import ctypes as C

import numpy as np

import lcd_testlib as T
from longcalld_b200 import synth
from test_oracle_digar import digar_cases


from longcalld_b200.check import to_md      # noqa: E402  (shared with smoke())


def convert(oracle, e, md_off, md):
    """the M + MD chunk -> =/X CIGARs through the oracle's restatement of the reference's MD walk"""
    oracle.lcd_oracle_md_to_eqx.restype = C.c_int64
    cig, off, cnt = [], [], []
    for r in range(e["n_reads"]):
        ops = np.ascontiguousarray(e["cigar"][int(e["cigar_off"][r]):int(e["cigar_off"][r]) + int(e["n_cigar"][r])], np.uint32)
        cap = int((ops >> 4).sum()) + len(ops) + 8
        out = np.zeros(cap, np.uint32)
        n = oracle.lcd_oracle_md_to_eqx(C.c_int(len(ops)), ops.ctypes.data_as(C.c_void_p), C.c_char_p(bytes(md[int(md_off[r]):]).split(b"\0")[0]),
                                        out.ctypes.data_as(C.c_void_p), C.c_int64(cap))
        assert n >= 0, (r, n)
        off.append(len(cig)); cnt.append(int(n)); cig.extend(out[:n].tolist())
    return dict(e, cigar=np.array(cig + [0], np.uint32), cigar_off=np.array(off + [0], np.int64), n_cigar=np.array(cnt + [0], np.int32))


def test_md_walk_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(71)
    n_x = n_d = 0
    for n, d in enumerate(digar_cases(73, 80)):
        e, md_off, md = to_md(d, rng)
        want = T.collect_digar(ref, "ref_collect_digar_md", e, mid_args=(md_off.ctypes.data_as(C.c_void_p), md.ctypes.data_as(C.c_void_p)), cap_like=d)
        x = convert(oracle, e, md_off, md)
        got = T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", x)
        assert got == want, n
        assert got == T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d), n          # and the same as the chunk's own =/X CIGARs give
        ops = np.asarray(e["cigar"], np.uint32) & 15
        n_x += sum(1 for rd in want["reads"].values() for ev in rd[3] if ev[1] == 8); n_d += int((ops == 2).sum())
    assert n_x > 5000 and n_d > 500, (n_x, n_d)


def test_md_walk_quirks(oracle):
    """the corners of the reference's walk: a run continuing over an insertion, zero-length runs, 0 after a mismatch / deletion, errors"""
    oracle.lcd_oracle_md_to_eqx.restype = C.c_int64
    def conv(cigar, md):
        ops = np.array([(ln << 4) | "MIDNSH=X".index(op) + (0 if op in "MIDNSH" else 1) for ln, op in cigar], np.uint32)
        out = np.zeros(64, np.uint32)
        n = oracle.lcd_oracle_md_to_eqx(C.c_int(len(ops)), ops.ctypes.data_as(C.c_void_p), C.c_char_p(md), out.ctypes.data_as(C.c_void_p), C.c_int64(64))
        return n if n < 0 else [(int(w >> 4), "MIDNSH.=X"[int(w & 15)]) for w in out[:n]]
    assert conv([(5, "M"), (2, "I"), (5, "M")], b"10") == [(5, "="), (2, "I"), (5, "=")]
    assert conv([(17, "M")], b"10A0C5") == [(10, "="), (1, "X"), (1, "X"), (5, "=")]
    assert conv([(3, "M"), (2, "D"), (4, "M")], b"3^AC0T3") == [(3, "="), (2, "D"), (1, "X"), (3, "=")]
    assert conv([(2, "M")], b"0A0C0") == [(1, "X"), (1, "X")]
    assert conv([(4, "S"), (6, "M"), (3, "N"), (6, "M"), (2, "H")], b"12") == [(4, "S"), (6, "="), (3, "N"), (6, "="), (2, "H")]
    assert conv([(5, "M")], b"2#2") == -2 and conv([(5, "=")], b"5") == -3
