"""CPU: pins oracle/digar_cs.c (the cs-tag variant of the difference-list pass, collect_digar_from_cs_tag, reference src/bam_utils.c:844-1001)
against the unmodified reference (oracle/_ref/libref_shim.so: ref_collect_digar_cs appends a cs tag to the bam1_t records it builds).  Groundwork
for the next K1 variant on the GPU: the cs variant takes its alt bases from the tag's letters, clips from the first / last CIGAR op only, and does
not advance over introns, so it is not a front end of the =/X kernels."""
import ctypes as C

import numpy as np

import lcd_testlib as T
from test_oracle_digar import digar_cases


def to_cs(d, rng):
    """A chunk with =/X CIGARs -> the same chunk with plain-M CIGARs + one short-form cs tag per read (":n", "*xy", "+seq", "-seq", "~xxNxx")."""
    cig = np.asarray(d["cigar"], np.uint32); seq = np.asarray(d["bseq"], np.uint8)
    code = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)
    new_cig, new_off, new_n, tags, tag_off = [], [], [], bytearray(), []
    for r in range(d["n_reads"]):
        ops = cig[int(d["cigar_off"][r]):int(d["cigar_off"][r]) + int(d["n_cigar"][r])]
        so = int(d["seq_off"][r])
        base = lambda qi: chr(int(code[(seq[so + (qi >> 1)] >> ((~qi & 1) << 2)) & 15])).lower()
        out, cs, run, m, qi = [], "", 0, 0, 0
        def flush():
            nonlocal cs, run
            if run: cs += ":" + str(run); run = 0
        for w in ops.tolist():
            op, ln = w & 15, w >> 4
            if op == 7: run += ln; m += ln; qi += ln
            elif op == 8:
                flush()
                for _ in range(ln):
                    cs += "*" + "acgt"[int(rng.integers(0, 4))] + base(qi); qi += 1
                m += ln
            else:
                if m: out.append((m << 4) | 0); m = 0
                if op == 1: flush(); cs += "+" + "".join(base(qi + j) for j in range(ln)); qi += ln
                elif op == 2: flush(); cs += "-" + "".join("acgt"[int(x)] for x in rng.integers(0, 4, ln))
                elif op == 3: flush(); cs += "~gt" + str(ln) + "ag"
                elif op == 4: qi += ln
                out.append(w)
        if m: out.append((m << 4) | 0)
        flush()
        new_off.append(len(new_cig)); new_n.append(len(out)); new_cig.extend(out)
        tag_off.append(len(tags)); tags += cs.encode() + b"\0"
    e = dict(d, cigar=np.array(new_cig + [0], np.uint32), cigar_off=np.array(new_off + [0], np.int64), n_cigar=np.array(new_n + [0], np.int32))
    return e, np.array(tag_off + [0], np.int64), np.frombuffer(bytes(tags), np.uint8).copy()


def test_oracle_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(83)
    n_x = n_same = 0
    for n, d in enumerate(digar_cases(85, 100)):
        e, off, cs = to_cs(d, rng)
        mid = (off.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p))
        want = T.collect_digar(ref, "ref_collect_digar_cs", e, mid_args=mid, cap_like=d)
        got = T.collect_digar(oracle, "lcd_oracle_collect_digar_cs", e, mid_args=mid, cap_like=d)
        assert got == want, n
        n_x += sum(1 for rd in want["reads"].values() for ev in rd[3] if ev[1] == 8)
        n_same += got == T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d)
    assert n_x > 5000
    print("chunks whose cs-variant result equals the =/X variant's:", n_same, "of 100")
