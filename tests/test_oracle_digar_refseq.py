"""CPU: pins the no-tag variant of the difference-list pass (collect_digar_from_ref_seq, reference src/bam_utils.c:1176-1290: plain-M reads without
cs / MD tags, every base compared with the chunk's reference window) in oracle/digar_cs.c (lcd_oracle_collect_digar_refseq) against the
unmodified reference (oracle/_ref/libref_shim.so: ref_collect_digar_refseq), including reads that hang over the ends of the window."""
import ctypes as C

import numpy as np

import lcd_testlib as T
from test_oracle_digar import digar_cases

ENC = np.array([1, 2, 4, 8], np.uint8)        # A C G T in BAM's 4-bit code


def to_refseq(d, rng, trim=True):
    """A chunk with =/X CIGARs -> the same chunk with plain-M CIGARs, a reference window, and SEQ rewritten so that the bases under '=' equal the
    reference and the bases under 'X' differ from it.  With trim the window is cut inside the span of the outermost reads."""
    cig = np.asarray(d["cigar"], np.uint32); seq = np.array(d["bseq"], np.uint8, copy=True)
    lo, hi = 1 << 62, 0
    spans = []
    for r in range(d["n_reads"]):
        ops = cig[int(d["cigar_off"][r]):int(d["cigar_off"][r]) + int(d["n_cigar"][r])].tolist()
        rl = sum(w >> 4 for w in ops if (w & 15) in (0, 2, 3, 7, 8))
        b = int(d["read_pos0"][r]) + 1; e = b + max(rl, 1) - 1
        spans.append((b, e)); lo = min(lo, b); hi = max(hi, e)
    ref = rng.integers(0, 4, hi - lo + 1).astype(np.uint8)
    def put(so, qi, code):
        byte = so + (qi >> 1); sh = ((~qi) & 1) << 2
        seq[byte] = (int(seq[byte]) & (0xff ^ (15 << sh))) | (int(code) << sh)
    new_cig, new_off, new_n = [], [], []
    for r in range(d["n_reads"]):
        ops = cig[int(d["cigar_off"][r]):int(d["cigar_off"][r]) + int(d["n_cigar"][r])].tolist()
        so = int(d["seq_off"][r]); pos = int(d["read_pos0"][r]) + 1; qi = 0; out = []; m = 0
        for w in ops:
            op, ln = w & 15, w >> 4
            if op in (7, 8):
                for j in range(ln):
                    rb = int(ref[pos - lo])
                    put(so, qi, ENC[rb] if op == 7 else ENC[(rb + 1 + int(rng.integers(0, 3))) & 3])
                    pos += 1; qi += 1
                m += ln
            else:
                if m: out.append((m << 4) | 0); m = 0
                out.append(w)
                if op in (2, 3): pos += ln
                elif op in (1, 4): qi += ln
        if m: out.append((m << 4) | 0)
        new_off.append(len(new_cig)); new_n.append(len(out)); new_cig.extend(out)
    ref_beg, ref_end = lo, hi
    if trim and hi - lo > 400:
        ref_beg = lo + int(rng.integers(0, 150)); ref_end = hi - int(rng.integers(0, 150))
    ascii_ref = np.frombuffer(b"ACGT", np.uint8)[ref[ref_beg - lo:ref_end - lo + 1]].copy()
    if len(ascii_ref) > 50:                                   # a few N and lower-case bases, as a FASTA has them
        for k in rng.integers(0, len(ascii_ref), 6).tolist(): ascii_ref[k] = ord("N") if k % 2 else ascii_ref[k] | 0x20
    e = dict(d, cigar=np.array(new_cig + [0], np.uint32), cigar_off=np.array(new_off + [0], np.int64), n_cigar=np.array(new_n + [0], np.int32), bseq=seq)
    return e, ascii_ref, ref_beg, ref_end


def refseq_args(ascii_ref, ref_beg, ref_end):
    return (ascii_ref.ctypes.data_as(C.c_void_p), C.c_int64(ref_beg), C.c_int64(ref_end))


def test_oracle_vs_live_reference(oracle, ref):
    rng = np.random.default_rng(91)
    n_x = n_out = 0
    for n, d in enumerate(digar_cases(93, 100)):
        e, ascii_ref, rb, re_ = to_refseq(d, rng, trim=n % 3 != 0)
        mid = refseq_args(ascii_ref, rb, re_)
        want = T.collect_digar(ref, "ref_collect_digar_refseq", e, mid_args=mid, cap_like=d, slack=2000)
        got = T.collect_digar(oracle, "lcd_oracle_collect_digar_refseq", e, mid_args=mid, cap_like=d, slack=2000)
        assert got == want, n
        n_x += sum(1 for rd in want["reads"].values() for ev in rd[3] if ev[1] == 8)
        n_out += sum(1 for rd in want["reads"].values() if rd[3] and (rd[1] < rb or rd[2] > re_))
    assert n_x > 5000 and n_out > 50, (n_x, n_out)
