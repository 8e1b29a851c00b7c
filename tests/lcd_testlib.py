"""Shared helpers for the test-suite: ctypes bindings for the oracle port (oracle/liblcd_oracle.so),
the reference shim (oracle/_ref/libref_shim.so, built from /root/reference where present) and the
product C-ABI (longcalld_b200/csrc/liblcd_gpu.so), plus seeded workload generators.

TEST INFRASTRUCTURE: nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblcd_oracle.so")
REF_SHIM_SO = os.path.join(ORACLE_DIR, "_ref", "libref_shim.so")


class WfaParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2", "affine2p", "heuristic",
        "min_wavefront_length", "max_distance_threshold", "zdrop", "steps_between_cutoffs")]


class WfaResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("status", "score", "n_ops", "end_v", "end_h")]


class EdlibResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("status", "edit_distance", "start_loc", "end_loc", "aln_len")]


HEUR_NONE, HEUR_ADAPTIVE, HEUR_ZDROP = 0, 1, 2


def wfa_params(heuristic=HEUR_NONE, affine2p=1, plen=0, tlen=0, x=6, o1=6, e1=2, o2=24, e2=1):
    """Parameter points longcallD uses (src/align.h:21-26, src/align.c:398-406)."""
    p = WfaParams(x, o1, e1, o2, e2, affine2p, heuristic, 10, 50, 0, 1)
    if heuristic == HEUR_ZDROP:
        p.zdrop = min(500, int(min(plen, tlen) * 0.1))
        p.steps_between_cutoffs = 100
    return p


def wfa_params_tuple(heuristic=HEUR_NONE, affine2p=1, plen=0, tlen=0, x=6, o1=6, e1=2, o2=24, e2=1):
    p = wfa_params(heuristic, affine2p, plen, tlen, x, o1, e1, o2, e2)
    return tuple(getattr(p, f) for f, _ in WfaParams._fields_)


def unrle(s):
    """'14M1I86M' -> b'MMMM...'"""
    import re
    return b"".join(op.encode() * int(n) for n, op in re.findall(r"(\d+)([MXID])", s))


def load_golden(name):
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_libs = {}


def _load(path):
    if path not in _libs:
        _libs[path] = C.CDLL(path)
    return _libs[path]


def oracle_lib():
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    return _load(ORACLE_SO)


def ref_lib():
    """The unmodified reference behind a flat-C shim; None when it was never built."""
    if not os.path.exists(REF_SHIM_SO):
        if os.path.isdir("/root/reference/src"):
            build_oracle()
        else:
            return None
    return _load(REF_SHIM_SO)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def wfa_align(lib, fn, pattern, text, par):
    p, pp = _u8(pattern)
    t, tp = _u8(text)
    ops = C.create_string_buffer(2 * (len(p) + len(t)) + 16)
    res = WfaResult()
    getattr(lib, fn)(pp, C.c_int(len(p)), tp, C.c_int(len(t)), C.byref(par), ops, C.byref(res))
    return res.status, res.score, ops.raw[:res.n_ops], res.end_v, res.end_h


def edlib_align(lib, fn, query, target, mode=0, want_path=1):
    q, qp = _u8(query)
    t, tp = _u8(target)
    aln = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
    res = EdlibResult()
    getattr(lib, fn)(qp, C.c_int(len(q)), tp, C.c_int(len(t)), C.c_int(mode), C.c_int(want_path),
                     aln.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(res))
    return res.status, res.edit_distance, res.start_loc, res.end_loc, bytes(aln[:res.aln_len])


# ----------------------------------------------------------------------------- workload generators
def mutate(rng, seq, sub=0.01, ins=0.01, dele=0.01, max_indel=1, sv=None):
    """Return a mutated copy of `seq` (uint8 codes 0..3).  `sv`=(pos, kind, length) plants one large
    indel.  Homopolymer-style indels repeat the neighbouring base (what HiFi errors look like)."""
    out = []
    i = 0
    n = len(seq)
    while i < n:
        if sv is not None and i == sv[0]:
            if sv[1] == "ins":
                out.extend(rng.integers(0, 4, sv[2]).tolist())
            else:
                i += sv[2]
                if i >= n:
                    break
        r = rng.random()
        if r < sub:
            out.append((int(seq[i]) + int(rng.integers(1, 4))) % 4)
            i += 1
        elif r < sub + ins:
            L = int(rng.integers(1, max_indel + 1))
            out.extend([int(seq[i])] * L if rng.random() < 0.7 else rng.integers(0, 4, L).tolist())
            out.append(int(seq[i]))
            i += 1
        elif r < sub + ins + dele:
            i += int(rng.integers(1, max_indel + 1))
        else:
            out.append(int(seq[i]))
            i += 1
    return np.array(out, dtype=np.uint8)


def random_pair(rng, n, **kw):
    a = rng.integers(0, 4, n).astype(np.uint8)
    return a, mutate(rng, a, **kw)


# ----------------------------------------------------------------------------- POA helpers
class PoaParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_ext1", C.c_int32),
                ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("sub_aln", C.c_int32), ("max_n_cons", C.c_int32)]


def poa_params(sub_aln=1, wb=10, wf=0.01):
    """longcallD's two abPOA set-ups (src/align.c:769-783 phased, :876-889 de-novo uses wb=-1)."""
    return PoaParams(2, 6, 6, 2, 24, 1, wb, wf, sub_aln, 1)


def poa(lib, fn, seqs, par):
    """-> (rc, consensus bytes, msa as (n_seq+1, msa_len) uint8 array)"""
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    off = np.zeros(n, dtype=np.int64)
    if n > 1:
        off[1:] = np.cumsum(lens[:-1])
    flat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if n else np.zeros(1, np.uint8)
    flat = np.ascontiguousarray(flat)
    tot = int(lens.sum())
    cons = np.zeros(tot + 8, dtype=np.uint8)
    cap = (n + 1) * (tot + 8)
    msa = np.zeros(cap, dtype=np.uint8)
    cl, ml = C.c_int32(0), C.c_int32(0)
    rc = getattr(lib, fn)(C.c_int(n), flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                          lens.ctypes.data_as(C.c_void_p), C.byref(par), cons.ctypes.data_as(C.c_void_p), C.byref(cl),
                          msa.ctypes.data_as(C.c_void_p), C.byref(ml), C.c_int32(min(cap, 2**31 - 1)))
    return rc, cons[:cl.value].tobytes(), msa[:(n + 1) * ml.value].reshape(n + 1, ml.value).copy()


def poa_ncons(lib, fn, seqs, par, min_freq=0.20):
    """abpoa_aln_msa_cons with max_n_cons consensus sequences -> (rc, [consensus bytes], read clusters, msa as (n_seq + n_cons, msa_len))"""
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    off = np.zeros(n, dtype=np.int64)
    if n > 1:
        off[1:] = np.cumsum(lens[:-1])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs] + [np.zeros(1, np.uint8)]))
    tot = int(lens.sum())
    cons = np.zeros(tot + 8, dtype=np.uint8)
    cap = (n + 2) * (tot + 8)
    msa = np.zeros(cap, dtype=np.uint8)
    clu = np.zeros(n + 1, dtype=np.uint8)
    cl = (C.c_int32 * 2)(0, 0)
    nc, ml = C.c_int32(0), C.c_int32(0)
    rc = getattr(lib, fn)(C.c_int(n), flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), C.byref(par),
                          C.c_double(min_freq), cons.ctypes.data_as(C.c_void_p), cl, C.byref(nc), clu.ctypes.data_as(C.c_void_p),
                          msa.ctypes.data_as(C.c_void_p), C.byref(ml), C.c_int32(min(cap, 2**31 - 1)))
    cs, at = [], 0
    for c in range(nc.value):
        cs.append(cons[at:at + cl[c]].tobytes()); at += cl[c]
    return rc, cs, clu[:n].copy(), msa[:(n + nc.value) * ml.value].reshape(n + nc.value, ml.value).copy()


def denovo_problems(mbp, tech, seed, max_len=600, max_reads=40):
    """Read sets as wfa_collect_noisy_aln_str_no_ps_hap hands them to abpoa_aln_msa_cons (src/align.c:1180): all fully covering reads of a region,
    both haplotypes mixed in input order."""
    from longcalld_b200 import synth
    for r in synth.make_regions(mbp, tech, seed=seed, with_reads=True):
        seqs = [s for s in r.reads if len(s)]
        if 2 <= len(seqs) <= max_reads and max(len(s) for s in seqs) <= max_len:
            yield seqs


def poa_sub(lib, fn, seqs, sub_beg, sub_end, par):
    """The sub-graph form (partially covering reads): sub_beg / sub_end per read as lcd_poa_sub_batch takes them."""
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    off = np.zeros(n, dtype=np.int64)
    if n > 1:
        off[1:] = np.cumsum(lens[:-1])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs] + [np.zeros(1, np.uint8)]))
    tot = int(lens.sum())
    cons = np.zeros(tot + 8, dtype=np.uint8)
    cap = (n + 1) * (tot + 8)
    msa = np.zeros(cap, dtype=np.uint8)
    cl, ml = C.c_int32(0), C.c_int32(0)
    sb, se = np.ascontiguousarray(sub_beg, np.int32), np.ascontiguousarray(sub_end, np.int32)
    rc = getattr(lib, fn)(C.c_int(n), flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p),
                          sb.ctypes.data_as(C.c_void_p), se.ctypes.data_as(C.c_void_p), C.byref(par), cons.ctypes.data_as(C.c_void_p), C.byref(cl),
                          msa.ctypes.data_as(C.c_void_p), C.byref(ml), C.c_int32(min(cap, 2**31 - 1)))
    return rc, cons[:cl.value].tobytes(), msa[:(n + 1) * ml.value].reshape(n + 1, ml.value).copy()


def partial_cover_problems(mbp, tech, seed, rng, max_len=1500):
    """(reads, sub_beg, sub_end) per (region, haplotype): some reads of every problem are cut to a prefix / suffix / inner piece of the region and
    anchored on the nodes of the first read they were cut at (what collect_partial_aln_beg_end + abpoa_subgraph_nodes's caller computes,
    src/align.c:795-803: beg_id = ref_beg + 1, end_id = ref_end + 1 with 1-based positions in the first read), some are left out (sub_beg < 0)."""
    from longcalld_b200 import synth
    for r in synth.make_regions(mbp, tech, seed=seed):
        for hap in (1, 2):
            seqs = [np.asarray(s, np.uint8) for s, h in zip(r.reads, r.read_hap) if h == hap and len(s) > 0]
            if len(seqs) < 3 or len(seqs[0]) < 40 or max(len(s) for s in seqs) > max_len:
                continue
            L0 = len(seqs[0]); out, sb, se = [seqs[0]], [0], [0]
            for s in seqs[1:]:
                u = rng.random()
                if u < 0.45 or len(s) < 30: out.append(s); sb.append(0); se.append(0); continue
                if u < 0.5: out.append(s); sb.append(-1); se.append(-1); continue
                a = int(rng.integers(0, len(s) // 2)) if u > 0.7 else 0                 # cut on the left
                b = int(rng.integers(len(s) // 2 + 8, len(s))) if u < 0.9 else len(s)    # cut on the right
                rb = min(max(1, int(a * L0 / len(s)) + 1 + int(rng.integers(-3, 4))), L0 - 1)     # 1-based positions in the first read
                re_ = min(max(rb + 1, int(b * L0 / len(s)) + int(rng.integers(-3, 4))), L0)
                out.append(s[a:b]); sb.append(rb + 1); se.append(re_ + 1)
            yield out, np.array(sb, np.int32), np.array(se, np.int32)


# ----------------------------------------------------------------------------- phasing (K4) helpers
class PhaseInput(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_vars", C.c_int32), ("target_var_cate", C.c_int32), ("is_ont", C.c_int32),
                ("ordered_read_ids", C.c_void_p), ("is_skipped", C.c_void_p), ("prof_start", C.c_void_p), ("prof_end", C.c_void_p),
                ("allele_off", C.c_void_p), ("alleles", C.c_void_p), ("var_cate", C.c_void_p), ("var_type", C.c_void_p),
                ("is_hp_indel", C.c_void_p), ("n_uniq_alles", C.c_void_p), ("alle_covs", C.c_void_p), ("total_cov", C.c_void_p),
                ("pos", C.c_void_p)]


class PhaseOutput(C.Structure):
    _fields_ = [("haps", C.c_void_p), ("phase_sets", C.c_void_p), ("hap_to_cons_alle", C.c_void_p), ("hap_to_alle_profile", C.c_void_p),
                ("var_phase_set", C.c_void_p), ("n_clean_agree_snps", C.c_void_p), ("n_clean_conflict_snps", C.c_void_p)]


from longcalld_b200.synth import CATE, CATE_CLEAN, CATE_GERMLINE, make_phase_chunk  # noqa: E402,F401  (the generator lives with the workloads)
PHASE_IN_FIELDS = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("prof_start", np.int32), ("prof_end", np.int32),
                   ("allele_off", np.int64), ("alleles", np.int8), ("var_cate", np.int32), ("var_type", np.int32),
                   ("is_hp_indel", np.int32), ("n_uniq_alles", np.int32), ("alle_covs", np.int32), ("total_cov", np.int32), ("pos", np.int64))


def phase(lib, fn, d, target, is_ont=0):
    """Run an assign-hap implementation with the flat-array interface -> dict of output arrays."""
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in PHASE_IN_FIELDS}
    inp = PhaseInput(d["n_reads"], d["n_vars"], target, is_ont, *[keep[k].ctypes.data for k, _ in PHASE_IN_FIELDS])
    nr, nv = d["n_reads"], d["n_vars"]
    out = dict(haps=np.full(nr + 1, -9, np.int32), phase_sets=np.full(nr + 1, -9, np.int64), hap_to_cons_alle=np.full(3 * nv + 3, -9, np.int32),
               hap_to_alle_profile=np.full(12 * nv + 12, -9, np.int32), var_phase_set=np.full(nv + 1, -9, np.int64),
               n_clean_agree_snps=np.full(nr + 1, -9, np.int32), n_clean_conflict_snps=np.full(nr + 1, -9, np.int32))
    o = PhaseOutput(*[out[k].ctypes.data for k, _ in PhaseOutput._fields_])
    rc = getattr(lib, fn)(C.byref(inp), C.byref(o))
    assert rc == 0
    return out


# ----------------------------------------------------------------------------- pileup (K2) helpers
PILEUP_IN_FIELDS = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("read_beg", np.int64), ("read_end", np.int64),
                    ("read_is_rev", np.uint8), ("digar_first", np.int64), ("n_digar", np.int32), ("qual_off", np.int64), ("qual", np.uint8),
                    ("digar_pos", np.int64), ("digar_type", np.int8), ("digar_len", np.int32), ("digar_qi", np.int32),
                    ("digar_low_qual", np.uint8), ("digar_alt_off", np.int64), ("digar_alt", np.uint8),
                    ("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("site_alt_len", np.int32),
                    ("site_alt_off", np.int64), ("site_alt", np.uint8))


class PileupInput(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_sites", C.c_int32), ("min_bq", C.c_int32), ("min_sv_len", C.c_int32)] + \
               [(k, C.c_void_p) for k, _ in PILEUP_IN_FIELDS]


class PileupOutput(C.Structure):
    _fields_ = [("site_counts", C.c_void_p)]


def pileup(lib, fn, d):
    """Run a per-site coverage implementation with the flat-array interface -> (n_sites, 8) counts."""
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in PILEUP_IN_FIELDS}
    inp = PileupInput(d["n_reads"], d["n_sites"], d["min_bq"], d["min_sv_len"], *[keep[k].ctypes.data for k, _ in PILEUP_IN_FIELDS])
    counts = np.full((d["n_sites"] + 1, 8), -9, np.int32)
    out = PileupOutput(counts.ctypes.data)
    rc = getattr(lib, fn)(C.byref(inp), C.byref(out))
    assert rc == 0
    return counts[:d["n_sites"]]


# ----------------------------------------------------------------------------- read x variant profile (K3) helpers
PROFILE_EX_FIELDS = (("var_cate", np.int32), ("nreg_first", np.int64), ("n_nreg", np.int32), ("nreg_beg", np.int64), ("nreg_end", np.int64))


class ProfileExtra(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _ in PROFILE_EX_FIELDS]


class ProfileOutput(C.Structure):
    _fields_ = [("prof_start", C.c_void_p), ("prof_end", C.c_void_p), ("allele_off", C.c_void_p), ("alleles", C.c_void_p), ("alt_qi", C.c_void_p),
                ("alleles_cap", C.c_int64), ("n_alleles", C.c_int64)]


def profile_capacity(d):
    """Entries that always suffice: per read the variants positioned inside its span, plus 2."""
    pos = np.asarray(d["site_pos"][:d["n_sites"]])
    lo = np.searchsorted(pos, np.asarray(d["read_beg"]) - 1, side="left"); hi = np.searchsorted(pos, np.asarray(d["read_end"]), side="right")
    return int((hi - lo + 2).sum()) + 8


def read_var_profile(lib, fn, d):
    """-> per read (start, end, alleles tuple, alt_qi tuple): the rows of read_var_profile_t."""
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in PILEUP_IN_FIELDS + PROFILE_EX_FIELDS}
    inp = PileupInput(d["n_reads"], d["n_sites"], d["min_bq"], d["min_sv_len"], *[keep[k].ctypes.data for k, _ in PILEUP_IN_FIELDS])
    ex = ProfileExtra(*[keep[k].ctypes.data for k, _ in PROFILE_EX_FIELDS])
    nr, cap = d["n_reads"], profile_capacity(d)
    ps, pe, ao = np.full(nr + 1, -7, np.int32), np.full(nr + 1, -7, np.int32), np.zeros(nr + 1, np.int64)
    al, qi = np.full(cap, -7, np.int8), np.full(cap, -7, np.int32)
    out = ProfileOutput(ps.ctypes.data, pe.ctypes.data, ao.ctypes.data, al.ctypes.data, qi.ctypes.data, cap, 0)
    rc = getattr(lib, fn)(C.byref(inp), C.byref(ex), C.byref(out))
    assert rc == 0, rc
    rows = []
    for r in range(nr):
        n = max(0, int(pe[r]) - int(ps[r]) + 1) if ps[r] >= 0 else 0
        rows.append((int(ps[r]), int(pe[r]), tuple(al[ao[r]:ao[r] + n].tolist()), tuple(qi[ao[r]:ao[r] + n].tolist())))
    return rows


# ----------------------------------------------------------------------------- difference lists from =/X CIGARs (K1) helpers
DIGAR_IN_FIELDS = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("read_pos0", np.int64), ("read_is_rev", np.uint8),
                   ("is_palindrome", np.uint8), ("n_cigar", np.int32), ("cigar_off", np.int64), ("cigar", np.uint32),
                   ("l_qseq", np.int32), ("seq_off", np.int64), ("bseq", np.uint8), ("qual_off", np.int64), ("qual", np.uint8))
DIGAR_SCALARS = (("n_reads", C.c_int32), ("min_bq", C.c_int32), ("noisy_reg_max_xgaps", C.c_int32), ("noisy_reg_slide_win", C.c_int32),
                 ("end_clip_reg", C.c_int32), ("end_clip_reg_flank_win", C.c_int32), ("max_noisy_frac_per_read", C.c_double),
                 ("max_var_ratio_per_read", C.c_double), ("whole_ref_len", C.c_int64), ("reg_beg", C.c_int64), ("reg_end", C.c_int64))


class DigarInput(C.Structure):
    _fields_ = list(DIGAR_SCALARS) + [(k, C.c_void_p) for k, _ in DIGAR_IN_FIELDS]


DIGAR_OUT_FIELDS = (("skip", np.uint8, "r"), ("read_beg", np.int64, "r"), ("read_end", np.int64, "r"), ("digar_first", np.int64, "r"), ("n_digar", np.int32, "r"),
                    ("digar_pos", np.int64, "d"), ("digar_type", np.int8, "d"), ("digar_len", np.int32, "d"), ("digar_qi", np.int32, "d"),
                    ("digar_low_qual", np.uint8, "d"), ("digar_alt_off", np.int64, "d"), ("digar_alt", np.uint8, "a"))


class DigarOutput(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _, _ in DIGAR_OUT_FIELDS] + [("digar_cap", C.c_int64), ("alt_cap", C.c_int64)] + \
               [(k, C.c_void_p) for k in ("nreg_first", "n_nreg", "nreg_beg", "nreg_end", "nreg_label")] + [("nreg_cap", C.c_int64)] + \
               [(k, C.c_void_p) for k in ("cnreg_beg", "cnreg_end", "cnreg_label")] + [("cnreg_cap", C.c_int64), ("n_cnreg", C.c_int64)] + \
               [("qual_counts", C.c_void_p), ("n_digar_total", C.c_int64), ("n_alt_total", C.c_int64), ("n_nreg_total", C.c_int64)]


def digar_capacity(d):
    """(digar1_t records, alt bases, intervals) that always suffice for a chunk: from the CIGAR words alone."""
    cig = np.asarray(d["cigar"], np.uint32); op = cig & 15; ln = (cig >> 4).astype(np.int64)
    n_dig = int(np.where(op == 8, ln, 1).sum()); n_alt = int(ln[(op == 8) | (op == 1)].sum())
    return n_dig + 8, n_alt + 8, n_dig + 2 * d["n_reads"] + 8


def digar_input(d):
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in DIGAR_IN_FIELDS}
    inp = DigarInput(*[d[k] for k, _ in DIGAR_SCALARS], *[keep[k].ctypes.data for k, _ in DIGAR_IN_FIELDS])
    return inp, keep


def collect_digar_raw(lib, fn, d, mid_args=(), cap_like=None, slack=0):
    """Run an implementation of the difference-list pass -> its flat output arrays (the layout lcd_digar_output_t / the GPU plan's fetch() use)."""
    inp, keep = digar_input(d)
    nr = d["n_reads"]; dc, ac, rc_ = digar_capacity(cap_like if cap_like is not None else d)        # (cap_like: a chunk whose CIGARs size the outputs)
    dc += slack; ac += slack; rc_ += slack
    size = {"r": nr + 1, "d": dc, "a": ac}
    buf = {k: np.full(size[w], 77, t) for k, t, w in DIGAR_OUT_FIELDS}
    nf, nn = np.zeros(nr + 1, np.int64), np.zeros(nr + 1, np.int32)
    nb, ne, nl = np.zeros(rc_, np.int64), np.zeros(rc_, np.int64), np.zeros(rc_, np.int32)
    cb, ce, cl = np.zeros(rc_, np.int64), np.zeros(rc_, np.int64), np.zeros(rc_, np.int32)
    qc = np.zeros(256, np.int64)
    out = DigarOutput(*[buf[k].ctypes.data for k, _, _ in DIGAR_OUT_FIELDS], dc, ac, nf.ctypes.data, nn.ctypes.data, nb.ctypes.data, ne.ctypes.data,
                      nl.ctypes.data, rc_, cb.ctypes.data, ce.ctypes.data, cl.ctypes.data, rc_, 0, qc.ctypes.data, 0, 0, 0)
    rc = getattr(lib, fn)(C.byref(inp), *mid_args, C.byref(out))
    assert rc == 0, rc
    o = dict(buf)
    o.update(nreg_first=nf, n_nreg=nn, nreg_beg=nb, nreg_end=ne, nreg_label=nl, cnreg_beg=cb, cnreg_end=ce, cnreg_label=cl, n_cnreg=int(out.n_cnreg), qual_counts=qc,
             n_digar_total=int(out.n_digar_total), n_alt_total=int(out.n_alt_total), n_nreg_total=int(out.n_nreg_total))
    return o


def collect_digar(lib, fn, d, mid_args=(), cap_like=None, slack=0):
    """Run an implementation of the =/X difference-list pass -> dict with per-read records (in read-id order, layout independent).
    mid_args: extra ctypes arguments between the input and the output struct (the MD-tag shim takes the tags there)."""
    o = collect_digar_raw(lib, fn, d, mid_args, cap_like, slack)
    nr = d["n_reads"]; buf = o
    nf, nn, nb, ne, nl, cb, ce, cl = (o[k] for k in ("nreg_first", "n_nreg", "nreg_beg", "nreg_end", "nreg_label", "cnreg_beg", "cnreg_end", "cnreg_label"))
    reads = {}
    for i in range(nr):
        r = int(d["ordered_read_ids"][i])
        if d["is_skipped"][r]: continue
        f, n = int(buf["digar_first"][r]), int(buf["n_digar"][r])
        ev = []
        for k in range(f, f + n):
            t, ln = int(buf["digar_type"][k]), int(buf["digar_len"][k])
            alt = bytes(buf["digar_alt"][buf["digar_alt_off"][k]:buf["digar_alt_off"][k] + ln]) if t in (1, 8) else b""
            ev.append((int(buf["digar_pos"][k]), t, ln, int(buf["digar_qi"][k]), int(buf["digar_low_qual"][k]), alt))
        iv = [(int(nb[k]), int(ne[k]), int(nl[k])) for k in range(int(nf[r]), int(nf[r]) + int(nn[r]))]
        reads[r] = (int(buf["skip"][r]), int(buf["read_beg"][r]), int(buf["read_end"][r]), ev, iv)
    chunk_iv = [(int(cb[k]), int(ce[k]), int(cl[k])) for k in range(o["n_cnreg"])]
    return dict(reads=reads, chunk_noisy=chunk_iv, qual_counts=o["qual_counts"].tolist(), totals=(o["n_digar_total"], o["n_alt_total"], o["n_nreg_total"]))


def digar_digest(res):
    """sha1 over the per-read records (read-id order) + histogram of a collect_digar() result: what the fixtures store."""
    import hashlib
    h = hashlib.sha1()
    for r in sorted(res["reads"]):
        h.update(repr((r, res["reads"][r])).encode())
    h.update(repr(res["qual_counts"]).encode())
    return h.hexdigest()


def digar_case_from_json(j):
    d = {k: j[k] for k, _ in DIGAR_SCALARS}
    for k, t in DIGAR_IN_FIELDS: d[k] = np.array(j[k], dtype=t)
    return d


def digar_case_to_json(d):
    j = {k: (float(d[k]) if isinstance(d[k], float) else int(d[k])) for k, _ in DIGAR_SCALARS}
    for k, t in DIGAR_IN_FIELDS: j[k] = np.asarray(d[k]).tolist()
    return j


# ----------------------------------------------------------------------------- candidate-site list (a3) helpers
class SitesOutput(C.Structure):
    _fields_ = [("site_pos", C.c_void_p), ("site_type", C.c_void_p), ("site_ref_len", C.c_void_p), ("site_alt_len", C.c_void_p), ("site_src", C.c_void_p),
                ("cap", C.c_int64), ("n_sites", C.c_int64)]


def collect_sites(lib, fn, d, reg_beg, reg_end, src_is_offset=False, raw=False):
    """Run an implementation of collect_all_cand_var_sites over a chunk in lcd_pileup_input_t layout (sites unused)
    -> [(pos, type, ref_len, alt_len, alt bytes)] in output order."""
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in PILEUP_IN_FIELDS}
    inp = PileupInput(d["n_reads"], 0, d["min_bq"], d["min_sv_len"], *[keep[k].ctypes.data for k, _ in PILEUP_IN_FIELDS])
    cap = int(np.asarray(d["n_digar"][:d["n_reads"]]).sum()) + 8
    pos, typ, rl, al, src = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int64)
    out = SitesOutput(pos.ctypes.data, typ.ctypes.data, rl.ctypes.data, al.ctypes.data, src.ctypes.data, cap, 0)
    getattr(lib, fn).argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    rc = getattr(lib, fn)(C.byref(inp), reg_beg, reg_end, C.byref(out))
    assert rc == 0, rc
    if raw:
        return dict(n_sites=int(out.n_sites), site_pos=pos, site_type=typ, site_ref_len=rl, site_alt_len=al, site_src=src)
    return sites_view(d, pos, typ, rl, al, src, out.n_sites, src_is_offset)


def sites_view(d, pos, typ, rl, al, src, n, src_is_offset=False):
    res = []
    for i in range(int(n)):
        a0 = int(src[i]) if src_is_offset else int(d["digar_alt_off"][int(src[i])])
        alt = bytes(np.asarray(d["digar_alt"][a0:a0 + int(al[i])], np.uint8)) if typ[i] != 2 else b""
        res.append((int(pos[i]), int(typ[i]), int(rl[i]), int(al[i]), alt))
    return res


def sites_case_from_json(c):
    d = {k: (np.array(v, dtype=dict(PILEUP_IN_FIELDS)[k]) if k in dict(PILEUP_IN_FIELDS) else v) for k, v in c["in"].items()}
    for k, t in PILEUP_IN_FIELDS:
        if k.startswith("site_"): d[k] = np.zeros(1, t)
    return d, c["reg"], [(p_, t_, r_, a_, bytes.fromhex(h)) for p_, t_, r_, a_, h in c["sites"]]


# ----------------------------------------------------------------------------- per-site category (K2b) helpers
CLASSIFY_FIELDS = (("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("site_alt_len", np.int32),
                   ("site_alt_off", np.int64), ("site_alt", np.uint8), ("site_counts", np.int32))


class ClassifyInput(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("min_dp", C.c_int32), ("min_alt_dp", C.c_int32), ("max_xgaps", C.c_int32), ("is_ont", C.c_int32), ("pad", C.c_int32),
                ("min_af", C.c_double), ("max_af", C.c_double), ("ref_beg", C.c_int64), ("ref_end", C.c_int64), ("ref_seq", C.c_void_p)] + \
               [(k, C.c_void_p) for k, _ in CLASSIFY_FIELDS]


def classify_input(d):
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in CLASSIFY_FIELDS}
    keep["ref_seq"] = np.ascontiguousarray(d["ref_seq"], dtype=np.uint8)
    inp = ClassifyInput(d["n_sites"], d["min_dp"], d["min_alt_dp"], d["max_xgaps"], d["is_ont"], 0, d["min_af"], d["max_af"], d["ref_beg"], d["ref_end"],
                        keep["ref_seq"].ctypes.data, *[keep[k].ctypes.data for k, _ in CLASSIFY_FIELDS])
    return inp, keep


def classify(lib, fn, d):
    """Run an implementation of the per-site category over one chunk -> int32 array [n_sites]."""
    inp, keep = classify_input(d)
    out = np.full(d["n_sites"] + 1, -7, np.int32)
    rc = getattr(lib, fn)(C.byref(inp), out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return out[:d["n_sites"]]


CLASSIFY_SCALARS = ("n_sites", "min_dp", "min_alt_dp", "max_xgaps", "is_ont", "min_af", "max_af", "ref_beg", "ref_end")


def classify_case_to_json(d):
    j = {k: (float(d[k]) if isinstance(d[k], float) else int(d[k])) for k in CLASSIFY_SCALARS}
    j["ref_seq"] = bytes(np.asarray(d["ref_seq"], np.uint8)).decode()
    for k, _ in CLASSIFY_FIELDS: j[k] = np.asarray(d[k]).reshape(-1).tolist()
    return j


def classify_case_from_json(j):
    d = {k: j[k] for k in CLASSIFY_SCALARS}
    d["ref_seq"] = np.frombuffer(j["ref_seq"].encode(), np.uint8)
    for k, t in CLASSIFY_FIELDS: d[k] = np.array(j[k], dtype=t)
    d["site_counts"] = d["site_counts"].reshape(-1, 8)
    return d


# ----------------------------------------------------------------------------- noisy-region set (a5, second half) helpers
NOISYREG_FIELDS = (("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("var_cate", np.int32))
NOISYREG_READ_FIELDS = (("is_skipped", np.uint8), ("read_beg", np.int64), ("read_end", np.int64), ("digar_first", np.int64), ("n_digar", np.int32),
                        ("digar_pos", np.int64), ("digar_type", np.int8), ("digar_len", np.int32), ("nreg_first", np.int64), ("n_nreg", np.int32),
                        ("nreg_beg", np.int64), ("nreg_end", np.int64))


class NoisyRegInput(C.Structure):
    _fields_ = [("reg_beg", C.c_int64), ("reg_end", C.c_int64), ("min_alt_dp", C.c_int32), ("noisy_reg_flank_len", C.c_int32), ("is_ont", C.c_int32), ("pad", C.c_int32),
                ("min_af", C.c_double), ("n_sites", C.c_int32), ("n_reads", C.c_int32)] + [(k, C.c_void_p) for k, _ in NOISYREG_FIELDS] + \
               [("n_cnreg", C.c_int64), ("cnreg_beg", C.c_void_p), ("cnreg_end", C.c_void_p), ("cnreg_label", C.c_void_p),
                ("n_low", C.c_int64), ("low_beg", C.c_void_p), ("low_end", C.c_void_p)] + [(k, C.c_void_p) for k, _ in NOISYREG_READ_FIELDS]


class NoisyRegOutput(C.Structure):
    _fields_ = [("var_cate", C.c_void_p), ("keep", C.c_void_p), ("reg_beg", C.c_void_p), ("reg_end", C.c_void_p), ("reg_label", C.c_void_p), ("reg_cap", C.c_int64), ("n_regs", C.c_int64)]


def noisyreg_input(d):
    keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in NOISYREG_FIELDS + NOISYREG_READ_FIELDS}
    for k, t in (("cnreg_beg", np.int64), ("cnreg_end", np.int64), ("cnreg_label", np.int32), ("low_beg", np.int64), ("low_end", np.int64)):
        keep[k] = np.ascontiguousarray(np.append(np.asarray(d[k]), 0), dtype=t)
    inp = NoisyRegInput(d["reg_beg"], d["reg_end"], d["min_alt_dp"], d["noisy_reg_flank_len"], d["is_ont"], 0, d["min_af"], d["n_sites"], d["n_reads"],
                        *[keep[k].ctypes.data for k, _ in NOISYREG_FIELDS], d["n_cnreg"], keep["cnreg_beg"].ctypes.data, keep["cnreg_end"].ctypes.data,
                        keep["cnreg_label"].ctypes.data, d["n_low"], keep["low_beg"].ctypes.data, keep["low_end"].ctypes.data,
                        *[keep[k].ctypes.data for k, _ in NOISYREG_READ_FIELDS])
    return inp, keep


def noisyreg_cap(d):
    return int(d["n_cnreg"]) + int(d["n_sites"]) + 8


def noisy_regs(lib, fn, d):
    """-> (kept sites [(pos, type, ref_len, cate)], regions [(st, en, label)], working categories)"""
    inp, keep = noisyreg_input(d)
    n, cap = d["n_sites"], noisyreg_cap(d)
    cate, kp = np.full(n + 1, -7, np.int32), np.zeros(n + 1, np.uint8)
    rb, re_, rl = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap, np.int32)
    out = NoisyRegOutput(cate.ctypes.data, kp.ctypes.data, rb.ctypes.data, re_.ctypes.data, rl.ctypes.data, cap, 0)
    rc = getattr(lib, fn)(C.byref(inp), C.byref(out))
    assert rc == 0, rc
    kept = [(int(d["site_pos"][i]), int(d["site_type"][i]), int(d["site_ref_len"][i]), int(cate[i])) for i in np.nonzero(kp[:n])[0]]
    return kept, [(int(rb[k]), int(re_[k]), int(rl[k])) for k in range(out.n_regs)], cate[:n].copy()


def ref_noisy_regs(ref, ci, d):
    """The unmodified pre_process_noisy_regs + classify_cand_vars through the shim -> (kept sites, regions)"""
    cinp, ckeep = classify_input(ci)
    inp, keep = noisyreg_input(d)
    n, cap = d["n_sites"], noisyreg_cap(d)
    kpos, ktyp, krl, kc = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    nk = C.c_int32(0)
    rb, re_, rl = np.zeros(cap, np.int64), np.zeros(cap, np.int64), np.zeros(cap, np.int32)
    out = NoisyRegOutput(None, None, rb.ctypes.data, re_.ctypes.data, rl.ctypes.data, cap, 0)
    rc = ref.ref_noisy_regs(C.byref(cinp), C.byref(inp), kpos.ctypes.data_as(C.c_void_p), ktyp.ctypes.data_as(C.c_void_p), krl.ctypes.data_as(C.c_void_p),
                            kc.ctypes.data_as(C.c_void_p), C.byref(nk), C.byref(out))
    assert rc == 0, rc
    return [(int(kpos[i]), int(ktyp[i]), int(krl[i]), int(kc[i])) for i in range(nk.value)], [(int(rb[k]), int(re_[k]), int(rl[k])) for k in range(out.n_regs)]


def noisyreg_case(orc, d, seed, is_ont=0, low_every=400, min_sv_len=50):
    """The inputs of the noisy-region set for the chunk `d` (a K1 input): K1 -> K1b -> K2 -> K2b run by the oracle, plus synthetic low-complexity
    intervals (sdust's output in the reference: here random short intervals, a third of them planted on candidate sites and noisy intervals).
    -> (noisy-region input dict, classify input dict)"""
    from longcalld_b200 import synth
    rng = np.random.default_rng(seed)
    o = collect_digar_raw(orc, "lcd_oracle_collect_digar_eqx", d)
    raw = collect_sites(orc, "lcd_oracle_collect_sites", synth.pileup_input_from_digar(d, o, synth.empty_site_list(min_sv_len)), int(d["reg_beg"]), int(d["reg_end"]), raw=True)
    st = synth.site_list_from_sites(o, raw, min_sv_len)
    pin = synth.pileup_input_from_digar(d, o, st)
    counts = pileup(orc, "lcd_oracle_collect_cand_vars", pin)
    ci = synth.classify_input_from_sites(d, st, counts, seed, is_ont=is_ont)
    cate = classify(orc, "lcd_oracle_classify_sites", ci)
    n = st["n_sites"]
    span = int(d["reg_end"]) - int(d["reg_beg"])
    k = max(1, span // low_every)
    lb = rng.integers(int(d["reg_beg"]) - 200, int(d["reg_end"]) + 200, k)
    if n: lb[::3] = np.asarray(st["site_pos"][:n])[rng.integers(0, n, len(lb[::3]))] - rng.integers(0, 6, len(lb[::3]))
    if o["n_cnreg"]: lb[1::5] = o["cnreg_beg"][:o["n_cnreg"]][rng.integers(0, o["n_cnreg"], len(lb[1::5]))] - rng.integers(-3, 12, len(lb[1::5]))
    lb = np.sort(lb); le = lb + rng.integers(4, 40, k)
    nr = d["n_reads"]
    case = dict(reg_beg=int(d["reg_beg"]), reg_end=int(d["reg_end"]), min_alt_dp=2, noisy_reg_flank_len=10, is_ont=int(is_ont), min_af=0.20, n_sites=n, n_reads=nr,
                site_pos=st["site_pos"], site_type=st["site_type"], site_ref_len=st["site_ref_len"], var_cate=np.append(cate, 0),
                n_cnreg=o["n_cnreg"], cnreg_beg=o["cnreg_beg"][:o["n_cnreg"]], cnreg_end=o["cnreg_end"][:o["n_cnreg"]], cnreg_label=o["cnreg_label"][:o["n_cnreg"]],
                n_low=k, low_beg=lb, low_end=le, is_skipped=np.maximum(np.asarray(d["is_skipped"][:nr]), o["skip"][:nr]),
                **{f: o[f] for f in ("read_beg", "read_end", "digar_first", "n_digar", "digar_pos", "digar_type", "digar_len", "nreg_first", "n_nreg", "nreg_beg", "nreg_end")})
    return case, ci


NOISYREG_SCALARS = ("reg_beg", "reg_end", "min_alt_dp", "noisy_reg_flank_len", "is_ont", "min_af", "n_sites", "n_reads", "n_cnreg", "n_low")
NOISYREG_ARRAYS = dict(list(NOISYREG_FIELDS + NOISYREG_READ_FIELDS) + [("cnreg_beg", np.int64), ("cnreg_end", np.int64), ("cnreg_label", np.int32), ("low_beg", np.int64), ("low_end", np.int64)])


def noisyreg_case_to_json(d):
    """Only what the stage reads: records other than X / I / D keep their place (the per-read lists are position-sorted walks)."""
    out = {k: (float(d[k]) if k == "min_af" else int(d[k])) for k in NOISYREG_SCALARS}
    nr = d["n_reads"]
    nd = int(max([int(d["digar_first"][r]) + int(d["n_digar"][r]) for r in range(nr)] + [0])); nn = int(max([int(d["nreg_first"][r]) + int(d["n_nreg"][r]) for r in range(nr)] + [0]))
    size = dict(site_pos=d["n_sites"], site_type=d["n_sites"], site_ref_len=d["n_sites"], var_cate=d["n_sites"], is_skipped=nr, read_beg=nr, read_end=nr, digar_first=nr, n_digar=nr,
                digar_pos=nd, digar_type=nd, digar_len=nd, nreg_first=nr, n_nreg=nr, nreg_beg=nn, nreg_end=nn, cnreg_beg=d["n_cnreg"], cnreg_end=d["n_cnreg"], cnreg_label=d["n_cnreg"],
                low_beg=d["n_low"], low_end=d["n_low"])
    for k, n in size.items():
        out[k] = np.asarray(d[k])[:n].astype(np.int64).tolist()
    return out


def noisyreg_case_from_json(j):
    d = {k: j[k] for k in NOISYREG_SCALARS}
    for k, t in NOISYREG_ARRAYS.items():
        d[k] = np.array(j[k] + [0], dtype=t)
    return d


# ----------------------------------------------------------------------------- sdust helpers
def sdust(lib, fn, seq, T=5, W=20):
    """-> [(beg, end)] 0-based half-open low-complexity intervals of an ASCII sequence"""
    s = np.ascontiguousarray(np.frombuffer(bytes(seq), np.uint8) if not isinstance(seq, np.ndarray) else seq, dtype=np.uint8)
    cap = len(s) // 2 + 16
    b, e = np.zeros(cap, np.int64), np.zeros(cap, np.int64)
    f = getattr(lib, fn); f.restype = C.c_int
    n = f(s.ctypes.data_as(C.c_void_p), C.c_int(len(s)), C.c_int(T), C.c_int(W), b.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    assert n >= 0, n
    return list(zip(b[:n].tolist(), e[:n].tolist()))


def sdust_fixture_cases():
    """tests/golden/sdust_lcd.json.gz: (sequence, T, W, intervals of the UNMODIFIED sdust()) -- generated by tests/golden/make_golden.py"""
    import base64, gzip, json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "sdust_lcd.json.gz")) as f:
        d = json.load(f)
    return [(np.frombuffer(base64.b64decode(c["seq"]), np.uint8).copy(), int(c["T"]), int(c["W"]), [tuple(x) for x in c["iv"]]) for c in d["cases"]]


def sdust_sequence(rng, n, lc_every=120, n_frac=0.002):
    """ASCII reference-like sequence with planted homopolymers / short tandem repeats / low-entropy stretches, a few N runs and lower-case bases"""
    s = rng.integers(0, 4, n).astype(np.uint8)
    p = int(rng.integers(0, lc_every))
    while p < n - 80:
        kind = int(rng.integers(0, 4))
        if kind == 0: L = int(rng.integers(4, 40)); s[p:p + L] = rng.integers(0, 4)
        elif kind == 1: u = int(rng.integers(2, 7)); c = int(rng.integers(3, 15)); s[p:p + u * c] = np.tile(rng.integers(0, 4, u).astype(np.uint8), c)[:max(0, min(u * c, n - p))]
        elif kind == 2: L = int(rng.integers(10, 60)); s[p:p + L] = rng.choice(np.array([0, 3], np.uint8), L)      # AT-rich
        p += int(rng.integers(5, 2 * lc_every))
    a = np.frombuffer(b"ACGT", np.uint8)[s].copy()
    for q in rng.integers(0, n, max(1, int(n * n_frac))): a[q:q + int(rng.integers(1, 30))] = ord("N")
    low = rng.random(n) < 0.05; a[low & (a != ord("N"))] |= 0x20
    return a
