"""Shared helpers for the test-suite: ctypes bindings for the oracle port (oracle/liblcd_oracle.so),
the reference shim (oracle/_ref/libref_shim.so, built from /root/reference where present) and the
product C-ABI (longcalld_b200/csrc/liblcd_gpu.so), plus seeded workload generators.

TEST INFRASTRUCTURE: nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


CATE = {"CLEAN_HET_SNP": 0x004, "CLEAN_HET_INDEL": 0x008, "CLEAN_HOM_VAR": 0x080, "NOISY_CAND_HET_VAR": 0x100,
        "NOISY_CAND_HOM_VAR": 0x200, "LOW_COV_VAR": 0x001, "CAND_SOMATIC_VAR": 0x040}
CATE_CLEAN = 0x004 | 0x008 | 0x080
CATE_GERMLINE = CATE_CLEAN | 0x100 | 0x200
PHASE_IN_FIELDS = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("prof_start", np.int32), ("prof_end", np.int32),
                   ("allele_off", np.int64), ("alleles", np.int8), ("var_cate", np.int32), ("var_type", np.int32),
                   ("is_hp_indel", np.int32), ("n_uniq_alles", np.int32), ("alle_covs", np.int32), ("total_cov", np.int32), ("pos", np.int64))


def make_phase_chunk(rng, n_vars=60, n_reads=200, err=0.02, tech="hifi", shuffle_order=False):
    """A synthetic chunk for the read -> haplotype assignment: a diploid truth over sorted candidate variants of mixed
    categories, reads sampled from the two haplotypes spanning contiguous variant ranges (some skipped, some empty,
    some carrying noise / low-quality calls), with coverage counts derived from the reads."""
    pos = np.sort(rng.choice(np.arange(1000, 1000 + 400 * max(n_vars, 1)), size=n_vars, replace=False)).astype(np.int64)
    cats = np.array([CATE["CLEAN_HET_SNP"], CATE["CLEAN_HET_INDEL"], CATE["CLEAN_HOM_VAR"], CATE["NOISY_CAND_HET_VAR"],
                     CATE["NOISY_CAND_HOM_VAR"], CATE["LOW_COV_VAR"], CATE["CAND_SOMATIC_VAR"]], dtype=np.int32)
    var_cate = rng.choice(cats, size=n_vars, p=[0.45, 0.12, 0.12, 0.12, 0.05, 0.08, 0.06]).astype(np.int32)
    var_type = np.where(var_cate == CATE["CLEAN_HET_SNP"], 8, rng.choice([8, 1, 2], size=n_vars)).astype(np.int32)
    var_type[var_cate == CATE["CLEAN_HET_INDEL"]] = rng.choice([1, 2], size=int((var_cate == CATE["CLEAN_HET_INDEL"]).sum()))
    is_hp = ((var_type != 8) & (rng.random(n_vars) < 0.3)).astype(np.int32)
    is_hom = (var_cate == CATE["CLEAN_HOM_VAR"]) | (var_cate == CATE["NOISY_CAND_HOM_VAR"])
    alt_hap = rng.integers(1, 3, n_vars)                                   # which haplotype carries the alt allele of a het variant
    n_uniq = np.where(rng.random(n_vars) < 0.1, 3, 2).astype(np.int32)
    starts, ends, haps_true = [], [], []
    for _ in range(n_reads):
        if n_vars == 0 or rng.random() < 0.05:
            starts.append(-1); ends.append(-2)
        else:
            s = int(rng.integers(0, n_vars)); L = int(np.clip(rng.poisson(12 if tech == "hifi" else 25), 1, n_vars))
            starts.append(s); ends.append(min(n_vars - 1, s + L - 1))
        haps_true.append(int(rng.integers(1, 3)))
    order = np.argsort(np.array(starts), kind="stable")                    # reads arrive position-sorted
    starts = np.array(starts, dtype=np.int32)[order]; ends = np.array(ends, dtype=np.int32)[order]; haps_true = np.array(haps_true)[order]
    allele_off = np.zeros(n_reads, dtype=np.int64); alle = []
    alle_covs = np.zeros((n_vars, 4), dtype=np.int32)
    is_skipped = (rng.random(n_reads) < 0.04).astype(np.uint8)
    for r in range(n_reads):
        allele_off[r] = len(alle)
        for v in range(starts[r], ends[r] + 1) if starts[r] >= 0 else ():
            a = 1 if (is_hom[v] or alt_hap[v] == haps_true[r]) else 0
            u = rng.random()
            if u < err: a = 1 - a
            elif u < err + 0.02: a = -1
            elif u < err + 0.03: a = -2
            elif u < err + 0.035 and n_uniq[v] == 3: a = 2
            alle.append(a)
            if a >= 0 and not is_skipped[r]: alle_covs[v, a] += 1
    ordered = np.arange(n_reads, dtype=np.int32)
    if shuffle_order: rng.shuffle(ordered)
    d = dict(n_reads=n_reads, n_vars=n_vars, ordered_read_ids=ordered, is_skipped=is_skipped, prof_start=starts, prof_end=ends,
             allele_off=allele_off, alleles=np.array(alle + [0], dtype=np.int8), var_cate=var_cate, var_type=var_type, is_hp_indel=is_hp,
             n_uniq_alles=n_uniq, alle_covs=np.ascontiguousarray(alle_covs), total_cov=alle_covs.sum(axis=1).astype(np.int32), pos=pos)
    return d


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
