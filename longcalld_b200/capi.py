"""ctypes binding of include/lcd_gpu.h (liblcd_gpu.so).  Plain pointers and sizes only.

Every wrapper raises LcdGpuError with lcd_gpu_last_error() on a non-zero return: the CUDA path is
the only path (no CPU fallback, no oracle import).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.environ.get("LCD_GPU_SO") or os.path.join(_CSRC, "liblcd_gpu.so")   # LCD_GPU_SO: debug builds (e.g. -DLCD_POA_TIMING)

HEUR_NONE, HEUR_ADAPTIVE, HEUR_ZDROP = 0, 1, 2
GAP_RIGHT_ALN, GAP_LEFT_ALN = 0, 1           # reference src/call_var_main.h (LONGCALLD_GAP_*_ALN)


class LcdGpuError(RuntimeError):
    pass


class WfaParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2", "affine2p", "heuristic",
        "min_wavefront_length", "max_distance_threshold", "zdrop", "steps_between_cutoffs")]


class WfaResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("status", "score", "n_ops", "end_v", "end_h")]


WFA_PARAMS_DTYPE = np.dtype([(n, np.int32) for n, _ in WfaParams._fields_])
WFA_RESULT_DTYPE = np.dtype([(n, np.int32) for n, _ in WfaResult._fields_])


def wfa_params(heuristic=HEUR_NONE, affine2p=1, plen=0, tlen=0, x=6, o1=6, e1=2, o2=24, e2=1):
    """The parameter points wfa_end2end_aln builds (reference src/align.c:379-408, src/align.h:21-26)."""
    p = (x, o1, e1, o2, e2, affine2p, heuristic, 10, 50, 0, 1)
    if heuristic == HEUR_ZDROP:
        p = (x, o1, e1, o2, e2, affine2p, heuristic, 10, 50, min(500, int(min(plen, tlen) * 0.1)), 100)
    return p


def lib_path():
    return _SO


def build_library(verbose=False):
    """Compile liblcd_gpu.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s" if not verbose else "-j1", "-C", _CSRC])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise LcdGpuError(f"{_SO} is missing: run `make -C longcalld_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(_SO)
    L.lcd_gpu_last_error.restype = C.c_char_p
    L.lcd_gpu_launch_count.restype = C.c_uint64
    L.lcd_gpu_init.argtypes = [C.c_int, C.c_size_t]
    L.lcd_gpu_stream.restype = C.c_void_p
    for fn in ("lcd_wfa_plan_create", "lcd_edlib_plan_create", "lcd_poa_plan_create", "lcd_phase_plan_create", "lcd_pileup_plan_create", "lcd_profile_plan_create", "lcd_digar_plan_create"):
        if hasattr(L, fn):
            getattr(L, fn).restype = C.c_void_p
    L.lcd_plan_run.argtypes = [C.c_void_p, C.c_void_p]
    L.lcd_plan_sync.argtypes = [C.c_void_p, C.c_void_p]
    L.lcd_plan_destroy.argtypes = [C.c_void_p]
    L.lcd_plan_work_units.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise LcdGpuError(f"{what} failed ({rc}): {lib().lcd_gpu_last_error().decode()}")


def init(device=0, pool_bytes=0):
    _check(lib().lcd_gpu_init(device, pool_bytes), "lcd_gpu_init")


def shutdown():
    lib().lcd_gpu_shutdown()


def stream():
    """The library stream as an integer cudaStream_t (wrap with torch.cuda.ExternalStream)."""
    return int(lib().lcd_gpu_stream() or 0)


def aux_stream():
    lib().lcd_gpu_aux_stream.restype = C.c_void_p
    return int(lib().lcd_gpu_aux_stream() or 0)


def set_thread_stream(s):
    """Default stream of the calling host thread's plans (0 / None: the library stream)."""
    lib().lcd_gpu_set_thread_stream(C.c_void_p(s or 0))


def reserve_plan_memory(n_bytes):
    """lcd_gpu_reserve_plan_memory: map a reserve for the plans' stream-ordered allocations once (hosts that create plans from several threads)"""
    lib().lcd_gpu_reserve_plan_memory.argtypes = [C.c_size_t]
    _check(lib().lcd_gpu_reserve_plan_memory(C.c_size_t(n_bytes)), "lcd_gpu_reserve_plan_memory")


def split_pool(lower_bytes):
    """POA plans carve from the lower `lower_bytes` of the workspace pool, WFA / edlib plans from the rest (lcd_gpu_split_pool)."""
    lib().lcd_gpu_split_pool.argtypes = [C.c_size_t]
    _check(lib().lcd_gpu_split_pool(lower_bytes), "lcd_gpu_split_pool")


def reserve_sms(n):
    """CTA slots of n SMs are left free by the persistent DP grids (lcd_gpu_reserve_sms)."""
    _check(lib().lcd_gpu_reserve_sms(C.c_int(n)), "lcd_gpu_reserve_sms")


def launch_count():
    return int(lib().lcd_gpu_launch_count())


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def pack_pairs(pairs):
    """[(pattern, text), ...] of uint8 arrays -> (seqs, a_off, a_len, b_off, b_len)."""
    n = len(pairs)
    a_len = np.fromiter((len(p) for p, _ in pairs), dtype=np.int32, count=n)
    b_len = np.fromiter((len(t) for _, t in pairs), dtype=np.int32, count=n)
    tot = a_len.astype(np.int64) + b_len
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(tot, out=start[1:])
    seqs = np.empty(max(int(start[-1]), 1), dtype=np.uint8)
    for i, (p, t) in enumerate(pairs):
        s = int(start[i])
        seqs[s:s + a_len[i]] = p
        seqs[s + a_len[i]:s + a_len[i] + b_len[i]] = t
    a_off = start[:-1].copy()
    b_off = a_off + a_len
    return seqs, a_off, a_len, b_off, b_len


class _Plan:
    def __init__(self, handle, n):
        if not handle:
            raise LcdGpuError("plan creation failed: " + lib().lcd_gpu_last_error().decode())
        self.h = C.c_void_p(handle)
        self.n = n

    def run(self, stream=None):
        _check(lib().lcd_plan_run(self.h, C.c_void_p(stream or 0)), "lcd_plan_run")

    def sync(self, stream=None):
        _check(lib().lcd_plan_sync(self.h, C.c_void_p(stream or 0)), "lcd_plan_sync")

    def work_units(self, stream=None):
        u = C.c_uint64(0)
        _check(lib().lcd_plan_work_units(self.h, C.c_void_p(stream or 0), C.byref(u)), "lcd_plan_work_units")
        return int(u.value)

    def destroy(self):
        if self.h:
            lib().lcd_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _params_array(params, n):
    arr = np.zeros(n, dtype=WFA_PARAMS_DTYPE)
    if isinstance(params, np.ndarray) and params.dtype == WFA_PARAMS_DTYPE:
        arr[:] = params
    elif isinstance(params, (list,)) and len(params) == n and not isinstance(params[0], (int, np.integer)):
        for i, p in enumerate(params):
            arr[i] = tuple(getattr(p, f) for f, _ in WfaParams._fields_) if isinstance(p, WfaParams) else tuple(p)
    else:
        arr[:] = tuple(getattr(params, f) for f, _ in WfaParams._fields_) if isinstance(params, WfaParams) else tuple(params)
    return arr


class WfaPlan(_Plan):
    """Inputs resident in HBM; run() re-executes the batch (what bench.py times)."""

    def __init__(self, seqs, pat_off, plen, txt_off, tlen, params):
        n = len(plen)
        self.seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self.pat_off = np.ascontiguousarray(pat_off, dtype=np.int64)
        self.txt_off = np.ascontiguousarray(txt_off, dtype=np.int64)
        self.plen = np.ascontiguousarray(plen, dtype=np.int32)
        self.tlen = np.ascontiguousarray(tlen, dtype=np.int32)
        self.params = _params_array(params, n)
        h = lib().lcd_wfa_plan_create(C.c_int(n), _ptr(self.seqs, C.c_uint8), C.c_size_t(self.seqs.size),
                                      _ptr(self.pat_off, C.c_int64), _ptr(self.plen, C.c_int32),
                                      _ptr(self.txt_off, C.c_int64), _ptr(self.tlen, C.c_int32),
                                      self.params.ctypes.data_as(C.c_void_p))
        super().__init__(h, n)

    def ops_layout(self):
        cap = 2 * (self.plen.astype(np.int64) + self.tlen) + 8
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(cap, out=off[1:])
        return off

    def fetch(self, stream=None, want_ops=True):
        res = np.zeros(self.n, dtype=WFA_RESULT_DTYPE)
        off = self.ops_layout()
        ops = np.zeros(max(int(off[-1]), 1), dtype=np.uint8) if want_ops else None
        rc = lib().lcd_wfa_plan_fetch(self.h, C.c_void_p(stream or 0),
                                      ops.ctypes.data_as(C.c_char_p) if want_ops else None,
                                      _ptr(off, C.c_int64) if want_ops else None,
                                      res.ctypes.data_as(C.c_void_p))
        _check(rc, "lcd_wfa_plan_fetch")
        return res, ops, off


def wfa_batch(pairs, params):
    """Drop-in batch call over HOST buffers (lcd_wfa_batch): H2D, kernels, D2H.
    Returns [(status, score, ops_bytes, end_v, end_h)] per problem."""
    n = len(pairs)
    seqs, po, pl, to, tl = pack_pairs(pairs)
    par = _params_array(params, n)
    res = np.zeros(n, dtype=WFA_RESULT_DTYPE)
    cap = 2 * (pl.astype(np.int64) + tl) + 8
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cap, out=off[1:])
    ops = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
    rc = lib().lcd_wfa_batch(C.c_int(n), _ptr(seqs, C.c_uint8), C.c_size_t(seqs.size),
                             _ptr(po, C.c_int64), _ptr(pl, C.c_int32), _ptr(to, C.c_int64), _ptr(tl, C.c_int32),
                             par.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_char_p),
                             _ptr(off, C.c_int64), res.ctypes.data_as(C.c_void_p))
    _check(rc, "lcd_wfa_batch")
    out = []
    for i in range(n):
        r = res[i]
        out.append((int(r["status"]), int(r["score"]), ops[off[i]:off[i] + r["n_ops"]].tobytes(),
                    int(r["end_v"]), int(r["end_h"])))
    return out


# ----------------------------------------------------------------------------- K7: edlib
MODE_NW, MODE_SHW, MODE_HW = 0, 1, 2          # EdlibAlignMode (reference edlib/include/edlib.h)
EDLIB_RESULT_DTYPE = np.dtype([(n, np.int32) for n in ("status", "edit_distance", "start_loc", "end_loc", "aln_len")])


def _edlib_modes(n, mode, want_path):
    m = np.empty(n, dtype=np.int32); m[:] = mode
    w = np.empty(n, dtype=np.int32); w[:] = want_path
    return m, w


class EdlibPlan(_Plan):
    """(query, target) pairs resident in HBM; run() re-executes the batch."""

    def __init__(self, seqs, q_off, qlen, t_off, tlen, mode=MODE_NW, want_path=1):
        n = len(qlen)
        self.seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self.q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        self.t_off = np.ascontiguousarray(t_off, dtype=np.int64)
        self.qlen = np.ascontiguousarray(qlen, dtype=np.int32)
        self.tlen = np.ascontiguousarray(tlen, dtype=np.int32)
        self.mode, self.want_path = _edlib_modes(n, mode, want_path)
        h = lib().lcd_edlib_plan_create(C.c_int(n), _ptr(self.seqs, C.c_uint8), C.c_size_t(self.seqs.size),
                                        _ptr(self.q_off, C.c_int64), _ptr(self.qlen, C.c_int32),
                                        _ptr(self.t_off, C.c_int64), _ptr(self.tlen, C.c_int32),
                                        _ptr(self.mode, C.c_int32), _ptr(self.want_path, C.c_int32))
        super().__init__(h, n)

    def fetch(self, stream=None, want_aln=True):
        res = np.zeros(self.n, dtype=EDLIB_RESULT_DTYPE)
        cap = self.qlen.astype(np.int64) + self.tlen + 2
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(cap, out=off[1:])
        aln = np.zeros(max(int(off[-1]), 1), dtype=np.uint8) if want_aln else None
        rc = lib().lcd_edlib_plan_fetch(self.h, C.c_void_p(stream or 0), _ptr(aln, C.c_uint8) if want_aln else None,
                                        _ptr(off, C.c_int64) if want_aln else None, res.ctypes.data_as(C.c_void_p))
        _check(rc, "lcd_edlib_plan_fetch")
        return res, aln, off


def edlib_batch(pairs, mode=MODE_NW, want_path=1):
    """Drop-in batch call over HOST buffers (lcd_edlib_batch) for [(query, target), ...].
    Returns [(status, edit_distance, start_loc, end_loc, path bytes)] per problem."""
    n = len(pairs)
    if n == 0:
        return []
    seqs, qo, ql, to, tl = pack_pairs(pairs)
    m, w = _edlib_modes(n, mode, want_path)
    res = np.zeros(n, dtype=EDLIB_RESULT_DTYPE)
    cap = ql.astype(np.int64) + tl + 2
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cap, out=off[1:])
    aln = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
    rc = lib().lcd_edlib_batch(C.c_int(n), _ptr(seqs, C.c_uint8), C.c_size_t(seqs.size), _ptr(qo, C.c_int64), _ptr(ql, C.c_int32),
                               _ptr(to, C.c_int64), _ptr(tl, C.c_int32), _ptr(m, C.c_int32), _ptr(w, C.c_int32),
                               _ptr(aln, C.c_uint8), _ptr(off, C.c_int64), res.ctypes.data_as(C.c_void_p))
    _check(rc, "lcd_edlib_batch")
    return [(int(r["status"]), int(r["edit_distance"]), int(r["start_loc"]), int(r["end_loc"]),
             aln[off[i]:off[i] + r["aln_len"]].tobytes()) for i, r in enumerate(res)]


def xgaps(path):
    """edlibAlignmentToXGAPS (reference src/align.c:189-208): mismatches + gap openings of an edlib path."""
    a = np.frombuffer(path, dtype=np.uint8)
    if a.size == 0:
        return 0
    gap = (a == 1) | (a == 2)
    opens = gap & np.concatenate(([True], a[1:] != a[:-1]))
    return int((a == 3).sum() + opens.sum())


# ----------------------------------------------------------------------------- K1: pileup scan, difference lists from =/X CIGARs
_DIGAR_SCALARS = (("n_reads", C.c_int32), ("min_bq", C.c_int32), ("noisy_reg_max_xgaps", C.c_int32), ("noisy_reg_slide_win", C.c_int32),
                  ("end_clip_reg", C.c_int32), ("end_clip_reg_flank_win", C.c_int32), ("max_noisy_frac_per_read", C.c_double),
                  ("max_var_ratio_per_read", C.c_double), ("whole_ref_len", C.c_int64), ("reg_beg", C.c_int64), ("reg_end", C.c_int64))
_DIGAR_IN = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("read_pos0", np.int64), ("read_is_rev", np.uint8),
             ("is_palindrome", np.uint8), ("n_cigar", np.int32), ("cigar_off", np.int64), ("cigar", np.uint32),
             ("l_qseq", np.int32), ("seq_off", np.int64), ("bseq", np.uint8), ("qual_off", np.int64), ("qual", np.uint8))
_DIGAR_OUT = (("skip", np.uint8, "r"), ("read_beg", np.int64, "r"), ("read_end", np.int64, "r"), ("digar_first", np.int64, "r"), ("n_digar", np.int32, "r"),
              ("digar_pos", np.int64, "d"), ("digar_type", np.int8, "d"), ("digar_len", np.int32, "d"), ("digar_qi", np.int32, "d"),
              ("digar_low_qual", np.uint8, "d"), ("digar_alt_off", np.int64, "d"), ("digar_alt", np.uint8, "a"))
_DIGAR_NREG = (("nreg_first", np.int64, "r"), ("n_nreg", np.int32, "r"), ("nreg_beg", np.int64, "n"), ("nreg_end", np.int64, "n"), ("nreg_label", np.int32, "n"))
_DIGAR_CNREG = (("cnreg_beg", np.int64), ("cnreg_end", np.int64), ("cnreg_label", np.int32))


class DigarInput(C.Structure):
    _fields_ = list(_DIGAR_SCALARS) + [(k, C.c_void_p) for k, _ in _DIGAR_IN]


class DigarOutput(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _, _ in _DIGAR_OUT] + [("digar_cap", C.c_int64), ("alt_cap", C.c_int64)] + \
               [(k, C.c_void_p) for k, _, _ in _DIGAR_NREG] + [("nreg_cap", C.c_int64)] + \
               [(k, C.c_void_p) for k, _ in _DIGAR_CNREG] + [("cnreg_cap", C.c_int64), ("n_cnreg", C.c_int64)] + \
               [("qual_counts", C.c_void_p), ("n_digar_total", C.c_int64), ("n_alt_total", C.c_int64), ("n_nreg_total", C.c_int64)]


def _digar_inputs(chunks):
    n = len(chunks)
    ins, keep = (DigarInput * max(n, 1))(), []
    for i, d in enumerate(chunks):
        arrs = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _DIGAR_IN}
        keep.append(arrs)
        ins[i] = DigarInput(*[d[k] for k, _ in _DIGAR_SCALARS], *[arrs[k].ctypes.data for k, _ in _DIGAR_IN])
    return ins, keep


def _digar_outputs(chunks, sizes):
    """sizes[i] = (records, alt bases, intervals) capacities of chunk i."""
    n = len(chunks)
    outs, results = (DigarOutput * max(n, 1))(), []
    for i, d in enumerate(chunks):
        cap = {"r": d["n_reads"] + 1, "d": sizes[i][0] + 1, "a": sizes[i][1] + 1, "n": sizes[i][2] + 1}
        o = {k: np.zeros(cap[w], dtype=t) for k, t, w in _DIGAR_OUT + _DIGAR_NREG}
        o.update({k: np.zeros(cap["n"], dtype=t) for k, t in _DIGAR_CNREG})
        o["qual_counts"] = np.zeros(256, np.int64)
        results.append(o)
        outs[i] = DigarOutput(*[o[k].ctypes.data for k, _, _ in _DIGAR_OUT], cap["d"], cap["a"], *[o[k].ctypes.data for k, _, _ in _DIGAR_NREG], cap["n"],
                              *[o[k].ctypes.data for k, _ in _DIGAR_CNREG], cap["n"], 0, o["qual_counts"].ctypes.data, 0, 0, 0)
    return outs, results


def _digar_finish(outs, results):
    for i, o in enumerate(results):
        o["n_cnreg"] = int(outs[i].n_cnreg); o["n_digar_total"] = int(outs[i].n_digar_total)
        o["n_alt_total"] = int(outs[i].n_alt_total); o["n_nreg_total"] = int(outs[i].n_nreg_total)
    return results


def digar_capacity(ins, n):
    sizes = []
    for i in range(n):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().lcd_digar_capacity(C.byref(ins[i]), C.byref(a), C.byref(b), C.byref(c)), "lcd_digar_capacity")
        sizes.append((a.value, b.value, c.value))
    return sizes


def digar_capacity_host(chunks):
    """lcd_digar_capacity in numpy (same formula, no library call): what callers that must not load liblcd_gpu.so -- the reference
    arm of bench.py -- size their output arrays with."""
    sizes = []
    for d in chunks:
        nr = int(d["n_reads"])
        off, n = np.asarray(d["cigar_off"][:nr], np.int64), np.asarray(d["n_cigar"][:nr], np.int64)
        tot = int(n.sum())
        idx = np.repeat(off - (np.cumsum(n) - n), n) + np.arange(tot)
        cg = np.asarray(d["cigar"], np.uint32)[idx]; op = cg & 15; ln = (cg >> 4).astype(np.int64)
        x, i_, d_ = int(ln[op == 8].sum()), int((op == 1).sum()), int((op == 2).sum())
        sizes.append((x + i_ + d_ + int(np.isin(op, (7, 4, 5)).sum()) + 1, x + int(ln[op == 1].sum()) + 1, x + i_ + d_ + 2 * nr + 1))
    return sizes


def digar_batch(chunks):
    """Drop-in batch call over HOST buffers (lcd_digar_batch).  chunks: dicts with the fields of lcd_digar_input_t.  Returns per
    chunk a dict with the arrays of lcd_digar_output_t (+ the n_* totals)."""
    ins, keep = _digar_inputs(chunks)
    outs, results = _digar_outputs(chunks, digar_capacity(ins, len(chunks)))
    _check(lib().lcd_digar_batch(C.c_int(len(chunks)), ins, outs), "lcd_digar_batch")
    return _digar_finish(outs, results)


class MdTags(C.Structure):
    _fields_ = [("md_off", C.c_void_p), ("md", C.c_void_p)]


def digar_md_batch(chunks, tags):
    """lcd_digar_md_batch: chunks whose reads carry plain-M CIGARs + MD tags.  tags[i] = (md_off int64 [n_reads] (< 0: the read's CIGAR is =/X
    already), md uint8 bytes: NUL-terminated strings)."""
    ins, keep = _digar_inputs(chunks)
    sizes = digar_capacity(ins, len(chunks))
    sizes = [(d + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()), a + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()), r + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()))
             for (d, a, r), c in zip(sizes, chunks)]                      # an M op expands into up to its length in records
    outs, results = _digar_outputs(chunks, sizes)
    tkeep = [(np.ascontiguousarray(o, np.int64), np.ascontiguousarray(m, np.uint8)) for o, m in tags]
    tarr = (MdTags * max(len(chunks), 1))(*[MdTags(o.ctypes.data, m.ctypes.data) for o, m in tkeep])
    _check(lib().lcd_digar_md_batch(C.c_int(len(chunks)), ins, tarr, outs), "lcd_digar_md_batch")
    return _digar_finish(outs, results)


class ReadTags(C.Structure):
    _fields_ = [("kind", C.c_void_p), ("off", C.c_void_p), ("text", C.c_void_p), ("ref_seq", C.c_void_p), ("ref_beg", C.c_int64), ("ref_end", C.c_int64)]


TAG_EQX, TAG_MD, TAG_CS, TAG_REFSEQ = -1, 0, 1, 2


def digar_tags_batch(chunks, tags):
    """lcd_digar_tags_batch: every read names the variant of the reference's pass it takes (TAG_EQX / TAG_MD / TAG_CS / TAG_REFSEQ).
    tags[i] = dict(kind int8 [n_reads], off int64 [n_reads] or None, text uint8 bytes or None, ref_seq uint8 ASCII or None, ref_beg, ref_end)."""
    ins, keep = _digar_inputs(chunks)
    sizes = digar_capacity(ins, len(chunks))
    sizes = [(d + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()), a + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()), r + int(np.asarray(c["l_qseq"][:c["n_reads"]]).sum()))
             for (d, a, r), c in zip(sizes, chunks)]                      # an M op expands into up to its length in records
    outs, results = _digar_outputs(chunks, sizes)
    tkeep, arr = [], []
    for t in tags:
        k = np.ascontiguousarray(t["kind"], np.int8)
        o = np.ascontiguousarray(t["off"], np.int64) if t.get("off") is not None else None
        x = np.ascontiguousarray(t["text"], np.uint8) if t.get("text") is not None else None
        r = np.ascontiguousarray(t["ref_seq"], np.uint8) if t.get("ref_seq") is not None else None
        tkeep.append((k, o, x, r))
        arr.append(ReadTags(k.ctypes.data, o.ctypes.data if o is not None else None, x.ctypes.data if x is not None else None, r.ctypes.data if r is not None else None,
                            int(t.get("ref_beg", 0)), int(t.get("ref_end", -1))))
    tarr = (ReadTags * max(len(chunks), 1))(*arr)
    _check(lib().lcd_digar_tags_batch(C.c_int(len(chunks)), ins, tarr, outs), "lcd_digar_tags_batch")
    return _digar_finish(outs, results)


class DigarPlan(_Plan):
    """Inputs (CIGAR words, packed SEQ, QUAL of every chunk) resident in HBM; run() = count + scan + fill + histogram."""
    def __init__(self, chunks):
        self.chunks = chunks
        self.ins, self.keep = _digar_inputs(chunks)
        super().__init__(lib().lcd_digar_plan_create(C.c_int(len(chunks)), self.ins), len(chunks))

    def sizes(self, stream=None):
        out = []
        for i in range(self.n):
            a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
            _check(lib().lcd_digar_plan_sizes(self.h, C.c_void_p(stream or 0), C.c_int(i), C.byref(a), C.byref(b), C.byref(c)), "lcd_digar_plan_sizes")
            out.append((a.value, b.value, c.value))
        return out

    def fetch(self, stream=None):
        outs, results = _digar_outputs(self.chunks, self.sizes(stream))
        _check(lib().lcd_digar_plan_fetch(self.h, C.c_void_p(stream or 0), outs), "lcd_digar_plan_fetch")
        return _digar_finish(outs, results)


# ----------------------------------------------------------------------------- K2: pileup scan, per-site coverage
_PILEUP_IN = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("read_beg", np.int64), ("read_end", np.int64),
              ("read_is_rev", np.uint8), ("digar_first", np.int64), ("n_digar", np.int32), ("qual_off", np.int64), ("qual", np.uint8),
              ("digar_pos", np.int64), ("digar_type", np.int8), ("digar_len", np.int32), ("digar_qi", np.int32),
              ("digar_low_qual", np.uint8), ("digar_alt_off", np.int64), ("digar_alt", np.uint8),
              ("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("site_alt_len", np.int32),
              ("site_alt_off", np.int64), ("site_alt", np.uint8))


class PileupInput(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_sites", C.c_int32), ("min_bq", C.c_int32), ("min_sv_len", C.c_int32)] + \
               [(k, C.c_void_p) for k, _ in _PILEUP_IN]


class PileupOutput(C.Structure):
    _fields_ = [("site_counts", C.c_void_p)]


def _pileup_structs(chunks):
    n = len(chunks)
    ins, outs, keep, results = (PileupInput * max(n, 1))(), (PileupOutput * max(n, 1))(), [], []
    for i, d in enumerate(chunks):
        arrs = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _PILEUP_IN}
        keep.append(arrs)
        ins[i] = PileupInput(d["n_reads"], d["n_sites"], d["min_bq"], d["min_sv_len"], *[arrs[k].ctypes.data for k, _ in _PILEUP_IN])
        counts = np.zeros((d["n_sites"] + 1, 8), dtype=np.int32)
        results.append(counts)
        outs[i] = PileupOutput(counts.ctypes.data)
    return ins, outs, keep, results


def pileup_batch(chunks):
    """Drop-in batch call over HOST buffers (lcd_pileup_batch): per chunk the (n_sites, 8) coverage counters
    [total_cov, low_qual_cov, alle_covs[0..1], strand_to_alle_covs[0..1][0..1]]."""
    ins, outs, keep, results = _pileup_structs(chunks)
    _check(lib().lcd_pileup_batch(C.c_int(len(chunks)), ins, outs), "lcd_pileup_batch")
    return [r[:d["n_sites"]] for r, d in zip(results, chunks)]


class PileupPlan(_Plan):
    def __init__(self, chunks):
        self.chunks = chunks
        self.ins, self.outs, self.keep, self.results = _pileup_structs(chunks)
        super().__init__(lib().lcd_pileup_plan_create(C.c_int(len(chunks)), self.ins), len(chunks))

    def fetch(self, stream=None):
        _check(lib().lcd_pileup_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_pileup_plan_fetch")
        return [r[:d["n_sites"]] for r, d in zip(self.results, self.chunks)]


# ----------------------------------------------------------------------------- K1 -> K2 / K3 in place (difference lists stay in HBM)
_SITE_LIST = (("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("site_alt_len", np.int32),
              ("site_alt_off", np.int64), ("site_alt", np.uint8))


class SiteList(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("min_sv_len", C.c_int32)] + [(k, C.c_void_p) for k, _ in _SITE_LIST] + [("var_cate", C.c_void_p)]


def _site_lists(sites, want_cate):
    n = len(sites)
    arr, keep = (SiteList * max(n, 1))(), []
    for i, d in enumerate(sites):
        a = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _SITE_LIST}
        a["var_cate"] = np.ascontiguousarray(d["var_cate"], dtype=np.int32) if want_cate else None
        keep.append(a)
        arr[i] = SiteList(d["n_sites"], d["min_sv_len"], *[a[k].ctypes.data for k, _ in _SITE_LIST], a["var_cate"].ctypes.data if want_cate else None)
    return arr, keep


class PileupOnDigarPlan(_Plan):
    """K2 on the difference lists a DigarPlan left in HBM; sites: per chunk a dict(n_sites, min_sv_len, site_*)."""
    def __init__(self, digar_plan, sites):
        self.sites, self.digar_plan = sites, digar_plan
        self.arr, self.keep = _site_lists(sites, False)
        lib().lcd_pileup_plan_create_on_digar.restype = C.c_void_p
        super().__init__(lib().lcd_pileup_plan_create_on_digar(digar_plan.h, C.c_int(len(sites)), self.arr), len(sites))
        self.results = [np.zeros((d["n_sites"] + 1, 8), dtype=np.int32) for d in sites]
        self.outs = (PileupOutput * max(self.n, 1))(*[PileupOutput(r.ctypes.data) for r in self.results])

    def fetch(self, stream=None):
        _check(lib().lcd_pileup_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_pileup_plan_fetch")
        return [r[:d["n_sites"]] for r, d in zip(self.results, self.sites)]


class ProfileOnDigarPlan(_Plan):
    """K3 on the difference lists (and per-read noisy intervals) a DigarPlan left in HBM; sites carry var_cate."""
    def __init__(self, digar_plan, sites, n_reads):
        self.sites, self.digar_plan, self.n_reads = sites, digar_plan, n_reads
        self.arr, self.keep = _site_lists(sites, True)
        L = lib()
        L.lcd_profile_plan_create_on_digar.restype = C.c_void_p
        L.lcd_profile_plan_capacity.restype = C.c_int64
        L.lcd_profile_plan_capacity.argtypes = [C.c_void_p, C.c_int]
        super().__init__(L.lcd_profile_plan_create_on_digar(digar_plan.h, C.c_int(len(sites)), self.arr), len(sites))
        self.res, self.outs = [], (ProfileOutput * max(self.n, 1))()
        for i in range(self.n):
            cap = int(L.lcd_profile_plan_capacity(self.h, i)); nr = n_reads[i]
            o = dict(prof_start=np.zeros(nr + 1, np.int32), prof_end=np.zeros(nr + 1, np.int32), allele_off=np.zeros(nr + 1, np.int64),
                     alleles=np.zeros(cap + 1, np.int8), alt_qi=np.zeros(cap + 1, np.int32))
            self.res.append(o)
            self.outs[i] = ProfileOutput(o["prof_start"].ctypes.data, o["prof_end"].ctypes.data, o["allele_off"].ctypes.data, o["alleles"].ctypes.data,
                                         o["alt_qi"].ctypes.data, cap, 0)

    def fetch(self, stream=None):
        _check(lib().lcd_profile_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_profile_plan_fetch")
        return self.res


# ----------------------------------------------------------------------------- K1b: candidate-site list
class SitesParams(C.Structure):
    _fields_ = [("reg_beg", C.c_int64), ("reg_end", C.c_int64), ("min_sv_len", C.c_int32), ("pad", C.c_int32)]


class SitesOutput(C.Structure):
    _fields_ = [("site_pos", C.c_void_p), ("site_type", C.c_void_p), ("site_ref_len", C.c_void_p), ("site_alt_len", C.c_void_p), ("site_src", C.c_void_p),
                ("cap", C.c_int64), ("n_sites", C.c_int64)]


def _sites_params(regs, min_sv_len):
    n = len(regs)
    return (SitesParams * max(n, 1))(*[SitesParams(int(b), int(e), int(m), 0) for (b, e), m in zip(regs, min_sv_len)])


def _sites_outputs(caps):
    outs, res = (SitesOutput * max(len(caps), 1))(), []
    for i, cap in enumerate(caps):
        o = dict(site_pos=np.zeros(cap + 1, np.int64), site_type=np.zeros(cap + 1, np.int32), site_ref_len=np.zeros(cap + 1, np.int32),
                 site_alt_len=np.zeros(cap + 1, np.int32), site_src=np.zeros(cap + 1, np.int64))
        res.append(o)
        outs[i] = SitesOutput(o["site_pos"].ctypes.data, o["site_type"].ctypes.data, o["site_ref_len"].ctypes.data, o["site_alt_len"].ctypes.data,
                              o["site_src"].ctypes.data, cap, 0)
    return outs, res


def _sites_finish(outs, res):
    return [{k: v[:outs[i].n_sites] for k, v in o.items()} | {"n_sites": int(outs[i].n_sites)} for i, o in enumerate(res)]


def _sites_inputs(chunks):
    z = {k: np.zeros(1, t) for k, t in _PILEUP_IN if k.startswith("site_")}
    return _pileup_structs([{**d, **z, "n_sites": 0} for d in chunks])


def sites_batch(chunks, regs):
    """Drop-in batch call over HOST buffers (lcd_sites_batch): chunks in lcd_pileup_input_t layout (site fields unused),
    regs[i] = (reg_beg, reg_end) -> per chunk dict(site_pos, site_type, site_ref_len, site_alt_len, site_src, n_sites)."""
    ins, _, keep, _ = _sites_inputs(chunks)
    par = _sites_params(regs, [d["min_sv_len"] for d in chunks])
    outs, res = _sites_outputs([int(np.isin(np.asarray(d["digar_type"]), (1, 2, 8)).sum()) for d in chunks])
    _check(lib().lcd_sites_batch(C.c_int(len(chunks)), ins, par, outs), "lcd_sites_batch")
    return _sites_finish(outs, res)


class SitesPlan(_Plan):
    """Resident plan: on host difference lists (chunks) or, with digar_plan, on the lists K1 left in HBM (chunks ignored)."""
    def __init__(self, chunks, regs, min_sv_len=None, digar_plan=None):
        L = lib()
        L.lcd_sites_plan_create.restype = C.c_void_p
        L.lcd_sites_plan_create_on_digar.restype = C.c_void_p
        self.digar_plan = digar_plan
        n = len(regs)
        self.par = _sites_params(regs, min_sv_len if min_sv_len is not None else [d["min_sv_len"] for d in chunks])
        if digar_plan is not None:
            h = L.lcd_sites_plan_create_on_digar(digar_plan.h, C.c_int(n), self.par)
        else:
            self.ins, _, self.keep, _ = _sites_inputs(chunks)
            h = L.lcd_sites_plan_create(C.c_int(n), self.ins, self.par)
        super().__init__(h, n)

    def sizes(self, stream=None):
        out = []
        for i in range(self.n):
            v = C.c_int64(0)
            _check(lib().lcd_sites_plan_sizes(self.h, C.c_void_p(stream or 0), C.c_int(i), C.byref(v)), "lcd_sites_plan_sizes")
            out.append(int(v.value))
        return out

    def fetch(self, stream=None):
        outs, res = _sites_outputs(self.sizes(stream))
        _check(lib().lcd_sites_plan_fetch(self.h, C.c_void_p(stream or 0), outs), "lcd_sites_plan_fetch")
        return _sites_finish(outs, res)


class PileupOnSitesPlan(_Plan):
    """K2 on the site lists a SitesPlan (created on the same DigarPlan, and run) left in HBM."""
    def __init__(self, digar_plan, sites_plan):
        self.digar_plan, self.sites_plan = digar_plan, sites_plan
        lib().lcd_pileup_plan_create_on_sites.restype = C.c_void_p
        super().__init__(lib().lcd_pileup_plan_create_on_sites(digar_plan.h, sites_plan.h), sites_plan.n)
        self.n_sites = sites_plan.sizes()
        self.results = [np.zeros((k + 1, 8), dtype=np.int32) for k in self.n_sites]
        self.outs = (PileupOutput * max(self.n, 1))(*[PileupOutput(r.ctypes.data) for r in self.results])

    def fetch(self, stream=None):
        _check(lib().lcd_pileup_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_pileup_plan_fetch")
        return [r[:k] for r, k in zip(self.results, self.n_sites)]


# ----------------------------------------------------------------------------- K2b: category of every candidate site
_CLASSIFY = (("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("site_alt_len", np.int32),
             ("site_alt_off", np.int64), ("site_alt", np.uint8), ("site_counts", np.int32))


class ClassifyInput(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("min_dp", C.c_int32), ("min_alt_dp", C.c_int32), ("max_xgaps", C.c_int32), ("is_ont", C.c_int32), ("pad", C.c_int32),
                ("min_af", C.c_double), ("max_af", C.c_double), ("ref_beg", C.c_int64), ("ref_end", C.c_int64), ("ref_seq", C.c_void_p)] + \
               [(k, C.c_void_p) for k, _ in _CLASSIFY]


class ClassifyOutput(C.Structure):
    _fields_ = [("var_cate", C.c_void_p)]


def _classify_structs(chunks):
    n = len(chunks)
    ins, outs, keep, res = (ClassifyInput * max(n, 1))(), (ClassifyOutput * max(n, 1))(), [], []
    for i, d in enumerate(chunks):
        a = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _CLASSIFY}
        a["ref_seq"] = np.ascontiguousarray(d["ref_seq"], dtype=np.uint8)
        keep.append(a)
        ins[i] = ClassifyInput(d["n_sites"], d["min_dp"], d["min_alt_dp"], d["max_xgaps"], d.get("is_ont", 0), 0, d["min_af"], d["max_af"], d["ref_beg"], d["ref_end"],
                               a["ref_seq"].ctypes.data, *[a[k].ctypes.data for k, _ in _CLASSIFY])
        r = np.full(d["n_sites"] + 1, -7, np.int32); res.append(r)
        outs[i] = ClassifyOutput(r.ctypes.data)
    return ins, outs, keep, res


def classify_batch(chunks):
    """Drop-in batch call over HOST buffers (lcd_classify_batch): per chunk the LONGCALLD_* category of every candidate site
    (dict keys: n_sites, min_dp, min_alt_dp, max_xgaps, min_af, max_af, ref_beg, ref_end, ref_seq (ASCII bytes), site_*, site_counts)."""
    ins, outs, keep, res = _classify_structs(chunks)
    _check(lib().lcd_classify_batch(C.c_int(len(chunks)), ins, outs), "lcd_classify_batch")
    return [r[:d["n_sites"]] for r, d in zip(res, chunks)]


class ClassifyPlan(_Plan):
    def __init__(self, chunks):
        self.chunks = chunks
        self.ins, self.outs, self.keep, self.res = _classify_structs(chunks)
        lib().lcd_classify_plan_create.restype = C.c_void_p
        super().__init__(lib().lcd_classify_plan_create(C.c_int(len(chunks)), self.ins), len(chunks))

    def fetch(self, stream=None):
        _check(lib().lcd_classify_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_classify_plan_fetch")
        return [r[:d["n_sites"]] for r, d in zip(self.res, self.chunks)]


class ClassifyParams(C.Structure):
    _fields_ = [("min_dp", C.c_int32), ("min_alt_dp", C.c_int32), ("max_xgaps", C.c_int32), ("is_ont", C.c_int32), ("min_af", C.c_double), ("max_af", C.c_double),
                ("ref_beg", C.c_int64), ("ref_end", C.c_int64), ("ref_seq", C.c_void_p)]


class ClassifyOnPileupPlan(_Plan):
    """K2b on the sites and counters a K2 plan (PileupPlan / PileupOnDigarPlan / PileupOnSitesPlan, run) holds in HBM; params: per chunk a
    dict(min_dp, min_alt_dp, max_xgaps, min_af, max_af, ref_beg, ref_end, ref_seq); n_sites: the K2 plan's site counts."""
    def __init__(self, pileup_plan, params, n_sites):
        self.pileup_plan, self.n_sites = pileup_plan, list(n_sites)
        n = len(params)
        self.keep = [np.ascontiguousarray(d["ref_seq"], dtype=np.uint8) for d in params]
        self.par = (ClassifyParams * max(n, 1))(*[ClassifyParams(d["min_dp"], d["min_alt_dp"], d["max_xgaps"], d.get("is_ont", 0), d["min_af"], d["max_af"], d["ref_beg"], d["ref_end"],
                                                                 r.ctypes.data) for d, r in zip(params, self.keep)])
        lib().lcd_classify_plan_create_on_pileup.restype = C.c_void_p
        super().__init__(lib().lcd_classify_plan_create_on_pileup(pileup_plan.h, C.c_int(n), self.par), n)
        self.res = [np.full(k + 1, -7, np.int32) for k in self.n_sites]
        self.outs = (ClassifyOutput * max(n, 1))(*[ClassifyOutput(r.ctypes.data) for r in self.res])

    def fetch(self, stream=None):
        _check(lib().lcd_classify_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_classify_plan_fetch")
        return [r[:k] for r, k in zip(self.res, self.n_sites)]


# ----------------------------------------------------------------------------- K3: pileup scan, read x variant profile
_PROFILE_EX = (("var_cate", np.int32), ("nreg_first", np.int64), ("n_nreg", np.int32), ("nreg_beg", np.int64), ("nreg_end", np.int64))


class ProfileExtra(C.Structure):
    _fields_ = [(k, C.c_void_p) for k, _ in _PROFILE_EX]


class ProfileOutput(C.Structure):
    _fields_ = [("prof_start", C.c_void_p), ("prof_end", C.c_void_p), ("allele_off", C.c_void_p), ("alleles", C.c_void_p), ("alt_qi", C.c_void_p),
                ("alleles_cap", C.c_int64), ("n_alleles", C.c_int64)]


def profile_batch(chunks):
    """Drop-in batch call over HOST buffers (lcd_profile_batch).  chunks: dicts with the fields of lcd_pileup_input_t and
    lcd_profile_extra_t.  Returns per chunk a dict(prof_start, prof_end, allele_off, alleles, alt_qi) -- the arrays
    lcd_phase_input_t takes."""
    n = len(chunks)
    if n == 0:
        return []
    L = lib()
    L.lcd_profile_capacity.restype = C.c_int64
    ins, _, keep, _ = _pileup_structs(chunks)
    exs, outs, res = (ProfileExtra * n)(), (ProfileOutput * n)(), []
    for i, d in enumerate(chunks):
        arrs = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _PROFILE_EX}
        keep.append(arrs)
        exs[i] = ProfileExtra(*[arrs[k].ctypes.data for k, _ in _PROFILE_EX])
        cap = int(L.lcd_profile_capacity(C.byref(ins[i])))
        nr = d["n_reads"]
        o = dict(prof_start=np.zeros(nr + 1, np.int32), prof_end=np.zeros(nr + 1, np.int32), allele_off=np.zeros(nr + 1, np.int64),
                 alleles=np.zeros(cap + 1, np.int8), alt_qi=np.zeros(cap + 1, np.int32))
        res.append(o)
        outs[i] = ProfileOutput(o["prof_start"].ctypes.data, o["prof_end"].ctypes.data, o["allele_off"].ctypes.data, o["alleles"].ctypes.data,
                                o["alt_qi"].ctypes.data, cap, 0)
    _check(L.lcd_profile_batch(C.c_int(n), ins, exs, outs), "lcd_profile_batch")
    return res


# ----------------------------------------------------------------------------- K4: read -> haplotype assignment / phasing
class PhaseInput(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_vars", C.c_int32), ("target_var_cate", C.c_int32), ("is_ont", C.c_int32),
                ("ordered_read_ids", C.c_void_p), ("is_skipped", C.c_void_p), ("prof_start", C.c_void_p), ("prof_end", C.c_void_p),
                ("allele_off", C.c_void_p), ("alleles", C.c_void_p), ("var_cate", C.c_void_p), ("var_type", C.c_void_p),
                ("is_hp_indel", C.c_void_p), ("n_uniq_alles", C.c_void_p), ("alle_covs", C.c_void_p), ("total_cov", C.c_void_p),
                ("pos", C.c_void_p)]


class PhaseOutput(C.Structure):
    _fields_ = [("haps", C.c_void_p), ("phase_sets", C.c_void_p), ("hap_to_cons_alle", C.c_void_p), ("hap_to_alle_profile", C.c_void_p),
                ("var_phase_set", C.c_void_p), ("n_clean_agree_snps", C.c_void_p), ("n_clean_conflict_snps", C.c_void_p)]


_PHASE_IN = (("ordered_read_ids", np.int32), ("is_skipped", np.uint8), ("prof_start", np.int32), ("prof_end", np.int32),
             ("allele_off", np.int64), ("alleles", np.int8), ("var_cate", np.int32), ("var_type", np.int32),
             ("is_hp_indel", np.int32), ("n_uniq_alles", np.int32), ("alle_covs", np.int32), ("total_cov", np.int32), ("pos", np.int64))
_PHASE_OUT = (("haps", np.int32, 1, "r"), ("phase_sets", np.int64, 1, "r"), ("hap_to_cons_alle", np.int32, 3, "v"),
              ("hap_to_alle_profile", np.int32, 12, "v"), ("var_phase_set", np.int64, 1, "v"),
              ("n_clean_agree_snps", np.int32, 1, "r"), ("n_clean_conflict_snps", np.int32, 1, "r"))


def _phase_structs(chunks, prefill):
    """chunks: [(dict of flat arrays (the fields of lcd_phase_input_t), target_var_cate, is_ont)] -> ctypes arrays + keep-alives."""
    n = len(chunks)
    ins, outs, keep, results = (PhaseInput * max(n, 1))(), (PhaseOutput * max(n, 1))(), [], []
    for i, (d, target, is_ont) in enumerate(chunks):
        arrs = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _PHASE_IN}
        keep.append(arrs)
        ins[i] = PhaseInput(d["n_reads"], d["n_vars"], target, is_ont, *[arrs[k].ctypes.data for k, _ in _PHASE_IN])
        out = {k: np.full(m * (d["n_reads"] if w == "r" else d["n_vars"]) + m, prefill, dtype=t) for k, t, m, w in _PHASE_OUT}
        results.append(out)
        outs[i] = PhaseOutput(*[out[k].ctypes.data for k, _, _, _ in _PHASE_OUT])
    return ins, outs, keep, results


def phase_batch(chunks, prefill=-9):
    """Drop-in batch call over HOST buffers (lcd_phase_batch): one entry per region chunk.  Returns the output arrays per
    chunk (entries the reference leaves untouched keep `prefill`)."""
    ins, outs, keep, results = _phase_structs(chunks, prefill)
    _check(lib().lcd_phase_batch(C.c_int(len(chunks)), ins, outs), "lcd_phase_batch")
    return results


class PhasePlan(_Plan):
    def __init__(self, chunks, prefill=-9):
        self.ins, self.outs, self.keep, self.results = _phase_structs(chunks, prefill)
        lib().lcd_phase_plan_create.restype = C.c_void_p
        super().__init__(lib().lcd_phase_plan_create(C.c_int(len(chunks)), self.ins, self.outs), len(chunks))

    def fetch(self, stream=None):
        _check(lib().lcd_phase_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_phase_plan_fetch")
        return self.results


# ----------------------------------------------------------------------------- K2c: the noisy-region set (a5, second half)
_NOISYREG_SITE = (("site_pos", np.int64), ("site_type", np.int32), ("site_ref_len", np.int32), ("var_cate", np.int32))
_NOISYREG_READ = (("is_skipped", np.uint8), ("read_beg", np.int64), ("read_end", np.int64), ("digar_first", np.int64), ("n_digar", np.int32),
                  ("digar_pos", np.int64), ("digar_type", np.int8), ("digar_len", np.int32), ("nreg_first", np.int64), ("n_nreg", np.int32),
                  ("nreg_beg", np.int64), ("nreg_end", np.int64))
_NOISYREG_IV = (("cnreg_beg", np.int64), ("cnreg_end", np.int64), ("cnreg_label", np.int32), ("low_beg", np.int64), ("low_end", np.int64))


class NoisyRegInput(C.Structure):
    _fields_ = [("reg_beg", C.c_int64), ("reg_end", C.c_int64), ("min_alt_dp", C.c_int32), ("noisy_reg_flank_len", C.c_int32), ("is_ont", C.c_int32), ("pad", C.c_int32),
                ("min_af", C.c_double), ("n_sites", C.c_int32), ("n_reads", C.c_int32)] + [(k, C.c_void_p) for k, _ in _NOISYREG_SITE] + \
               [("n_cnreg", C.c_int64), ("cnreg_beg", C.c_void_p), ("cnreg_end", C.c_void_p), ("cnreg_label", C.c_void_p),
                ("n_low", C.c_int64), ("low_beg", C.c_void_p), ("low_end", C.c_void_p)] + [(k, C.c_void_p) for k, _ in _NOISYREG_READ]


class NoisyRegOutput(C.Structure):
    _fields_ = [("var_cate", C.c_void_p), ("keep", C.c_void_p), ("reg_beg", C.c_void_p), ("reg_end", C.c_void_p), ("reg_label", C.c_void_p), ("reg_cap", C.c_int64), ("n_regs", C.c_int64)]


def _noisyreg_structs(chunks):
    n = len(chunks)
    ins, outs, keep, res = (NoisyRegInput * max(n, 1))(), (NoisyRegOutput * max(n, 1))(), [], []
    for i, d in enumerate(chunks):
        a = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in _NOISYREG_SITE + _NOISYREG_READ}
        for k, t in _NOISYREG_IV:
            a[k] = np.ascontiguousarray(np.append(np.asarray(d[k]), 0), dtype=t)
        keep.append(a)
        ins[i] = NoisyRegInput(d["reg_beg"], d["reg_end"], d["min_alt_dp"], d["noisy_reg_flank_len"], d.get("is_ont", 0), 0, d["min_af"], d["n_sites"], d["n_reads"],
                               *[a[k].ctypes.data for k, _ in _NOISYREG_SITE], d["n_cnreg"], a["cnreg_beg"].ctypes.data, a["cnreg_end"].ctypes.data, a["cnreg_label"].ctypes.data,
                               d["n_low"], a["low_beg"].ctypes.data, a["low_end"].ctypes.data, *[a[k].ctypes.data for k, _ in _NOISYREG_READ])
        cap = int(d["n_cnreg"]) + int(d["n_sites"]) + 8
        r = dict(var_cate=np.full(d["n_sites"] + 1, -7, np.int32), keep=np.zeros(d["n_sites"] + 1, np.uint8), reg_beg=np.zeros(cap, np.int64), reg_end=np.zeros(cap, np.int64),
                 reg_label=np.zeros(cap, np.int32))
        res.append(r)
        outs[i] = NoisyRegOutput(r["var_cate"].ctypes.data, r["keep"].ctypes.data, r["reg_beg"].ctypes.data, r["reg_end"].ctypes.data, r["reg_label"].ctypes.data, cap, 0)
    return ins, outs, keep, res


def _noisyreg_results(chunks, outs, res):
    out = []
    for i, (d, r) in enumerate(zip(chunks, res)):
        k = int(outs[i].n_regs); ns = d["n_sites"]
        out.append(dict(var_cate=r["var_cate"][:ns], keep=r["keep"][:ns], n_regs=k, reg_beg=r["reg_beg"][:k], reg_end=r["reg_end"][:k], reg_label=r["reg_label"][:k]))
    return out


def noisyreg_batch(chunks):
    """Drop-in batch call over HOST buffers (lcd_noisyreg_batch): per chunk the noisy-region set and the candidate sites that stay clean-region
    candidates (pre_process_noisy_regs + classify_cand_vars after classify_var_cate).  dict keys: lcd_noisyreg_input_t's fields.
    -> [dict(var_cate, keep, n_regs, reg_beg, reg_end, reg_label)]"""
    ins, outs, keep, res = _noisyreg_structs(chunks)
    _check(lib().lcd_noisyreg_batch(C.c_int(len(chunks)), ins, outs), "lcd_noisyreg_batch")
    return _noisyreg_results(chunks, outs, res)


class NoisyRegPlan(_Plan):
    def __init__(self, chunks):
        self.chunks = chunks
        self.ins, self.outs, self.keep, self.res = _noisyreg_structs(chunks)
        lib().lcd_noisyreg_plan_create.restype = C.c_void_p
        super().__init__(lib().lcd_noisyreg_plan_create(C.c_int(len(chunks)), self.ins), len(chunks))

    def fetch(self, stream=None):
        _check(lib().lcd_noisyreg_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_noisyreg_plan_fetch")
        return _noisyreg_results(self.chunks, self.outs, self.res)


class NoisyRegParams(C.Structure):
    _fields_ = [("min_alt_dp", C.c_int32), ("noisy_reg_flank_len", C.c_int32), ("is_ont", C.c_int32), ("pad", C.c_int32), ("min_af", C.c_double),
                ("n_low", C.c_int64), ("low_beg", C.c_void_p), ("low_end", C.c_void_p)]


class NoisyRegOnClassifyPlan(_Plan):
    """K2c on what a DigarPlan (run) and a ClassifyOnPileupPlan hold in HBM; params: per chunk a dict(min_alt_dp, noisy_reg_flank_len, is_ont, min_af,
    n_low, low_beg, low_end); n_sites: the classify plan's site counts; n_nreg: the digar plan's noisy intervals per chunk (DigarPlan.sizes())."""
    def __init__(self, digar_plan, classify_plan, params, n_sites, n_nreg, sdust_plan=None):
        """sdust_plan: a SdustPlan (one window per chunk, intervals in the chunk's coordinates) whose intervals K2c reads in HBM instead of params' low_beg / low_end
        (lcd_noisyreg_plan_create_on_sdust)"""
        self.digar_plan, self.classify_plan, self.sdust_plan, self.n_sites = digar_plan, classify_plan, sdust_plan, list(n_sites)
        n = len(params)
        if sdust_plan is not None: params = [dict(d, n_low=0, low_beg=np.zeros(0, np.int64), low_end=np.zeros(0, np.int64)) for d in params]
        self.keep = [(np.ascontiguousarray(np.append(np.asarray(d["low_beg"]), 0), np.int64), np.ascontiguousarray(np.append(np.asarray(d["low_end"]), 0), np.int64)) for d in params]
        self.par = (NoisyRegParams * max(n, 1))(*[NoisyRegParams(d["min_alt_dp"], d["noisy_reg_flank_len"], d.get("is_ont", 0), 0, d["min_af"], d["n_low"], a.ctypes.data, b.ctypes.data)
                                                  for d, (a, b) in zip(params, self.keep)])
        if sdust_plan is None:
            lib().lcd_noisyreg_plan_create_on_classify.restype = C.c_void_p
            super().__init__(lib().lcd_noisyreg_plan_create_on_classify(digar_plan.h, classify_plan.h, C.c_int(n), self.par), n)
        else:
            lib().lcd_noisyreg_plan_create_on_sdust.restype = C.c_void_p
            super().__init__(lib().lcd_noisyreg_plan_create_on_sdust(digar_plan.h, classify_plan.h, sdust_plan.h, C.c_int(n), self.par), n)
        self.res, outs = [], []
        for k, m in zip(self.n_sites, n_nreg):
            cap = int(k) + int(m) + 8
            r = dict(var_cate=np.full(k + 1, -7, np.int32), keep=np.zeros(k + 1, np.uint8), reg_beg=np.zeros(cap, np.int64), reg_end=np.zeros(cap, np.int64), reg_label=np.zeros(cap, np.int32))
            self.res.append(r)
            outs.append(NoisyRegOutput(r["var_cate"].ctypes.data, r["keep"].ctypes.data, r["reg_beg"].ctypes.data, r["reg_end"].ctypes.data, r["reg_label"].ctypes.data, cap, 0))
        self.outs = (NoisyRegOutput * max(n, 1))(*outs)

    def fetch(self, stream=None):
        _check(lib().lcd_noisyreg_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_noisyreg_plan_fetch")
        return _noisyreg_results([dict(n_sites=k) for k in self.n_sites], self.outs, self.res)


# ----------------------------------------------------------------------------- K0: low-complexity intervals (symmetric DUST)
class SdustInput(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("l_seq", C.c_int32), ("T", C.c_int32), ("W", C.c_int32), ("pad", C.c_int32), ("base", C.c_int64)]


class SdustOutput(C.Structure):
    _fields_ = [("beg", C.c_void_p), ("end", C.c_void_p), ("cap", C.c_int64), ("n", C.c_int64)]


def _sdust_structs(seqs, T, W, base):
    n = len(seqs)
    keep = [np.ascontiguousarray(np.frombuffer(bytes(s), np.uint8) if not isinstance(s, np.ndarray) else s, dtype=np.uint8) for s in seqs]
    keep = [np.append(k, np.uint8(0)) for k in keep]
    bases = np.broadcast_to(np.asarray(base, np.int64), (n,))
    ins = (SdustInput * max(n, 1))(*[SdustInput(k.ctypes.data, len(k) - 1, T, W, 0, int(b)) for k, b in zip(keep, bases)])
    res = [(np.zeros((len(k) - 1) // 4 + 16, np.int64), np.zeros((len(k) - 1) // 4 + 16, np.int64)) for k in keep]
    outs = (SdustOutput * max(n, 1))(*[SdustOutput(b.ctypes.data, e.ctypes.data, len(b), 0) for b, e in res])
    return ins, outs, keep, res


def sdust_batch(seqs, T=5, W=20, base=0):
    """Drop-in batch call over HOST buffers (lcd_sdust_batch): per ASCII sequence the low-complexity intervals sdust() returns, `base` added
    -> [[(beg, end), ...]]"""
    ins, outs, keep, res = _sdust_structs(seqs, T, W, base)
    _check(lib().lcd_sdust_batch(C.c_int(len(seqs)), ins, outs), "lcd_sdust_batch")
    return [list(zip(b[:outs[i].n].tolist(), e[:outs[i].n].tolist())) for i, (b, e) in enumerate(res)]


class SdustPlan(_Plan):
    def __init__(self, seqs, T=5, W=20, base=0):
        self.ins, self.outs, self.keep, self.res = _sdust_structs(seqs, T, W, base)
        lib().lcd_sdust_plan_create.restype = C.c_void_p
        super().__init__(lib().lcd_sdust_plan_create(C.c_int(len(seqs)), self.ins), len(seqs))

    def fetch(self, stream=None):
        _check(lib().lcd_sdust_plan_fetch(self.h, C.c_void_p(stream or 0), self.outs), "lcd_sdust_plan_fetch")
        return [list(zip(b[:self.outs[i].n].tolist(), e[:self.outs[i].n].tolist())) for i, (b, e) in enumerate(self.res)]


# ----------------------------------------------------------------------------- K5: POA
class PoaParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_ext1", C.c_int32),
                ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("sub_aln", C.c_int32), ("max_n_cons", C.c_int32)]


POA_PARAMS_DTYPE = np.dtype([("match", np.int32), ("mismatch", np.int32), ("gap_open1", np.int32), ("gap_ext1", np.int32),
                             ("gap_open2", np.int32), ("gap_ext2", np.int32), ("wb", np.int32), ("wf", np.float32),
                             ("sub_aln", np.int32), ("max_n_cons", np.int32)])
POA_RESULT_DTYPE = np.dtype([(n, np.int32) for n in ("status", "cons_len", "msa_len", "n_nodes")])


def poa_params(sub_aln=1, wb=10, wf=0.01):
    """longcallD's abPOA set-ups (reference src/align.c:769-783 phased; :876-889 de-novo uses wb=-1)."""
    return (2, 6, 6, 2, 24, 1, wb, wf, sub_aln, 1)


def pack_poa(problems):
    """[[read, read, ...], ...] -> (seqs, first_read, n_reads, read_off, read_len)"""
    reads = [r for p in problems for r in p]
    n_reads = np.fromiter((len(p) for p in problems), dtype=np.int32, count=len(problems))
    first = np.zeros(len(problems), dtype=np.int32)
    if len(problems) > 1:
        first[1:] = np.cumsum(n_reads[:-1])
    read_len = np.fromiter((len(r) for r in reads), dtype=np.int32, count=len(reads))
    read_off = np.zeros(len(reads), dtype=np.int64)
    if len(reads) > 1:
        read_off[1:] = np.cumsum(read_len[:-1].astype(np.int64))
    seqs = np.concatenate([np.asarray(r, dtype=np.uint8) for r in reads]) if reads else np.zeros(1, np.uint8)
    return np.ascontiguousarray(seqs), first, n_reads, read_off, read_len


def _poa_params_array(params, n):
    arr = np.zeros(n, dtype=POA_PARAMS_DTYPE)
    if isinstance(params, list) and len(params) == n and isinstance(params[0], (tuple, list)):
        for i, p in enumerate(params):
            arr[i] = tuple(p)
    else:
        arr[:] = tuple(params)
    return arr


class PoaPlan(_Plan):
    def __init__(self, seqs, first_read, n_reads, read_off, read_len, params):
        n = len(n_reads)
        self.seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        self.first = np.ascontiguousarray(first_read, dtype=np.int32)
        self.n_reads = np.ascontiguousarray(n_reads, dtype=np.int32)
        self.read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        self.read_len = np.ascontiguousarray(read_len, dtype=np.int32)
        self.params = _poa_params_array(params, n)
        h = lib().lcd_poa_plan_create(C.c_int(n), _ptr(self.seqs, C.c_uint8), C.c_size_t(self.seqs.size),
                                      _ptr(self.first, C.c_int32), _ptr(self.n_reads, C.c_int32),
                                      _ptr(self.read_off, C.c_int64), _ptr(self.read_len, C.c_int32),
                                      C.c_int(len(self.read_len)), self.params.ctypes.data_as(C.c_void_p))
        super().__init__(h, n)

    def layout(self):
        """Host output layout: consensus capacity = sum of read lengths; MSA capacity (n_reads+1) x (2*max_len+64)."""
        sum_len = np.add.reduceat(self.read_len.astype(np.int64), self.first) if self.n else np.zeros(0, np.int64)
        max_len = np.maximum.reduceat(self.read_len, self.first).astype(np.int64) if self.n else np.zeros(0, np.int64)
        cons_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(sum_len, out=cons_off[1:])
        msa_cap = (self.n_reads.astype(np.int64) + 1) * (2 * max_len + 64)
        msa_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(msa_cap, out=msa_off[1:])
        return cons_off, msa_off, msa_cap

    def fetch(self, stream=None, want_msa=True):
        cons_off, msa_off, msa_cap = self.layout()
        cons = np.zeros(max(int(cons_off[-1]), 1), dtype=np.uint8)
        msa = np.zeros(max(int(msa_off[-1]), 1), dtype=np.uint8) if want_msa else None
        res = np.zeros(self.n, dtype=POA_RESULT_DTYPE)
        rc = lib().lcd_poa_plan_fetch(self.h, C.c_void_p(stream or 0), _ptr(cons, C.c_uint8), _ptr(cons_off, C.c_int64),
                                      _ptr(msa, C.c_uint8) if want_msa else None, _ptr(msa_off, C.c_int64) if want_msa else None,
                                      _ptr(msa_cap, C.c_int64) if want_msa else None, res.ctypes.data_as(C.c_void_p))
        _check(rc, "lcd_poa_plan_fetch")
        return res, cons, cons_off, msa, msa_off


def poa_batch(problems, params, want_msa=True, sub=None):
    """Drop-in batch call over HOST buffers (lcd_poa_batch; lcd_poa_sub_batch when sub = [(sub_beg, sub_end) per problem] names the reads that
    cover their region only partially).  -> [(status, consensus bytes, msa (n+1, msa_len))]"""
    n = len(problems)
    if n == 0:
        return []
    seqs, first, n_reads, read_off, read_len = pack_poa(problems)
    sb = np.ascontiguousarray(np.concatenate([np.asarray(b, np.int32) for b, _ in sub])) if sub is not None else None
    se = np.ascontiguousarray(np.concatenate([np.asarray(e, np.int32) for _, e in sub])) if sub is not None else None
    par = _poa_params_array(params, n)
    sum_len = np.add.reduceat(read_len.astype(np.int64), first)
    max_len = np.maximum.reduceat(read_len, first).astype(np.int64)
    cons_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sum_len, out=cons_off[1:])
    msa_cap = (n_reads.astype(np.int64) + 1) * (2 * max_len + 64)
    msa_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(msa_cap, out=msa_off[1:])
    cons = np.zeros(max(int(cons_off[-1]), 1), dtype=np.uint8)
    msa = np.zeros(max(int(msa_off[-1]), 1), dtype=np.uint8) if want_msa else None
    res = np.zeros(n, dtype=POA_RESULT_DTYPE)
    tail = (par.ctypes.data_as(C.c_void_p), _ptr(cons, C.c_uint8), _ptr(cons_off, C.c_int64),
            _ptr(msa, C.c_uint8) if want_msa else None, _ptr(msa_off, C.c_int64) if want_msa else None,
            _ptr(msa_cap, C.c_int64) if want_msa else None, res.ctypes.data_as(C.c_void_p))
    head = (C.c_int(n), _ptr(seqs, C.c_uint8), C.c_size_t(seqs.size), _ptr(first, C.c_int32), _ptr(n_reads, C.c_int32), _ptr(read_off, C.c_int64),
            _ptr(read_len, C.c_int32), C.c_int(len(read_len)))
    if sub is None:
        rc = lib().lcd_poa_batch(*head, *tail)
    else:
        rc = lib().lcd_poa_sub_batch(*head, _ptr(sb, C.c_int32), _ptr(se, C.c_int32), *tail)
    _check(rc, "lcd_poa_batch")
    out = []
    for i in range(n):
        r = res[i]
        c = cons[cons_off[i]:cons_off[i] + r["cons_len"]].tobytes()
        m = None
        if want_msa:
            rows = int(n_reads[i]) + 1
            m = msa[msa_off[i]:msa_off[i] + rows * int(r["msa_len"])].reshape(rows, int(r["msa_len"])).copy()
        out.append((int(r["status"]), c, m))
    return out


def poa_ncons_batch(problems, params, min_freq=0.20):
    """De-novo POA with up to two consensus sequences (lcd_poa_ncons_batch: abpoa_aln_msa_cons, max_n_cons = 2) over HOST buffers.
    -> [(status, [consensus bytes per cluster], read clusters (uint8 per read), msa (n_reads + n_cons, msa_len))]"""
    n = len(problems)
    if n == 0:
        return []
    seqs, first, n_reads, read_off, read_len = pack_poa(problems)
    par = _poa_params_array(params, n)
    mf = np.ascontiguousarray(np.broadcast_to(np.asarray(min_freq, dtype=np.float64), (n,)))
    sum_len = np.add.reduceat(read_len.astype(np.int64), first)
    max_len = np.maximum.reduceat(read_len, first).astype(np.int64)
    cons_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sum_len, out=cons_off[1:])
    msa_cap = (n_reads.astype(np.int64) + 2) * (2 * max_len + 64)
    msa_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(msa_cap, out=msa_off[1:])
    cons = np.zeros(max(int(cons_off[-1]), 1), dtype=np.uint8)
    msa = np.zeros(max(int(msa_off[-1]), 1), dtype=np.uint8)
    res = np.zeros(n, dtype=POA_RESULT_DTYPE)
    n_cons, len2, clu = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(max(len(read_len), 1), np.uint8)
    rc = lib().lcd_poa_ncons_batch(C.c_int(n), _ptr(seqs, C.c_uint8), C.c_size_t(seqs.size), _ptr(first, C.c_int32), _ptr(n_reads, C.c_int32),
                                   _ptr(read_off, C.c_int64), _ptr(read_len, C.c_int32), C.c_int(len(read_len)), par.ctypes.data_as(C.c_void_p),
                                   _ptr(mf, C.c_double), _ptr(cons, C.c_uint8), _ptr(cons_off, C.c_int64), _ptr(msa, C.c_uint8), _ptr(msa_off, C.c_int64),
                                   _ptr(msa_cap, C.c_int64), res.ctypes.data_as(C.c_void_p), _ptr(n_cons, C.c_int32), _ptr(len2, C.c_int32), _ptr(clu, C.c_uint8))
    _check(rc, "lcd_poa_ncons_batch")
    out = []
    for i in range(n):
        r = res[i]
        l1, l2, nc = int(r["cons_len"]), int(len2[i]), int(n_cons[i])
        cs = [cons[cons_off[i]:cons_off[i] + l1].tobytes()] + ([cons[cons_off[i] + l1:cons_off[i] + l1 + l2].tobytes()] if nc == 2 else [])
        rows = int(n_reads[i]) + nc
        m = msa[msa_off[i]:msa_off[i] + rows * int(r["msa_len"])].reshape(rows, int(r["msa_len"])).copy()
        out.append((int(r["status"]), cs, clu[first[i]:first[i] + n_reads[i]].copy(), m))
    return out
