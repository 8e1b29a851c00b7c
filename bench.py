#!/usr/bin/env python3
"""bench.py -- throughput of the B200 re-alignment hot path on synthetic 30x long-read noisy regions.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mbp M] [--tech hifi|ont]

A *step* is one pass of the hot path over one batch: all noisy regions of `--mbp` megabases of
reference per GPU (default 50 Mb: BASELINE.json configs[1], "synthetic HiFi 30x, 50 Mb ref, 1 GPU").
The metric is reference megabases called per second.  `value` is the device-resident number (inputs in
HBM, CUDA events on the library stream); `e2e` goes through the host-buffer C-ABI batch calls with
H2D/D2H inside the timed region.  One JSON line on stdout (rank 0).

`--impl reference` times the UNMODIFIED reference libraries (oracle/_ref/libref_shim.so: WFA2-lib /
edlib / abPOA compiled from /root/reference) on the same workload with all host threads, each step a
bounded sample of the batch.  That is the only place this file touches oracle/ (as the baseline, never
as the product path).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ref_Mbp_per_s_called"
UNIT = "Mbp/s"
WFA_BYTES_PER_CELL = 48          # 5 components written + 7 read, int32
POA_BYTES_PER_CELL = 16          # 5 int16 planes written + 3 predecessor planes read
EDLIB_BYTES_PER_BLOCKCOL = 28    # Peq word in + (P, M, score) stored for the traceback
PHASE_BYTES_PER_PAIR = 24        # allele + variant state in, counts out, per (read, variant) pair and pass


NCU_CAPTURES = ("r2_poa_full_v4.raw.csv", "r2_poa_full_v2.raw.csv", "r1_poa_full_v5.raw.csv", "r1_poa_full_v3.raw.csv")       # newest first


def ncu_capture():
    for f in NCU_CAPTURES:
        if os.path.exists(os.path.join(ROOT, "profiles", f)):
            return f
    return None


def ncu_traffic(kernel="poa_kernel"):
    """DRAM bytes (read + write) of one launch of the dominant kernel from the committed `ncu --set full` capture of this
    same workload (profiles/r2_poa_full_v*.raw.csv: tools/poa_prof.py 50 = the step's POA batch); None when the file is missing."""
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", ncu_capture()))))
        h, u, v = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(name)
            tot += float(v[i].replace(",", "")) * scale[u[i]]
        return tot if kernel in v[h.index("Kernel Name")] else None
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ workload
class Workload:
    """All noisy regions of `mbp` Mb: per (region, haplotype) one POA problem (its reads) and one
    ref-vs-consensus WFA problem whose text is that POA's consensus (wfa_collect_aln_str, align.c:565)."""

    def __init__(self, mbp, tech, seed):
        from longcalld_b200 import synth
        from longcalld_b200.capi import pack_poa
        self.mbp = mbp
        regions = synth.make_regions(mbp, tech, seed=seed, with_reads=True)
        self.n_regions = len(regions)
        self.problems, self.refs, self.region_of = [], [], []
        for ri, r in enumerate(regions):
            for hap in (1, 2):
                reads = [s for s, h in zip(r.reads, r.read_hap) if h == hap and len(s) > 0]
                if reads:
                    self.problems.append(reads); self.refs.append(r.ref); self.region_of.append(ri)
        self.region_of = np.asarray(self.region_of)
        # K4: the 500 kb chunks of the batch, phased twice per step (clean mask, then germline mask: collect_var.c:2942,2975)
        self.phase = [(d, mask, int(tech == "ont")) for d in synth.phase_chunks(mbp, tech, seed) for mask in (synth.CATE_CLEAN, synth.CATE_GERMLINE)]
        # K7: the sampling filter's read-vs-first-read NW paths (align.c:722-733)
        from longcalld_b200.capi import pack_pairs
        self.edlib = pack_pairs(synth.edlib_pairs(regions, seed=seed, mbp=mbp))
        self.seqs, self.first, self.n_reads, self.read_off, self.read_len = pack_poa(self.problems)
        self.n_poa = len(self.problems)
        self.n_poa_reads = int(self.n_reads.sum())
        self.sum_len = np.add.reduceat(self.read_len.astype(np.int64), self.first)
        self.cons_off = np.zeros(self.n_poa + 1, dtype=np.int64)
        np.cumsum(self.sum_len, out=self.cons_off[1:])

    def wfa_layout(self):
        """Zero-copy hand-over POA -> WFA: one host buffer [region references | consensus slots]; lcd_poa_batch writes the
        consensus of problem i into its slot and lcd_wfa_batch reads (reference, consensus) pairs by offset."""
        if not hasattr(self, "_layout"):
            seen, parts, off, ref_off = {}, [], 0, np.zeros(self.n_poa, dtype=np.int64)
            for i, r in enumerate(self.refs):
                if id(r) not in seen:
                    seen[id(r)] = off; parts.append(np.asarray(r, dtype=np.uint8)); off += len(r)
                ref_off[i] = seen[id(r)]
            ref_len = np.fromiter((len(r) for r in self.refs), dtype=np.int32, count=self.n_poa)
            buf = np.zeros(off + int(self.cons_off[-1]) + 16, dtype=np.uint8)
            buf[:off] = np.concatenate(parts)
            self._layout = (buf, off, ref_off, ref_len, np.ascontiguousarray(off + self.cons_off[:-1]))
        return self._layout

    def wfa_inputs(self, cons, cons_len, idx=None):
        """(ref, consensus) pairs packed for the reference arm; consensus i = cons[cons_off[i] : +cons_len[i]]."""
        from longcalld_b200.capi import pack_pairs
        idx = range(self.n_poa) if idx is None else idx
        pairs = [(self.refs[i], cons[self.cons_off[i]:self.cons_off[i] + cons_len[i]]) for i in idx]
        return pack_pairs(pairs)

    def subset(self, regs):
        """Indices of the problems of the given (sorted) region ids."""
        return np.nonzero(np.isin(self.region_of, regs))[0]


class PileupStage:
    """K1 -> K1b -> K2 -> K3 of the same batch: mbp / 0.5 region chunks (500 kb, 30x, ~1 050 reads of ~15 kb with =/X CIGARs), tiled
    from a template of <= 10 distinct chunks.  The classification (a5) is not on the GPU yet: the classified variants K3 runs on are
    derived once at set-up from K2's counters (synth.classify_sites) and that step is not timed in either arm."""
    N_TEMPLATE = 10

    def __init__(self, mbp, tech, seed, digar_fn, sites_fn, pileup_fn, classify_fn, noisyreg_fn, pin=False):
        from longcalld_b200 import synth
        self.n_chunks = max(1, int(round(mbp / 0.5)))
        nt = min(self.N_TEMPLATE, self.n_chunks)
        self.template = synth.digar_chunks_30x(nt, tech, seed)
        if pin:                                     # the loader's buffers: page-locked, so the e2e H2D copies run at PCIe speed
            import torch
            for d in self.template:
                for k in ("cigar", "bseq", "qual"):
                    t = torch.from_numpy(d[k]).pin_memory(); d["_pin_" + k] = t; d[k] = t.numpy()
        outs = digar_fn(self.template)
        self.regs = [(int(d["reg_beg"]), int(d["reg_end"])) for d in self.template]
        bare = [synth.pileup_input_from_digar(d, o, synth.empty_site_list()) for d, o in zip(self.template, outs)]      # K1b's input: the lists without sites
        raw = sites_fn(bare, outs, self.regs)
        piles = [synth.pileup_input_from_digar(d, o, s) for d, o, s in zip(self.template, outs, raw)]
        counts = pileup_fn(piles)
        self.cls = [synth.classify_input_from_sites(d, s, c, seed + 977 * i, is_ont=int(tech == "ont")) for i, (d, s, c) in enumerate(zip(self.template, raw, counts))]      # K2b's input
        cates = classify_fn(self.cls)
        nreg = [synth.noisyreg_input_from(d, o, s, ct, seed + 311 * i, is_ont=int(tech == "ont")) for i, (d, o, s, ct) in enumerate(zip(self.template, outs, raw, cates))]      # K2c's input
        kept = noisyreg_fn(nreg, self.cls)                                   # K2c: the sites that stay, with their categories
        var = [synth.kept_site_list(s, k["keep"], k["var_cate"]) for s, k in zip(raw, kept)]       # chunk->cand_vars after classify_cand_vars: K3's input
        profs = [synth.pileup_input_from_digar(d, o, s) for d, o, s in zip(self.template, outs, var)]
        tile = lambda xs: [xs[i % nt] for i in range(self.n_chunks)]
        self.chunks, self.raw_sites, self.var_sites, self.piles, self.profs = tile(self.template), tile(raw), tile(var), tile(piles), tile(profs)
        self.bare, self.regs, self.cls, self.nreg = tile(bare), tile(self.regs), tile(self.cls), tile(nreg)
        self.n_reads = [d["n_reads"] for d in self.chunks]
        self.read_bases = int(sum(int(d["l_qseq"].sum()) for d in self.chunks))
        self.cigar_ops = int(sum(int(d["n_cigar"].sum()) for d in self.chunks))
        self.records = int(sum(int(outs[i % nt]["n_digar_total"]) for i in range(self.n_chunks)))
        self.n_raw_sites = int(sum(s["n_sites"] for s in self.raw_sites)); self.n_vars = int(sum(s["n_sites"] for s in self.var_sites))
        self.h2d = int(sum(d["cigar"].nbytes + d["bseq"].nbytes + d["qual"].nbytes + 46 * d["n_reads"] for d in self.chunks) +
                       sum(sum(s[k].nbytes for k in ("site_pos", "site_type", "site_ref_len", "site_alt_len", "site_alt_off", "site_alt")) for s in self.var_sites))
        # algorithmic bytes: K1 1.5 B per read base + 4 B per CIGAR op in, 32 B per record out; K2 / K3 32 B per record + 24 B per site (+ 8 B per profile entry)
        self.k1_bytes = int(1.5 * self.read_bases + 4 * self.cigar_ops + 32 * self.records)
        self.k1b_bytes = int(14 * self.records + 36 * self.n_raw_sites)      # K1b: position, type, length and quality flag of every record in, a 36-byte site record out
        self.k2_bytes = int(32 * self.records + 24 * self.n_raw_sites)
        self.k2b_bytes = int((32 + 24 + 4) * self.n_raw_sites)              # K2b: the site's counters and record in, its category out (+ ~36 reference bases for small indels)
        self.k3_bytes = int(32 * self.records + 24 * self.n_vars)
        # K2c: a site record + category in, category + keep flag out; per read its span and noisy intervals; the interval lists (12 B each); records only where asked
        self.n_low = int(sum(x["n_low"] for x in self.nreg)); self.n_cnreg = int(sum(x["n_cnreg"] for x in self.nreg))
        self.k2c_bytes = int((20 + 5) * self.n_raw_sites + 24 * sum(self.n_reads) + 16 * self.n_low + 20 * self.n_cnreg)
        self.k2c_h2d = int(sum(np.asarray(x["low_beg"]).nbytes + np.asarray(x["low_end"]).nbytes + 40 for x in self.nreg))      # K2c in place: options + low-complexity intervals


# ------------------------------------------------------------------------------------------ reference arm
def ref_shim():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_shim.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            return None
    return C.CDLL(path)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def reference_step(lib, wl, idx, n_threads, keep=None):
    """The reference's CPU implementation (abPOA then WFA2-lib) of the same stages on problems `idx`.
    keep: dict that receives the reference's outputs (consensus, CIGAR ops, phasing, edlib paths) for the parity check of the bench line."""
    from longcalld_b200.capi import WFA_PARAMS_DTYPE, WFA_RESULT_DTYPE, POA_PARAMS_DTYPE, wfa_params, poa_params
    idx = np.asarray(idx, dtype=np.int64)
    n = len(idx)
    first, n_reads = np.ascontiguousarray(wl.first[idx]), np.ascontiguousarray(wl.n_reads[idx])
    cons_off = np.ascontiguousarray(wl.cons_off[idx])
    ppar = np.zeros(n, dtype=POA_PARAMS_DTYPE); ppar[:] = poa_params()
    cons = np.zeros(int(wl.cons_off[-1]) + 16, dtype=np.uint8)
    cons_len = np.zeros(n, dtype=np.int32)
    t0 = time.perf_counter()
    lib.ref_poa_batch(C.c_int(n), _vp(wl.seqs), _vp(first), _vp(n_reads), _vp(wl.read_off), _vp(wl.read_len), _vp(ppar),
                      _vp(cons), _vp(cons_off), _vp(cons_len), C.c_int(n_threads))
    t1 = time.perf_counter()
    full_len = np.zeros(wl.n_poa, dtype=np.int32); full_len[idx] = cons_len
    seqs, po, pl, to, tl = wl.wfa_inputs(cons, full_len, idx)
    wpar = np.zeros(n, dtype=WFA_PARAMS_DTYPE); wpar[:] = wfa_params()
    cap = 2 * (pl.astype(np.int64) + tl) + 8
    off = np.zeros(n + 1, dtype=np.int64); np.cumsum(cap, out=off[1:])
    ops = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
    res = np.zeros(n, dtype=WFA_RESULT_DTYPE)
    t2 = time.perf_counter()
    lib.ref_wfa_batch(C.c_int(n), _vp(seqs), _vp(po), _vp(pl), _vp(to), _vp(tl), _vp(wpar), _vp(ops), _vp(off), _vp(res), C.c_int(n_threads))
    t3 = time.perf_counter()
    # K4 / K7 on the same fraction of the batch
    from longcalld_b200.capi import _phase_structs, EDLIB_RESULT_DTYPE
    frac = n / max(1, wl.n_poa)
    ph = wl.phase[:max(2, 2 * int(round(frac * len(wl.phase) / 2)))]
    ins, outs, ph_keep_, ph_results = _phase_structs(ph, -9)
    t4 = time.perf_counter()
    lib.ref_phase_batch(C.c_int(len(ph)), ins, outs, C.c_int(n_threads))
    t5 = time.perf_counter()
    eseqs, qo, ql, to_, tl_ = wl.edlib
    ne = max(1, int(round(frac * len(ql))))
    m = np.zeros(ne, np.int32); w = np.ones(ne, np.int32); eres = np.zeros(ne, dtype=EDLIB_RESULT_DTYPE)
    ecap = ql[:ne].astype(np.int64) + tl_[:ne] + 2
    eoff = np.zeros(ne + 1, dtype=np.int64); np.cumsum(ecap, out=eoff[1:])
    ealn = np.zeros(int(eoff[-1]) + 1, dtype=np.uint8)
    t6 = time.perf_counter()
    lib.ref_edlib_batch(C.c_int(ne), _vp(eseqs), _vp(qo), _vp(ql), _vp(to_), _vp(tl_), _vp(m), _vp(w), _vp(ealn), _vp(eoff), _vp(eres), C.c_int(n_threads))
    t7 = time.perf_counter()
    if keep is not None:
        keep.update(idx=idx, cons=cons, cons_len=cons_len, wfa_ops=ops, wfa_off=off, wfa_res=res, phase=ph_results,
                    n_edlib=ne, edlib_res=eres, edlib_aln=ealn, edlib_off=eoff)
    return (t1 - t0) + (t3 - t2) + (t5 - t4) + (t7 - t6), (t1 - t0), (t3 - t2), (t5 - t4), (t7 - t6)


def parity_check(ref, pile_ref, wl, ps, gpu):
    """The reference arm's outputs of the bounded sample against the GPU outputs of the same batch (the last e2e batch): every byte
    of every consensus, every CIGAR op, haplotype / phase set / per-variant consensus allele, edlib path, difference-list record, site,
    counter, category and profile row.  Raises on the first difference; returns what was compared."""
    idx = ref["idx"]
    n_cons = n_ops = 0
    for j, i in enumerate(idx.tolist()):
        o, l = int(wl.cons_off[i]), int(ref["cons_len"][j])
        if l != int(gpu["cons_len"][i]) or not np.array_equal(ref["cons"][o:o + l], gpu["cons"][o:o + l]):
            raise AssertionError(f"parity: consensus of POA problem {i} differs from the reference's")
        rr, gr = ref["wfa_res"][j], gpu["wfa_res"][i]
        if tuple(rr) != tuple(gr):
            raise AssertionError(f"parity: WFA result record of problem {i} differs: {tuple(rr)} vs {tuple(gr)}")
        k = int(rr["n_ops"]); a, b = int(ref["wfa_off"][j]), int(gpu["wfa_off"][i])
        if not np.array_equal(ref["wfa_ops"][a:a + k], gpu["wfa_ops"][b:b + k]):
            raise AssertionError(f"parity: CIGAR of WFA problem {i} differs from the reference's")
        n_cons += l; n_ops += k
    for c, (r, g) in enumerate(zip(ref["phase"], gpu["phase"])):
        for key in r:
            if not np.array_equal(r[key], g[key]):
                raise AssertionError(f"parity: phasing output {key} of chunk pass {c} differs from the reference's")
    ne = ref["n_edlib"]
    if not np.array_equal(ref["edlib_res"], gpu["edlib_res"][:ne]):
        raise AssertionError("parity: edlib result records differ from the reference's")
    for e in range(ne):
        k = int(ref["edlib_res"]["aln_len"][e]); a, b = int(ref["edlib_off"][e]), int(gpu["edlib_off"][e])
        if not np.array_equal(ref["edlib_aln"][a:a + k], gpu["edlib_aln"][b:b + k]):
            raise AssertionError(f"parity: edlib path {e} differs from the reference's")
    k = pile_ref["k"]
    for c in range(k):
        rs, gs = pile_ref["sites"][c], gpu["sites"][c]
        if rs["n_sites"] != gs["n_sites"] or any(not np.array_equal(rs[key], gs[key]) for key in ("site_pos", "site_type", "site_ref_len", "site_alt_len")):
            raise AssertionError(f"parity: candidate sites of chunk {c} differ from the reference's")
        if not np.array_equal(pile_ref["counts"][c][:rs["n_sites"]], gpu["counts"][c][:rs["n_sites"]]):
            raise AssertionError(f"parity: coverage counters of chunk {c} differ from the reference's")
        if not np.array_equal(pile_ref["cate"][c][:rs["n_sites"]], gpu["cate"][c][:rs["n_sites"]]):
            raise AssertionError(f"parity: site categories of chunk {c} differ from the reference's")
        # K2c: the compacted cand_vars (position, type, ref_len, category of the kept sites, in order) and chunk_noisy_regs
        gn, rn, rk = gpu["nreg"][c], pile_ref["nreg_regs"][c], pile_ref["nreg_kept"][c]
        x = ps.nreg[c]; kidx = np.nonzero(gn["keep"])[0]
        if not (np.array_equal(np.asarray(x["site_pos"])[kidx], rk[0]) and np.array_equal(np.asarray(x["site_type"])[kidx], rk[1]) and np.array_equal(np.asarray(x["site_ref_len"])[kidx], rk[2])
                and np.array_equal(gn["var_cate"][kidx], rk[3])):
            raise AssertionError(f"parity: the candidate sites kept after classify_cand_vars of chunk {c} differ from the reference's")
        if gn["n_regs"] != rn["n_regs"] or any(not np.array_equal(gn[key], rn[key]) for key in ("reg_beg", "reg_end", "reg_label")):
            raise AssertionError(f"parity: noisy regions of chunk {c} differ from the reference's")
        rp, gp = pile_ref["prof"][c], gpu["prof"][c]
        nr = ps.n_reads[c]
        if not (np.array_equal(rp[0][:nr], gp["prof_start"][:nr]) and np.array_equal(rp[1][:nr], gp["prof_end"][:nr])):
            raise AssertionError(f"parity: profile row ranges of chunk {c} differ from the reference's")
        for r in range(nr):
            m = int(rp[1][r]) - int(rp[0][r]) + 1
            if m <= 0: continue
            a, b = int(rp[2][r]), int(gp["allele_off"][r])
            if not (np.array_equal(rp[3][a:a + m], gp["alleles"][b:b + m]) and np.array_equal(rp[4][a:a + m], gp["alt_qi"][b:b + m])):
                raise AssertionError(f"parity: profile row of read {r} of chunk {c} differs from the reference's")
        from longcalld_b200.check import digar_view, digar_same
        digar_same(digar_view(ps.chunks[c], gpu["digar"][c]), digar_view(ps.chunks[c], pile_ref["digar"][c]), f"parity: difference lists of chunk {c}")
    return {"poa_consensus": len(idx), "consensus_bases": int(n_cons), "wfa_alignments": len(idx), "cigar_ops": int(n_ops), "phase_chunk_passes": len(ref["phase"]),
            "edlib_paths": int(ne), "pileup_chunks": int(k), "sites": int(sum(s["n_sites"] for s in pile_ref["sites"])), "what": "reference arm (unmodified reference functions) == GPU e2e outputs of the same batch, bit for bit"}


def ref_pileup_fns(lib, n_threads):
    """K1 / K2 / K3 of the unmodified reference (oracle/_ref/libref_shim.so) behind the same dict interface as longcalld_b200.capi."""
    from longcalld_b200 import capi

    def digar(chunks):
        ins, keep = capi._digar_inputs(chunks)
        outs, results = capi._digar_outputs(chunks, capi.digar_capacity_host(chunks))
        if lib.ref_digar_batch(C.c_int(len(chunks)), ins, outs, C.c_int(n_threads), None): raise RuntimeError("ref_digar_batch failed")
        return capi._digar_finish(outs, results)

    def sites(bare, digar_outs, regs):
        from longcalld_b200 import synth
        ins, _, keep, _ = capi._pileup_structs(bare)
        souts, res = capi._sites_outputs([int(np.isin(np.asarray(d["digar_type"]), (1, 2, 8)).sum()) for d in bare])
        reg = np.array(regs, np.int64).reshape(-1)
        if lib.ref_sites_batch(C.c_int(len(bare)), ins, _vp(reg), souts, C.c_int(n_threads), None): raise RuntimeError("ref_sites_batch failed")
        return [synth.site_list_from_sites(o, st, src_is_offset=True) for o, st in zip(digar_outs, capi._sites_finish(souts, res))]

    def pileup(piles):
        ins, outs, keep, results = capi._pileup_structs(piles)
        if lib.ref_pileup_batch(C.c_int(len(piles)), ins, outs, C.c_int(n_threads)): raise RuntimeError("ref_pileup_batch failed")
        return [r[:d["n_sites"]] for r, d in zip(results, piles)]

    def classify(cls):
        cins, _, ckeep, cres = capi._classify_structs(cls)
        cptr = (C.c_void_p * max(len(cls), 1))(*[r.ctypes.data for r in cres])
        if lib.ref_classify_batch(C.c_int(len(cls)), cins, cptr, C.c_int(n_threads)): raise RuntimeError("ref_classify_batch failed")
        return [r[:d["n_sites"]] for r, d in zip(cres, cls)]

    def noisyreg(nreg, cls):
        k = len(nreg)
        cins, _, ckeep, _ = capi._classify_structs(cls)
        nins, nouts, nkeep, nres = capi._noisyreg_structs(nreg)
        kp = [[np.zeros(x["n_sites"] + 1, t) for t in (np.int64, np.int32, np.int32, np.int32)] for x in nreg]
        kptr = [(C.c_void_p * max(k, 1))(*[kp[i][j].ctypes.data for i in range(k)]) for j in range(4)]
        nkept = np.zeros(k + 1, np.int32)
        if lib.ref_noisyreg_batch(C.c_int(k), cins, nins, kptr[0], kptr[1], kptr[2], kptr[3], _vp(nkept), nouts, C.c_int(n_threads)): raise RuntimeError("ref_noisyreg_batch failed")
        return capi._noisyreg_results(nreg, nouts, nres)
    return digar, sites, pileup, classify, noisyreg


def reference_pileup_step(lib, ps, k, n_threads, out=None):
    """K1 + K1b + K2 + K2b + K2c + K3 of the reference on the first k chunks of the batch; returns seconds (total, K1, K2, K3, K1b, K2b, K2c).
    (K2c = pre_process_noisy_regs + classify_cand_vars, which runs the reference's classify_var_cate loop again inside: K2b's 3 ms count twice.)"""
    from longcalld_b200 import capi
    chunks, piles, profs = ps.chunks[:k], ps.piles[:k], ps.profs[:k]
    ins, keep = capi._digar_inputs(chunks)
    outs, results = capi._digar_outputs(chunks, capi.digar_capacity_host(chunks))
    pins, pouts, pkeep, pres = capi._pileup_structs(piles)
    fins, _, fkeep, _ = capi._pileup_structs(profs)
    bins, _, bkeep, _ = capi._pileup_structs(ps.bare[:k])
    souts, sres = capi._sites_outputs([int(s["n_sites"]) + 8 for s in ps.raw_sites[:k]])
    reg = np.array(ps.regs[:k], np.int64).reshape(-1)
    cins, _, ckeep, cres = capi._classify_structs(ps.cls[:k])
    cptr = (C.c_void_p * k)(*[r.ctypes.data for r in cres])
    exs, fouts, fres = (capi.ProfileExtra * k)(), (capi.ProfileOutput * k)(), []
    for i, d in enumerate(profs):
        arrs = {kk: np.ascontiguousarray(d[kk], dtype=t) for kk, t in capi._PROFILE_EX}; fkeep.append(arrs)
        exs[i] = capi.ProfileExtra(*[arrs[kk].ctypes.data for kk, _ in capi._PROFILE_EX])
        # rows the profile can need: per read the sites whose span can touch it (numpy upper bound; the reference arm never loads the product library)
        nsv = int(d["n_sites"]); sp = np.asarray(d["site_pos"][:nsv], np.int64); mrl = int(np.asarray(d["site_ref_len"][:nsv]).max()) + 2 if nsv else 2
        nrd = int(d["n_reads"])
        cap = int((np.searchsorted(sp, np.asarray(d["read_end"][:nrd], np.int64) + 2, "right") - np.searchsorted(sp, np.asarray(d["read_beg"][:nrd], np.int64) - mrl, "left") + 2).sum()) + 8
        o = [np.zeros(d["n_reads"] + 1, np.int32), np.zeros(d["n_reads"] + 1, np.int32), np.zeros(d["n_reads"] + 1, np.int64), np.zeros(cap, np.int8), np.zeros(cap, np.int32)]
        fres.append(o); fouts[i] = capi.ProfileOutput(*[a.ctypes.data for a in o], cap, 0)
    core = C.c_double(0.0)        # K1: the reference's own calls only (the shim's bam1_t construction and copy-out are not the reference's work)
    rc = lib.ref_digar_batch(C.c_int(k), ins, outs, C.c_int(n_threads), C.byref(core))
    score = C.c_double(0.0)       # K1b: collect_all_cand_var_sites alone (the shim's digar_t construction is not the reference's work)
    rc |= lib.ref_sites_batch(C.c_int(k), bins, _vp(reg), souts, C.c_int(n_threads), C.byref(score))
    t1 = time.perf_counter()
    rc |= lib.ref_pileup_batch(C.c_int(k), pins, pouts, C.c_int(n_threads))
    t2 = time.perf_counter()
    rc |= lib.ref_classify_batch(C.c_int(k), cins, cptr, C.c_int(n_threads))
    t2b0 = time.perf_counter()
    nins, nouts, nkeep, nres = capi._noisyreg_structs(ps.nreg[:k])
    kp = [[np.zeros(x["n_sites"] + 1, t) for t in (np.int64, np.int32, np.int32, np.int32)] for x in ps.nreg[:k]]
    kptr = [(C.c_void_p * k)(*[kp[i][j].ctypes.data for i in range(k)]) for j in range(4)]
    nkept = np.zeros(k + 1, np.int32)
    t2c0 = time.perf_counter()
    rc |= lib.ref_noisyreg_batch(C.c_int(k), cins, nins, kptr[0], kptr[1], kptr[2], kptr[3], _vp(nkept), nouts, C.c_int(n_threads))
    t2b = time.perf_counter()
    rc |= lib.ref_profile_batch(C.c_int(k), fins, exs, fouts, C.c_int(n_threads))
    t3 = time.perf_counter()
    if rc: raise RuntimeError("reference K1-K3 failed")
    if out is not None:
        out.update(k=k, digar=capi._digar_finish(outs, results), sites=capi._sites_finish(souts, sres), counts=[r for r in pres], cate=cres, prof=fres,
                   nreg_kept=[[a[:int(nkept[i])].copy() for a in kp[i]] for i in range(k)], nreg_regs=capi._noisyreg_results(ps.nreg[:k], nouts, nres))
    return core.value + score.value + (t3 - t1) - (t2c0 - t2b0), core.value, t2 - t1, t3 - t2b, score.value, t2b0 - t2, t2b - t2c0


def region_sample(wl, frac, seed=1):
    """Whole regions, uniformly sampled: a bounded slice of the batch."""
    rng = np.random.default_rng(seed)
    k = max(1, int(round(wl.n_regions * frac)))
    regs = np.sort(rng.choice(wl.n_regions, size=k, replace=False))
    return regs, wl.subset(regs)


def calibrate_sample(lib, wl, n_threads, target_s):
    frac = min(1.0, 150.0 / wl.n_regions)
    regs, idx = region_sample(wl, frac)
    dt = reference_step(lib, wl, idx, n_threads)[0]
    per_region = dt / len(regs)
    return float(min(1.0, max(frac, target_s / max(per_region, 1e-9) / wl.n_regions)))


def whole_program(args, with_gpu):
    """`longcallD call` itself, BAM + FASTA -> VCF, on a synthetic BAM of the same shape (tools/synth_bam.c, tools/whole_program.py): the
    unmodified reference binary on all host threads and -- with_gpu -- the same program with the GPU drop-in preloaded; wall seconds from
    the tool's own `Real time` line, VCF bodies compared by md5.  Bounded: one run each on args.mbp Mb."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import whole_program as wp
        if not wp.available():
            return {"unavailable": "oracle/_ref binaries, tools/_build/synth_bam or the drop-in were not built"}
        if not with_gpu:
            wd = tempfile.mkdtemp(prefix="lcd_wp_")
            fa, bam = wp.make_bam(os.path.join(wd, "synth"), args.mbp, args.tech, args.seed)
            nt = os.cpu_count() or 1
            r = wp.call(os.path.join(wp.REF_DIR, "longcallD_ref"), fa, bam, args.tech, nt)
            shutil.rmtree(wd, ignore_errors=True)
            return {"what": "oracle/_ref/longcallD_ref call, BAM + FASTA -> VCF", "mb": args.mbp, "threads": nt, "real_s": r["real_s"], "cpu_s": r["cpu_s"],
                    "mbp_per_s": args.mbp / r["real_s"], "vcf_md5": r["vcf_md5"], "vcf_records": r["vcf_records"]}
        o = wp.run(args.mbp, args.tech, None, "engines", None, args.seed, reps=2)          # best of two runs each: the boxes' host side is noisy
        return {"what": "longcallD call, BAM + FASTA -> VCF: the unmodified reference binary vs the same program with liblcd_dropin.so preloaded (K5 - K7 batched over regions and threads on the GPU; LCD_DROPIN_STAGES=engines)",
                "mb": o["mb"], "threads": o["threads"], "reference_real_s": o["reference"]["real_s"], "gpu_real_s": o["gpu"]["real_s"],
                "reference_mbp_per_s": o["reference_mbp_s"], "gpu_mbp_per_s": o["gpu_mbp_s"], "speedup": o["speedup"], "vcf_md5_equal": o["vcf_md5_equal"],
                "vcf_records": o["gpu"]["vcf_records"], "reference_cpu_s": o["reference"]["cpu_s"], "gpu_cpu_s": o["gpu"]["cpu_s"], "dropin": o["gpu"]["dropin"]}
    except Exception as e:                      # a measurement that could not be made is reported, not hidden
        return {"error": repr(e)[:300]}


def run_reference(args, rank):
    if rank != 0:
        return
    lib = ref_shim()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_shim.so missing and /root/reference absent"}))
        return
    n_threads = os.cpu_count() or 1
    wl = Workload(args.mbp, args.tech, args.seed)
    ps = PileupStage(args.mbp, args.tech, args.seed, *ref_pileup_fns(lib, n_threads))
    frac = calibrate_sample(lib, wl, n_threads, target_s=max(2.0, 90.0 / max(1, args.steps + args.warmup)))
    regs, idx = region_sample(wl, frac)
    mbp_sample = args.mbp * len(regs) / wl.n_regions
    k_chunks = min(ps.n_chunks, max(1, int(round(ps.n_chunks * len(regs) / wl.n_regions))))
    pile_scale = (len(regs) / wl.n_regions) / (k_chunks / ps.n_chunks)          # K1-K3 run on whole chunks; their time is scaled to the region sample
    for _ in range(args.warmup):
        reference_step(lib, wl, idx, n_threads); reference_pileup_step(lib, ps, k_chunks, n_threads)
    t, tp = [], []
    for _ in range(args.steps):
        t.append(reference_step(lib, wl, idx, n_threads)); tp.append(reference_pileup_step(lib, ps, k_chunks, n_threads))
    total = sum(x[0] for x in t) + pile_scale * sum(x[0] for x in tp)
    value = mbp_sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16/int32", "data": "synthetic",
            "config": workload_config(args, wl, ps),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "reference",
                             "sample": f"{len(regs)} of {wl.n_regions} regions ({mbp_sample:.3f} Mb) per step; K1-K3 on {k_chunks} of {ps.n_chunks} chunks, time scaled by {pile_scale:.4f}",
                             "poa_s": sum(x[1] for x in t) / args.steps, "wfa_s": sum(x[2] for x in t) / args.steps,
                             "phase_s": sum(x[3] for x in t) / args.steps, "edlib_s": sum(x[4] for x in t) / args.steps,
                             "digar_s": pile_scale * sum(x[1] for x in tp) / args.steps, "pileup_s": pile_scale * sum(x[2] for x in tp) / args.steps,
                             "profile_s": pile_scale * sum(x[3] for x in tp) / args.steps, "sites_s": pile_scale * sum(x[4] for x in tp) / args.steps, "classify_s": pile_scale * sum(x[5] for x in tp) / args.steps,
                             "noisyreg_s": pile_scale * sum(x[6] for x in tp) / args.steps},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "whole_program": None if args.no_whole_program else whole_program(args, with_gpu=False)}
    print(json.dumps(line))


def workload_config(args, wl, ps=None):
    return {"workload": f"synthetic {args.tech.upper()} 30x noisy-region re-alignment, {args.mbp:g} Mb ref per GPU "
                        f"(BASELINE configs[1] shape), {wl.n_regions} regions",
            "stages": ["K1 difference lists + noisy intervals + quality histogram from =/X CIGARs per 500 kb chunk (bam_utils.c:701)",
                       "K1b candidate-site list: sorted distinct X/I/D records incl. the large-insertion merge, on K1's lists in HBM (collect_var.c:1209)",
                       "K2 per-site coverage of the candidate sites, on K1's lists and K1b's sites in HBM (collect_var.c:238)",
                       "K2b category of every candidate site: depth / allele-fraction thresholds, homopolymer and repeat context of small indels (collect_var.c:413)",
                       "K2c noisy-region set + the sites that stay clean-region candidates: low-complexity extension, label-window merges, read votes, site-overlap / noisy-read-ratio rules, flank sweep (collect_var.c:557,902)",
                       "K3 read x variant profile of the classified variants, on K1's lists in HBM (collect_var.c:1389)",
                       "K4 read->haplotype assignment + phasing per 500 kb chunk, clean then germline mask (assign_hap.c:473)",
                       "K7 edlib NW path read-vs-first-read sampling filter (align.c:722)",
                       "K5 abPOA consensus+MSA per (region, haplotype) (align.c:762)",
                       "K6 WFA gap-affine-2p ref-vs-consensus (align.c:565)"],
            "stages_not_yet_on_gpu": ["noisy-region orchestration (a8: the reference's host code, run as coroutines by the drop-in), vars from MSA (a13: merge_var_profile restructured in the drop-in, the rest host code), somatic chain (a14)"],
            "not_in_this_step": "regions with partially covering / sampled reads (sub-graph POA, lcd_poa_sub_batch), de-novo regions with two consensus sequences (lcd_poa_ncons_batch) and cs / MD / untagged plain-M reads (K1's tag front end) run on the GPU in the tests and in whole_program; the synthetic step holds full-cover regions and =/X CIGARs",
            "pileup": (None if ps is None else {"chunks": ps.n_chunks, "distinct_chunks": min(ps.N_TEMPLATE, ps.n_chunks), "reads": int(sum(ps.n_reads)),
                                                 "read_bases": ps.read_bases, "cigar_ops": ps.cigar_ops, "records": ps.records,
                                                 "candidate_sites": ps.n_raw_sites, "classified_variants": ps.n_vars}),
            "n_phase_chunk_passes": len(wl.phase), "n_edlib": int(len(wl.edlib[1])),
            "n_poa": wl.n_poa, "n_poa_reads": wl.n_poa_reads, "n_wfa": wl.n_poa,
            "l2": "flushed between timed steps (256 MiB write)", "seed": args.seed,
            "pipeline": ("K6 / K7 of batch k - 1 overlap K5 of batch k: own stream and window of the workspace pool (all batches are the same synthetic batch; the e2e run drains the last batch inside the timed region)" if getattr(args, "pipeline", False) else "none"),
            "streams": f"K5 (-> K7 -> K6 without the pipeline) on the library stream, K1 -> K1b -> K2 -> K2b -> K2c -> K3 -> K4 on the auxiliary stream (CTA slots of {getattr(args, 'reserve_sms', 0)} SMs left free by the persistent DP grids); the step is timed fork to join"}


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    import longcalld_b200 as lcd
    from longcalld_b200.capi import WFA_PARAMS_DTYPE, WFA_RESULT_DTYPE, POA_PARAMS_DTYPE, POA_RESULT_DTYPE
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL may print its version banner on stdout when it initialises; stdout carries exactly one JSON line, so the
        # process group is brought up (first collective included) with fd 1 pointed at stderr
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    lcd.init(local, (64 << 30) if args.pipeline else 0)
    if args.pipeline: lcd.split_pool(32 << 30)          # K5 in the lower 32 GiB of the workspace pool, K6 / K7 in the upper
    stream = torch.cuda.ExternalStream(lcd.stream(), device=local)
    aux_h = lcd.aux_stream()
    lcd.reserve_sms(args.reserve_sms)              # room for the pileup / phasing kernels next to the persistent DP grids
    lcd.reserve_plan_memory(12 << 30)              # the e2e pass creates plans from four host threads while the POA grid is resident: no pool growth under it
    aux = torch.cuda.ExternalStream(aux_h, device=local)
    shard_seed = args.seed + (rank if args.distinct_shards else 0)
    wl = Workload(args.mbp, args.tech, shard_seed)            # weak scaling: one shard per GPU (the same synthetic shard on every rank unless --distinct-shards: the step's length is its longest POA problem, so shards drawn with different seeds measure the draw, not the scaling)
    _pin = torch.from_numpy(wl.seqs).pin_memory(); wl.seqs = _pin.numpy()      # the loader's read buffer: page-locked, as the K1 inputs are
    L = lcd.lib()
    n = wl.n_poa
    ppar = np.zeros(n, dtype=POA_PARAMS_DTYPE); ppar[:] = lcd.poa_params()
    wpar = np.zeros(n, dtype=WFA_PARAMS_DTYPE); wpar[:] = lcd.wfa_params()
    buf, R, ref_off, ref_len, txt_off = wl.wfa_layout()
    cons = buf[R:]
    pres = np.zeros(n, dtype=POA_RESULT_DTYPE)
    wres = np.zeros(n, dtype=WFA_RESULT_DTYPE)

    from longcalld_b200.capi import _phase_structs, EDLIB_RESULT_DTYPE
    ph_ins, ph_outs, ph_keep, ph_res = _phase_structs(wl.phase, -9)
    eseqs, eqo, eql, eto, etl = wl.edlib
    ne = len(eql)
    emode = np.zeros(ne, np.int32); ewant = np.ones(ne, np.int32); eres = np.zeros(ne, dtype=EDLIB_RESULT_DTYPE)
    eoff = np.zeros(ne + 1, dtype=np.int64); np.cumsum(eql.astype(np.int64) + etl + 2, out=eoff[1:])
    ealn = np.zeros(int(eoff[-1]) + 1, dtype=np.uint8)

    # the consensus slots of two batches in flight: POA of batch k writes one while WFA of batch k - 1 reads the other
    bufs = [buf, buf.copy()]; press = [pres, pres.copy()]; wress = [wres, wres.copy()]
    dp_stream = torch.cuda.Stream(device=local)          # K6 / K7 (window 1 of the workspace pool) when they overlap the next batch's K5
    dp_h = dp_stream.cuda_stream
    stage_err = {}
    pile_gpu_done = threading.Event()

    L.lcd_poa_plan_create.restype = C.c_void_p

    stage_stream = torch.cuda.Stream(device=local)

    def poa_create():
        """host reads -> POA plan (packing + H2D), on a stream of its own: the library stream is busy with the batch before, and
        uploads queued behind its grid would wait for it (and hold up every other thread's copies of pageable memory meanwhile)"""
        lcd.set_thread_stream(stage_stream.cuda_stream)
        try:
            return _poa_create()
        finally:
            lcd.set_thread_stream(0)

    def _poa_create():
        h = L.lcd_poa_plan_create(C.c_int(n), _vp(wl.seqs), C.c_size_t(wl.seqs.size), _vp(wl.first), _vp(wl.n_reads),
                                  _vp(wl.read_off), _vp(wl.read_len), C.c_int(len(wl.read_len)), _vp(ppar))
        if not h:
            raise RuntimeError(L.lcd_gpu_last_error().decode())
        return C.c_void_p(h)

    def poa_stage(b, h, after_launch=None, during_launch=None):
        """POA plan -> launch -> consensus of every (region, haplotype) in host memory (slot b).  lcd_poa_batch spelled out in its
        C-ABI calls, so that the other engines' host threads can be released once the persistent grid is in flight (lcd_plan_run
        only enqueues) and the next batch's plan can be staged while this one's grid works."""
        tm = [time.perf_counter()]
        rc = L.lcd_plan_run(h, None)
        if after_launch: after_launch()
        nxt = during_launch() if during_launch else None
        tm.append(time.perf_counter())
        rc = rc or L.lcd_poa_plan_fetch(h, None, _vp(bufs[b][R:]), _vp(wl.cons_off), None, None, None, _vp(press[b]))
        tm.append(time.perf_counter())
        L.lcd_plan_destroy(h)
        stage_err["t_poa"] = [round(1e3 * (y - x), 1) for x, y in zip(tm, tm[1:])]
        if rc:
            raise RuntimeError(L.lcd_gpu_last_error().decode())
        return nxt

    def wfa_stage(b, own_stream=False):
        """K7 through its host-buffer batch call, then consensus in host memory (slot b) -> lcd_wfa_batch -> host CIGAR ops"""
        try:
            tw = [time.perf_counter()]
            if own_stream: lcd.set_thread_stream(dp_h)
            # K7 first: it depends on nothing the other threads produce, and its one-thread-per-problem grid is light enough to run beside the short kernels
            if L.lcd_edlib_batch(C.c_int(ne), _vp(eseqs), C.c_size_t(eseqs.size), _vp(eqo), _vp(eql), _vp(eto), _vp(etl), _vp(emode), _vp(ewant),
                                 _vp(ealn), _vp(eoff), _vp(eres)):
                raise RuntimeError(L.lcd_gpu_last_error().decode())
            tw.append(time.perf_counter())
            if own_stream: pile_gpu_done.wait()            # K6's persistent grid would keep the reserved CTA slots to itself: it follows the short kernels
            tw.append(time.perf_counter())
            tl = np.ascontiguousarray(press[b]["cons_len"])
            cap = 2 * (ref_len.astype(np.int64) + tl) + 8
            off = np.zeros(n + 1, dtype=np.int64); np.cumsum(cap, out=off[1:])
            ops = np.empty(int(off[-1]) + 1, dtype=np.uint8)
            rc = L.lcd_wfa_batch(C.c_int(n), _vp(bufs[b]), C.c_size_t(bufs[b].size), _vp(ref_off), _vp(ref_len), _vp(txt_off), _vp(tl),
                                 _vp(wpar), ops.ctypes.data_as(C.c_char_p), _vp(off), _vp(wress[b]))
            if rc:
                raise RuntimeError(L.lcd_gpu_last_error().decode())
            tw.append(time.perf_counter())
            stage_err["tl"] = tl; stage_err["t_wfa"] = [round(1e3 * (y - x), 1) for x, y in zip(tw, tw[1:])]
            stage_err["ops"] = (ops, off, b)
        except Exception as e:                                       # re-raised by the main thread
            stage_err["error"] = e

    def e2e_run(steps):
        """`steps` batches end to end through the host-buffer C-ABI calls, all copies inside.  Per batch: the pileup stages (host BAM
        fields of every chunk, 2.4 GB in; K1 -> K1b -> K2 -> K3, site lists, counters and profile rows out) and the phasing passes (K4)
        run from a second host thread on the auxiliary stream (their plans own their buffers); K5 runs on the main thread; K6 + K7 of
        the batch run from a third host thread on their own stream and pool window WHILE the next batch's K5 runs (two-stage software
        pipeline over the batches, --no-pipeline: in sequence); the last batch's K6 + K7 drain at the end, inside the timed region."""
        prev = None
        h_next = poa_create()
        for k in range(steps):
            b = k & 1
            pile_gpu_done.clear()
            pile_thread = threading.Thread(target=pileup_stage_thread); pile_thread.start()
            ph_thread = threading.Thread(target=phase_stage_thread); ph_thread.start()
            wt = threading.Thread(target=wfa_stage, args=(prev, True)) if (prev is not None and args.pipeline) else None
            t_it = time.perf_counter()
            # K6 / K7 of the batch before are released once K5 is in flight; the next batch's POA plan is staged meanwhile
            h_next = poa_stage(b, h_next, wt.start if wt is not None else None, poa_create if k + 1 < steps else None)
            t_poa_done = time.perf_counter()
            if not args.pipeline: wfa_stage(b)
            if wt is not None: wt.join()
            pile_thread.join(); ph_thread.join()
            if os.environ.get("LCD_BENCH_VERBOSE"):
                print(f"[e2e step {k}] {1e3 * (time.perf_counter() - t_it):.1f} ms; main: [launch + next plan, fetch] {stage_err.get('t_poa')}; "
                      f"K7/K6 thread [edlib batch, wait, wfa batch] {stage_err.get('t_wfa')}; pileup thread {pile_res.get('t')}; K4 batch {pile_res.get('t_phase')}", file=sys.stderr)
            for d in (pile_res, stage_err):
                if "error" in d: raise d.pop("error")
            if world > 1 and (prev is not None or not args.pipeline):
                gather_step(prev if args.pipeline else b)
            prev = b
        if args.pipeline:
            wfa_stage(prev)
            if "error" in stage_err: raise stage_err.pop("error")
            if world > 1: gather_step(prev)
        return bufs[prev], ref_off, ref_len, txt_off, stage_err["tl"]

    # N > 1: the one exchange of the path -- per region chunk (500 kb) the result records
    # go to rank 0, which would stitch and write the VCF; NCCL gather over NVLink, inside the e2e region
    from longcalld_b200 import shard
    reg_per_chunk = max(1, int(round(0.5 * wl.n_regions / max(args.mbp, 1e-9))))
    k_local = -(-wl.n_regions // reg_per_chunk)
    k_pad = -(-k_local // shard.CHUNK_BLOCK) * shard.CHUNK_BLOCK
    my_chunks = shard.deal_chunks(k_pad * world, world)[rank] if world > 1 else None
    chunk_of_problem = wl.region_of // reg_per_chunk
    chunk_first = np.searchsorted(chunk_of_problem, np.arange(k_pad + 1))
    gathered = {"bytes": 0}

    def gather_step(b):
        pres, wres = press[b], wress[b]
        blobs = []
        for j in range(k_pad):
            a, b = int(chunk_first[j]), int(chunk_first[j + 1])
            if a == b:
                blobs.append(b"")
                continue
            blobs.append(pres[a:b].tobytes() + wres[a:b].tobytes())
        out = shard.gather_chunk_results(my_chunks, blobs, k_pad * world, dst=0, device=torch.device("cuda", local))
        if rank == 0:
            gathered["bytes"] = sum(len(x) for x in out)

    from longcalld_b200 import synth as _synth
    gpu_sites = lambda bare, outs, regs: [_synth.site_list_from_sites(o, st) for o, st in zip(outs, lcd.sites_batch(bare, regs))]
    ps = PileupStage(args.mbp, args.tech, shard_seed, lcd.digar_batch, gpu_sites, lcd.pileup_batch, lcd.classify_batch, lambda nreg, cls: lcd.noisyreg_batch(nreg), pin=True)
    min_sv = [50] * ps.n_chunks
    pile_res = {}

    def pileup_e2e_step(dp, t_plan):
        """K1 plan (its H2D copies were issued by the staging thread) -> K1 -> K1b -> K2 / K3 on the lists in HBM -> site lists, coverage counters and profile rows on the host"""
        t = [time.perf_counter()]
        dp.run(); dp.sync(); t.append(time.perf_counter())
        if pile_res.get("want_digar"): pile_res["digar"] = dp.fetch()
        k1b = lcd.SitesPlan(None, ps.regs, min_sv_len=min_sv, digar_plan=dp); k1b.run(); pile_res["sites"] = k1b.fetch()
        k2 = lcd.PileupOnSitesPlan(dp, k1b); t.append(time.perf_counter()); k2.run(); pile_res["counts"] = k2.fetch(); t.append(time.perf_counter())
        k2b = lcd.ClassifyOnPileupPlan(k2, ps.cls, k2.n_sites); k2b.run(); pile_res["cate"] = k2b.fetch()      # K2b on K2's sites and counters in HBM (the reference windows come from the host)
        k2c = lcd.NoisyRegOnClassifyPlan(dp, k2b, ps.nreg, k2.n_sites, [s_[2] for s_ in dp.sizes()]); k2c.run(); pile_res["nreg"] = k2c.fetch(); k2c.destroy(); k2b.destroy()      # K2c on K1's lists and K2b's categories in HBM (options + low-complexity intervals come from the host)
        k3 = lcd.ProfileOnDigarPlan(dp, ps.var_sites, ps.n_reads); t.append(time.perf_counter()); k3.run(); pile_res["prof"] = k3.fetch(); t.append(time.perf_counter())
        for x in (k3, k2, k1b, dp): x.destroy()
        t.append(time.perf_counter())
        # ms (second host thread, overlapped with the POA launch): K1 plan incl. H2D, K1 run, K1b plan+run+fetch and K2 plan, K2 run+fetch, K2b + K2c batches and K3 plan, K3 run+fetch, destroy
        pile_res["t"] = [round(1e3 * t_plan, 2)] + [round(1e3 * (b - a), 2) for a, b in zip(t, t[1:])]

    def pileup_stage_thread():
        try:
            lcd.set_thread_stream(aux_h)                             # everything this thread does goes to the auxiliary stream
            t0 = time.perf_counter()
            dp = lcd.DigarPlan(ps.chunks)
            pileup_e2e_step(dp, time.perf_counter() - t0)
            pile_gpu_done.set()
        except Exception as e:                                       # re-raised by the main thread
            pile_res["error"] = e; pile_gpu_done.set()

    ph_stream = torch.cuda.Stream(device=local)

    def phase_stage_thread():
        """K4 through its host-buffer batch call, on its own stream: it depends on nothing the other threads of the batch produce"""
        try:
            lcd.set_thread_stream(ph_stream.cuda_stream)
            t1 = time.perf_counter()
            if L.lcd_phase_batch(C.c_int(len(wl.phase)), ph_ins, ph_outs):
                raise RuntimeError(L.lcd_gpu_last_error().decode())
            pile_res["t_phase"] = round(1e3 * (time.perf_counter() - t1), 2)
        except Exception as e:
            pile_res["error"] = e

    wseqs, po, pl, to, tl = e2e_run(1)                              # also yields the consensus sequences for the WFA plan
    digar_plan = lcd.DigarPlan(ps.chunks); digar_plan.run(); digar_plan.sync()
    sites_plan = lcd.SitesPlan(None, ps.regs, min_sv_len=min_sv, digar_plan=digar_plan); sites_plan.run(); sites_plan.sync()
    assert sum(sites_plan.sizes()) == ps.n_raw_sites
    k2_plan = lcd.PileupOnSitesPlan(digar_plan, sites_plan)
    k3_plan = lcd.ProfileOnDigarPlan(digar_plan, ps.var_sites, ps.n_reads)
    k2b_plan = lcd.ClassifyOnPileupPlan(k2_plan, ps.cls, k2_plan.n_sites)
    k2c_plan = lcd.NoisyRegOnClassifyPlan(digar_plan, k2b_plan, ps.nreg, k2_plan.n_sites, [s_[2] for s_ in digar_plan.sizes()])
    poa_plan = lcd.PoaPlan(wl.seqs, wl.first, wl.n_reads, wl.read_off, wl.read_len, lcd.poa_params())
    wfa_plan = lcd.WfaPlan(wseqs, po, pl, to, tl, lcd.wfa_params())
    phase_plan = lcd.PhasePlan(wl.phase)
    edlib_plan = lcd.EdlibPlan(eseqs, eqo, eql, eto, etl, lcd.MODE_NW, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(overlap):
        """One pass over the batch, inputs resident, bracketed by ev[9] .. ev[10] on the library stream.
        overlap=False: every stage on the library stream, one after the other -- the per-kernel times of the roofline come from here.
        overlap=True (what `value` is): the pool-free stages (K1 -> K1b -> K2 -> K3, K4) on the auxiliary stream, K5 on the library
        stream and, under the pipeline, K7 -> K6 -- which then stand for the batch before, whose consensus is resident -- on a third
        stream with their own window of the workspace pool; the library stream waits for the other two before ev[10]."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(15)]
        with torch.cuda.stream(stream):
            flush.zero_()
            ev[9].record(stream)
        sa, sa_h = (aux, aux_h) if overlap else (stream, None)
        sd, sd_h = (dp_stream, dp_h) if (overlap and args.pipeline) else (stream, None)

        def dp_tail():
            ev[1].record(sd); edlib_plan.run(sd_h); ev[4].record(sd); wfa_plan.run(sd_h); ev[11].record(sd)      # K7 (light, independent) first: the stream's tail is K6 alone

        def pile_chain():
            ev[5].record(sa); digar_plan.run(sa_h)
            ev[8].record(sa); sites_plan.run(sa_h)
            ev[6].record(sa); k2_plan.run(sa_h)
            ev[13].record(sa); k2b_plan.run(sa_h)
            ev[14].record(sa); k2c_plan.run(sa_h)
            ev[7].record(sa); k3_plan.run(sa_h)
            ev[2].record(sa); phase_plan.run(sa_h)
            ev[3].record(sa)
        if not overlap:
            pile_chain()
        ev[0].record(stream)
        poa_plan.run()                      # enqueues the persistent grid (the statuses are looked at by sync() below)
        ev[12].record(stream)
        if overlap:                         # the other engines are issued once K5 is in flight: they fill the reserved CTA slots and K5's tail
            aux.wait_event(ev[9])
            with torch.cuda.stream(aux): torch.cuda._sleep(600_000)                # ~0.3 ms: K5's grid is resident before the others ask for SMs
            pile_chain()
            if args.pipeline:               # K6's persistent grid would keep the reserved slots to itself: it follows the short kernels
                dp_stream.wait_event(ev[3])
                dp_tail()
        if not (overlap and args.pipeline): dp_tail()
        poa_plan.sync()                     # K5's statuses (a rescue launch, had a problem outgrown its workspace, would run here)
        if overlap:
            stream.wait_event(ev[11]); stream.wait_event(ev[3])
        ev[10].record(stream)
        return ev

    for _ in range(args.warmup):
        device_step(False); device_step(True)
    barrier()
    seq = [device_step(False) for _ in range(args.steps)]          # per-kernel times, one stage at a time
    barrier()
    launches0 = lcd.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [device_step(True) for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()
    launches = lcd.launch_count() - launches0
    poa_ms = sum(e[0].elapsed_time(e[12]) for e in seq)
    wfa_ms = sum(e[4].elapsed_time(e[11]) for e in seq)
    phase_ms = sum(e[2].elapsed_time(e[3]) for e in seq)
    edlib_ms = sum(e[1].elapsed_time(e[4]) for e in seq)
    k1_ms = sum(e[5].elapsed_time(e[8]) for e in seq); k1b_ms = sum(e[8].elapsed_time(e[6]) for e in seq); k2_ms = sum(e[6].elapsed_time(e[13]) for e in seq); k2b_ms = sum(e[13].elapsed_time(e[14]) for e in seq); k2c_ms = sum(e[14].elapsed_time(e[7]) for e in seq); k3_ms = sum(e[7].elapsed_time(e[2]) for e in seq)
    seq_ms = sum(e[9].elapsed_time(e[10]) for e in seq)
    dev_ms = sum(e[9].elapsed_time(e[10]) for e in evs)            # whole step: all streams, fork at ev[9], join at ev[10]
    poa_ovl_ms = sum(e[0].elapsed_time(e[12]) for e in evs)
    phase_pairs = phase_plan.work_units()
    edlib_units = edlib_plan.work_units()
    poa_cells = poa_plan.work_units()
    wfa_cells = wfa_plan.work_units()
    r1 = poa_plan.fetch(want_msa=False)[0]
    r2 = wfa_plan.fetch(want_ops=False)[0]
    assert (r1["status"] == 0).all() and (r2["status"] == 0).all()

    e2e_run(min(args.warmup, 1) + (1 if args.pipeline else 0))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    phase_bytes = sum(a.nbytes for k in ph_keep for a in k.values())
    pile_d2h = int(sum(sum(a.nbytes for a in st.values() if hasattr(a, "nbytes")) for st in pile_res["sites"]) + sum(c.nbytes for c in pile_res["counts"]) + sum(sum(a.nbytes for a in o.values()) for o in pile_res["prof"]))
    cls_h2d = int(sum(d["ref_seq"].nbytes for d in ps.cls)) + ps.k2c_h2d; pile_d2h += int(sum(c.nbytes for c in pile_res["cate"]))
    pile_d2h += int(sum(g["var_cate"].nbytes + g["keep"].nbytes + 20 * g["n_regs"] for g in pile_res["nreg"]))
    h2d = int(ps.h2d + cls_h2d + phase_bytes + eseqs.size + 24 * ne + wl.seqs.size + 12 * len(wl.read_len) + 64 * n + (pl.astype(np.int64) + tl + 56).sum() + 96 * n)
    d2h = int(pile_d2h + sum(a.nbytes for r in ph_res for a in r.values()) + eres.nbytes + int(eres["aln_len"].sum()) + pres["cons_len"].sum() + 32 * n + wres.nbytes + 2 * (pl.astype(np.int64) + tl + 4).sum())

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_rank_ms = [[round(float(x[0]) / args.steps, 3), round(float(x[1]) / args.steps, 3)] for x in per_rank]       # [device, e2e] ms per step of every rank
    dev_ms_max, e2e_ms_max = t.tolist()
    total_mbp = args.mbp * world
    value = total_mbp * args.steps / (dev_ms_max / 1e3)
    e2e_value = total_mbp * args.steps / (e2e_ms_max / 1e3)

    cpu_baseline = None; parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lib = ref_shim()
        if lib is not None:
            nt = os.cpu_count() or 1
            frac = calibrate_sample(lib, wl, nt, target_s=15.0)
            regs, idx = region_sample(wl, frac)
            ref_out, pile_out = {}, {}
            dt, dt_poa, dt_wfa, dt_phase, dt_edlib = reference_step(lib, wl, idx, nt, keep=ref_out)
            mbp_sample = args.mbp * len(regs) / wl.n_regions
            k_chunks = min(ps.n_chunks, max(1, int(round(ps.n_chunks * len(regs) / wl.n_regions))))
            pile_scale = (len(regs) / wl.n_regions) / (k_chunks / ps.n_chunks)
            dtp = reference_pileup_step(lib, ps, k_chunks, nt, out=pile_out)
            # parity at bench scale: one more e2e batch whose difference lists are fetched too, compared with the reference arm's outputs
            pile_res["want_digar"] = True
            e2e_run(2 if args.pipeline else 1)
            ops_g, off_g, b_g = stage_err["ops"]
            parity = parity_check(ref_out, pile_out, wl, ps, dict(cons=bufs[b_g][R:], cons_len=press[b_g]["cons_len"], wfa_res=wress[b_g], wfa_ops=ops_g, wfa_off=off_g,
                                                                  phase=ph_res, edlib_res=eres, edlib_aln=ealn, edlib_off=eoff,
                                                                  sites=pile_res["sites"], counts=pile_res["counts"], cate=pile_res["cate"], prof=pile_res["prof"], digar=pile_res["digar"], nreg=pile_res["nreg"]))
            cpu_baseline = {"value": mbp_sample / (dt + pile_scale * dtp[0]), "unit": UNIT, "cores": nt, "kind": "reference",
                            "sample": f"{len(regs)} of {wl.n_regions} regions ({mbp_sample:.3f} Mb): the unmodified reference's own functions via oracle/_ref; "
                                      f"K1-K3 on {k_chunks} of {ps.n_chunks} chunks, time scaled by {pile_scale:.4f}",
                            "poa_s": dt_poa, "wfa_s": dt_wfa, "phase_s": dt_phase, "edlib_s": dt_edlib,
                            "digar_s": pile_scale * dtp[1], "pileup_s": pile_scale * dtp[2], "profile_s": pile_scale * dtp[3], "sites_s": pile_scale * dtp[4], "classify_s": pile_scale * dtp[5], "noisyreg_s": pile_scale * dtp[6]}
    widening = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.synchronize()
        widening = {"sdust_kernels": sdust_widening(args, ps, ref_shim())}
    wp_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_whole_program:
        torch.cuda.synchronize()
        wp_line = whole_program(args, with_gpu=True)
    if rank == 0:
        peak, which = load_peaks()
        poa_gbs = poa_cells * POA_BYTES_PER_CELL / (poa_ms / args.steps / 1e3) / 1e9
        wfa_gbs = wfa_cells * WFA_BYTES_PER_CELL / (wfa_ms / args.steps / 1e3) / 1e9
        dominant_is_poa = poa_ms >= wfa_ms
        achieved = poa_gbs if dominant_is_poa else wfa_gbs
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "ms_per_step_one_stream": seq_ms / args.steps, "poa_ms_overlapped": poa_ovl_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16/int32", "data": "synthetic", "config": workload_config(args, wl, ps),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "e2e_pileup_ms": pile_res.get("t"), "e2e_phase_ms": pile_res.get("t_phase"),
                "ms_per_step_per_rank": per_rank_ms, "shards": ("seeded by rank" if args.distinct_shards else "the same synthetic shard on every rank"),
                "gpu_launches": int(launches), "clocks": clocks,
                "gather": (None if world == 1 else {"what": "per-chunk POA/WFA result records to rank 0 (NCCL gather, inside e2e)",
                                                    "bytes_per_step": int(gathered["bytes"])}),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic() if dominant_is_poa else None,
                             "traffic_source": f"profiles/{ncu_capture()} (ncu --set full of one poa_kernel launch over this step's POA batch: tools/poa_prof.py 50)",
                             "kernel": "poa_kernel" if dominant_is_poa else "wfa_kernel<32>+wfa_kernel<256>",
                             "algorithmic": (f"{poa_cells} banded POA cells x {POA_BYTES_PER_CELL} B" if dominant_is_poa
                                             else f"{wfa_cells} wavefront cells x {WFA_BYTES_PER_CELL} B"),
                             "peak_source": which,
                             "per_kernel": {"poa_kernel": {"ms": poa_ms / args.steps, "cells": poa_cells, "GBps": poa_gbs, "frac": poa_gbs / peak},
                                            "wfa_kernels": {"ms": wfa_ms / args.steps, "cells": wfa_cells, "GBps": wfa_gbs, "frac": wfa_gbs / peak},
                                            "phase_kernel": {"ms": phase_ms / args.steps, "read_var_pairs": phase_pairs,
                                                             "GBps": phase_pairs * PHASE_BYTES_PER_PAIR / (phase_ms / args.steps / 1e3) / 1e9,
                                                             "note": "one pass of the pair list; the kernel makes 2-11 passes (seed + iterations), latency-bound"},
                                            "digar_kernels": {"ms": k1_ms / args.steps, "read_bases": ps.read_bases, "records": ps.records,
                                                              "GBps": ps.k1_bytes / (k1_ms / args.steps / 1e3) / 1e9, "frac": ps.k1_bytes / (k1_ms / args.steps / 1e3) / 1e9 / peak,
                                                              "note": "count + scan + fill + histogram; 1.5 B per read base + 4 B per CIGAR op in, 32 B per record out"},
                                            "sites_kernels": {"ms": k1b_ms / args.steps, "records": ps.records, "sites": ps.n_raw_sites,
                                                              "GBps": ps.k1b_bytes / (k1b_ms / args.steps / 1e3) / 1e9,
                                                              "note": "count + scan + scatter + group + scan + emit; 14 B per record in, 36 B per site out"},
                                            "pileup_kernel": {"ms": k2_ms / args.steps, "records": ps.records, "sites": ps.n_raw_sites,
                                                              "GBps": ps.k2_bytes / (k2_ms / args.steps / 1e3) / 1e9},
                                            "classify_kernel": {"ms": k2b_ms / args.steps, "sites": ps.n_raw_sites, "GBps": ps.k2b_bytes / (k2b_ms / args.steps / 1e3) / 1e9},
                                            "noisyreg_kernel": {"ms": k2c_ms / args.steps, "sites": ps.n_raw_sites, "noisy_intervals": ps.n_cnreg, "low_complexity_intervals": ps.n_low,
                                                                "GBps": ps.k2c_bytes / (k2c_ms / args.steps / 1e3) / 1e9,
                                                                "note": "one CTA per chunk; interval merges and the flank sweep run on one thread of it (order-dependent in the reference): latency-bound by design"},
                                            "profile_kernel": {"ms": k3_ms / args.steps, "records": ps.records, "variants": ps.n_vars,
                                                               "GBps": ps.k3_bytes / (k3_ms / args.steps / 1e3) / 1e9},
                                            "edlib_kernel": {"ms": edlib_ms / args.steps, "block_columns": edlib_units,
                                                             "GBps": edlib_units * EDLIB_BYTES_PER_BLOCKCOL / (edlib_ms / args.steps / 1e3) / 1e9}}},
                "cpu_baseline": cpu_baseline, "parity_checked": parity, "whole_program": wp_line, "widening": widening}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sdust_widening(args, ps, lib):
    """K0, outside the timed step: the low-complexity intervals of the chunks' reference windows (sdust, T = 5, W = 20) -- one window of the
    chunk's length per chunk, ten distinct ones tiled like the chunks themselves -- timed with CUDA events on the library's stream, and compared interval
    for interval with the unmodified reference's sdust() (oracle/_ref) on the distinct windows."""
    import torch
    import longcalld_b200 as lcd
    from longcalld_b200 import synth
    n_chunks = ps.n_chunks; L = int(round(args.mbp * 1e6 / n_chunks))
    rng = np.random.default_rng(args.seed + 4242)
    tmpl = [synth.sdust_window(rng, L) for _ in range(min(10, n_chunks))]
    plan = lcd.SdustPlan([tmpl[i % len(tmpl)] for i in range(n_chunks)], 5, 20)
    st = torch.cuda.ExternalStream(lcd.stream())
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); plan.run(); e1.record(st); plan.sync()
        ms.append(e0.elapsed_time(e1))
    out = plan.fetch()
    res = {"ms": float(np.median(ms[1:])), "windows": n_chunks, "window_bases": L, "intervals": int(sum(len(x) for x in out)), "launches_per_run": 6,
           "Mbp_per_s": n_chunks * L / 1e6 / (float(np.median(ms[1:])) / 1e3), "in_timed_step": False, "parity_checked": None, "reference_s_per_window": None}
    if lib is not None and hasattr(lib, "ref_sdust"):
        lib.ref_sdust.restype = C.c_int
        dt = 0.0
        for k, seq in enumerate(tmpl):
            cap = len(seq) // 2 + 16
            b, e = np.zeros(cap, np.int64), np.zeros(cap, np.int64)
            t0 = time.perf_counter()
            n = lib.ref_sdust(_vp(seq), C.c_int(len(seq)), C.c_int(5), C.c_int(20), _vp(b), _vp(e), C.c_int64(cap))
            dt += time.perf_counter() - t0
            got = np.asarray(out[k], np.int64).reshape(-1, 2)
            if n != len(got) or not (np.array_equal(got[:, 0], b[:n]) and np.array_equal(got[:, 1], e[:n])):
                raise SystemExit(f"K0: window {k} differs from the reference's sdust() ({n} vs {len(got)} intervals)")
        res["parity_checked"] = f"{len(tmpl)} distinct windows ({sum(len(x) for x in out[:len(tmpl)])} intervals) identical to the reference's sdust()"
        res["reference_s_per_window"] = dt / len(tmpl)
        # the same batch on all host cores: the distinct windows, a thread each, as often as the batch tiles them
        from concurrent.futures import ThreadPoolExecutor
        nt = os.cpu_count() or 1; reps = [tmpl[i % len(tmpl)] for i in range(min(n_chunks, 2 * nt))]

        def one(seq):
            cap = len(seq) // 2 + 16; b, e = np.zeros(cap, np.int64), np.zeros(cap, np.int64)
            return lib.ref_sdust(_vp(seq), C.c_int(len(seq)), C.c_int(5), C.c_int(20), _vp(b), _vp(e), C.c_int64(cap))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nt) as ex: list(ex.map(one, reps))
        res["reference_all_cores"] = {"cores": nt, "windows": len(reps), "s": time.perf_counter() - t0, "Mbp_per_s": len(reps) * L / 1e6 / (time.perf_counter() - t0)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mbp", type=float, default=50.0, help="reference megabases per GPU per step")
    ap.add_argument("--tech", default="hifi", choices=["hifi", "ont"])
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--distinct-shards", action="store_true", help="seed every rank's shard differently (default: the same synthetic shard on every rank)")
    ap.add_argument("--no-whole-program", action="store_true", help="skip the BAM -> VCF run of `longcallD call` (reference binary vs GPU drop-in)")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false", help="K6 / K7 after K5 on one stream instead of overlapping the next batch's K5")
    ap.add_argument("--reserve-sms", type=int, default=12, help="SMs whose CTA slots the persistent DP grids leave to the concurrently running pileup / phasing kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
