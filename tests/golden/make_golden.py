#!/usr/bin/env python3
"""Regenerates tests/golden/*.json.gz.  Needs /root/reference (this container only).

  wfa_utest.json.gz   WFA2-lib's own regression vectors (WFA2-lib/tests/wfa.utest.seq + the
                      match==0 goldens in tests/wfa.utest.check/: affine, affine2p, p0-p2,
                      wfapt0/1) -- pins the recurrence, the backtrace tie-breaks and wf-adaptive.
  digar_lcd.json.gz   digests of the UNMODIFIED reference =/X difference-list pass (collect_digar_from_eqx_cigar) on seeded chunks.
  sites_lcd.json.gz   candidate-site lists of the UNMODIFIED reference (collect_all_cand_var_sites) on seeded chunks.
  classify_lcd.json.gz per-site categories of the UNMODIFIED reference (classify_var_cate) on seeded sites / reference windows.
  pileup_lcd.json.gz  outputs of the UNMODIFIED reference per-site coverage pass (collect_cand_vars) on seeded chunks.
  phase_lcd.json.gz   outputs of the UNMODIFIED reference read->haplotype assignment / phasing on seeded chunks.
  poa_ncons_lcd.json.gz outputs of the UNMODIFIED abpoa_aln_msa_cons (two consensus sequences by read clustering) on seeded de-novo regions.
  noisyreg_lcd.json.gz  kept sites + chunk_noisy_regs of the UNMODIFIED pre_process_noisy_regs + classify_cand_vars on seeded chunks.
  sdust_lcd.json.gz   low-complexity intervals of the UNMODIFIED sdust() (src/sdust.c:184) on seeded reference windows (T = 5, W = 20 and two more settings).
  edlib_lcd.json.gz   outputs of the UNMODIFIED reference edlib (NW / HW, path) on seeded inputs.
  wfa_lcd.json.gz     outputs of the UNMODIFIED reference WFA2-lib (oracle/_ref/libref_shim.so)
                      at longcallD's own parameter points (src/align.h:21-26, src/align.c:398-406)
                      on seeded inputs: end2end 2p/no-heuristic, 1p/wf-adaptive, 2p/z-drop.
"""
import gzip
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import lcd_testlib as T  # noqa: E402

REF = "/root/reference"


def rle(ops: bytes) -> str:
    out, i = [], 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append(f"{j - i}{chr(ops[i])}")
        i = j
    return "".join(out)


def wfa_utest():
    seqs = open(f"{REF}/WFA2-lib/tests/wfa.utest.seq").read().split("\n")
    pairs = [(seqs[i][1:], seqs[i + 1][1:]) for i in range(0, len(seqs) - 1, 2) if seqs[i].startswith(">")]
    sets = {  # name -> (x, o1, e1, o2, e2, affine2p, heuristic, min_wf, max_dist, zdrop, steps)
        "affine": (4, 6, 2, 0, 0, 0, 0, 0, 0, 0, 1),
        "affine2p": (4, 6, 2, 24, 1, 1, 0, 0, 0, 0, 1),
        "affine.p0": (1, 2, 1, 0, 0, 0, 0, 0, 0, 0, 1),
        "affine.p1": (3, 1, 4, 0, 0, 0, 0, 0, 0, 0, 1),
        "affine.p2": (5, 3, 2, 0, 0, 0, 0, 0, 0, 0, 1),
        "affine.wfapt0": (4, 6, 2, 0, 0, 0, 1, 10, 50, 0, 1),
        "affine.wfapt1": (4, 6, 2, 0, 0, 0, 1, 10, 50, 0, 10),
    }
    golden = {}
    for name in sets:
        rows = open(f"{REF}/WFA2-lib/tests/wfa.utest.check/test.{name}.alg").read().strip().split("\n")
        assert len(rows) == len(pairs), (name, len(rows), len(pairs))
        golden[name] = [(int(r.split("\t")[0]), r.split("\t")[1]) for r in rows]
    return {"pairs": pairs, "params": sets, "golden": golden}


def wfa_lcd():
    ref = T.ref_lib()
    assert ref is not None, "reference shim not built: make -C oracle ref"
    rng = np.random.default_rng(20261017)
    cases = []

    def add(p, t, par):
        st, score, ops, ev, eh = T.wfa_align(ref, "ref_wfa_align", p, t, T.WfaParams(*par))
        cases.append({"p": "".join(map(str, p.tolist())), "t": "".join(map(str, t.tolist())), "par": list(par),
                      "status": st, "score": score, "ops": rle(ops), "end_v": ev, "end_h": eh})

    for n in (0, 1, 2, 5, 17, 33, 64, 65, 127, 200, 333, 512, 800, 1500):
        for rep in range(6):
            a = rng.integers(0, 4, n).astype(np.uint8)
            b = T.mutate(rng, a, sub=0.02, ins=0.01, dele=0.01, max_indel=3) if n else a.copy()
            if rep == 4 and n > 64:   # one SV
                b = T.mutate(rng, a, sub=0.005, ins=0.002, dele=0.002, sv=(n // 3, "ins" if n % 2 else "del", n // 4))
            if rep == 5:
                b = rng.integers(0, 4, max(0, n + int(rng.integers(-3, 4)))).astype(np.uint8)   # unrelated
            add(a, b, T.wfa_params_tuple(T.HEUR_NONE, 1))
            add(a, b, T.wfa_params_tuple(T.HEUR_ADAPTIVE, 0))
            add(a, b, T.wfa_params_tuple(T.HEUR_ZDROP, 1, len(a), len(b)))
    # partial-read shaped z-drop cases: text is a prefix of the pattern plus junk (src/align.c:667-707)
    for n in (300, 900, 2500):
        for rep in range(4):
            a = rng.integers(0, 4, n).astype(np.uint8)
            cut = int(n * (0.3 + 0.15 * rep))
            b = np.concatenate([T.mutate(rng, a[:cut], sub=0.01, ins=0.01, dele=0.01), rng.integers(0, 4, n - cut).astype(np.uint8)])
            add(a, b, T.wfa_params_tuple(T.HEUR_ZDROP, 1, len(a), len(b)))
            add(b, a, T.wfa_params_tuple(T.HEUR_ZDROP, 1, len(b), len(a)))
    # low-complexity (homopolymer / STR) pairs: tie-break heavy
    for unit in ("0", "01", "012", "0011"):
        for n in (40, 150, 600):
            a = np.array([int(c) for c in (unit * n)[:n]], dtype=np.uint8)
            for d in (-7, -1, 1, 2, 30):
                m = max(0, n + d)
                b = np.array([int(c) for c in (unit * (m + 1))[:m]], dtype=np.uint8)
                add(a, b, T.wfa_params_tuple(T.HEUR_NONE, 1))
                add(a, b, T.wfa_params_tuple(T.HEUR_ADAPTIVE, 0))
    return {"cases": cases}


def poa_lcd():
    """Outputs of the UNMODIFIED abPOA (AVX-512BW build of oracle/_ref) driven as longcallD drives it
    (src/align.c:762-870 phased / :872-953 de-novo set-up, max_n_cons = 1) on seeded noisy regions."""
    import hashlib
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    ref = T.ref_lib()
    cases = []
    for tech, mbp, seed in (("hifi", 0.25, 21), ("ont", 0.03, 22)):
        for r in synth.make_regions(mbp, tech, seed=seed):
            for hap in (1, 2):
                seqs = [s for s, h in zip(r.reads, r.read_hap) if h == hap]
                if not seqs or min(len(s) for s in seqs) == 0 or max(len(s) for s in seqs) > 1200:
                    continue
                for sub, wb in ((1, 10), (0, -1)):
                    if wb < 0 and len(cases) % 3:
                        continue
                    rc, cons, msa = T.poa(ref, "ref_poa", seqs, T.poa_params(sub, wb))
                    assert rc == 0
                    cases.append({"seqs": ["".join(map(str, s.tolist())) for s in seqs], "sub_aln": sub, "wb": wb,
                                  "cons": "".join(map(str, cons)), "msa_shape": list(msa.shape),
                                  "msa_sha1": hashlib.sha1(msa.tobytes()).hexdigest()})
    return {"cases": cases}


def edlib_lcd():
    """Outputs of the UNMODIFIED edlib (oracle/_ref/libref_shim.so) as longcallD calls it (src/align.c:210-275:
    k = -1, NW or HW, TASK_PATH) on seeded read-vs-consensus / infix shaped inputs, incl. one Hirschberg case."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_oracle_edlib import edlib_cases
    ref = T.ref_lib()
    rng = np.random.default_rng(20261018)
    cases = []
    for q, t, mode in edlib_cases(rng, 330) + edlib_cases(rng, 2, big=True)[:1]:
        st, ed, s, e, aln = T.edlib_align(ref, "ref_edlib_align", q, t, mode, 1)
        assert st == 0
        cases.append({"q": "".join(map(str, q.tolist())), "t": "".join(map(str, t.tolist())), "mode": mode,
                      "ed": ed, "start": s, "end": e, "aln": "".join(map(str, aln))})
    return {"cases": cases}


def phase_lcd():
    """Outputs of the UNMODIFIED assign_hap_based_on_germline_het_vars_kmeans (src/assign_hap.c:473, via
    oracle/_ref/libref_shim.so: ref_assign_hap) on seeded synthetic chunks (read x variant allele profiles)."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_oracle_phase import phase_cases, CMP
    ref = T.ref_lib()
    cases = []
    for d, target, is_ont in phase_cases(20261019, 120):
        if d["n_reads"] > 200 or d["n_vars"] > 150:
            continue
        out = T.phase(ref, "ref_assign_hap", d, target, is_ont)
        cases.append({"in": {k: (np.asarray(v).reshape(-1).tolist() if hasattr(v, "tolist") else v) for k, v in d.items()},
                      "target": target, "is_ont": is_ont, "out": {k: out[k].tolist() for k in CMP}})
    return {"cases": cases}


def pileup_lcd():
    """Outputs of the UNMODIFIED collect_cand_vars (src/collect_var.c:238, via oracle/_ref/libref_shim.so) on seeded chunks."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    ref = T.ref_lib()
    rng = np.random.default_rng(20261020)
    cases = []
    for it in range(24):
        d = synth.make_pileup_chunk(rng, ref_len=int(rng.choice([600, 3000])), n_reads=int(rng.choice([1, 8, 40])),
                                    read_len=(200, 500) if it % 3 == 0 else (800, 2500), var_every=int(rng.choice([40, 150])), err_every=int(rng.choice([60, 800])))
        counts = T.pileup(ref, "ref_collect_cand_vars", d)
        cases.append({"in": {k: (np.asarray(v).reshape(-1).tolist() if hasattr(v, "tolist") else v) for k, v in d.items()}, "counts": counts.tolist()})
    return {"cases": cases}


def digar_lcd():
    """Per-read records (sha1 digest, T.digar_digest) + chunk noisy list of the UNMODIFIED collect_digar_from_eqx_cigar
    (src/bam_utils.c:701, via oracle/_ref/libref_shim.so: ref_collect_digar_eqx) on seeded chunks."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    ref = T.ref_lib()
    rng = np.random.default_rng(20261021)
    cases = []
    for it in range(16):
        d = synth.make_digar_chunk(rng, n_reads=int(rng.choice([1, 5, 14])), read_len=(60, 300) if it % 4 == 0 else (400, 1500),
                                   err_every=int(rng.choice([15, 80, 300])), tech="ont" if it % 3 == 0 else "hifi", low_qual_frac=float(rng.choice([0.0, 0.05, 0.4])))
        res = T.collect_digar(ref, "ref_collect_digar_eqx", d)
        cases.append({"in": T.digar_case_to_json(d), "digest": T.digar_digest(res), "chunk_noisy": res["chunk_noisy"],
                      "n_skip": sum(v[0] for v in res["reads"].values()), "n_intervals": sum(len(v[4]) for v in res["reads"].values())})
    return {"cases": cases}


def sites_lcd():
    """Sorted unique candidate sites of the UNMODIFIED collect_all_cand_var_sites (src/collect_var.c:1209, via oracle/_ref/libref_shim.so)
    on seeded chunks: [pos, type, ref_len, alt_len, alt bytes (hex)] per site."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    ref = T.ref_lib()
    rng = np.random.default_rng(20261022)
    cases = []
    for it in range(20):
        d = synth.make_pileup_chunk(rng, ref_len=int(rng.choice([600, 3000])), n_reads=int(rng.choice([1, 8, 40])),
                                    read_len=(200, 500) if it % 3 == 0 else (800, 2500), var_every=int(rng.choice([40, 150])), err_every=int(rng.choice([60, 800])),
                                    min_sv_len=int(rng.choice([30, 50])))
        lo, hi = int(d["read_beg"].min()), int(d["read_end"].max())
        reg = [-1, -1] if it % 4 == 0 else [lo + (hi - lo) // 6, hi - (hi - lo) // 6]
        sites = T.collect_sites(ref, "ref_collect_sites", d, reg[0], reg[1], src_is_offset=True)
        keys = [k for k, _ in T.PILEUP_IN_FIELDS if not k.startswith("site_")]
        cases.append({"in": {**{k: np.asarray(d[k]).reshape(-1).tolist() for k in keys}, "n_reads": d["n_reads"], "min_bq": d["min_bq"], "min_sv_len": d["min_sv_len"]},
                      "reg": reg, "sites": [[p_, t_, r_, a_, alt.hex()] for p_, t_, r_, a_, alt in sites]})
    return {"cases": cases}


def classify_lcd():
    """Per-site categories of the UNMODIFIED classify_var_cate (src/collect_var.c:413, via oracle/_ref/libref_shim.so) on seeded sites,
    counters and reference windows with planted homopolymers / tandem repeats."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    ref = T.ref_lib()
    rng = np.random.default_rng(20261023)
    cases = []
    for it in range(16):
        d = synth.make_classify_chunk(rng, ref_len=int(rng.choice([600, 2000])), n_sites=int(rng.choice([1, 30, 120])), max_xgaps=int(rng.choice([5, 3])),
                                      ref0=int(rng.choice([1, 100000])))
        cases.append({"in": T.classify_case_to_json(d), "cate": T.classify(ref, "ref_classify_sites", d).tolist()})
    return {"cases": cases}


def poa_ncons_lcd():
    """Outputs of the UNMODIFIED abpoa_aln_msa_cons (src/align.c:872-953: wb = -1, max_n_cons = 2, min_freq = 0.20, via oracle/_ref/libref_shim.so) on the
    mixed-haplotype reads of seeded noisy regions: number of clusters, the consensus sequences, every read's cluster, sha1 of the MSA."""
    import hashlib
    par = T.poa_params(0, -1); par.max_n_cons = 2
    ref = T.ref_lib()
    cases = []
    for tech, mbp, seed in (("hifi", 0.25, 23), ("ont", 0.06, 24)):
        for seqs in T.denovo_problems(mbp, tech, seed, max_len=500):
            rc, cons, clu, msa = T.poa_ncons(ref, "ref_poa_ncons", seqs, par, 0.20)
            assert rc == 0
            cases.append({"seqs": ["".join(map(str, np.asarray(s).tolist())) for s in seqs], "cons": ["".join(map(str, c)) for c in cons], "clu": "".join(map(str, clu.tolist())),
                          "msa_shape": list(msa.shape), "msa_sha1": hashlib.sha1(msa.tobytes()).hexdigest()})
    return {"cases": cases}


def noisyreg_lcd():
    """Outputs of the UNMODIFIED pre_process_noisy_regs + classify_cand_vars (src/collect_var.c:557,902, via oracle/_ref/libref_shim.so) on seeded chunks:
    the kept sites with their categories and chunk_noisy_regs.  The inputs are K1 / K1b / K2 / K2b results of the same chunks (the oracle's, pinned separately)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from longcalld_b200 import synth
    orc, ref = T.oracle_lib(), T.ref_lib()
    rng = np.random.default_rng(20261101)
    cases = []
    for it in range(8):
        d = synth.make_digar_chunk(rng, n_reads=int(rng.integers(40, 90)), read_len=(1500, 5000), err_every=int(rng.choice([150, 400])), ref_len=12000, tech="ont" if it % 4 == 3 else "hifi")
        case, ci = T.noisyreg_case(orc, d, 700 + it, is_ont=int(it % 4 == 3), low_every=int(rng.choice([150, 400])))
        kept, regs = T.ref_noisy_regs(ref, ci, case)
        cases.append({"in": T.noisyreg_case_to_json(case), "kept": kept, "regs": regs})
    return {"cases": cases}


def sdust_lcd():
    """Outputs of the UNMODIFIED sdust() (src/sdust.c:184, via oracle/_ref/libref_shim.so) on seeded reference-like windows: random sequence with planted
    homopolymers / tandem repeats / AT-rich stretches, N runs, lower case, long repeats, the codes 0 .. 3 and stray bytes."""
    import base64
    ref = T.ref_lib()
    rng = np.random.default_rng(20261102)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = [T.sdust_sequence(rng, int(n), lc_every=int(e), n_frac=float(f)) for n, e, f in ((20000, 400, 0.002), (6000, 60, 0.0), (3000, 120, 0.05), (45, 30, 0.0), (21, 30, 0.0))]
    for _ in range(3):      # long tandem repeats with mutations: many perfect intervals alive at once
        u, reps = int(rng.integers(1, 8)), int(rng.integers(100, 400))
        s = np.tile(rng.integers(0, 4, u), reps); hit = rng.random(len(s)) < 0.03; s[hit] = rng.integers(0, 4, int(hit.sum()))
        seqs.append(np.concatenate([acgt[rng.integers(0, 4, 50)], acgt[s], acgt[rng.integers(0, 4, 50)]]))
    s = np.frombuffer(b"ACGTacgt\x00\x01\x02\x03", np.uint8)[rng.integers(0, 12, 2500)].copy(); hit = rng.random(len(s)) < 0.03; s[hit] = rng.integers(0, 256, int(hit.sum())); s[800:860] = s[800]
    seqs.append(s)
    cases = []
    for q in seqs:
        q = np.ascontiguousarray(q, np.uint8)
        for Tt, W in ((5, 20), (8, 16), (4, 24)):
            cases.append({"seq": base64.b64encode(q.tobytes()).decode(), "T": Tt, "W": W, "iv": [list(x) for x in T.sdust(ref, "ref_sdust", q, Tt, W)]})
    return {"cases": cases}


def main():
    only = sys.argv[1:]
    for name, fn in (("wfa_utest", wfa_utest), ("wfa_lcd", wfa_lcd), ("poa_lcd", poa_lcd), ("edlib_lcd", edlib_lcd), ("phase_lcd", phase_lcd), ("pileup_lcd", pileup_lcd), ("digar_lcd", digar_lcd), ("sites_lcd", sites_lcd), ("classify_lcd", classify_lcd), ("poa_ncons_lcd", poa_ncons_lcd), ("noisyreg_lcd", noisyreg_lcd), ("sdust_lcd", sdust_lcd)):
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(fn(), separators=(",", ":")).encode())
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
