#!/usr/bin/env python
"""MEASUREMENT: `longcallD call` itself, BAM + FASTA -> VCF, on a synthetic BAM of BASELINE.json's shape (tools/synth_bam.c):
the unmodified reference (oracle/_ref/longcallD_ref -t <threads>) against the same program with the GPU drop-in preloaded
(oracle/_ref/longcallD_so + longcalld_b200/dropin/liblcd_dropin.so).  Prints one JSON object: wall seconds from the tool's own
`Real time` line (src/call_var_main.c:1030), Mbp/s, md5 of the non-header VCF lines of both runs and whether they are equal.

    python tools/whole_program.py --mb 50 --tech hifi [--threads N] [--stages all|engines] [--keep DIR]
Nothing here reads /root/reference: the binaries and the generator were built in-tree (oracle/_ref, tools/_build) and travel to the GPU box."""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(ROOT, "longcalld_b200", "dropin", "liblcd_dropin.so")
SYNTH = os.path.join(ROOT, "tools", "_build", "synth_bam")


def available():
    return all(os.path.exists(p) for p in (SYNTH, DROPIN, os.path.join(REF_DIR, "longcallD_ref"), os.path.join(REF_DIR, "longcallD_so")))


def make_bam(prefix, mb, tech, seed=11, coverage=30, mosaic=0):
    if not os.path.exists(prefix + ".bam.bai"):
        subprocess.check_call([SYNTH, prefix, str(mb), tech, str(seed), str(coverage), "1", str(mosaic)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return prefix + ".fa", prefix + ".bam"


def call(exe, fa, bam, tech, threads, env_extra=None, extra_args=()):
    cmd = [exe, "call", "--ont" if tech == "ont" else "--hifi", *extra_args, fa, bam, "-t", str(threads)]
    env = dict(os.environ, **(env_extra or {}))
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True)
    wall = time.perf_counter() - t0
    err = r.stderr.decode(errors="replace")
    if r.returncode != 0:
        raise RuntimeError("longcallD call failed: " + err[-2000:])
    body = b"".join(l + b"\n" for l in r.stdout.split(b"\n") if l and not l.startswith(b"#"))
    m = re.search(r"Real time: ([0-9.]+) sec; CPU: ([0-9.]+) sec", err)
    rep = [l[l.index("[lcd_dropin]"):] for l in err.splitlines() if "[lcd_dropin] GPU calls" in l]
    return {"real_s": float(m.group(1)) if m else wall, "cpu_s": float(m.group(2)) if m else None, "process_wall_s": wall,
            "vcf_md5": hashlib.md5(body).hexdigest(), "vcf_records": body.count(b"\n"), "dropin": rep[-1] if rep else None}


def run(mb=50.0, tech="hifi", threads=None, stages="all", workdir=None, seed=11, reps=1, skip_reference=False, device=None):
    threads = threads or os.cpu_count()
    own = workdir is None
    workdir = workdir or tempfile.mkdtemp(prefix="lcd_wp_")
    os.makedirs(workdir, exist_ok=True)
    fa, bam = make_bam(os.path.join(workdir, f"synth_{tech}_{mb:g}mb_s{seed}"), mb, tech, seed)
    out = {"mb": mb, "tech": tech, "threads": threads, "stages": stages, "bam_bytes": os.path.getsize(bam)}
    if not skip_reference:
        ref = min((call(os.path.join(REF_DIR, "longcallD_ref"), fa, bam, tech, threads) for _ in range(reps)), key=lambda d: d["real_s"])
        out["reference"] = ref
        out["reference_mbp_s"] = mb / ref["real_s"]
    env = {"LD_PRELOAD": DROPIN, "LCD_DROPIN_VERBOSE": "1", "LCD_DROPIN_STAGES": stages}
    if device is not None:
        env["LCD_DROPIN_DEVICE"] = str(device)
    gpu = min((call(os.path.join(REF_DIR, "longcallD_so"), fa, bam, tech, threads, env) for _ in range(reps)), key=lambda d: d["real_s"])
    out["gpu"] = gpu
    out["gpu_mbp_s"] = mb / gpu["real_s"]
    if not skip_reference:
        out["vcf_md5_equal"] = gpu["vcf_md5"] == out["reference"]["vcf_md5"]
        out["speedup"] = out["reference"]["real_s"] / gpu["real_s"]
    if own:
        for f in os.listdir(workdir):
            os.unlink(os.path.join(workdir, f))
        os.rmdir(workdir)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=50.0)
    ap.add_argument("--tech", default="hifi", choices=["hifi", "ont"])
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--stages", default="all", choices=["all", "engines"])
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--keep", default=None, help="work directory to keep the BAM in")
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "oracle/_ref, tools/_build/synth_bam or the drop-in were not built"})); sys.exit(0)
    print(json.dumps(run(a.mb, a.tech, a.threads, a.stages, a.keep, a.seed, a.reps, a.skip_reference)))
