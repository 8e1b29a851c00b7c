"""GPU, end to end: `longcallD call` on the reference's bundled HiFi / ONT data with the UNMODIFIED reference linked as a shared
library (oracle/_ref/longcallD_so) and longcalld_b200/dropin/liblcd_dropin.so preloaded, so that the =/X difference lists (K1), the candidate-site list (K1b), the per-site coverage pass
(K2), the read x variant profile (K3), the read->haplotype assignment / phasing (K4), the POA consensus of full-cover regions (K5), every WFA alignment (K6) and every edlib call (K7) execute on
the B200 through the C-ABI.  The VCF body must be the reference's own (md5 of the non-header lines)."""
import hashlib
import os
import subprocess

import pytest

import lcd_testlib as T

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(T.ROOT, "oracle", "_ref")
DROPIN = os.path.join(T.ROOT, "longcalld_b200", "dropin", "liblcd_dropin.so")
GOLDEN = {"hifi": "dcbd4523c01ab37cce5dd88d5e56b564", "ont": "71f0e1aa2ee7667ad2a1f31e2eace81d",
          "mosaic": "ea77d40096193eb9ad4497c297c21fd7"}
FORWARDED_POA = {"hifi": 0, "ont": 0, "mosaic": 0}             # mosaic: HiFi with -s -T <TE consensus> (BASELINE configs[4] on the bundled data)
PARTIAL_POA = {"hifi": 23, "ont": 34, "mosaic": 23}            # POA problems with partially covering / sampled reads: sub-graph alignment on the GPU


def _run(tech, preload, threads=4):
    exe = os.path.join(REF_DIR, "longcallD_so")
    data = os.path.join(REF_DIR, "test_data")
    if not (os.path.exists(exe) and os.path.exists(DROPIN) and os.path.exists(os.path.join(data, "chr11_2M.fa"))):
        pytest.skip("oracle/_ref/longcallD_so, the drop-in or the bundled test data were not built in this checkout")
    env = dict(os.environ, LCD_DROPIN_VERBOSE="1")
    if preload:
        env["LD_PRELOAD"] = DROPIN
    bam = "ont" if tech == "ont" else "hifi"
    extra = ["-s", "-T", os.path.join(data, "AluY_L1_SVA_cons_noPA.fa")] if tech == "mosaic" else []
    cmd = [exe, "call", "--ont" if tech == "ont" else "--hifi"] + extra + [os.path.join(data, "chr11_2M.fa"),
           os.path.join(data, f"HG002_chr11_{bam}_test.bam"), "-t", str(threads)]
    r = subprocess.run(cmd, env=env, capture_output=True, timeout=1500)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    body = b"".join(l + b"\n" for l in r.stdout.split(b"\n") if l and not l.startswith(b"#"))
    return hashlib.md5(body).hexdigest(), r.stderr.decode()


@pytest.mark.parametrize("tech", ["hifi", "ont", "mosaic"])
def test_vcf_identical_with_gpu_dropin(tech):
    md5, err = _run(tech, preload=True)
    calls = [l[l.index("[lcd_dropin] GPU calls"):] for l in err.splitlines() if "[lcd_dropin] GPU calls" in l]
    assert calls, "the drop-in was not loaded"
    counts = dict((k, int(v)) for k, v in __import__("re").findall(r"(digar|sites|pileup|profile|phase|edlib|wfa|poa) (\d+)", calls[-1].split("(library time")[0]))
    need = ("sites", "pileup", "phase", "wfa", "poa") if tech == "mosaic" else ("sites", "pileup", "profile", "phase", "wfa", "poa")   # -s: the profile takes the reference's somatic path
    need += ("digar",)
    fwd = int(__import__("re").search(r"digar \d+ \(forwarded: (\d+)\)", calls[-1]).group(1))
    assert fwd == 0, calls[-1]                      # (the bundled ONT BAM carries plain-M CIGARs + MD tags: K1's MD front end)
    assert all(counts[k] > 0 for k in need), calls[-1]                                             # the kernels really ran on the GPU
    # no POA call of abpoa_partial_aln_msa_cons goes to the reference's abPOA any more: a kernel that starts refusing problems (a silent
    # CPU fallback) fails here, not in the md5; the problems with partially covering / sampled reads are counted
    m_fwd = __import__("re").search(r"forwarded to abPOA: (\d+) \+ (\d+)", calls[-1])
    fwd_poa, fwd_denovo = int(m_fwd.group(1)), int(m_fwd.group(2))
    part_poa = int(__import__("re").search(r"with partially covering reads: (\d+)", calls[-1]).group(1))
    two_cons = int(__import__("re").search(r"max_n_cons = 2: (\d+)", calls[-1]).group(1))
    assert fwd_poa == FORWARDED_POA[tech] and part_poa == PARTIAL_POA[tech], calls[-1]
    # the de-novo POA of regions without a usable phase set (abpoa_aln_msa_cons, two consensus sequences from the read clustering) runs on the GPU too
    assert fwd_denovo == 0 and (two_cons > 0 if tech == "ont" else True), calls[-1]
    # the noisy-region set of every chunk (K2b + K2c) ran on the GPU; -s (mosaic) keeps the reference's own classification (somatic candidates: a14)
    nrs = __import__("re").search(r"noisy-region set (\d+) \(forwarded: (\d+)\)", calls[-1])
    assert (int(nrs.group(1)) > 0 and int(nrs.group(2)) == 0) if tech != "mosaic" else int(nrs.group(1)) == 0, calls[-1]
    print(calls[-1])
    assert md5 == GOLDEN[tech], (md5, calls[-1])


@pytest.mark.parametrize("style", ["m", "md", "cs"])
def test_vcf_identical_for_every_cigar_flavour(style, tmp_path):
    """The reference picks one of four difference-list variants per read (src/collect_var.c:1072-1080).  A synthetic 1 Mb HiFi BAM whose reads carry
    plain-M CIGARs without tags / with MD tags / with cs tags (tools/synth_bam.c) goes through the unmodified reference and through the GPU drop-in
    (K1's tag front end on the device): same VCF body, no chunk forwarded to the reference's own pass."""
    synth = os.path.join(T.ROOT, "tools", "_build", "synth_bam")
    exe, ref_exe = os.path.join(REF_DIR, "longcallD_so"), os.path.join(REF_DIR, "longcallD_ref")
    if not all(os.path.exists(p) for p in (synth, exe, ref_exe, DROPIN)):
        pytest.skip("tools/_build/synth_bam, oracle/_ref or the drop-in were not built in this checkout")
    prefix = str(tmp_path / f"s_{style}")
    subprocess.check_call([synth, prefix, "1", "hifi", "11", "30", "1", "0", style], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    body = lambda out: hashlib.md5(b"".join(l + b"\n" for l in out.split(b"\n") if l and not l.startswith(b"#"))).hexdigest()
    want = subprocess.run([ref_exe, "call", "--hifi", prefix + ".fa", prefix + ".bam", "-t", "4"], capture_output=True, timeout=900)
    assert want.returncode == 0
    got = subprocess.run([exe, "call", "--hifi", prefix + ".fa", prefix + ".bam", "-t", "4"], capture_output=True, timeout=900,
                         env=dict(os.environ, LD_PRELOAD=DROPIN, LCD_DROPIN_VERBOSE="1", LCD_DROPIN_STAGES="all"))
    assert got.returncode == 0, got.stderr.decode()[-2000:]
    line = [l for l in got.stderr.decode().splitlines() if "[lcd_dropin] GPU calls" in l][-1]
    m = __import__("re").search(r"digar (\d+) \(forwarded: (\d+)\)", line)
    assert int(m.group(1)) > 0 and int(m.group(2)) == 0, line
    assert body(got.stdout) == body(want.stdout) and want.stdout.count(b"\n") > 500, line
