"""CPU, end to end: the reference-side binding (longcalld_b200/dropin/lcd_dropin.c: marshalling of the reference's structures, the coroutine
region driver that runs the pending noisy regions of a chunk side by side, the cross-thread combiner that turns their engine calls into
one batch per engine) preloaded into the UNMODIFIED reference, with the C-ABI answered by an oracle-backed test double
(tests/emu/fake_lcd_gpu.c) instead of liblcd_gpu.so -- so the host logic is checked where there is no GPU: the VCF body must be the
reference's own (md5s of SURVEY.md section 6) at any thread count, and the engine calls must really have been batched."""
import hashlib
import os
import re
import subprocess

import pytest

import lcd_testlib as T

REF_DIR = os.path.join(T.ROOT, "oracle", "_ref")
EMU_DIR = os.path.join(T.ROOT, "tests", "emu")
GOLDEN = {"hifi": "dcbd4523c01ab37cce5dd88d5e56b564", "ont": "71f0e1aa2ee7667ad2a1f31e2eace81d", "mosaic": "ea77d40096193eb9ad4497c297c21fd7"}


@pytest.fixture(scope="module")
def dropin_cpu():
    if not (os.path.exists(os.path.join(REF_DIR, "longcallD_so")) and os.path.isdir("/root/reference/src")):
        pytest.skip("needs the reference built by oracle/Makefile (oracle/_ref) and its headers")
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "liblcd_dropin_cpu.so"])
    return os.path.join(EMU_DIR, "liblcd_dropin_cpu.so")


def run(so, tech, threads, extra_env=None):
    data = os.path.join(REF_DIR, "test_data")
    bam = "ont" if tech == "ont" else "hifi"
    extra = ["-s", "-T", os.path.join(data, "AluY_L1_SVA_cons_noPA.fa")] if tech == "mosaic" else []
    cmd = [os.path.join(REF_DIR, "longcallD_so"), "call", "--ont" if tech == "ont" else "--hifi"] + extra + \
          [os.path.join(data, "chr11_2M.fa"), os.path.join(data, f"HG002_chr11_{bam}_test.bam"), "-t", str(threads)]
    r = subprocess.run(cmd, env=dict(os.environ, LD_PRELOAD=so, LCD_DROPIN_VERBOSE="1", **(extra_env or {})), capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    body = b"".join(l + b"\n" for l in r.stdout.split(b"\n") if l and not l.startswith(b"#"))
    line = [l for l in r.stderr.decode().splitlines() if "[lcd_dropin] GPU calls" in l][-1]
    return hashlib.md5(body).hexdigest(), line


@pytest.mark.parametrize("tech,threads", [("hifi", 1), ("hifi", 8), ("ont", 4), ("mosaic", 4)])
def test_vcf_identical_through_the_batched_driver(dropin_cpu, tech, threads):
    md5, line = run(dropin_cpu, tech, threads)
    assert md5 == GOLDEN[tech], line
    c = {k: int(v) for k, v in re.findall(r"(edlib|wfa|poa) (\d+)", line.split("(library time")[0])}
    batches = int(re.search(r"in (\d+) engine batches", line).group(1))
    assert c["poa"] > 100 and c["wfa"] > 100
    fwd = re.search(r"forwarded to abPOA: (\d+) \+ (\d+)", line)
    assert fwd and int(fwd.group(1)) == 0 and int(fwd.group(2)) == 0, line          # partial-cover and de-novo (two-consensus) POA included
    if tech == "ont":
        assert int(re.search(r"max_n_cons = 2: (\d+)", line).group(1)) > 0, line
    nrs = re.search(r"noisy-region set (\d+) \(forwarded: (\d+)\)", line)
    assert (int(nrs.group(1)) > 0 and int(nrs.group(2)) == 0) if tech != "mosaic" else int(nrs.group(1)) == 0, line      # -s keeps the reference's own classification
    if tech != "mosaic":                 # -s rewrites the difference lists region by region: one region at a time there
        assert batches * 4 < c["poa"] + c["wfa"] + c["edlib"], line


def test_serial_regions_give_the_same_vcf(dropin_cpu):
    """LCD_DROPIN_SERIAL=1 runs the regions one after the other (every engine call a batch of one): same records."""
    md5, line = run(dropin_cpu, "hifi", 2, {"LCD_DROPIN_SERIAL": "1"})
    assert md5 == GOLDEN["hifi"], line


@pytest.mark.parametrize("style", ["m", "md", "cs"])
def test_every_cigar_flavour_through_the_binding(dropin_cpu, style, tmp_path):
    """Plain-M reads without tags / with MD tags / with cs tags (tools/synth_bam.c): the binding hands every read's variant to lcd_digar_tags_batch
    (the double answers with the oracle's restatements of the reference's three other passes); same VCF as the unmodified reference, no chunk forwarded."""
    synth = os.path.join(T.ROOT, "tools", "_build", "synth_bam")
    if not os.path.exists(synth):
        pytest.skip("tools/_build/synth_bam was not built")
    prefix = str(tmp_path / f"s_{style}")
    subprocess.check_call([synth, prefix, "0.6", "hifi", "11", "30", "1", "0", style], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    body = lambda out: hashlib.md5(b"".join(l + b"\n" for l in out.split(b"\n") if l and not l.startswith(b"#"))).hexdigest()
    cmd = ["call", "--hifi", prefix + ".fa", prefix + ".bam", "-t", "4"]
    want = subprocess.run([os.path.join(REF_DIR, "longcallD_ref")] + cmd, capture_output=True, timeout=900)
    got = subprocess.run([os.path.join(REF_DIR, "longcallD_so")] + cmd, capture_output=True, timeout=900, env=dict(os.environ, LD_PRELOAD=dropin_cpu, LCD_DROPIN_VERBOSE="1"))
    assert want.returncode == 0 and got.returncode == 0, got.stderr.decode()[-2000:]
    line = [l for l in got.stderr.decode().splitlines() if "[lcd_dropin] GPU calls" in l][-1]
    m = re.search(r"digar (\d+) \(forwarded: (\d+)\)", line)
    assert int(m.group(1)) > 0 and int(m.group(2)) == 0, line
    assert body(got.stdout) == body(want.stdout) and want.stdout.count(b"\n") > 300, line


@pytest.mark.parametrize("tech", ["hifi", "ont"])
def test_noisy_region_set_self_check_on_the_bundled_data(dropin_cpu, tech):
    """LCD_DROPIN_CHECK_K2C=1: per chunk the binding asks the library (here the oracle-backed double) for the noisy-region set, then runs the unmodified
    pre_process_noisy_regs + classify_cand_vars on the same chunk and compares both stages -- the restatement pinned on real reads."""
    md5, line = run(dropin_cpu, tech, 2, {"LCD_DROPIN_CHECK_K2C": "1"})
    assert md5 == GOLDEN[tech]
    data = os.path.join(REF_DIR, "test_data")
    r = subprocess.run([os.path.join(REF_DIR, "longcallD_so"), "call", "--" + tech, os.path.join(data, "chr11_2M.fa"), os.path.join(data, f"HG002_chr11_{tech}_test.bam"), "-t", "2"],
                       env=dict(os.environ, LD_PRELOAD=dropin_cpu, LCD_DROPIN_CHECK_K2C="1"), capture_output=True, timeout=900)
    checks = [l for l in r.stderr.decode().splitlines() if l.startswith("[k2c check]")]
    big = [l for l in checks if re.search(r"chunk \d+: (\d+) sites", l) and int(re.search(r"chunk \d+: (\d+) sites", l).group(1)) > 1000]
    assert big and all("identical" in l and "DIFFERENT" not in l for l in checks), [l for l in checks if "DIFFERENT" in l][:3]
