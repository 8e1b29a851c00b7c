"""CPU: bench.py's reference arm (the unmodified reference's own functions via oracle/_ref, all host threads) runs on a small workload and
prints the contract's JSON line -- every timed stage of the B200 arm has its reference counterpart in it."""
import json
import os
import subprocess
import sys

import pytest

import lcd_testlib as T


def test_reference_arm_line():
    if T.ref_lib() is None:
        pytest.skip("oracle/_ref/libref_shim.so not built (no /root/reference here)")
    # the reference arm must never load the product library: point it at a file that does not exist
    r = subprocess.run([sys.executable, os.path.join(T.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--mbp", "1"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, LCD_GPU_SO="/nonexistent/liblcd_gpu.so"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                         # stdout carries exactly one JSON line
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "ref_Mbp_per_s_called" and j["unit"] == "Mbp/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1
    for k in ("poa_s", "wfa_s", "phase_s", "edlib_s", "digar_s", "sites_s", "pileup_s", "classify_s", "noisyreg_s", "profile_s"):
        assert cb[k] > 0, k
    assert len(j["config"]["stages"]) == 10 and j["config"]["pileup"]["candidate_sites"] > 0
