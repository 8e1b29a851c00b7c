"""CPU: the product's K5 device logic (longcalld_b200/csrc/poa_device.cuh) compiled for the host with
32-element arrays standing in for warp lanes (tests/emu) against the oracle: consensus and full MSA."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lcd_testlib as T

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libpoa_emu.so"])
    return C.CDLL(os.path.join(EMU_DIR, "libpoa_emu.so"))


# mode bits: 1 = thread-per-problem lane policy (packed int16x2), 2 = tight first-attempt workspace budgets,
#            4 = on-chip previous-row cache (warp policy), 8 = two-phase rows (CTA-per-problem policy)
@pytest.mark.parametrize("mode", [0, 3, 4, 8])
@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 0.12, 41), ("ont", 0.03, 42)])
def test_emu_poa_vs_oracle(emu, oracle, tech, mbp, seed, mode):
    from longcalld_b200 import synth
    emu.emu_poa_mode(mode)
    n = 0
    for r in synth.make_regions(mbp, tech, seed=seed):
        for hap in (1, 2):
            seqs = [s for s, h in zip(r.reads, r.read_hap) if h == hap]
            if not seqs or min(len(s) for s in seqs) == 0:
                continue
            for sub, wb in ((1, 10), (0, -1)):
                if wb < 0 and max(len(s) for s in seqs) > 500:
                    continue
                par = T.poa_params(sub, wb)
                a = T.poa(oracle, "lcd_oracle_poa", seqs, par)
                b = T.poa(emu, "emu_poa", seqs, par)
                assert a[0] == b[0] == 0 and a[1] == b[1] and a[2].shape == b[2].shape and (a[2] == b[2]).all(), (n, sub, wb)
                n += 1
    assert n > 60


def test_emu_poa_vs_fixtures(emu):
    import hashlib
    emu.emu_poa_mode(0)
    g = T.load_golden("poa_lcd")
    for i, c in enumerate(g["cases"][::3]):
        seqs = [np.array([int(x) for x in s], dtype=np.uint8) for s in c["seqs"]]
        rc, cons, msa = T.poa(emu, "emu_poa", seqs, T.poa_params(c["sub_aln"], c["wb"]))
        assert rc == 0 and "".join(map(str, cons)) == c["cons"], i
        assert list(msa.shape) == c["msa_shape"] and hashlib.sha1(msa.tobytes()).hexdigest() == c["msa_sha1"], i


@pytest.mark.parametrize("mode", [0, 8])
def test_emu_poa_two_consensus_vs_oracle(emu, oracle, mode):
    """max_n_cons = 2 (de-novo clustering on the MSA + one consensus per cluster): the device code on the host against the oracle"""
    emu.emu_poa_mode(mode)
    par = T.poa_params(0, -1); par.max_n_cons = 2
    n = two = 0
    for tech, mbp, seed in (("hifi", 0.2, 43), ("ont", 0.08, 44)):
        for seqs in T.denovo_problems(mbp, tech, seed, max_len=400):
            for mf in ((0.2, 0.34) if mode == 0 else (0.2,)):
                a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, par, mf)
                b = T.poa_ncons(emu, "emu_poa_ncons", seqs, par, mf)
                assert a[0] == b[0] == 0 and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3].shape == b[3].shape and (a[3] == b[3]).all(), (n, mf)
            n += 1; two += len(a[1]) == 2
    assert n >= 40 and two >= 10


def test_emu_poa_two_consensus_vs_fixtures(emu):
    from test_oracle_poa_ncons import ncons_fixture_cases, same_as_fixture
    emu.emu_poa_mode(0)
    par = T.poa_params(0, -1); par.max_n_cons = 2
    for i, want in enumerate(list(ncons_fixture_cases())[::3]):
        assert same_as_fixture(T.poa_ncons(emu, "emu_poa_ncons", want[0], par), want), i
