"""CPU: the WARP policy of the product's K5 device logic (longcalld_b200/csrc/poa_device.cuh: packed int16x2 chain segments with
the previous row in registers, strip rows, ballot backtrack, parallel fusion) run under the one-warp SIMT emulator of
tests/emu/simt_emu.h (32 fibers, every *_sync collective a rendezvous) against the oracle: consensus and full MSA."""
import ctypes as C
import os
import subprocess

import pytest

import lcd_testlib as T

EMU_DIR = os.path.join(T.ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libpoa_simt_emu.so"])
    lib = C.CDLL(os.path.join(EMU_DIR, "libpoa_simt_emu.so"))
    lib.emu_poa_stat.restype = C.c_long
    return lib


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 0.15, 41), ("ont", 0.03, 42)])
def test_simt_poa_vs_oracle(emu, oracle, tech, mbp, seed):
    from longcalld_b200 import synth
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
    # the packed chain segments must be what computed (almost) all rows, in all three strip widths
    seg, gen = emu.emu_poa_stat(0), emu.emu_poa_stat(1)
    assert seg > 3 * gen and all(emu.emu_poa_stat(i) > 0 for i in (3, 4, 6)), (seg, gen)


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 0.5, 61), ("ont", 0.1, 62)])
def test_simt_poa_partial_cover_vs_oracle(emu, oracle, tech, mbp, seed):
    """Partially covering reads: the device-side BFS node index, abpoa_subgraph_nodes, the sub-graph view of the alignment (rows, restricted
    in-edge lists with the reference's path-score indexing), fusion between two inner nodes and the span update -- against the oracle that
    tests/test_oracle_poa_sub.py pins to the unmodified abPOA."""
    import numpy as np
    rng = np.random.default_rng(seed)
    par = T.poa_params(1, 10)
    n = n_sub = 0
    for seqs, sb, se in T.partial_cover_problems(mbp, tech, seed, rng, max_len=900):
        a = T.poa_sub(oracle, "lcd_oracle_poa_sub", seqs, sb, se, par)
        b = T.poa_sub(emu, "emu_poa_sub", seqs, sb, se, par)
        assert a[0] == b[0] == 0 and a[1] == b[1] and a[2].shape == b[2].shape and (a[2] == b[2]).all(), (n, sb.tolist(), se.tolist())
        n += 1; n_sub += int((sb > 0).sum())
    assert n > 40 and n_sub > 150, (n, n_sub)


def test_simt_poa_two_consensus_vs_oracle(emu, oracle):
    """max_n_cons = 2 under the one-warp SIMT emulator (lanes dealt over columns / candidates / read pairs)"""
    import numpy as np
    emu.emu_poa_mode(0)
    par = T.poa_params(0, -1); par.max_n_cons = 2
    n = two = 0
    for seqs in T.denovo_problems(0.5, "hifi", 45, max_len=260, max_reads=24):
        a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, par)
        b = T.poa_ncons(emu, "emu_poa_ncons", seqs, par)
        assert a[0] == b[0] == 0 and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3].shape == b[3].shape and (a[3] == b[3]).all(), n
        n += 1; two += len(a[1]) == 2
        if n >= 20:
            break
    assert n >= 10 and two >= 3
