"""CPU: pins oracle/poa.c (scalar restatement of abPOA's convex-gap adaptive-banded POA as longcallD
drives it) against committed outputs of the unmodified abPOA and, where oracle/_ref exists, against the
live reference on seeded noisy regions -- consensus AND the full row-column MSA, bit for bit."""
import hashlib

import numpy as np
import pytest

import lcd_testlib as T


def _case_seqs(c):
    return [np.array([int(x) for x in s], dtype=np.uint8) for s in c["seqs"]]


def test_oracle_poa_vs_reference_fixtures(oracle):
    g = T.load_golden("poa_lcd")
    assert len(g["cases"]) > 100
    for i, c in enumerate(g["cases"]):
        rc, cons, msa = T.poa(oracle, "lcd_oracle_poa", _case_seqs(c), T.poa_params(c["sub_aln"], c["wb"]))
        assert rc == 0
        assert "".join(map(str, cons)) == c["cons"], i
        assert list(msa.shape) == c["msa_shape"] and hashlib.sha1(msa.tobytes()).hexdigest() == c["msa_sha1"], i


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 0.4, 31), ("ont", 0.06, 32)])
def test_oracle_poa_vs_live_reference(oracle, ref, tech, mbp, seed):
    from longcalld_b200 import synth
    n = 0
    for r in synth.make_regions(mbp, tech, seed=seed):
        for hap in (1, 2):
            seqs = [s for s, h in zip(r.reads, r.read_hap) if h == hap]
            if not seqs or min(len(s) for s in seqs) == 0:
                continue
            for sub, wb in ((1, 10), (0, -1)):
                if wb < 0 and max(len(s) for s in seqs) > 800:
                    continue
                par = T.poa_params(sub, wb)
                a = T.poa(oracle, "lcd_oracle_poa", seqs, par)
                b = T.poa(ref, "ref_poa", seqs, par)
                assert a[0] == b[0] == 0 and a[1] == b[1] and a[2].shape == b[2].shape and (a[2] == b[2]).all(), (n, sub, wb)
                n += 1
    assert n > 100


def test_oracle_poa_edge_cases(oracle, ref):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 4, 50).astype(np.uint8)
    cases = [[a], [a, a], [a, a[:25]], [a[:1], a[:1], a[:2]], [a, T.mutate(rng, a, sub=0.3), T.mutate(rng, a, ins=0.2)],
             [np.zeros(40, np.uint8)] * 5 + [np.zeros(37, np.uint8)] * 4,
             [np.concatenate([a[:20], np.full(3, 4, np.uint8), a[23:]]), a, a]]   # N bases score 0 against everything
    for seqs in cases:
        for sub, wb in ((1, 10), (0, -1)):
            par = T.poa_params(sub, wb)
            x = T.poa(oracle, "lcd_oracle_poa", seqs, par)
            y = T.poa(ref, "ref_poa", seqs, par)
            assert x[0] == y[0] and x[1] == y[1] and x[2].shape == y[2].shape and (x[2] == y[2]).all()
