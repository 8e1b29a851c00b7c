"""CPU: pins the 2-consensus de-novo POA of the oracle (oracle/poa_cluster.c + lcd_oracle_poa_ncons: abPOA's k-medoids read clustering on
the row-column MSA and one most-frequent consensus per cluster) against the UNMODIFIED reference's own abpoa_aln_msa_cons
(src/align.c:872-953 -> abPOA/src/abpoa_output.c:676-1180), called through oracle/_ref/libref_shim.so."""
import numpy as np
import pytest
import lcd_testlib as T


def _par(max_n_cons=2):
    p = T.poa_params(0, -1)
    p.max_n_cons = max_n_cons
    return p


def _same(a, b):
    return a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3].shape == b[3].shape and (a[3] == b[3]).all()


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 0.6, 21), ("ont", 0.4, 22), ("hifi", 0.5, 23)])
def test_oracle_ncons_vs_live_reference(oracle, ref, tech, mbp, seed):
    n = two = 0
    for seqs in T.denovo_problems(mbp, tech, seed):
        a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, _par())
        b = T.poa_ncons(ref, "ref_poa_ncons", seqs, _par())
        assert b[0] == 0 and _same(a, b), f"problem {n}: oracle differs from abpoa_aln_msa_cons"
        n += 1; two += len(b[1]) == 2
    assert n >= 100 and two >= 30 and two < n, (n, two)


def test_oracle_ncons_other_min_freq_and_one_cluster(oracle, ref):
    rng = np.random.default_rng(5)
    n = 0
    for seqs in T.denovo_problems(0.1, "ont", 31):
        for mf in (0.1, 0.5):
            a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, _par(), mf)
            b = T.poa_ncons(ref, "ref_poa_ncons", seqs, _par(), mf)
            assert _same(a, b), (n, mf)
        # max_n_cons = 1 through the same entry point is the single-consensus path
        a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, _par(1))
        b = T.poa(oracle, "lcd_oracle_poa", seqs, T.poa_params(0, -1))
        assert a[0] == b[0] == 0 and a[1] == [b[1]] and (a[3] == b[2]).all()
        n += 1
    assert n >= 10
    # reads of one haplotype only, identical reads, two reads
    base = rng.integers(0, 4, 300).astype(np.uint8)
    for seqs in ([base] * 6, [base, T.mutate(rng, base)], [T.mutate(rng, base) for _ in range(12)]):
        a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, _par())
        b = T.poa_ncons(ref, "ref_poa_ncons", seqs, _par())
        assert _same(a, b)


def ncons_fixture_cases():
    import hashlib
    for c in T.load_golden("poa_ncons_lcd")["cases"]:
        seqs = [np.array([int(x) for x in s], dtype=np.uint8) for s in c["seqs"]]
        yield seqs, [bytes(int(x) for x in s) for s in c["cons"]], np.array([int(x) for x in c["clu"]], np.uint8), c["msa_shape"], c["msa_sha1"]


def same_as_fixture(got, want):
    import hashlib
    _, cons, clu, shape, sha = want
    return got[0] == 0 and got[1] == cons and np.array_equal(got[2], clu) and list(got[3].shape) == shape and hashlib.sha1(got[3].tobytes()).hexdigest() == sha


def test_oracle_ncons_vs_reference_fixtures(oracle):
    """committed outputs of the unmodified abpoa_aln_msa_cons (tests/golden/make_golden.py poa_ncons_lcd): runs where /root/reference does not exist"""
    n = two = 0
    for want in ncons_fixture_cases():
        assert same_as_fixture(T.poa_ncons(oracle, "lcd_oracle_poa_ncons", want[0], _par()), want), n
        n += 1; two += len(want[1]) == 2
    assert n >= 40 and two >= 10
