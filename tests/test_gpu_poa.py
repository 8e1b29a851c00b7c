"""GPU: liblcd_gpu.so's POA kernel (through the C-ABI) against the oracle / golden fixtures, bit-exact on
consensus and on every cell of the row-column MSA."""
import hashlib

import numpy as np
import pytest

import lcd_testlib as T

pytestmark = pytest.mark.gpu


def _problems(tech, mbp, seed, max_len=None):
    from longcalld_b200 import synth
    out = []
    for r in synth.make_regions(mbp, tech, seed=seed):
        for hap in (1, 2):
            seqs = [s for s, h in zip(r.reads, r.read_hap) if h == hap]
            if not seqs or min(len(s) for s in seqs) == 0:
                continue
            if max_len and max(len(s) for s in seqs) > max_len:
                continue
            out.append(seqs)
    return out


def _check(gpu, oracle, problems, sub, wb):
    got = gpu.poa_batch(problems, gpu.poa_params(sub, wb))
    par = T.poa_params(sub, wb)
    bad = []
    for i, seqs in enumerate(problems):
        rc, cons, msa = T.poa(oracle, "lcd_oracle_poa", seqs, par)
        g = got[i]
        if not (g[0] == rc == 0 and g[1] == cons and g[2].shape == msa.shape and (g[2] == msa).all()):
            bad.append(i)
    assert not bad, (len(bad), bad[:10])


def test_gpu_poa_vs_reference_fixtures(gpu):
    g = T.load_golden("poa_lcd")
    for sub, wb in ((1, 10), (0, -1)):
        cases = [c for c in g["cases"] if c["sub_aln"] == sub and c["wb"] == wb]
        problems = [[np.array([int(x) for x in s], dtype=np.uint8) for s in c["seqs"]] for c in cases]
        got = gpu.poa_batch(problems, gpu.poa_params(sub, wb))
        for i, (c, (st, cons, msa)) in enumerate(zip(cases, got)):
            assert st == 0 and "".join(map(str, cons)) == c["cons"], (sub, wb, i)
            assert list(msa.shape) == c["msa_shape"] and hashlib.sha1(msa.tobytes()).hexdigest() == c["msa_sha1"], (sub, wb, i)


def test_gpu_poa_vs_oracle_hifi(gpu, oracle):
    _check(gpu, oracle, _problems("hifi", 0.6, 51), 1, 10)


def test_gpu_poa_vs_oracle_ont(gpu, oracle):
    _check(gpu, oracle, _problems("ont", 0.12, 52), 1, 10)


def test_gpu_poa_vs_oracle_unbanded(gpu, oracle):
    _check(gpu, oracle, _problems("hifi", 0.3, 53, max_len=700), 0, -1)


def test_gpu_poa_edge_cases_and_rerun(gpu, oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 4, 50).astype(np.uint8)
    problems = [[a], [a, a], [a, a[:25]], [a[:1], a[:1], a[:2]], [a, T.mutate(rng, a, sub=0.3), T.mutate(rng, a, ins=0.2)],
                [np.zeros(40, np.uint8)] * 5 + [np.zeros(37, np.uint8)] * 4]
    _check(gpu, oracle, problems, 1, 10)
    _check(gpu, oracle, problems, 0, -1)
    assert gpu.poa_batch([], gpu.poa_params()) == []
    seqs, first, n_reads, read_off, read_len = gpu.pack_poa(problems)
    plan = gpu.PoaPlan(seqs, first, n_reads, read_off, read_len, gpu.poa_params())
    ref = None
    for _ in range(3):
        plan.run(); plan.sync()
        res, cons, cons_off, msa, msa_off = plan.fetch()
        cur = [(int(r["status"]), cons[cons_off[i]:cons_off[i] + r["cons_len"]].tobytes()) for i, r in enumerate(res)]
        ref = ref or cur
        assert cur == ref
    assert plan.work_units() > 0


def test_gpu_poa_full_size_properties(gpu):
    """BASELINE-sized slice (5 Mb of regions; too slow for the scalar oracle): every MSA row, gaps removed,
    must spell its read; the consensus row, gaps removed, must equal the consensus."""
    problems = _problems("hifi", 5.0, 54)
    got = gpu.poa_batch(problems, gpu.poa_params())
    for seqs, (st, cons, msa) in zip(problems, got):
        assert st == 0
        for r, s in enumerate(seqs):
            row = msa[r]
            assert row[row != 5].tobytes() == np.asarray(s, np.uint8).tobytes()
        assert msa[len(seqs)][msa[len(seqs)] != 5].tobytes() == cons


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 1.5, 71), ("ont", 0.3, 72)])
def test_gpu_poa_partial_cover_reads(gpu, oracle, tech, mbp, seed):
    """Reads that cover their region only partly (abpoa_partial_aln_msa_cons, src/align.c:790-812): lcd_poa_sub_batch aligns them against the
    sub-graph abpoa_subgraph_nodes finds between their anchor nodes, leaves out the ones marked so, and mixes such problems with plain ones in one
    launch; consensus and every MSA cell against the oracle that tests/test_oracle_poa_sub.py pins to the unmodified abPOA."""
    rng = np.random.default_rng(seed)
    cases = list(T.partial_cover_problems(mbp, tech, seed, rng, max_len=3000))
    plain = _problems(tech, mbp / 8, seed + 1, max_len=1500)
    problems = [c[0] for c in cases] + plain
    sub = [(c[1], c[2]) for c in cases] + [(np.zeros(len(p), np.int32), np.zeros(len(p), np.int32)) for p in plain]
    got = gpu.poa_batch(problems, gpu.poa_params(1, 10), sub=sub)
    par = T.poa_params(1, 10)
    bad = []
    for i, (seqs, (sb, se)) in enumerate(zip(problems, sub)):
        rc, cons, msa = T.poa_sub(oracle, "lcd_oracle_poa_sub", seqs, sb, se, par)
        g = got[i]
        if not (g[0] == rc == 0 and g[1] == cons and g[2].shape == msa.shape and (g[2] == msa).all()):
            bad.append(i)
    assert not bad and len(cases) > 300 and sum(int((c[1] > 0).sum()) for c in cases) > 1000, (len(bad), bad[:10], len(cases))
    # anchors outside the first read are refused, loudly
    seqs, sb, se = cases[0]
    sb = sb.copy(); se = se.copy(); k = int(np.argmax(sb > 0)); sb[k] = 10 ** 6; se[k] = 10 ** 6 + 5
    with pytest.raises(gpu.LcdGpuError):
        gpu.poa_batch([seqs], gpu.poa_params(1, 10), sub=[(sb, se)])


def test_gpu_poa_two_consensus_vs_oracle(gpu, oracle):
    """lcd_poa_ncons_batch (abpoa_aln_msa_cons, max_n_cons = 2): number of clusters, both consensus sequences, every read's cluster and the
    (n_reads + n_cons)-row MSA, bit-exact; warp kernel (reads <= 224 bp) and CTA kernel (longer reads, unbanded) in one batch; problems
    with max_n_cons = 1 in the same batch; another min_freq."""
    problems = list(T.denovo_problems(0.5, "hifi", 51, max_len=700)) + list(T.denovo_problems(0.12, "ont", 52, max_len=500))
    assert len(problems) >= 80 and any(max(len(s) for s in p) > 224 for p in problems) and any(max(len(s) for s in p) <= 224 for p in problems)
    par2 = list(gpu.poa_params(0, -1)); par2[9] = 2
    par = T.poa_params(0, -1); par.max_n_cons = 2
    two = 0
    for mf in (0.20, 0.34):
        got = gpu.poa_ncons_batch(problems, tuple(par2), mf)
        bad = []
        for i, seqs in enumerate(problems):
            a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, par, mf)
            g = got[i]
            if not (g[0] == a[0] == 0 and g[1] == a[1] and np.array_equal(g[2], a[2]) and g[3].shape == a[3].shape and (g[3] == a[3]).all()):
                bad.append(i)
            two += len(a[1]) == 2
        assert not bad, (mf, len(bad), bad[:10])
    assert two >= 30
    # mixed batch: every other problem asks for one consensus only
    pars = [tuple(par2) if i % 2 else gpu.poa_params(0, -1) for i in range(len(problems))]
    got = gpu.poa_ncons_batch(problems, pars, 0.20)
    p1 = T.poa_params(0, -1)
    for i, seqs in enumerate(problems[:60]):
        if i % 2:
            a = T.poa_ncons(oracle, "lcd_oracle_poa_ncons", seqs, par, 0.20)
            assert got[i][1] == a[1] and np.array_equal(got[i][2], a[2]) and (got[i][3] == a[3]).all(), i
        else:
            rc, cons, msa = T.poa(oracle, "lcd_oracle_poa", seqs, p1)
            assert got[i][0] == rc == 0 and got[i][1] == [cons] and (got[i][3] == msa).all() and not got[i][2].any(), i
    # lcd_poa_batch keeps refusing max_n_cons = 2 (it cannot return the clusters)
    with pytest.raises(gpu.LcdGpuError):
        gpu.poa_batch(problems[:2], tuple(par2))


def test_gpu_poa_two_consensus_vs_reference_fixtures(gpu):
    """committed outputs of the unmodified abpoa_aln_msa_cons: no oracle, no /root/reference in this comparison"""
    from test_oracle_poa_ncons import ncons_fixture_cases, same_as_fixture
    cases = list(ncons_fixture_cases())
    par2 = list(gpu.poa_params(0, -1)); par2[9] = 2
    got = gpu.poa_ncons_batch([c[0] for c in cases], tuple(par2), 0.20)
    bad = [i for i, (g, c) in enumerate(zip(got, cases)) if not same_as_fixture(g, c)]
    assert not bad and len(cases) >= 40, bad[:10]
