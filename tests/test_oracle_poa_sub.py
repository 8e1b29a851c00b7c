"""CPU: pins the sub-graph alignment of partially covering reads (abpoa_subgraph_nodes abPOA/src/abpoa_graph.c:595-680, the index_map / pre_index
restriction of simd_abpoa_align_sequence_to_subgraph abpoa_align_simd.c:1250-1332, abpoa_add_subgraph_alignment between two inner nodes, the
span update over the BFS index range abpoa_graph.c:559-571) in oracle/poa.c (lcd_oracle_poa_sub) against the unmodified abPOA driven as
abpoa_partial_aln_msa_cons drives it (src/align.c:790-812; oracle/_ref/libref_shim.so: ref_poa_sub): consensus and every MSA cell."""
import numpy as np
import pytest

import lcd_testlib as T


@pytest.mark.parametrize("tech,mbp,seed", [("hifi", 1.2, 51), ("ont", 0.25, 52)])
def test_oracle_vs_live_abpoa(oracle, ref, tech, mbp, seed):
    rng = np.random.default_rng(seed)
    n = n_sub = n_skip = 0
    par = T.poa_params(1, 10)
    for seqs, sb, se in T.partial_cover_problems(mbp, tech, seed, rng):
        a = T.poa_sub(ref, "ref_poa_sub", seqs, sb, se, par)
        b = T.poa_sub(oracle, "lcd_oracle_poa_sub", seqs, sb, se, par)
        assert a[0] == b[0] == 0 and a[1] == b[1] and a[2].shape == b[2].shape and (a[2] == b[2]).all(), n
        n += 1; n_sub += int((sb > 0).sum()); n_skip += int((sb < 0).sum())
    assert n > 300 and n_sub > 1000 and n_skip > 50, (n, n_sub, n_skip)
