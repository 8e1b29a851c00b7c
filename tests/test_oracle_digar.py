"""CPU: pins oracle/digar.c (difference lists from =/X CIGARs: collect_digar_from_eqx_cigar + push_xid_size_queue_win,
reference src/bam_utils.c:701-841, :161-205, driven as collect_digars_from_bam does, src/collect_var.c:1063-1082) against
the unmodified reference (oracle/_ref/libref_shim.so: ref_collect_digar_eqx builds bam1_t records with htslib's bam_set1)
and against committed reference outputs."""
import numpy as np

import lcd_testlib as T
from longcalld_b200 import synth


def digar_cases(seed, n):
    rng = np.random.default_rng(seed)
    for it in range(n):
        yield synth.make_digar_chunk(rng, n_reads=int(rng.choice([1, 6, 40, 120])), read_len=(60, 400) if it % 5 == 0 else (800, 5000),
                                     err_every=int(rng.choice([15, 80, 300, 1500])), tech="ont" if it % 3 == 0 else "hifi",
                                     low_qual_frac=float(rng.choice([0.0, 0.05, 0.4])))


def test_oracle_vs_live_reference(oracle, ref):
    n_iv = n_skip = n_ev = n_civ = 0
    for n, d in enumerate(digar_cases(21, 150)):
        a = T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d)
        b = T.collect_digar(ref, "ref_collect_digar_eqx", d)
        assert a["qual_counts"] == b["qual_counts"], n
        assert a["totals"] == b["totals"], (n, a["totals"], b["totals"])
        for r in b["reads"]:
            assert a["reads"][r] == b["reads"][r], (n, r, [x for x, y in zip(a["reads"][r], b["reads"][r]) if x != y][:1])
        assert a["chunk_noisy"] == b["chunk_noisy"], n
        n_iv += sum(len(v[4]) for v in b["reads"].values()); n_skip += sum(v[0] for v in b["reads"].values())
        n_ev += sum(len(v[3]) for v in b["reads"].values()); n_civ += len(b["chunk_noisy"])
    assert n_iv > 3000 and n_skip > 100 and n_ev > 100000 and n_civ > 500, (n_iv, n_skip, n_ev, n_civ)


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("digar_lcd")
    assert len(g["cases"]) >= 12
    for c in g["cases"]:
        d = T.digar_case_from_json(c["in"])
        a = T.collect_digar(oracle, "lcd_oracle_collect_digar_eqx", d)
        assert T.digar_digest(a) == c["digest"]
        assert a["chunk_noisy"] == [tuple(x) for x in c["chunk_noisy"]]
