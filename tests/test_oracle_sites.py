"""CPU: pins oracle/sites.c (sorted unique candidate sites: collect_all_cand_var_sites, reference src/collect_var.c:1209, with
exact_comp_var_site / exact_comp_var_site_ins :1878-1935) against the unmodified reference (oracle/_ref/libref_shim.so:
ref_collect_sites builds digar_t records around the same flat arrays)."""
import numpy as np

import lcd_testlib as T
from longcalld_b200 import synth


def sites_cases(seed, n):
    """Chunks with shared variants (incl. large insertions with length-perturbed copies: the fuzzy merge) and random errors."""
    rng = np.random.default_rng(seed)
    for it in range(n):
        d = synth.make_pileup_chunk(rng, ref_len=int(rng.choice([600, 3000, 20000])), n_reads=int(rng.choice([1, 8, 60, 250])),
                                    read_len=(200, 500) if it % 4 == 0 else (1500, 6000), var_every=int(rng.choice([40, 150, 600])),
                                    err_every=int(rng.choice([60, 800])), min_sv_len=int(rng.choice([30, 50])))
        lo, hi = int(d["read_beg"].min()), int(d["read_end"].max())
        reg = (-1, -1) if it % 5 == 0 else (lo + (hi - lo) // 5, hi - (hi - lo) // 5)
        yield d, reg


def test_oracle_vs_live_reference(oracle, ref):
    tot = fuzzy = 0
    for n, (d, (b, e)) in enumerate(sites_cases(41, 150)):
        a = T.collect_sites(oracle, "lcd_oracle_collect_sites", d, b, e)
        r = T.collect_sites(ref, "ref_collect_sites", d, b, e, src_is_offset=True)
        assert a == r, (n, len(a), len(r), [x for x, y in zip(a, r) if x != y][:2])
        tot += len(a); fuzzy += sum(1 for s in a if s[1] == 1 and s[3] >= d["min_sv_len"])
    assert tot > 20000 and fuzzy > 200, (tot, fuzzy)


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("sites_lcd")
    assert len(g["cases"]) >= 16
    for c in g["cases"]:
        d, reg, want = T.sites_case_from_json(c)
        assert T.collect_sites(oracle, "lcd_oracle_collect_sites", d, reg[0], reg[1]) == want
