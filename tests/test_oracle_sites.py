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


def test_site_list_from_sites_matches_the_records(oracle):
    """bench / drop-in helper: the lcd_site_list_t arrays built from a site list carry the alt bases of the records the sites stand for."""
    for n, (d, (b, e)) in enumerate(sites_cases(53, 12)):
        keep = {k: np.ascontiguousarray(d[k], dtype=t) for k, t in T.PILEUP_IN_FIELDS}
        inp = T.PileupInput(d["n_reads"], 0, d["min_bq"], d["min_sv_len"], *[keep[k].ctypes.data for k, _ in T.PILEUP_IN_FIELDS])
        cap = int(np.asarray(d["n_digar"][:d["n_reads"]]).sum()) + 8
        pos, typ, rl, al, src = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int64)
        out = T.SitesOutput(pos.ctypes.data, typ.ctypes.data, rl.ctypes.data, al.ctypes.data, src.ctypes.data, cap, 0)
        assert oracle.lcd_oracle_collect_sites(T.C.byref(inp), T.C.c_int64(b), T.C.c_int64(e), T.C.byref(out)) == 0
        ns = int(out.n_sites)
        st = dict(n_sites=ns, site_pos=pos[:ns], site_type=typ[:ns], site_ref_len=rl[:ns], site_alt_len=al[:ns], site_src=src[:ns])
        sl = synth.site_list_from_sites(d, st, min_sv_len=d["min_sv_len"])
        want = T.sites_view(d, pos, typ, rl, al, src, ns)
        got = [(int(sl["site_pos"][i]), int(sl["site_type"][i]), int(sl["site_ref_len"][i]), int(sl["site_alt_len"][i]),
                bytes(sl["site_alt"][int(sl["site_alt_off"][i]):int(sl["site_alt_off"][i]) + (0 if sl["site_type"][i] == 2 else int(sl["site_alt_len"][i]))])) for i in range(ns)]
        assert got == want, n
        assert sl["n_sites"] == ns and len(sl["site_pos"]) == ns + 1
