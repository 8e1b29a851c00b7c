"""CPU: pins oracle/classify.c (per-site category: classify_var_cate, reference src/collect_var.c:413, with var_is_homopolymer :306 and
var_is_repeat_region :361) against the unmodified reference (oracle/_ref/libref_shim.so: ref_classify_sites builds cand_var_t records
around the same flat arrays and calls classify_var_cate as the first loop of classify_cand_vars does)."""
import collections

import numpy as np

import lcd_testlib as T
from longcalld_b200 import synth


def classify_cases(seed, n, is_ont=0):
    rng = np.random.default_rng(seed)
    for it in range(n):
        yield synth.make_classify_chunk(rng, ref_len=int(rng.choice([600, 4000, 20000])), n_sites=int(rng.choice([1, 40, 400])),
                                        max_xgaps=int(rng.choice([5, 5, 3, 8])), ref0=int(rng.choice([1, 100000])), is_ont=is_ont)


def test_oracle_vs_live_reference(oracle, ref):
    seen = collections.Counter()
    for n, d in enumerate(classify_cases(61, 150)):
        a = T.classify(oracle, "lcd_oracle_classify_sites", d)
        r = T.classify(ref, "ref_classify_sites", d)
        assert np.array_equal(a, r), (n, np.nonzero(a != r)[0][:5], a[a != r][:5], r[a != r][:5])
        seen.update(a.tolist())
    # every category the HiFi path can return, the context tests included
    assert all(seen[c] > 100 for c in (0x001, 0x400, 0x080, 0x010, 0x004, 0x008)), seen


def test_oracle_vs_live_reference_ont(oracle, ref):
    """ONT chunks: var_is_strand_bias (src/collect_var.c:270) -> fisher_exact_test with the lgamma cache (src/math_utils.c:6-170); the
    `p < 0.01` decision of every site is compared (deep sites beyond the cache's 500 entries included)."""
    seen = collections.Counter()
    for n, d in enumerate(classify_cases(67, 120, is_ont=1)):
        if n % 10 == 0:          # depth beyond the lgamma cache (LONGCALLD_LGAMMA_MAX_I = 500): lgamma() itself
            d["site_counts"][:, :] *= 9
        a = T.classify(oracle, "lcd_oracle_classify_sites", d)
        r = T.classify(ref, "ref_classify_sites", d)
        assert np.array_equal(a, r), (n, np.nonzero(a != r)[0][:5], a[a != r][:5], r[a != r][:5])
        seen.update(a.tolist())
    assert seen[0x002] > 500 and all(seen[c] > 100 for c in (0x001, 0x400, 0x080, 0x010, 0x004, 0x008)), seen


def test_oracle_vs_reference_fixtures(oracle):
    g = T.load_golden("classify_lcd")
    assert len(g["cases"]) >= 12
    for c in g["cases"]:
        d = T.classify_case_from_json(c["in"])
        assert T.classify(oracle, "lcd_oracle_classify_sites", d).tolist() == c["cate"]
