/* oracle/classify.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the per-site category of longcallD's pileup scan: classify_var_cate (reference src/collect_var.c:413-432)
 * as called by the first loop of classify_cand_vars (:902-925), with var_is_homopolymer (:306-358: a 1-6 bp unit repeated three
 * times right of the variant, or left of it) and var_is_repeat_region (:361-400: the deleted / inserted bases repeated three times
 * in the reference).  ONT's strand-bias test (var_is_strand_bias :270 -> fisher_exact_test, src/math_utils.c:119, with the
 * lgamma cache of :6-17) is restated below.  Pinned against the unmodified reference (oracle/_ref/libref_shim.so: ref_classify_sites) in
 * tests/test_oracle_classify.py.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include "lcd_oracle.h"

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
enum { NON_VAR = 0x800, LOW_COV_VAR = 0x001, STRAND_BIAS_VAR = 0x002, LOW_AF_VAR = 0x400, CLEAN_HET_SNP = 0x004, CLEAN_HET_INDEL = 0x008,
       REP_HET_VAR = 0x010, CLEAN_HOM_VAR = 0x080 };

static int nt4(char c) {                                                  /* nst_nt4_table, src/seq.c */
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

/* var_is_homopolymer, src/collect_var.c:306-358 */
static int is_homopolymer(const lcd_classify_input_t *in, int i) {
    const int64_t pos = in->site_pos[i]; const int type = in->site_type[i];
    int64_t start_pos, end_pos;
    if (type == CDIFF) { start_pos = pos - 1; end_pos = pos + 1; }
    else if (type == CINS) { if (in->site_alt_len[i] > in->max_xgaps) return 0; start_pos = pos - 1; end_pos = pos; }
    else { if (in->site_ref_len[i] > in->max_xgaps) return 0; start_pos = pos + in->site_ref_len[i] - 1; end_pos = pos; }
    const char *ref = in->ref_seq; const int64_t rb = in->ref_beg;
    int unit[6], hp = 1;
    for (int k = 0; k < 6; ++k) unit[k] = nt4(ref[end_pos + k - rb]);
    for (int len = 1; len <= 6; ++len) {
        hp = 1;
        for (int c = 1; c < 3 && hp; ++c)
            for (int j = 0; j < len; ++j) if (nt4(ref[end_pos - rb + c * len + j]) != unit[j]) { hp = 0; break; }
        if (hp) break;
    }
    if (hp) return 1;
    for (int k = 0; k < 6; ++k) unit[k] = nt4(ref[start_pos - rb - k]);
    for (int len = 1; len <= 6; ++len) {
        hp = 1;
        for (int c = 1; c < 3 && hp; ++c)
            for (int j = 0; j < len; ++j) if (nt4(ref[start_pos - rb - c * len - j]) != unit[j]) { hp = 0; break; }
        if (hp) break;
    }
    return hp;
}

/* var_is_repeat_region, src/collect_var.c:361-400 */
static int is_repeat_region(const lcd_classify_input_t *in, int i) {
    const int64_t pos = in->site_pos[i]; const char *ref = in->ref_seq + (pos - in->ref_beg);
    if (in->site_type[i] == CDEL) {
        const int del_len = in->site_ref_len[i];
        if (del_len > in->max_xgaps) return 0;
        const int len = del_len * 3;
        if (pos < in->ref_beg || pos + del_len + len >= in->ref_end) return 0;
        for (int k = 0; k < len; ++k) if (nt4(ref[k]) != nt4(ref[del_len + k])) return 0;
        return 1;
    }
    const int ins_len = in->site_alt_len[i];
    if (ins_len > in->max_xgaps) return 0;
    const int len = ins_len * 3;
    if (pos < in->ref_beg || pos + len >= in->ref_end) return 0;
    const uint8_t *alt = in->site_alt + in->site_alt_off[i];
    /* the reference builds alt_bseq = ref window, shifts it right by ins_len in place (so the first unit ends up repeated) and
       overwrites the first unit with the inserted bases: inserted bases, then twice the reference's first unit */
    for (int k = 0; k < len; ++k) {
        const int a = k < ins_len ? alt[k] : nt4(ref[(k - ins_len) % ins_len]);
        if (nt4(ref[k]) != a) return 0;
    }
    return 1;
}

/* fast_lgamma / log_hypergeometric / fisher_exact_test, src/math_utils.c:13-18,101-170 */
#define LGAMMA_MAX_I 500
static double lgamma_cache[LGAMMA_MAX_I + 1]; static int lgamma_ready = 0;
static double fast_lgamma(int x) { return (x >= 0 && x <= LGAMMA_MAX_I) ? lgamma_cache[x] : lgamma(x); }
static double log_hypergeometric(int a, int b, int c, int d) {
    const int n1 = a + b, n2 = c + d, m1 = a + c, m2 = b + d, N = n1 + n2;
    if (n1 > n2) return log_hypergeometric(c, d, a, b);
    if (m1 > m2) return log_hypergeometric(b, a, d, c);
    return fast_lgamma(n1 + 1) + fast_lgamma(n2 + 1) + fast_lgamma(m1 + 1) + fast_lgamma(m2 + 1) -
           (fast_lgamma(a + 1) + fast_lgamma(b + 1) + fast_lgamma(c + 1) + fast_lgamma(d + 1) + fast_lgamma(N + 1));
}
static double fisher_exact_test(int a, int b, int c, int d) {
    const double p_observed = exp(log_hypergeometric(a, b, c, d));
    double total_p = 0.0;
    const int min_a = (0 > (a + c) - (a + b + c + d)) ? 0 : (a + c) - (b + d);
    const int max_a = (a + b) < (a + c) ? (a + b) : (a + c);
    const int mode_a = (int)((a + b) * (a + c) / (double)(a + b + c + d));
    for (int delta = 0; delta <= max_a - min_a; delta++) {
        int ca = mode_a + delta;
        if (ca <= max_a) {
            const int cb = (a + b) - ca, cc = (a + c) - ca, cd = (b + d) - cb;
            if (cb >= 0 && cc >= 0 && cd >= 0) { const double p = exp(log_hypergeometric(ca, cb, cc, cd)); if (p <= p_observed + DBL_EPSILON) total_p += p; }
        }
        if (delta > 0) {
            ca = mode_a - delta;
            if (ca >= min_a) {
                const int cb = (a + b) - ca, cc = (a + c) - ca, cd = (b + d) - cb;
                if (cb >= 0 && cc >= 0 && cd >= 0) { const double p = exp(log_hypergeometric(ca, cb, cc, cd)); if (p <= p_observed + DBL_EPSILON) total_p += p; }
            }
        }
    }
    return total_p;
}
/* var_is_strand_bias, src/collect_var.c:270-284 (strand_bias_pval = LONGCALLD_STRAND_BIAS_PVAL_ONT, src/call_var_main.h:74) */
static int is_strand_bias(int for_alt_cov, int rev_alt_cov) {
    const int expected = (for_alt_cov + rev_alt_cov) / 2;
    if (expected == 0) return 0;
    const float fisher_p = fisher_exact_test(for_alt_cov, rev_alt_cov, expected, expected);
    const float pval = 0.01;
    return fisher_p < pval;
}

int lcd_oracle_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate) {
    if (!lgamma_ready) { for (int i = 0; i <= LGAMMA_MAX_I; ++i) lgamma_cache[i] = lgamma(i); lgamma_ready = 1; }
    for (int i = 0; i < in->n_sites; ++i) {
        const int32_t *c = in->site_counts + 8 * (int64_t)i;
        const int total_cov = c[0], low_qual_cov = c[1], alt_dp = c[3], type = in->site_type[i];
        int cate;
        if (total_cov + low_qual_cov < in->min_dp) cate = LOW_COV_VAR;
        else {
            const double alt_af = (double)alt_dp / total_cov;
            if (alt_dp < in->min_alt_dp) cate = LOW_COV_VAR;
            else if (in->is_ont && is_strand_bias(c[5], c[7])) cate = STRAND_BIAS_VAR;
            else if (alt_af < in->min_af) cate = LOW_AF_VAR;
            else if (alt_af > in->max_af) cate = CLEAN_HOM_VAR;
            else if ((type == CINS || type == CDEL) && (is_homopolymer(in, i) || is_repeat_region(in, i))) cate = REP_HET_VAR;
            else cate = type == CDIFF ? CLEAN_HET_SNP : CLEAN_HET_INDEL;
        }
        var_cate[i] = cate;
    }
    return 0;
}
