// classify_device.cuh -- device-side logic of K2b: the category of every candidate site, replacing classify_var_cate (reference
// src/collect_var.c:413-432) as the first loop of classify_cand_vars (:915-918) calls it, with var_is_homopolymer (:306-358) and
// var_is_repeat_region (:361-400).
//
// B200 design: ONE THREAD PER SITE over all chunks of a batch (~870 k sites per 50 Mb).  A site reads its 8 counters (one 32-byte
// record, K2's output where it lies), its type / lengths and -- for the ~10 % of sites that are small indels between the allele-fraction
// thresholds -- up to 36 reference bases around it (L2 resident: neighbouring sites share them) and its inserted bases.  The double
// division of the allele fraction is IEEE on both sides, so the thresholds cut exactly where the reference's do.
// ONT chunks add the strand-bias test (a two-tailed Fisher exact test on the strand counts of the alternative allele).
// The file compiles for the host as well (tests/emu).
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace classify {

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
enum { NON_VAR = 0x800, LOW_COV_VAR = 0x001, STRAND_BIAS_VAR = 0x002, LOW_AF_VAR = 0x400, CLEAN_HET_SNP = 0x004, CLEAN_HET_INDEL = 0x008, REP_HET_VAR = 0x010, CLEAN_HOM_VAR = 0x080 };

struct __align__(16) Chunk {
    int32_t min_dp, min_alt_dp, max_xgaps, is_ont;
    double min_af, max_af;
    long long ref_beg, ref_end;       // chunk->ref_beg / ref_end
    long long ref_off;                // first byte of the chunk's reference window in the concatenated ref array
    long long alt_base;               // first site_alt byte of the chunk (site_alt_off is relative to it)
};

struct KernelArgs {
    const Chunk *chunks; long long n_sites_total;
    const int32_t *site_chunk;
    const long long *site_pos; const int32_t *site_type, *site_ref_len, *site_alt_len; const long long *site_alt_off; const uint8_t *site_alt;
    const int32_t *site_counts;       // [n_sites_total][8]
    const char *ref;
    int32_t *var_cate;
    const double *lgamma_cache;       // lgamma(0 .. LGAMMA_MAX_I) as the host's libm returns them (opt->lgamma_cache, src/math_utils.c:6-11)
    int32_t *status;                  // optional: set to 1 when a small indel lies within REF_MARGIN bases of its reference window's ends
};
constexpr int REF_MARGIN = 24;       // bases of reference the context tests may read beyond a site (the reference reads them unchecked)

__device__ __forceinline__ int nt4(char c) {                              // nst_nt4_table (src/seq.c): upper / lower case ACGT -> 0..3, anything else 4
    const int u = c & 0xdf;
    return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4;
}

// a unit of 1..6 bases starting at `at` and running in direction dir (+1 / -1) is followed by two more copies of itself
__device__ __forceinline__ bool unit_repeats(const char *ref, long long at, int dir) {
    int unit[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) unit[k] = nt4(ref[at + dir * k]);
    for (int len = 1; len <= 6; ++len) {
        bool hp = true;
        for (int c = 1; c < 3 && hp; ++c)
            for (int j = 0; j < len; ++j) if (nt4(ref[at + dir * (c * len + j)]) != unit[j]) { hp = false; break; }
        if (hp) return true;
    }
    return false;
}

// var_is_homopolymer, src/collect_var.c:306-358 (called for insertions and deletions only)
__device__ __forceinline__ bool is_homopolymer(const Chunk &ch, const char *ref, long long pos, int type, int ref_len, int alt_len) {
    long long start_pos, end_pos;
    if (type == CINS) { if (alt_len > ch.max_xgaps) return false; start_pos = pos - 1; end_pos = pos; }
    else { if (ref_len > ch.max_xgaps) return false; start_pos = pos + ref_len - 1; end_pos = pos; }
    return unit_repeats(ref, end_pos - ch.ref_beg, 1) || unit_repeats(ref, start_pos - ch.ref_beg, -1);
}

// var_is_repeat_region, src/collect_var.c:361-400
__device__ __forceinline__ bool is_repeat_region(const Chunk &ch, const char *ref, long long pos, int type, int ref_len, int alt_len, const uint8_t *alt) {
    const char *r = ref + (pos - ch.ref_beg);
    if (type == CDEL) {
        if (ref_len > ch.max_xgaps) return false;
        const int len = ref_len * 3;
        if (pos < ch.ref_beg || pos + ref_len + len >= ch.ref_end) return false;
        for (int k = 0; k < len; ++k) if (nt4(r[k]) != nt4(r[ref_len + k])) return false;
        return true;
    }
    if (alt_len > ch.max_xgaps) return false;
    const int len = alt_len * 3;
    if (pos < ch.ref_beg || pos + len >= ch.ref_end) return false;
    // the reference compares the window with: the inserted bases, then twice the window's own first unit
    for (int k = 0; k < len; ++k) {
        const int a = k < alt_len ? (int)alt[k] : nt4(r[(k - alt_len) % alt_len]);
        if (nt4(r[k]) != a) return false;
    }
    return true;
}

// ---- ONT strand bias: var_is_strand_bias (src/collect_var.c:270-284) -> fisher_exact_test (src/math_utils.c:119-170), two-tailed, on
// the table {forward alt, reverse alt, expected, expected}.  Double arithmetic in the reference's order of operations; lgamma comes from
// the host-built cache (bit-identical to the reference's), so only exp() can differ from the host's libm -- by an ulp of a p-value that
// is then summed, rounded to float and compared with 0.01f.
constexpr int LGAMMA_MAX_I = 500;                                            // LONGCALLD_LGAMMA_MAX_I, src/call_var_main.h:86
constexpr float STRAND_BIAS_PVAL_ONT = 0.01f;                                // LONGCALLD_STRAND_BIAS_PVAL_ONT, src/call_var_main.h:74
__device__ __forceinline__ double fast_lgamma(const double *cache, int x) { return (x >= 0 && x <= LGAMMA_MAX_I) ? cache[x] : lgamma((double)x); }
__device__ inline double log_hypergeometric(const double *lg, int a, int b, int c, int d) {       // src/math_utils.c:101-116 (the recursion unrolled)
    for (;;) {
        const int n1 = a + b, n2 = c + d, m1 = a + c, m2 = b + d;
        if (n1 > n2) { int t = a; a = c; c = t; t = b; b = d; d = t; continue; }                  // -> (c, d, a, b)
        if (m1 > m2) { int t = a; a = b; b = t; t = c; c = d; d = t; continue; }                  // -> (b, a, d, c)
        const int N = n1 + n2;
        return fast_lgamma(lg, n1 + 1) + fast_lgamma(lg, n2 + 1) + fast_lgamma(lg, m1 + 1) + fast_lgamma(lg, m2 + 1) -
               (fast_lgamma(lg, a + 1) + fast_lgamma(lg, b + 1) + fast_lgamma(lg, c + 1) + fast_lgamma(lg, d + 1) + fast_lgamma(lg, N + 1));
    }
}
__device__ inline double fisher_exact_test(const double *lg, int a, int b, int c, int d) {        // src/math_utils.c:119-170
    const double p_observed = exp(log_hypergeometric(lg, a, b, c, d));
    double total_p = 0.0;
    const int min_a = (0 > (a + c) - (a + b + c + d)) ? 0 : (a + c) - (b + d);
    const int max_a = (a + b) < (a + c) ? (a + b) : (a + c);
    const int mode_a = (int)((a + b) * (a + c) / (double)(a + b + c + d));
    for (int delta = 0; delta <= max_a - min_a; delta++) {
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && delta == 0) continue;
            const int ca = side == 0 ? mode_a + delta : mode_a - delta;
            if (side == 0 ? ca > max_a : ca < min_a) continue;
            const int cb = (a + b) - ca, cc = (a + c) - ca, cd = (b + d) - cb;
            if (cb >= 0 && cc >= 0 && cd >= 0) {
                const double p = exp(log_hypergeometric(lg, ca, cb, cc, cd));
                if (p <= p_observed + 2.2204460492503131e-16) total_p += p;
            }
        }
    }
    return total_p;
}
__device__ inline bool is_strand_bias(const double *lg, int for_alt_cov, int rev_alt_cov) {
    const int expected = (for_alt_cov + rev_alt_cov) / 2;
    if (expected == 0) return false;
    const float fisher_p = (float)fisher_exact_test(lg, for_alt_cov, rev_alt_cov, expected, expected);
    return fisher_p < STRAND_BIAS_PVAL_ONT;
}

__device__ void classify_site(const KernelArgs &a, long long s) {
    const Chunk ch = a.chunks[a.site_chunk[s]];
    const int4 c = *reinterpret_cast<const int4 *>(a.site_counts + 8 * s);        // total_cov, low_qual_cov, alle_covs[0], alle_covs[1]
    const int total_cov = c.x, low_qual_cov = c.y, alt_dp = c.w, type = a.site_type[s];
    int cate;
    if (total_cov + low_qual_cov < ch.min_dp) cate = LOW_COV_VAR;
    else {
        const double alt_af = (double)alt_dp / total_cov;
        if (alt_dp < ch.min_alt_dp) cate = LOW_COV_VAR;
        else if (ch.is_ont && is_strand_bias(a.lgamma_cache, a.site_counts[8 * s + 5], a.site_counts[8 * s + 7])) cate = STRAND_BIAS_VAR;   // strand_to_alle_covs[0][1], [1][1]
        else if (alt_af < ch.min_af) cate = LOW_AF_VAR;
        else if (alt_af > ch.max_af) cate = CLEAN_HOM_VAR;
        else if (type == CDIFF) cate = CLEAN_HET_SNP;
        else {
            const char *ref = a.ref + ch.ref_off; const long long pos = a.site_pos[s]; const int rl = a.site_ref_len[s], al = a.site_alt_len[s];
            const int len = type == CINS ? al : rl;
            if (a.status && len <= ch.max_xgaps && (pos - REF_MARGIN < ch.ref_beg || pos + len + REF_MARGIN > ch.ref_end)) { *a.status = 1; a.var_cate[s] = NON_VAR; return; }
            const bool rep = is_homopolymer(ch, ref, pos, type, rl, al) || is_repeat_region(ch, ref, pos, type, rl, al, a.site_alt + ch.alt_base + a.site_alt_off[s]);
            cate = rep ? REP_HET_VAR : CLEAN_HET_INDEL;
        }
    }
    a.var_cate[s] = cate;
}

} // namespace classify
} // namespace lcd
