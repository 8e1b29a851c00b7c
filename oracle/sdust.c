/* oracle/sdust.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Symmetric DUST as the reference vendors it (src/sdust.c, H. Li's implementation of Morgulis et al. 2006), which the chunk loader runs over the chunk's
 * reference window to fill chunk->low_comp_cr (src/bam_utils.c:1574-1583, T = LONGCALLD_SDUST_T = 5, W = LONGCALLD_SDUST_W = 20): the word window, its
 * running triplet counts, the longest suffix whose words all stay at or below 2T / 10 occurrences, the list of perfect intervals and the merged output
 * intervals, restated with flat arrays (the deque is an array indexed from a moving front).  Pinned against the unmodified sdust() in
 * tests/test_oracle_sdust.py. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

#define WLEN 3
#define WTOT 64
#define WMSK 63
typedef struct { int start, finish, r, l; } perf_t;

static int nt4(uint8_t c) { if (c < 4) return c;      /* seq_nt4_table maps the codes 0..3 to themselves too */
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }

/* beg / end: 0-based half-open output intervals (r[i] >> 32, (uint32_t)r[i] of sdust()); returns their number, or -1 when they exceed cap */
int lcd_oracle_sdust(const uint8_t *seq, int l_seq, int T, int W, int64_t *beg, int64_t *end, int64_t cap) {
    int *w = (int*)malloc(sizeof(int) * ((size_t)l_seq + 8)); int front = 0, cnt = 0;            /* the window's words: w[front .. front + cnt) */
    perf_t *P = (perf_t*)malloc(sizeof(perf_t) * ((size_t)W + 8) * 4); int np = 0, mp = (W + 8) * 4;
    int cv[WTOT], cw[WTOT], rv = 0, rw = 0, L = 0, l = 0; unsigned t = 0; int64_t n_res = 0; int over = 0;
    memset(cv, 0, sizeof(cv)); memset(cw, 0, sizeof(cw));
#define SAVE(start_) do { /* save_masked_regions :87-103 */ \
        if (np > 0 && P[np - 1].start < (start_)) { \
            const perf_t *p_ = &P[np - 1]; int saved_ = 0; \
            if (n_res) { const int64_t f_ = end[n_res - 1]; if (p_->start <= f_) { saved_ = 1; if (p_->finish > f_) end[n_res - 1] = p_->finish; } } \
            if (!saved_) { if (n_res < cap) { beg[n_res] = p_->start; end[n_res] = p_->finish; n_res++; } else over = 1; } \
            int i_ = np - 1; while (i_ >= 0 && P[i_].start < (start_)) --i_; \
            np = i_ + 1; \
        } } while (0)
    for (int i = 0; i <= l_seq; ++i) {
        const int b = i < l_seq ? nt4(seq[i]) : 4;
        if (b < 4) {
            ++l; t = (t << 2 | (unsigned)b) & WMSK;
            if (l >= WLEN) {
                const int start = (l - W > 0 ? l - W : 0) + (i + 1 - l);
                SAVE(start);
                /* shift_window :66-85 */
                if (cnt >= W - WLEN + 1) {
                    const int s = w[front]; ++front; --cnt;
                    rw -= --cw[s];
                    if (L > cnt) { --L; rv -= --cv[s]; }
                }
                w[front + cnt] = (int)t; ++cnt;
                ++L;
                rw += cw[t]++;
                rv += cv[t]++;
                if (cv[t] * 10 > T << 1) {
                    int s;
                    do { s = w[front + cnt - L]; rv -= --cv[s]; --L; } while (s != (int)t);
                }
                if (rw * 10 > L * T) {
                    /* find_perfect :105-131 */
                    int c[WTOT], r = rv, max_r = 0, max_l = 0;
                    memcpy(c, cv, sizeof(c));
                    for (int k = cnt - L - 1; k >= 0; --k) {
                        const int tt = w[front + k];
                        r += c[tt]++;
                        const int new_r = r, new_l = cnt - k - 1;
                        if (new_r * 10 > T * new_l) {
                            int j;
                            for (j = 0; j < np && P[j].start >= k + start; ++j) {
                                const perf_t *p = &P[j];
                                if (max_r == 0 || p->r * max_l > max_r * p->l) { max_r = p->r; max_l = p->l; }
                            }
                            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                                max_r = new_r; max_l = new_l;
                                if (np == mp) { mp *= 2; P = (perf_t*)realloc(P, sizeof(perf_t) * (size_t)mp); }
                                memmove(&P[j + 1], &P[j], (size_t)(np - j) * sizeof(perf_t));
                                ++np;
                                P[j].start = k + start; P[j].finish = cnt + (WLEN - 1) + start; P[j].r = new_r; P[j].l = new_l;
                            }
                        }
                    }
                }
            }
        } else {
            int start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            while (np) { SAVE(start); ++start; }
            l = 0; t = 0;
        }
    }
#undef SAVE
    free(w); free(P);
    return over ? -1 : (int)n_res;
}
