/* oracle/wfa.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar restatement of WFA2-lib (vendored @42f8bf3) gap-affine / gap-affine-2p end-to-end
 * alignment in memory_high mode, exactly as longcallD configures it (src/align.c:374-460):
 * match = 0, heuristic in {none, wf-adaptive, z-drop}, full (non-piggyback) backtrace.
 *
 * Reference call graph restated here:
 *   wavefront_unialign              WFA2-lib/wavefront/wavefront_unialign.c:242-275
 *   wavefront_extend_end2end        wavefront_extend.c:90-128
 *   wavefront_termination_end2end   wavefront_termination.c:37-66
 *   wavefront_heuristic_cufoff      wavefront_heuristic.c:509-570 (+wfadaptive :257, zdrop :400)
 *   wavefront_compute_affine2p      wavefront_compute_affine2p.c:334-369 (kernels :45-106,
 *                                   wavefront_compute_affine.c:45-80)
 *   limits / allocate / trim        wavefront_compute.c:40-86, 407-494, 579-613
 *   wavefront_backtrace_affine      wavefront_backtrace.c:320-539
 *   wavefront_unialign_terminate    wavefront_unialign.c:146-236
 *   cigar_maxtrim_gap_affine[2p]    alignment/cigar.c:476-600
 *
 * A wavefront component is stored as [lo,hi] + offsets; the reference guarantees (init_ends,
 * wavefront_compute.c:528-578) that every read outside [lo,hi] yields WAVEFRONT_OFFSET_NULL, and
 * that a NULL / ->null wavefront reads as all-NULL, which is what wf_get() states directly.
 */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "lcd_oracle.h"

#define OFF_NULL (INT32_MIN/2)           /* wavefront_offset.h:44 */
#define DIAG_NULL INT_MAX                /* wavefront_offset.h:54 */
#define MAX2(a,b) ((a)>(b)?(a):(b))
#define MIN2(a,b) ((a)<(b)?(a):(b))

enum { C_M = 0, C_I1, C_D1, C_I2, C_D2, N_COMP };

typedef struct {
    int exists;      /* pointer != NULL in the reference */
    int lo, hi;      /* current limits (after trim / heuristics) */
    int base, n;     /* off[0] is diagonal `base`; n allocated */
    int32_t *off;
} wf_t;
typedef struct { wf_t c[N_COMP]; } wfset_t;

typedef struct {
    const uint8_t *pattern, *text; int plen, tlen;
    lcd_wfa_params_t par;
    int n_comp, max_score_scope;
    wfset_t *sets; int n_sets, m_sets;
    int num_null_steps;
    /* heuristic state (wavefront_heuristic.h) */
    int steps_wait, max_sw_score, max_wf_score, max_sw_score_k, max_sw_score_offset;
    /* alignment end */
    int end_score, end_k, end_offset;
} wfa_t;

static int wf_is_null(const wfa_t *a, int comp, int s) {
    if (s < 0 || s >= a->n_sets) return 1;
    const wf_t *w = &a->sets[s].c[comp];
    return !w->exists || w->lo > w->hi;
}
static inline int32_t wf_get(const wfa_t *a, int comp, int s, int k) {
    if (s < 0 || s >= a->n_sets) return OFF_NULL;
    const wf_t *w = &a->sets[s].c[comp];
    if (!w->exists || w->lo > w->hi || k < w->lo || k > w->hi) return OFF_NULL;
    return w->off[k - w->base];
}
static wfset_t *wfa_new_set(wfa_t *a, int s) {
    while (s >= a->m_sets) {
        int m = a->m_sets ? a->m_sets * 2 : 64;
        a->sets = (wfset_t*)realloc(a->sets, (size_t)m * sizeof(wfset_t));
        memset(a->sets + a->m_sets, 0, (size_t)(m - a->m_sets) * sizeof(wfset_t));
        a->m_sets = m;
    }
    if (s >= a->n_sets) a->n_sets = s + 1;
    return &a->sets[s];
}
static void wf_alloc(wf_t *w, int lo, int hi) {
    w->exists = 1; w->lo = lo; w->hi = hi; w->base = lo; w->n = hi - lo + 1;
    w->off = (int32_t*)malloc((size_t)MAX2(w->n, 1) * sizeof(int32_t));
}
/* wavefront_compute_trim_ends, wavefront_compute.c:579-613 */
static void wf_trim(const wfa_t *a, wf_t *w) {
    int k;
    for (k = w->hi; k >= w->lo; --k) {
        int32_t off = w->off[k - w->base];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
    }
    w->hi = k;
    for (k = w->lo; k <= w->hi; ++k) {
        int32_t off = w->off[k - w->base];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= (uint32_t)a->tlen && v <= (uint32_t)a->plen) break;
    }
    w->lo = k;
}

/* wavefront_compute_affine2p, wavefront_compute_affine2p.c:334-369 (and the 1p twin,
 * wavefront_compute_affine.c:210-240) */
static void wfa_compute(wfa_t *a, int score) {
    const lcd_wfa_params_t *p = &a->par;
    const int two = p->affine2p;
    const int s_x = score - p->mismatch;
    const int s_o1 = score - p->gap_open1 - p->gap_ext1, s_e1 = score - p->gap_ext1;
    const int s_o2 = score - p->gap_open2 - p->gap_ext2, s_e2 = score - p->gap_ext2;
    const int n_misms = wf_is_null(a, C_M, s_x), n_open1 = wf_is_null(a, C_M, s_o1);
    const int n_i1 = wf_is_null(a, C_I1, s_e1), n_d1 = wf_is_null(a, C_D1, s_e1);
    const int n_open2 = two ? wf_is_null(a, C_M, s_o2) : 1;
    const int n_i2 = two ? wf_is_null(a, C_I2, s_e2) : 1, n_d2 = two ? wf_is_null(a, C_D2, s_e2) : 1;
    wfset_t *out = wfa_new_set(a, score); /* may realloc a->sets: fetch inputs afterwards via wf_get */
    memset(out, 0, sizeof(*out));
    if (n_misms && n_open1 && n_i1 && n_d1 && n_open2 && n_i2 && n_d2) {
        a->num_null_steps++;              /* wavefront_compute_allocate_output_null */
        return;
    }
    a->num_null_steps = 0;
    /* wavefront_compute_limits_input, wavefront_compute.c:40-86; a null wavefront has lo=1,hi=-1 */
#define LO(comp,s,isnull) ((isnull) ? 1 : a->sets[s].c[comp].lo)
#define HI(comp,s,isnull) ((isnull) ? -1 : a->sets[s].c[comp].hi)
    int lo = LO(C_M, s_x, n_misms), hi = HI(C_M, s_x, n_misms);
    lo = MIN2(lo, LO(C_M, s_o1, n_open1) - 1);  hi = MAX2(hi, HI(C_M, s_o1, n_open1) + 1);
    lo = MIN2(lo, LO(C_I1, s_e1, n_i1) + 1);    hi = MAX2(hi, HI(C_I1, s_e1, n_i1) + 1);
    lo = MIN2(lo, LO(C_D1, s_e1, n_d1) - 1);    hi = MAX2(hi, HI(C_D1, s_e1, n_d1) - 1);
    if (two) {
        lo = MIN2(lo, LO(C_M, s_o2, n_open2) - 1);  hi = MAX2(hi, HI(C_M, s_o2, n_open2) + 1);
        lo = MIN2(lo, LO(C_I2, s_e2, n_i2) + 1);    hi = MAX2(hi, HI(C_I2, s_e2, n_i2) + 1);
        lo = MIN2(lo, LO(C_D2, s_e2, n_d2) - 1);    hi = MAX2(hi, HI(C_D2, s_e2, n_d2) - 1);
    }
#undef LO
#undef HI
    /* wavefront_compute_allocate_output, wavefront_compute.c:407-494 */
    wf_alloc(&out->c[C_M], lo, hi);
    if (!n_open1 || !n_i1) wf_alloc(&out->c[C_I1], lo, hi);
    if (!n_open1 || !n_d1) wf_alloc(&out->c[C_D1], lo, hi);
    if (two && (!n_open2 || !n_i2)) wf_alloc(&out->c[C_I2], lo, hi);
    if (two && (!n_open2 || !n_d2)) wf_alloc(&out->c[C_D2], lo, hi);
    const int full2p = two && !(n_open2 && n_i2 && n_d2); /* dispatcher, affine2p.c:282-300 */
    for (int k = lo; k <= hi; ++k) {
        int32_t ins1 = MAX2(wf_get(a, C_M, s_o1, k-1), wf_get(a, C_I1, s_e1, k-1)) + 1;
        int32_t del1 = MAX2(wf_get(a, C_M, s_o1, k+1), wf_get(a, C_D1, s_e1, k+1));
        int32_t misms = wf_get(a, C_M, s_x, k) + 1;
        int32_t max;
        if (out->c[C_I1].exists) out->c[C_I1].off[k - lo] = ins1;
        if (out->c[C_D1].exists) out->c[C_D1].off[k - lo] = del1;
        if (full2p) {
            int32_t ins2 = MAX2(wf_get(a, C_M, s_o2, k-1), wf_get(a, C_I2, s_e2, k-1)) + 1;
            int32_t del2 = MAX2(wf_get(a, C_M, s_o2, k+1), wf_get(a, C_D2, s_e2, k+1));
            if (out->c[C_I2].exists) out->c[C_I2].off[k - lo] = ins2;
            if (out->c[C_D2].exists) out->c[C_D2].off[k - lo] = del2;
            int32_t ins = MAX2(ins1, ins2), del = MAX2(del1, del2);
            max = MAX2(del, MAX2(misms, ins));
        } else {
            max = MAX2(del1, MAX2(misms, ins1));
        }
        uint32_t h = (uint32_t)max, v = (uint32_t)(max - k);
        if (h > (uint32_t)a->tlen) max = OFF_NULL;
        if (v > (uint32_t)a->plen) max = OFF_NULL;
        out->c[C_M].off[k - lo] = max;
    }
    for (int c = 0; c < N_COMP; ++c) if (out->c[c].exists) wf_trim(a, &out->c[c]);
}

/* wavefront_heuristic_wfadaptive, wavefront_heuristic.c:232-292 */
static void heur_wfadaptive(wfa_t *a, wf_t *w) {
    if (a->steps_wait > 0) return;
    const int base_lo = w->lo, base_hi = w->hi;
    if (base_hi - base_lo + 1 < a->par.min_wavefront_length) return;
    int *dist = (int*)malloc((size_t)(base_hi - base_lo + 1) * sizeof(int));
    int min_distance = MAX2(a->plen, a->tlen);
    for (int k = base_lo; k <= base_hi; ++k) {
        int32_t off = w->off[k - w->base];
        int left_v = a->plen - (off - k), left_h = a->tlen - off;
        int d = (off >= 0) ? MAX2(left_v, left_h) : -OFF_NULL;
        dist[k - base_lo] = d;
        min_distance = MIN2(min_distance, d);
    }
    const int alignment_k = a->tlen - a->plen, thr = a->par.max_distance_threshold;
    const int top_limit = MIN2(alignment_k, w->hi);
    int lo_reduced = w->lo;
    for (int k = w->lo; k < top_limit; ++k) {
        if (dist[k - base_lo] - min_distance <= thr) break;
        ++lo_reduced;
    }
    w->lo = lo_reduced;
    const int bottom_limit = MAX2(alignment_k, w->lo);
    int hi_reduced = w->hi;
    for (int k = w->hi; k > bottom_limit; --k) {
        if (dist[k - base_lo] - min_distance <= thr) break;
        --hi_reduced;
    }
    w->hi = hi_reduced;
    free(dist);
    a->steps_wait = a->par.steps_between_cutoffs;
}
/* wavefront_heuristic_zdrop, wavefront_heuristic.c:400-452 (+ compute_sw_scores :297-331) */
static int heur_zdrop(wfa_t *a, wf_t *w, int score) {
    if (a->steps_wait > 0) return 0;
    const int swg_match = -1; /* penalties.match == 0 -> -1 (wavefront_heuristic.c:307) */
    int cmax = INT_MIN, cmax_k = 0, cmax_off = 0;
    for (int k = w->lo; k <= w->hi; ++k) {
        int32_t off = w->off[k - w->base];
        if (off < 0) continue;
        int v = off - k, h = off;
        int sw = (swg_match * (v + h) - score) / 2; /* WF_SCORE_TO_SW_SCORE, C truncation */
        if (cmax < sw) { cmax = sw; cmax_k = k; cmax_off = off; }
    }
    if (a->max_sw_score_k != DIAG_NULL) {
        if (cmax > a->max_sw_score) {
            a->max_sw_score = cmax; a->max_wf_score = score; a->max_sw_score_k = cmax_k; a->max_sw_score_offset = cmax_off;
        } else if (a->max_sw_score - cmax > a->par.zdrop) {
            a->end_score = a->max_wf_score; a->end_k = a->max_sw_score_k; a->end_offset = a->max_sw_score_offset;
            return 1;
        }
    } else {
        a->max_sw_score = cmax; a->max_wf_score = score; a->max_sw_score_k = cmax_k; a->max_sw_score_offset = cmax_off;
    }
    a->steps_wait = a->par.steps_between_cutoffs;
    return 0;
}
/* wavefront_heuristic_cufoff, wavefront_heuristic.c:509-570 */
static int heur_cutoff(wfa_t *a, int score) {
    wfset_t *set = &a->sets[score];
    wf_t *m = &set->c[C_M];
    if (!m->exists || m->lo > m->hi) return 0;
    --a->steps_wait;
    const int lo_base = m->lo, hi_base = m->hi;
    if (a->par.heuristic == LCD_WFA_HEUR_ADAPTIVE) heur_wfadaptive(a, m);
    else if (a->par.heuristic == LCD_WFA_HEUR_ZDROP) { if (heur_zdrop(a, m, score)) return 1; }
    if (lo_base == m->lo && hi_base == m->hi) return 0;
    for (int c = 1; c < N_COMP; ++c) {    /* wf_heuristic_equate :161-172 */
        wf_t *d = &set->c[c];
        if (!d->exists) continue;
        if (m->lo > d->lo) d->lo = m->lo;
        if (m->hi < d->hi) d->hi = m->hi;
    }
    return 0;
}

/* ---- backtrace, wavefront_backtrace.c:65-222 (candidates) and :320-539 (walk) ---------- */
enum { BT_I1_OPEN = 1, BT_I1_EXT, BT_I2_OPEN, BT_I2_EXT, BT_D1_OPEN, BT_D1_EXT, BT_D2_OPEN, BT_D2_EXT, BT_M };
static int64_t bt_cand(const wfa_t *a, int comp, int s, int k, int add, int type) {
    if (s < 0 || s >= a->n_sets) return OFF_NULL;
    const wf_t *w = &a->sets[s].c[comp];
    if (w->exists && w->lo <= k && k <= w->hi)
        return (int64_t)(((int64_t)(w->off[k - w->base] + add)) * 16) | type; /* <<4 | type (arithmetic) */
    return OFF_NULL;
}
typedef struct { char *ops; int max_ops, begin, end; int score, end_v, end_h; } cigar_t;

static void bt_push(cigar_t *c, char op) { if (c->begin >= 0 && c->begin < c->max_ops) c->ops[c->begin] = op; c->begin--; }
static void bt_matches(cigar_t *c, int n) {   /* wavefront_backtrace_matches :80-101 */
    int b = c->begin;
    c->begin -= n;
    for (int i = 0; i < n; ++i) { if (b - i >= 0 && b - i < c->max_ops) c->ops[b - i] = 'M'; }
}
static void wfa_backtrace(wfa_t *a, cigar_t *c, int alignment_score, int alignment_k, int alignment_offset) {
    const lcd_wfa_params_t *p = &a->par;
    c->end = c->max_ops - 1; c->begin = c->max_ops - 2;
    c->ops[c->end] = '\0';
    int matrix = C_M, score = alignment_score, k = alignment_k;
    int h = alignment_offset, v = alignment_offset - alignment_k, offset = alignment_offset;
    if (v < a->plen) for (int i = a->plen - v; i > 0; --i) bt_push(c, 'D');
    if (h < a->tlen) for (int i = a->tlen - h; i > 0; --i) bt_push(c, 'I');
    while (v > 0 && h > 0 && score > 0) {
        const int mismatch = score - p->mismatch;
        const int gap_open1 = score - p->gap_open1 - p->gap_ext1, gap_extend1 = score - p->gap_ext1;
        const int gap_open2 = score - p->gap_open2 - p->gap_ext2, gap_extend2 = score - p->gap_ext2;
        int64_t max_all;
        if (matrix == C_M) {
            int64_t misms = bt_cand(a, C_M, mismatch, k, 1, BT_M);
            int64_t i1o = bt_cand(a, C_M, gap_open1, k-1, 1, BT_I1_OPEN), i1e = bt_cand(a, C_I1, gap_extend1, k-1, 1, BT_I1_EXT);
            int64_t d1o = bt_cand(a, C_M, gap_open1, k+1, 0, BT_D1_OPEN), d1e = bt_cand(a, C_D1, gap_extend1, k+1, 0, BT_D1_EXT);
            int64_t mi = MAX2(i1o, i1e), md = MAX2(d1o, d1e);
            if (p->affine2p) {
                int64_t i2o = bt_cand(a, C_M, gap_open2, k-1, 1, BT_I2_OPEN), i2e = bt_cand(a, C_I2, gap_extend2, k-1, 1, BT_I2_EXT);
                int64_t d2o = bt_cand(a, C_M, gap_open2, k+1, 0, BT_D2_OPEN), d2e = bt_cand(a, C_D2, gap_extend2, k+1, 0, BT_D2_EXT);
                mi = MAX2(mi, MAX2(i2o, i2e)); md = MAX2(md, MAX2(d2o, d2e));
            }
            max_all = MAX2(misms, MAX2(mi, md));
        } else if (matrix == C_I1) max_all = MAX2(bt_cand(a, C_M, gap_open1, k-1, 1, BT_I1_OPEN), bt_cand(a, C_I1, gap_extend1, k-1, 1, BT_I1_EXT));
        else if (matrix == C_I2)   max_all = MAX2(bt_cand(a, C_M, gap_open2, k-1, 1, BT_I2_OPEN), bt_cand(a, C_I2, gap_extend2, k-1, 1, BT_I2_EXT));
        else if (matrix == C_D1)   max_all = MAX2(bt_cand(a, C_M, gap_open1, k+1, 0, BT_D1_OPEN), bt_cand(a, C_D1, gap_extend1, k+1, 0, BT_D1_EXT));
        else                       max_all = MAX2(bt_cand(a, C_M, gap_open2, k+1, 0, BT_D2_OPEN), bt_cand(a, C_D2, gap_extend2, k+1, 0, BT_D2_EXT));
        if (max_all < 0) break;
        if (matrix == C_M) {
            const int max_offset = (int)(max_all >> 4);
            bt_matches(c, offset - max_offset);
            offset = max_offset;
            v = offset - k; h = offset;
            if (v <= 0 || h <= 0) break;
        }
        const int type = (int)(max_all & 0xF);
        switch (type) {
            case BT_M:       score = mismatch;    matrix = C_M;  break;
            case BT_I1_OPEN: score = gap_open1;   matrix = C_M;  break;
            case BT_I1_EXT:  score = gap_extend1; matrix = C_I1; break;
            case BT_I2_OPEN: score = gap_open2;   matrix = C_M;  break;
            case BT_I2_EXT:  score = gap_extend2; matrix = C_I2; break;
            case BT_D1_OPEN: score = gap_open1;   matrix = C_M;  break;
            case BT_D1_EXT:  score = gap_extend1; matrix = C_D1; break;
            case BT_D2_OPEN: score = gap_open2;   matrix = C_M;  break;
            default:         score = gap_extend2; matrix = C_D2; break;
        }
        if (type == BT_M) { bt_push(c, 'X'); --offset; }
        else if (type <= BT_I2_EXT) { bt_push(c, 'I'); --k; --offset; }
        else { bt_push(c, 'D'); ++k; }
        v = offset - k; h = offset;
    }
    if (matrix == C_M) {
        if (v > 0 && h > 0) { int n = MIN2(v, h); bt_matches(c, n); v -= n; h -= n; }
        while (v > 0) { bt_push(c, 'D'); --v; }
        while (h > 0) { bt_push(c, 'I'); --h; }
    }
    ++c->begin;
    c->score = alignment_score;
}
static void cigar_clear(cigar_t *c) { c->begin = c->end = 0; c->score = INT32_MIN; c->end_v = c->end_h = -1; }

/* cigar_maxtrim_gap_affine, alignment/cigar.c:476-528 */
static int maxtrim_affine(cigar_t *c, const lcd_wfa_params_t *p) {
    const int b = c->begin, e = c->end;
    int max_score = 0, max_off = b, max_v = 0, max_h = 0, score = 0, ev = 0, eh = 0;
    char last = '\0';
    for (int i = b; i < e; ++i) {
        switch (c->ops[i]) {
            case 'M': score -= -1; ++ev; ++eh; break;
            case 'X': score -= p->mismatch; ++ev; ++eh; break;
            case 'I': score -= p->gap_ext1 + ((last == 'I') ? 0 : p->gap_open1); ++eh; break;
            case 'D': score -= p->gap_ext1 + ((last == 'D') ? 0 : p->gap_open1); ++ev; break;
        }
        last = c->ops[i];
        if (max_score < score) { max_score = score; max_off = i; max_v = ev; max_h = eh; }
    }
    const int trimmed = (max_off != e - 1);
    if (max_score == 0) cigar_clear(c);
    else { c->end = max_off + 1; c->score = max_score; c->end_v = max_v; c->end_h = max_h; }
    return trimmed;
}
/* cigar_maxtrim_gap_affine2p, alignment/cigar.c:529-600 */
static int score_op_2p(char op, int len, const lcd_wfa_params_t *p, int *ev, int *eh) {
    int s1 = p->gap_open1 + p->gap_ext1 * len, s2 = p->gap_open2 + p->gap_ext2 * len;
    switch (op) {
        case 'M': *ev += len; *eh += len; return -1 * len;
        case 'X': *ev += len; *eh += len; return p->mismatch * len;
        case 'D': *ev += len; return MIN2(s1, s2);
        default:  *eh += len; return MIN2(s1, s2);
    }
}
static int maxtrim_affine2p(cigar_t *c, const lcd_wfa_params_t *p) {
    const int b = c->begin, e = c->end;
    if (b >= e) return 0;
    int max_score = 0, max_off = b, max_v = 0, max_h = 0, score = 0, ev = 0, eh = 0, op_len = 0;
    char last = '\0';
    for (int i = b; i < e; ++i) {
        const char op = c->ops[i];
        if (op != last && last != '\0') {
            score -= score_op_2p(last, op_len, p, &ev, &eh);
            op_len = 0;
            if (max_score < score) { max_score = score; max_off = i - 1; max_v = ev; max_h = eh; }
        }
        last = op; ++op_len;
    }
    score -= score_op_2p(last, op_len, p, &ev, &eh);
    if (max_score < score) { max_score = score; max_off = e - 1; max_v = ev; max_h = eh; }
    const int trimmed = (max_off != e - 1);
    if (max_score == 0) cigar_clear(c);
    else { c->end = max_off + 1; c->score = max_score; c->end_v = max_v; c->end_h = max_h; }
    return trimmed;
}

int lcd_oracle_wfa_align(const uint8_t *pattern, int plen, const uint8_t *text, int tlen,
                         const lcd_wfa_params_t *par, char *ops, lcd_wfa_result_t *res) {
    wfa_t a; memset(&a, 0, sizeof(a));
    a.pattern = pattern; a.text = text; a.plen = plen; a.tlen = tlen; a.par = *par;
    const int scope_indel = par->affine2p ? MAX2(par->gap_open1 + par->gap_ext1, par->gap_open2 + par->gap_ext2)
                                          : par->gap_open1 + par->gap_ext1;
    a.max_score_scope = MAX2(scope_indel, par->mismatch) + 1;  /* wavefront_components.c:81-124 */
    a.steps_wait = par->steps_between_cutoffs;                 /* wavefront_heuristic_clear :115-122 */
    a.max_sw_score = 0; a.max_sw_score_offset = OFF_NULL; a.max_sw_score_k = DIAG_NULL;
    a.end_score = -1; a.end_k = DIAG_NULL; a.end_offset = OFF_NULL;
    wfset_t *s0 = wfa_new_set(&a, 0);                          /* wavefront_aligner_init_wf_m :251-310 */
    wf_alloc(&s0->c[C_M], 0, 0); s0->c[C_M].off[0] = 0;
    int score = 0, unreachable = 0;
    const int alignment_k = tlen - plen;
    for (;;) {                                                 /* wavefront_unialign :242-275 */
        wf_t *m = &a.sets[score].c[C_M];
        if (!m->exists) {                                      /* wavefront_extend_end2end :90-128 */
            if (a.num_null_steps > a.max_score_scope) { unreachable = 1; break; }
        } else {
            for (int k = m->lo; k <= m->hi; ++k) {             /* extend_matches_packed_end2end */
                int32_t off = m->off[k - m->base];
                if (off == OFF_NULL) continue;
                int v = off - k, h = off;
                while (v < plen && h < tlen && pattern[v] == text[h]) { ++v; ++h; ++off; }
                m->off[k - m->base] = off;
            }
            if (m->lo <= alignment_k && alignment_k <= m->hi && m->off[alignment_k - m->base] >= tlen) {
                a.end_score = score; a.end_k = alignment_k; a.end_offset = tlen;   /* termination :46-57 */
                break;
            }
            if (par->heuristic != LCD_WFA_HEUR_NONE && heur_cutoff(&a, score)) { unreachable = 1; break; }
        }
        ++score;
        wfa_compute(&a, score);
    }
    /* wavefront_unialign_terminate :146-236 (compute_alignment scope) */
    cigar_t c; c.max_ops = 2 * (plen + tlen); c.ops = (char*)malloc((size_t)c.max_ops + 8);
    cigar_clear(&c);
    if (a.end_offset != OFF_NULL) wfa_backtrace(&a, &c, score, a.end_k, a.end_offset);
    int status;
    if (unreachable) {
        if (par->affine2p) maxtrim_affine2p(&c, par); else maxtrim_affine(&c, par);
        status = LCD_WFA_STATUS_PARTIAL;
    } else {
        c.end_v = a.end_offset - a.end_k; c.end_h = a.end_offset;
        c.score = -score;                                      /* wavefront_compute_classic_score, match==0 */
        status = LCD_WFA_STATUS_COMPLETED;
    }
    int n = c.end - c.begin; if (n < 0) n = 0;
    if (ops) { memcpy(ops, c.ops + c.begin, (size_t)n); ops[n] = '\0'; }
    res->status = status; res->score = c.score; res->n_ops = n; res->end_v = c.end_v; res->end_h = c.end_h;
    for (int s = 0; s < a.n_sets; ++s) for (int cc = 0; cc < N_COMP; ++cc) free(a.sets[s].c[cc].off);
    free(a.sets); free(c.ops);
    return 0;
}
