/* oracle/noisyreg.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The chunk's noisy-region set and the candidate sites that stay in the clean regions: pre_process_noisy_regs
 * (src/collect_var.c:557-643) followed by classify_cand_vars after its first loop (:925-1033) for out_somatic = 0, with the helpers
 * cited at each function and the reference's own interval algebra (src/cgranges.c: cr_index order, cr_overlap, cr_cluster0 / cr_merge /
 * cr_merge2 :225-335, cr_is_contained :512-530), restated on flat (st, en, label) lists.
 */
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
enum { NON_VAR = 0x800, LOW_COV_VAR = 0x001, STRAND_BIAS_VAR = 0x002, LOW_AF_VAR = 0x400, REP_HET_VAR = 0x010 };
#define NOT_CAND (NON_VAR | LOW_COV_VAR | STRAND_BIAS_VAR)

typedef struct { int32_t st, en, label; } iv_t;
typedef struct { iv_t *a; int64_t n, m; } ivs_t;

static void iv_add(ivs_t *s, int32_t st, int32_t en, int32_t label) {
    if (s->n == s->m) { s->m = s->m ? s->m * 2 : 16; s->a = (iv_t*)realloc(s->a, (size_t)s->m * sizeof(iv_t)); }
    s->a[s->n].st = st; s->a[s->n].en = en; s->a[s->n].label = label; s->n++;
}
/* cr_add packs x = (uint64_t)st << 32 | en; cr_index sorts by x (src/cgranges.c:112-138) */
static uint64_t iv_key(const iv_t *v) { return ((uint64_t)(int64_t)v->st << 32) | (uint64_t)(int64_t)v->en; }
static int iv_cmp(const void *a, const void *b) { const uint64_t x = iv_key((const iv_t*)a), y = iv_key((const iv_t*)b); return x < y ? -1 : x > y; }
static void iv_index(ivs_t *s) { if (s->n > 1) qsort(s->a, (size_t)s->n, sizeof(iv_t), iv_cmp); }
static int iv_overlaps(const ivs_t *s, int32_t st, int32_t en) {              /* cr_overlap's count */
    int n = 0;
    for (int64_t i = 0; i < s->n; ++i) if (s->a[i].st < en && st < s->a[i].en) ++n;
    return n;
}
/* cr_is_contained (:512-530): only the LAST interval whose start is <= st is looked at (and those after it with the same start: none) */
static int iv_is_contained(const ivs_t *s, int32_t st, int32_t en) {
    int64_t left = 0, right = s->n;
    while (right > left) { const int64_t mid = left + ((right - left) >> 1); if (s->a[mid].st <= st) left = mid + 1; else right = mid; }
    if (left == 0) return 0;
    int n = 0;
    for (int64_t i = left - 1; i < s->n; ++i) {
        if (s->a[i].st >= en) break;
        if (s->a[i].st <= st && s->a[i].en >= en) ++n;
    }
    return n;
}
/* cr_cluster0 (:225-270): one pass; an interval absorbs every later one that starts within merge_win of its (growing) end */
static ivs_t iv_cluster0(ivs_t *s, int32_t fixed_win) {
    ivs_t o; memset(&o, 0, sizeof(o));
    uint8_t *merged = (uint8_t*)calloc((size_t)s->n + 1, 1);
    for (int64_t j = 0; j < s->n; ++j) {
        if (merged[j]) continue;
        uint64_t ms = (uint64_t)(int64_t)s->a[j].st, me = (uint64_t)(int64_t)s->a[j].en; int32_t ml = s->a[j].label;
        for (int64_t k = j + 1; k < s->n; ++k) {
            if (merged[k]) continue;
            const uint64_t ns = (uint64_t)(int64_t)s->a[k].st, ne = (uint64_t)(int64_t)s->a[k].en; const int32_t nl = s->a[k].label;
            const int win = fixed_win < 0 ? (ml < nl ? ml : nl) : fixed_win;
            if (me + win >= ns) { ml = ml > nl ? ml : nl; ms = ms < ns ? ms : ns; me = me > ne ? me : ne; merged[k] = 1; }
        }
        iv_add(&o, (int32_t)ms, (int32_t)me, ml);
    }
    free(merged); free(s->a); s->a = NULL; s->n = s->m = 0;
    iv_index(&o);
    return o;
}
static ivs_t iv_merge(ivs_t s, int32_t fixed_win) {                             /* cr_merge (:290-301) */
    int64_t cur = s.n;
    for (;;) { s = iv_cluster0(&s, fixed_win); if (s.n == cur) break; cur = s.n; }
    return s;
}

/* var_noisy_reads_ratio (:718-751) over the caches of build_var_noisy_reads_ratio_cache (:657-716) */
static float noisy_reads_ratio(const lcd_noisyreg_input_t *in, int64_t var_start, int64_t var_end) {
    const int32_t qs = (int32_t)(var_start - 1), qe = (int32_t)var_end;
    int total = 0, noisy = 0;
    for (int r = 0; r < in->n_reads; ++r) {
        if (in->is_skipped[r] || in->n_digar[r] <= 0 || in->read_beg[r] > in->read_end[r]) continue;
        if ((int32_t)(in->read_beg[r] - 1) < qe && qs < (int32_t)in->read_end[r]) ++total;
        /* the read's error intervals: overlapping X / I / D records joined (cur_start < noisy_end) */
        int have = 0, hit = 0; int64_t ns = -1, ne = -1;
        for (int j = 0; j < in->n_digar[r]; ++j) {
            const int64_t d = in->digar_first[r] + j; const int t = in->digar_type[d];
            if (t != CDIFF && t != CINS && t != CDEL) continue;
            const int64_t cs = in->digar_pos[d] - 1; int64_t ce = in->digar_pos[d];
            if (t == CDIFF || t == CDEL) ce += in->digar_len[d] - 1;
            if (!have) { ns = cs; ne = ce; have = 1; continue; }
            if (cs < ne) { if (ce > ne) ne = ce; continue; }
            if ((int32_t)ns < qe && qs < (int32_t)ne) hit = 1;
            ns = cs; ne = ce;
        }
        if (have && (int32_t)ns < qe && qs < (int32_t)ne) hit = 1;
        noisy += hit;
    }
    if (total == 0) return 0.0f;
    return (float)((float)noisy / (total + 0.0));
}

/* cr_add_var_cr (:754-778) */
static void add_var_cr(const lcd_noisyreg_input_t *in, const ivs_t *low, ivs_t *var_cr, int i, int check_ratio) {
    int64_t vs = in->site_pos[i], ve = in->site_type[i] == CINS ? in->site_pos[i] : in->site_pos[i] + in->site_ref_len[i] - 1;
    const int32_t qs = (int32_t)(vs - 1), qe = (int32_t)ve;
    for (int64_t j = 0; j < low->n; ++j) if (low->a[j].st < qe && qs < low->a[j].en) {
        const int32_t s = low->a[j].st + 1, e = low->a[j].en;
        if (s < vs) vs = s;
        if (e > ve) ve = e;
    }
    if (!check_ratio || noisy_reads_ratio(in, vs, ve) >= in->min_af) iv_add(var_cr, (int32_t)(vs - 1), (int32_t)ve, 1);
}

int lcd_oracle_noisy_regs(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) {
    const int n = in->n_sites;
    ivs_t R, low, var_pos, noisy_var; memset(&R, 0, sizeof(R)); memset(&low, 0, sizeof(low)); memset(&var_pos, 0, sizeof(var_pos)); memset(&noisy_var, 0, sizeof(noisy_var));
    for (int64_t i = 0; i < in->n_low; ++i) iv_add(&low, (int32_t)in->low_beg[i], (int32_t)in->low_end[i], 0);
    iv_index(&low);
    for (int64_t i = 0; i < in->n_cnreg; ++i) iv_add(&R, (int32_t)in->cnreg_beg[i], (int32_t)in->cnreg_end[i], in->cnreg_label[i]);
    /* ---- pre_process_noisy_regs (:557-643) */
    if (R.n > 0) {
        iv_index(&R);
        if (low.n > 0) {                                                            /* cr_extend_noisy_regs_with_low_comp (:538-553) */
            ivs_t E; memset(&E, 0, sizeof(E));
            for (int64_t i = 0; i < R.n; ++i) {
                const int32_t start = R.a[i].st + 1, end = R.a[i].en; int32_t ns = start, ne = end;
                for (int64_t j = 0; j < low.n; ++j) if (low.a[j].st < end && start - 1 < low.a[j].en) {       /* low_comp_cr_start_end (:466-479) */
                    if (low.a[j].st + 1 < ns) ns = low.a[j].st + 1;
                    if (low.a[j].en > ne) ne = low.a[j].en;
                }
                iv_add(&E, ns - 1, ne, R.a[i].label);
            }
            iv_index(&E); free(R.a); R = E;
        }
        R = iv_merge(R, -1);
        R = iv_merge(R, -1);
        int *tot = (int*)calloc((size_t)R.n + 1, sizeof(int)), *noi = (int*)calloc((size_t)R.n + 1, sizeof(int));
        for (int r = 0; r < in->n_reads; ++r) {
            if (in->is_skipped[r]) continue;
            const int32_t qs = (int32_t)(in->read_beg[r] - 1), qe = (int32_t)in->read_end[r];
            for (int64_t k = 0; k < R.n; ++k) {
                if (!(R.a[k].st < qe && qs < R.a[k].en)) continue;
                tot[k]++;
                int hit = 0;
                for (int x = 0; x < in->n_nreg[r] && !hit; ++x) {
                    const int64_t q = in->nreg_first[r] + x;
                    if ((int32_t)in->nreg_beg[q] < R.a[k].en && R.a[k].st < (int32_t)in->nreg_end[q]) hit = 1;
                }
                noi[k] += hit;
            }
        }
        const float min_ratio = (float)in->min_af;
        ivs_t K; memset(&K, 0, sizeof(K));
        for (int64_t k = 0; k < R.n; ++k) {
            if (noi[k] < in->min_alt_dp || (float)noi[k] / tot[k] < min_ratio) continue;
            iv_add(&K, R.a[k].st, R.a[k].en, R.a[k].label);
        }
        free(tot); free(noi); free(R.a); R = K;
    }
    /* ---- classify_cand_vars after classify_var_cate (:917-983); not called for a chunk without candidate sites (:2923-2924) */
    int32_t *c = out->var_cate;
    if (n == 0) goto done;
    for (int i = 0; i < n; ++i) {
        c[i] = in->var_cate[i]; out->keep[i] = 0;
        if (c[i] == LOW_COV_VAR) continue;
        if (in->is_ont && c[i] == STRAND_BIAS_VAR) continue;
        if (in->site_type[i] == CINS) iv_add(&var_pos, (int32_t)(in->site_pos[i] - 1), (int32_t)in->site_pos[i], 1);
        else iv_add(&var_pos, (int32_t)(in->site_pos[i] - 1), (int32_t)(in->site_pos[i] + in->site_ref_len[i] - 1), 1);
    }
    iv_index(&var_pos);
    for (int i = 0; i < n; ++i) {
        const int cate = c[i];
        if (cate == NON_VAR || cate == STRAND_BIAS_VAR) continue;
        const int32_t qs = (int32_t)(in->site_pos[i] - 1), qe = (int32_t)(in->site_type[i] == CINS ? in->site_pos[i] : in->site_pos[i] + in->site_ref_len[i] - 1);
        if (R.n > 0 && iv_overlaps(&R, qs, qe) > 0) { c[i] = NON_VAR; continue; }
        if (cate == LOW_COV_VAR) continue;
        const int in_reg = in->site_pos[i] >= in->reg_beg && in->site_pos[i] <= in->reg_end;
        if (cate == REP_HET_VAR) { if (in_reg) add_var_cr(in, &low, &noisy_var, i, 0); continue; }
        if (iv_overlaps(&var_pos, qs, qe) > 1 && in_reg) add_var_cr(in, &low, &noisy_var, i, 1);
        if (cate == LOW_AF_VAR) c[i] = LOW_COV_VAR;
    }
    if (noisy_var.n > 0) {                                                          /* cr_merge2 (:303-335) */
        for (int64_t k = 0; k < noisy_var.n; ++k) iv_add(&R, noisy_var.a[k].st, noisy_var.a[k].en, noisy_var.a[k].label);
        iv_index(&R);
        R = iv_merge(R, -1);
    }
    /* ---- post_process_noisy_regs + collect_noisy_reg_start_end (:482-535,646-655) */
    {
        const int nr = (int)R.n, flank = in->noisy_reg_flank_len;
        int *max_left = (int*)malloc(((size_t)nr + 1) * sizeof(int)), *min_right = (int*)malloc(((size_t)nr + 1) * sizeof(int));
        for (int k = 0; k < nr; ++k) max_left[k] = min_right[k] = -1;
        for (int k = 0, v = 0; k < nr && v < n;) {
            if (c[v] & NOT_CAND) { v++; continue; }
            const int32_t vs = (int32_t)in->site_pos[v], ve = (int32_t)(in->site_pos[v] + in->site_ref_len[v] - 1), rs = R.a[k].st + 1, re = R.a[k].en;
            if (vs > re) { if (min_right[k] == -1) min_right[k] = v; k++; }
            else if (ve < rs) { max_left[k] = v; v++; }
            else v++;
        }
        ivs_t P; memset(&P, 0, sizeof(P));
        for (int k = 0; k < nr; ++k) {
            if (max_left[k] == -1) max_left[k] = n - 1 < 0 ? n - 1 : 0;
            if (min_right[k] == -1) min_right[k] = n - 1 > 0 ? n - 1 : 0;
            const int32_t os = R.a[k].st + 1, oe = R.a[k].en;
            int32_t cs = os - flank, ce = oe + flank;
            for (int v = max_left[k]; v >= 0; --v) {
                if (c[v] & NOT_CAND) continue;
                const int32_t vs = (int32_t)in->site_pos[v], ve = (int32_t)(in->site_pos[v] + in->site_ref_len[v] - 1);
                if (ve < cs - 1) break;
                else if (vs - flank < cs) cs = vs - flank;
            }
            for (int v = min_right[k]; v < n; ++v) {
                if (c[v] & NOT_CAND) continue;
                const int32_t vs = (int32_t)in->site_pos[v], ve = (int32_t)(in->site_pos[v] + in->site_ref_len[v] - 1);
                if (vs > ce + 1) break;
                else if (ve + flank > ce) ce = ve + flank;
            }
            iv_add(&P, cs, ce, R.a[k].label);
        }
        free(max_left); free(min_right);
        iv_index(&P); free(R.a); R = iv_merge(P, 0);
    }
    /* ---- the sites that stay (:1012-1026) */
    for (int i = 0; i < n; ++i) {
        if (c[i] & NOT_CAND) continue;
        if (R.n > 0 && iv_is_contained(&R, (int32_t)(in->site_pos[i] - 1), (int32_t)(in->site_pos[i] + in->site_ref_len[i])) > 0) { c[i] = NON_VAR; continue; }
        out->keep[i] = 1;
    }
done:;
    int rc = 0;
    if (R.n > out->reg_cap) rc = -5;
    else for (int64_t k = 0; k < R.n; ++k) { out->reg_beg[k] = R.a[k].st; out->reg_end[k] = R.a[k].en; out->reg_label[k] = R.a[k].label; }
    out->n_regs = R.n;
    free(R.a); free(low.a); free(var_pos.a); free(noisy_var.a);
    return rc;
}
