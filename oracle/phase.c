/* oracle/phase.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of longcallD's read -> haplotype assignment and variant phasing
 * (reference src/assign_hap.c:16-547, assign_hap_based_on_germline_het_vars_kmeans) over flat arrays:
 * per read the span [start_var, end_var] of candidate variants it covers and its allele at each of them
 * (read_var_profile_t, src/collect_var.h:98-104), per variant its category, type, coverage
 * (cand_var_t, src/collect_var.h:71-95).  Integer arithmetic only (one 0.67 majority test for ONT
 * homopolymer indels).  What has to be reproduced bit for bit:
 *   - the greedy seed pass visits the reads covering a variant in the order cgranges returns them
 *     (src/cgranges.c: in-place MSD radix sort of the intervals by start, then an in-order tree walk),
 *     and every assignment updates the variant consensus the next read is scored against;
 *   - read_to_cons_allele_score fills in an unknown haplotype consensus in place (:139-143);
 *   - ties: haplotype 1 wins ('>' comparisons), the reference allele wins coverage ties.
 * Pinned against the unmodified reference (oracle/_ref/libref_shim.so: ref_assign_hap builds a bam_chunk_t
 * around the same arrays and calls the reference function) in tests/test_oracle_phase.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

enum { CLEAN_HET_SNP = 0x004, CLEAN_HET_INDEL = 0x008, CLEAN_HOM_VAR = 0x080, NOISY_CAND_HET_VAR = 0x100, NOISY_CAND_HOM_VAR = 0x200 };
#define GERMLINE_CLEAN (CLEAN_HET_SNP | CLEAN_HET_INDEL | CLEAN_HOM_VAR)
#define BAM_CDIFF 8

/* ---- the order cgranges hands intervals back in: cr_index = radix_sort_cr_intv (cgranges.c:13-64, in-place MSD
 *      radix sort, 8 bits a pass from the top byte of the 64-bit key, buckets <= 64 finished by insertion sort;
 *      NOT stable) followed by an in-order walk of the implicit tree (cr_overlap_int :449-490), i.e. array order. */
typedef struct { uint64_t key; int32_t label; } Intv;
static void intv_insertion_sort(Intv *beg, Intv *end) {
    for (Intv *i = beg + 1; i < end; ++i)
        if (i->key < (i - 1)->key) {
            Intv tmp = *i, *j;
            for (j = i; j > beg && tmp.key < (j - 1)->key; --j) *j = *(j - 1);
            *j = tmp;
        }
}
static void intv_flag_sort(Intv *beg, Intv *end, int shift) {
    Intv *bb[256], *be[256];
    size_t cnt[256];
    memset(cnt, 0, sizeof(cnt));
    for (Intv *i = beg; i != end; ++i) cnt[(i->key >> shift) & 255]++;
    Intv *p = beg;
    for (int k = 0; k < 256; ++k) { bb[k] = p; p += cnt[k]; be[k] = p; }
    for (int k = 0; k < 256;) {                 /* cycle-leader permutation into the buckets */
        if (bb[k] != be[k]) {
            int l = (int)((bb[k]->key >> shift) & 255);
            if (l != k) {
                Intv tmp = *bb[k], swap;
                do { swap = tmp; tmp = *bb[l]; *bb[l]++ = swap; l = (int)((tmp.key >> shift) & 255); } while (l != k);
                *bb[k]++ = tmp;
            } else ++bb[k];
        } else ++k;
    }
    if (shift) {
        const int next = shift > 8 ? shift - 8 : 0;
        Intv *b0 = beg;
        for (int k = 0; k < 256; ++k) {
            Intv *e0 = be[k];
            if (e0 - b0 > 64) intv_flag_sort(b0, e0, next);
            else if (e0 - b0 > 1) intv_insertion_sort(b0, e0);
            b0 = e0;
        }
    }
}
void lcd_oracle_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out) {
    Intv *v = (Intv*)malloc(sizeof(Intv) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { v[i].key = (uint64_t)(uint32_t)start[i]; v[i].label = label[i]; }
    if (n <= 64) intv_insertion_sort(v, v + n); else intv_flag_sort(v, v + n, 56);
    for (int i = 0; i < n; ++i) order_out[i] = v[i].label;
    free(v);
}

typedef struct {
    const lcd_phase_input_t *in;
    int32_t *cons;        /* [n_vars][3]  hap_to_cons_alle */
    int32_t *prof;        /* [n_vars][3][4] hap_to_alle_profile */
    int64_t *var_ps;      /* [n_vars] */
    int32_t *haps;        /* [n_reads] */
    int32_t *agree, *conflict;   /* n_clean_agree_snps / n_clean_conflict_snps per read */
} Ctx;

static inline int allele_of(const lcd_phase_input_t *in, int read, int var) { return in->alleles[in->allele_off[read] + (var - in->prof_start[read])]; }

/* get_var_init_max_cov_allele :22-35 */
static int init_max_cov_allele(const lcd_phase_input_t *in, int v) {
    if (in->is_ont == 1 && in->is_hp_indel[v]) return -1;
    int max_cov = 0, best = -1;
    for (int i = 0; i < in->n_uniq_alles[v]; ++i) if (in->alle_covs[4 * v + i] > max_cov) { max_cov = in->alle_covs[4 * v + i]; best = i; }
    return best;
}

/* read_to_cons_allele_score :127-147 (mutates the consensus of a half-known variant) */
static int cons_allele_score(Ctx *c, int hap, int v, int cate, int allele) {
    const int w = (cate == CLEAN_HET_SNP || cate == CLEAN_HET_INDEL) ? 2 : 1;
    int32_t *ca = c->cons + 3 * v;
    if (ca[hap] == -1 && ca[3 - hap] == -1) return 0;
    if (ca[hap] == -1) ca[hap] = 1 - ca[3 - hap];
    if (ca[3 - hap] == -1) ca[3 - hap] = 1 - ca[hap];
    if (ca[hap] == allele) return w;
    if (ca[hap] == -1) return 0;
    return -w;
}

/* init_assign_read_hap_based_on_cons_alle :151-197 */
static int assign_read(Ctx *c, int r, int target) {
    const lcd_phase_input_t *in = c->in;
    int score[3] = {0, 0, 0}, used[3] = {0, 0, 0}, ag[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
    c->agree[r] = c->conflict[r] = 0;
    for (int v = in->prof_start[r]; v <= in->prof_end[r]; ++v) {
        const int cate = in->var_cate[v];
        if ((cate & target) == 0) continue;
        if (in->is_hp_indel[v] == 1 || cate == NOISY_CAND_HOM_VAR) continue;
        const int al = allele_of(in, r, v);
        if (al < 0) continue;
        for (int hap = 1; hap <= 2; ++hap) {
            const int s = cons_allele_score(c, hap, v, cate, al);
            if (s != 0) {
                if (cate != CLEAN_HOM_VAR) used[hap]++;
                if ((cate & GERMLINE_CLEAN) > 0 && in->var_type[v] == BAM_CDIFF) { if (s > 0) ag[hap]++; else cf[hap]++; }
            }
            if (cate != CLEAN_HOM_VAR) score[hap] += s;
        }
    }
    int max_hap = 0, max_s = 0, min_hap = 0, min_s = 0;
    for (int hap = 1; hap <= 2; ++hap) {
        if (score[hap] > max_s) { max_hap = hap; max_s = score[hap]; }
        else if (score[hap] < min_s) { min_hap = hap; min_s = score[hap]; }
    }
    if (used[1] == 0 && used[2] == 0) return -1;
    if (max_s == 0 && min_s == 0) return 0;
    if (max_s > 0) { c->agree[r] = ag[max_hap]; c->conflict[r] = cf[max_hap]; return max_hap; }
    return 3 - min_hap;
}

/* update_var_hap_to_cons_alle :244-268 */
static void update_cons(Ctx *c, int v, int hap) {
    const lcd_phase_input_t *in = c->in;
    if (hap == 0) return;
    int max_cov = 0, best = -1, total = 0;
    const int32_t *p = c->prof + 12 * v + 4 * hap;
    for (int i = 0; i < in->n_uniq_alles[v]; ++i) { total += p[i]; if (p[i] > max_cov) { max_cov = p[i]; best = i; } }
    if (in->is_ont && in->is_hp_indel[v] == 1 && max_cov < total * 0.67) best = -1;
    c->cons[3 * v + hap] = best;
}

/* check_agree_haps :307-320 */
static int agree_haps(Ctx *c, int r, int hap, int v1, int v2) {
    const lcd_phase_input_t *in = c->in;
    if (v1 < in->prof_start[r] || v2 > in->prof_end[r]) return -1;
    if (hap == 0) return -1;
    const int a1 = allele_of(in, r, v1), a2 = allele_of(in, r, v2);
    if (a1 < 0 || a2 < 0) return -1;
    const int32_t *c1 = c->cons + 3 * v1, *c2 = c->cons + 3 * v2;
    if (c1[hap] == a1 && c2[hap] == a2) return 1;
    if (c1[hap] == a1 && c2[3 - hap] == a2) return 0;
    return -1;
}

int lcd_oracle_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out) {
    const int nr = in->n_reads, nv = in->n_vars, target = in->target_var_cate;
    Ctx c; c.in = in; c.cons = out->hap_to_cons_alle; c.prof = out->hap_to_alle_profile; c.var_ps = out->var_phase_set;
    c.haps = out->haps; c.agree = out->n_clean_agree_snps; c.conflict = out->n_clean_conflict_snps;
    int *valid = (int*)malloc(sizeof(int) * (nv > 0 ? nv : 1)), n_valid = 0;
    for (int v = 0; v < nv; ++v) if (in->var_cate[v] & target) valid[n_valid++] = v;
    if (n_valid == 0) { free(valid); return 0; }                                  /* :483-486: nothing is touched */
    /* read_init_hap_phase_set :16-20, var_init_hap_profile_cons_allele :39-63 */
    for (int r = 0; r < nr; ++r) { out->haps[r] = 0; out->phase_sets[r] = -1; }
    for (int k = 0; k < n_valid; ++k) {
        const int v = valid[k];
        memset(c.prof + 12 * v, 0, sizeof(int32_t) * 12);
        c.cons[3 * v] = init_max_cov_allele(in, v);
        c.cons[3 * v + 1] = c.cons[3 * v + 2] = (in->var_cate[v] == NOISY_CAND_HOM_VAR || in->var_cate[v] == CLEAN_HOM_VAR) ? 1 : -1;
    }
    /* read_var_cr as collect_read_var_profile builds it (src/collect_var.c:1407-1431) and the order it is walked in */
    int32_t *cr_start = (int32_t*)malloc(sizeof(int32_t) * (nr + 1)), *cr_label = (int32_t*)malloc(sizeof(int32_t) * (nr + 1));
    int32_t *cr_order = (int32_t*)malloc(sizeof(int32_t) * (nr + 1));
    int n_cr = 0;
    for (int i = 0; i < nr; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r] || in->prof_start[r] < 0 || in->prof_end[r] < 0) continue;
        cr_start[n_cr] = in->prof_start[r]; cr_label[n_cr] = r; n_cr++;
    }
    lcd_oracle_cr_order(n_cr, cr_start, cr_label, cr_order);
    /* select_init_var :94-125 */
    int init_k = -1;
    {
        int best[4] = {-1, -1, -1, -1}, depth[4] = {0, 0, 0, 0};
        for (int k = 0; k < n_valid; ++k) {
            const int v = valid[k], cate = in->var_cate[v];
            int cls = -1;
            if (cate == CLEAN_HET_SNP) cls = 0;
            else if (cate == CLEAN_HET_INDEL) cls = 1;
            else if (cate == NOISY_CAND_HET_VAR) { if (in->var_type[v] == BAM_CDIFF) cls = 2; else if (in->is_hp_indel[v] == 0) cls = 3; }
            if (cls >= 0 && (best[cls] == -1 || depth[cls] < in->total_cov[v])) { best[cls] = k; depth[cls] = in->total_cov[v]; }
        }
        for (int cls = 0; cls < 4 && init_k < 0; ++cls) init_k = best[cls];
    }
    /* seed pass :499-527: variants from the seed outwards (left side first, then right), reads in cgranges order */
    if (init_k != -1) {
        for (int step = 0; step < n_valid; ++step) {
            const int k = step == 0 ? init_k : (step <= init_k ? init_k - step : step);
            const int v = valid[k];
            if (in->var_cate[v] == NOISY_CAND_HOM_VAR || in->var_cate[v] == CLEAN_HOM_VAR) continue;
            for (int x = 0; x < n_cr; ++x) {
                const int r = cr_order[x];
                if (!(in->prof_start[r] < v + 1 && v < in->prof_end[r] + 1)) continue;      /* overlaps [v, v+1) */
                if (in->is_skipped[r] || c.haps[r] != 0) continue;
                int hap = assign_read(&c, r, target);
                if (hap == -1) hap = 1;
                c.haps[r] = hap;
                /* update_var_hap_profile_cons_alle_based_on_read_hap :270-290 (hap is 1 or 2 here, or 0 when tied) */
                for (int u = in->prof_start[r]; u <= in->prof_end[r]; ++u) {
                    if ((in->var_cate[u] & target) == 0) continue;
                    const int al = allele_of(in, r, u);
                    if (al < 0) continue;
                    if (hap == 0) { for (int h = 1; h <= 2; ++h) { c.prof[12 * u + 4 * h + al] += 1; update_cons(&c, u, h); } }
                    else { c.prof[12 * u + 4 * hap + al] += 1; update_cons(&c, u, hap); }
                }
            }
        }
    }
    /* iterations :530-542 */
    int *is_het = (int*)malloc(sizeof(int) * n_valid), *het = (int*)malloc(sizeof(int) * n_valid);
    int *n_agree = (int*)malloc(sizeof(int) * n_valid), *n_conf = (int*)malloc(sizeof(int) * n_valid);
    int32_t *snap = (int32_t*)malloc(sizeof(int32_t) * 2 * n_valid);
    for (int iter = 0; iter < 10; ++iter) {
        /* iter_update_var_hap_cons_phase_set :345-422 */
        int n_het = 0, changed1 = 0;
        for (int k = 0; k < n_valid; ++k) {
            const int v = valid[k]; const int32_t *ca = c.cons + 3 * v;
            is_het[k] = (ca[1] != -1 && ca[2] != -1 && ca[1] != ca[2] && in->is_hp_indel[v] == 0);
            if (is_het[k]) het[n_het++] = k;
            n_agree[k] = n_conf[k] = 0;
        }
        for (int h = 1; h < n_het; ++h) {
            const int k = het[h], v = valid[k], pv = valid[het[h - 1]];
            for (int x = 0; x < n_cr; ++x) {
                const int r = cr_order[x];
                if (!(in->prof_start[r] < v + 1 && pv < in->prof_end[r] + 1)) continue;      /* overlaps [pv, v+1) */
                if (in->is_skipped[r]) continue;
                const int a = agree_haps(&c, r, c.haps[r], pv, v);
                if (a > 0) n_agree[k]++; else if (a == 0) n_conf[k]++;
            }
        }
        int flip = 0; int64_t ps = -1;
        for (int k = 0; k < n_valid; ++k) {
            const int v = valid[k];
            const int64_t own = in->var_type[v] == BAM_CDIFF ? in->pos[v] : in->pos[v] - 1;
            if (k == 0) { ps = own; c.var_ps[v] = ps; continue; }
            if (is_het[k]) {
                if (n_agree[k] < 2 && n_conf[k] < 2) ps = own;
                else if (n_conf[k] > n_agree[k]) flip ^= 1;
                if (flip) {
                    changed1 = 1;
                    /* the reference's swap loop runs for hap = 1 and hap = 2, i.e. it swaps twice (:409-413) */
                    for (int hap = 1; hap <= 2; ++hap) { const int32_t t = c.cons[3 * v + hap]; c.cons[3 * v + hap] = c.cons[3 * v + 3 - hap]; c.cons[3 * v + 3 - hap] = t; }
                }
            }
            c.var_ps[v] = ps;
        }
        /* iter_update_var_hap_to_cons_alle :425-467 */
        for (int k = 0; k < n_valid; ++k) { snap[2 * k] = c.cons[3 * valid[k] + 1]; snap[2 * k + 1] = c.cons[3 * valid[k] + 2]; }
        for (int k = 0; k < n_valid; ++k) memset(c.prof + 12 * valid[k], 0, sizeof(int32_t) * 12);
        for (int i = 0; i < nr; ++i) {
            const int r = in->ordered_read_ids[i];
            if (in->is_skipped[r]) continue;
            int hap = assign_read(&c, r, target);
            if (hap == -1) hap = 0;
            c.haps[r] = hap;
            for (int u = in->prof_start[r]; u <= in->prof_end[r] && in->prof_start[r] >= 0; ++u) {
                if ((in->var_cate[u] & target) == 0) continue;
                const int al = allele_of(in, r, u);
                if (al < 0) continue;
                if (hap == 0) { c.prof[12 * u + 4 + al] += 1; c.prof[12 * u + 8 + al] += 1; }
                else c.prof[12 * u + 4 * hap + al] += 1;
            }
        }
        int changed2 = 0;
        for (int k = 0; k < n_valid; ++k) { update_cons(&c, valid[k], 1); update_cons(&c, valid[k], 2); }
        for (int k = 0; k < n_valid; ++k) if (c.cons[3 * valid[k] + 1] != snap[2 * k] || c.cons[3 * valid[k] + 2] != snap[2 * k + 1]) { changed2 = 1; break; }
        if (!changed1 && !changed2) break;
    }
    /* update_read_phase_set :322-339 */
    for (int i = 0; i < nr; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r] || in->prof_start[r] == -1) continue;
        int64_t ps = -1;
        for (int v = in->prof_start[r]; v <= in->prof_end[r]; ++v) {
            if ((in->var_cate[v] & target) == 0) continue;
            const int32_t *ca = c.cons + 3 * v;
            if (ca[1] != -1 && ca[2] != -1 && ca[1] != ca[2]) ps = c.var_ps[v];
            if (ps != -1) break;
        }
        out->phase_sets[r] = ps;
    }
    free(valid); free(cr_start); free(cr_label); free(cr_order); free(is_het); free(het); free(n_agree); free(n_conf); free(snap);
    return 0;
}
