/* oracle/sites.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the candidate-site list of longcallD's pileup scan, collect_all_cand_var_sites (reference
 * src/collect_var.c:1209-1254): every collectible record (is_collectible_var_digar :1153-1160: X / I / D, not low quality,
 * starting inside [reg_beg, reg_end]) of every kept read becomes a var_site_t (make_var_site_from_digar :1113-1121); the list
 * is sorted with exact_comp_var_site (:1878-1898: position with indels anchored one base left, type, ref_len, alt_len, alt
 * bytes) and de-duplicated with exact_comp_var_site_ins (:1901-1935), which also merges large insertions (alt_len >=
 * min_sv_len) at the same anchor whose shorter length is >= 0.8 x the longer one into the first (shortest) of them.
 * Pinned against the unmodified reference (oracle/_ref/libref_shim.so: ref_collect_sites) in tests/test_oracle_sites.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
typedef struct { int64_t pos; int32_t type, ref_len, alt_len; const uint8_t *alt; int64_t src; } Site;

static int cmp_exact(const void *a_, const void *b_) {                     /* exact_comp_var_site */
    const Site *a = (const Site*)a_, *b = (const Site*)b_;
    const int64_t pa = a->type == CDIFF ? a->pos : a->pos - 1, pb = b->type == CDIFF ? b->pos : b->pos - 1;
    if (pa != pb) return pa < pb ? -1 : 1;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    if (a->ref_len != b->ref_len) return a->ref_len < b->ref_len ? -1 : 1;
    if (a->alt_len != b->alt_len) return a->alt_len < b->alt_len ? -1 : 1;
    if (a->type == CDIFF || a->type == CINS) return memcmp(a->alt, b->alt, (size_t)a->alt_len);
    return 0;
}
static int cmp_ins(const Site *a, const Site *b, int min_sv_len) {          /* exact_comp_var_site_ins */
    const int64_t pa = a->type == CDIFF ? a->pos : a->pos - 1, pb = b->type == CDIFF ? b->pos : b->pos - 1;
    if (pa != pb) return pa < pb ? -1 : 1;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    if (a->ref_len != b->ref_len) return a->ref_len < b->ref_len ? -1 : 1;
    if (a->type == CDIFF || (a->type == CINS && a->alt_len < min_sv_len)) {
        if (a->alt_len != b->alt_len) return a->alt_len < b->alt_len ? -1 : 1;
        return memcmp(a->alt, b->alt, (size_t)a->alt_len);
    } else if (a->type == CINS) {
        const int mn = a->alt_len < b->alt_len ? a->alt_len : b->alt_len, mx = a->alt_len > b->alt_len ? a->alt_len : b->alt_len;
        if (mn >= mx * 0.8) return 0;
        return a->alt_len - b->alt_len;
    }
    return 0;
}

int lcd_oracle_collect_sites(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end, lcd_sites_output_t *out) {
    int64_t n = 0, m = 0;
    for (int i = 0; i < in->n_reads; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) m += in->n_digar[r]; }
    Site *s = (Site*)malloc(sizeof(Site) * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        for (int64_t d = in->digar_first[r]; d < in->digar_first[r] + in->n_digar[r]; ++d) {
            const int t = in->digar_type[d];
            if ((reg_beg != -1 && in->digar_pos[d] < reg_beg) || (reg_end != -1 && in->digar_pos[d] > reg_end)) continue;
            if (in->digar_low_qual[d] || (t != CDIFF && t != CINS && t != CDEL)) continue;
            Site *x = s + n++;
            x->pos = in->digar_pos[d]; x->type = t; x->ref_len = t == CINS ? 0 : (t == CDEL ? in->digar_len[d] : 1); x->alt_len = t == CDEL ? 0 : in->digar_len[d];
            x->alt = in->digar_alt + in->digar_alt_off[d]; x->src = d;
        }
    }
    qsort(s, (size_t)n, sizeof(Site), cmp_exact);
    int64_t w = n > 0 ? 1 : 0;
    for (int64_t i = 1; i < n; ++i) { if (cmp_ins(s + w - 1, s + i, in->min_sv_len) == 0) continue; s[w++] = s[i]; }
    out->n_sites = w;
    if (w > out->cap) { free(s); return -3; }
    for (int64_t i = 0; i < w; ++i) { out->site_pos[i] = s[i].pos; out->site_type[i] = s[i].type; out->site_ref_len[i] = s[i].ref_len; out->site_alt_len[i] = s[i].alt_len; out->site_src[i] = s[i].src; }
    free(s);
    return 0;
}
