/* oracle/pileup.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the per-site coverage pass of longcallD's pileup scan: collect_cand_vars
 * (reference src/collect_var.c:238-249) = for every kept read, update_cand_vars_from_digar
 * (src/bam_utils.c:287-329): a merge-join of the read's difference list (digar1_t: X / I / D events, sorted by
 * position) against the chunk's sorted candidate sites (var_site_t), which counts per site the reads carrying the
 * reference allele, the alternative allele (split by strand) and low-quality observations.
 * Comparator: exact_comp_var_site_ins (src/collect_var.c:1901-1935; large insertions match when their lengths are
 * within 20 %), first site of a read: get_var_site_start (src/bam_utils.c:229-241), event quality:
 * get_digar_ave_qual (src/bam_utils.c:258-280).
 * Pinned against the unmodified reference (oracle/_ref/libref_shim.so: ref_collect_cand_vars assembles a bam_chunk_t
 * with digar_t / var_site_t records around the same flat arrays) in tests/test_oracle_pileup.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

enum { CINS = 1, CDEL = 2, CEQUAL = 7, CDIFF = 8 };

typedef struct { int64_t pos; int type, ref_len, alt_len; const uint8_t *alt; } Site;

/* make_var_site_from_digar, src/collect_var.c:1113-1121 */
static Site site_of_digar(const lcd_pileup_input_t *in, int64_t d) {
    Site s; s.pos = in->digar_pos[d]; s.type = in->digar_type[d]; s.ref_len = 1; s.alt_len = in->digar_len[d];
    s.alt = in->digar_alt + in->digar_alt_off[d];
    if (s.type == CINS) s.ref_len = 0;
    else if (s.type == CDEL) { s.ref_len = in->digar_len[d]; s.alt_len = 0; }
    return s;
}
static Site site_of(const lcd_pileup_input_t *in, int i) {
    Site s; s.pos = in->site_pos[i]; s.type = in->site_type[i]; s.ref_len = in->site_ref_len[i]; s.alt_len = in->site_alt_len[i];
    s.alt = in->site_alt + in->site_alt_off[i];
    return s;
}
/* exact_comp_var_site_ins, src/collect_var.c:1901-1935 */
static int comp_site_ins(const lcd_pileup_input_t *in, const Site *a, const Site *b) {
    const int64_t pa = a->type == CDIFF ? a->pos : a->pos - 1, pb = b->type == CDIFF ? b->pos : b->pos - 1;
    if (pa < pb) return -1;
    if (pa > pb) return 1;
    if (a->type < b->type) return -1;
    if (a->type > b->type) return 1;
    if (a->ref_len < b->ref_len) return -1;
    if (a->ref_len > b->ref_len) return 1;
    if (a->type == CDIFF || (a->type == CINS && a->alt_len < in->min_sv_len)) {
        if (a->alt_len < b->alt_len) return -1;
        if (a->alt_len > b->alt_len) return 1;
        return memcmp(a->alt, b->alt, a->alt_len);
    } else if (a->type == CINS) {
        const int mn = a->alt_len < b->alt_len ? a->alt_len : b->alt_len, mx = a->alt_len > b->alt_len ? a->alt_len : b->alt_len;
        if (mn >= mx * 0.8) return 0;
        return a->alt_len - b->alt_len;
    }
    return 0;
}
/* get_digar_ave_qual, src/bam_utils.c:258-280 */
static int digar_ave_qual(const lcd_pileup_input_t *in, int r, int64_t d) {
    if (in->digar_low_qual[d]) return 0;
    const int qi = in->digar_qi[d];
    if (qi < 0) return 0;
    int q0, q1;
    if (in->digar_type[d] == CDEL) { if (qi == 0) { q0 = q1 = 0; } else { q0 = qi - 1; q1 = qi; } }
    else { q0 = qi; q1 = qi + in->digar_len[d] - 1; }
    int sum = 0;
    const uint8_t *qual = in->qual + in->qual_off[r];
    for (int i = q0; i <= q1; ++i) sum += qual[i];
    return sum / (q1 - q0 + 1);
}
/* update_var_site_with_allele, src/bam_utils.c:234-243 */
static void count(lcd_pileup_output_t *out, int site, int low_qual, int strand, int allele) {
    int32_t *c = out->site_counts + 8 * site;     /* total, low_qual, alle[2], strand[2][2] */
    if (low_qual) { c[1]++; return; }
    c[0]++; c[2 + allele]++; c[4 + 2 * strand + allele]++;
}

int lcd_oracle_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out) {
    const int ns = in->n_sites;
    memset(out->site_counts, 0, sizeof(int32_t) * 8 * (size_t)ns);
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        const int64_t beg = in->read_beg[r], end = in->read_end[r];
        const int strand = in->read_is_rev[r];
        /* get_var_site_start(var_sites, 0, n, beg), src/bam_utils.c:229-241 */
        int s;
        {
            const int64_t target = beg > 0 ? beg - 1 : beg;
            int left = 0, right = ns;
            while (left < right) {
                const int mid = left + (right - left) / 2;
                const int64_t mp = in->site_type[mid] == CDIFF ? in->site_pos[mid] : in->site_pos[mid] - 1;
                if (mp < target) left = mid + 1; else right = mid;
            }
            while (left < ns && in->site_pos[left] < beg) left++;
            s = left;
        }
        int64_t d = in->digar_first[r];
        const int64_t d_end = d + in->n_digar[r];
        while (s < ns && d < d_end) {
            if (in->digar_type[d] == CEQUAL) { d++; continue; }
            const Site ds = site_of_digar(in, d), vs = site_of(in, s);
            const int aq = digar_ave_qual(in, r, d);
            const int ret = comp_site_ins(in, &vs, &ds);
            if (ret < 0) { count(out, s, 0, strand, 0); s++; }
            else if (ret == 0) { count(out, s, in->digar_low_qual[d] || aq < in->min_bq, strand, 1); s++; }
            else d++;
        }
        for (; s < ns; ++s) {
            if (in->site_pos[s] > end) break;
            count(out, s, 0, strand, 0);
        }
    }
    return 0;
}
