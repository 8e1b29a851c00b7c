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

/* ---------------------------------------------------------------------------------------------------------
 * Read x variant profile: collect_read_var_profile (src/collect_var.c:1389-1431) = for every kept read,
 * update_read_vs_all_var_profile_from_digar (src/bam_utils.c:446-552), germline categories (a chunk holding
 * LONGCALLD_CAND_SOMATIC_VAR candidates, only produced with -s, is outside the restated path).
 * Comparator: comp_ovlp_var_site = ovlp_var_site (src/collect_var.c:79-95) + exact_comp_var_site (:1878-1898). */
enum { NON_VAR = 0x800, CAND_SOMATIC_VAR = 0x040 };

static int ovlp_site(const Site *a, const Site *b) {
    const int beg1 = (int)a->pos, end1 = (int)a->pos + a->ref_len, beg2 = (int)b->pos, end2 = (int)b->pos + b->ref_len;
    if (a->ref_len == 0 && b->ref_len == 0) return beg1 == beg2;
    if (a->ref_len == 0) return beg1 > beg2 && end1 < end2;
    if (b->ref_len == 0) return beg2 > beg1 && end2 < end1;
    return !(beg1 >= end2 || beg2 >= end1);
}
static int comp_site_exact(const Site *a, const Site *b) {
    const int64_t pa = a->type == CDIFF ? a->pos : a->pos - 1, pb = b->type == CDIFF ? b->pos : b->pos - 1;
    if (pa < pb) return -1;
    if (pa > pb) return 1;
    if (a->type < b->type) return -1;
    if (a->type > b->type) return 1;
    if (a->ref_len < b->ref_len) return -1;
    if (a->ref_len > b->ref_len) return 1;
    if (a->alt_len < b->alt_len) return -1;
    if (a->alt_len > b->alt_len) return 1;
    if (a->type == CDIFF || a->type == CINS) return memcmp(a->alt, b->alt, a->alt_len);
    return 0;
}

int lcd_oracle_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out) {
    const int nv = in->n_sites;
    for (int v = 0; v < nv; ++v) if (ex->var_cate[v] == CAND_SOMATIC_VAR) return -2;
    int64_t top = 0;
    for (int r = 0; r < in->n_reads; ++r) { out->prof_start[r] = -1; out->prof_end[r] = -2; out->allele_off[r] = 0; }
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        const int64_t beg = in->read_beg[r], end = in->read_end[r];
        /* get_var_start, src/bam_utils.c:215-227 (same search as get_var_site_start) */
        int v;
        {
            const int64_t target = beg > 0 ? beg - 1 : beg;
            int left = 0, right = nv;
            while (left < right) {
                const int mid = left + (right - left) / 2;
                const int64_t mp = in->site_type[mid] == CDIFF ? in->site_pos[mid] : in->site_pos[mid] - 1;
                if (mp < target) left = mid + 1; else right = mid;
            }
            while (left < nv && in->site_pos[left] < beg) left++;
            v = left;
        }
        /* capacity of this read's row: every variant from the first candidate up to the last one not right of the read */
        int cap_end = v;
        while (cap_end < nv && in->site_pos[cap_end] <= end) cap_end++;
        /* a variant right of `end` may still be visited by the merge loop (events at the read's last bases): one more */
        int64_t row0 = top; const int v0 = v; const int64_t cap = (int64_t)(cap_end - v0) + 2;
        if (top + cap > out->alleles_cap) return -3;
        for (int64_t k = 0; k < cap; ++k) { out->alleles[row0 + k] = -1; out->alt_qi[row0 + k] = -1; }
        top += cap;
        int start = -1, last = -2;
#define SET_PROFILE(var_i, al, qi_) do { if ((var_i) - v0 >= cap) return -4; if (start == -1) start = (var_i); last = (var_i); \
            out->alleles[row0 + ((var_i) - v0)] = (int8_t)(al); out->alt_qi[row0 + ((var_i) - v0)] = (qi_); } while (0)
        int64_t d = in->digar_first[r];
        const int64_t d_end = d + in->n_digar[r];
        while (v < nv && d < d_end) {
            if (ex->var_cate[v] == NON_VAR) { v++; continue; }
            if (in->digar_type[d] == CEQUAL) { d++; continue; }
            const Site vs = site_of(in, v), ds = site_of_digar(in, d);
            const int aq = digar_ave_qual(in, r, d);
            const int is_ovlp = ovlp_site(&vs, &ds), ret = comp_site_exact(&vs, &ds);
            if (!is_ovlp) {
                if (ret < 0) { SET_PROFILE(v, 0, -1); v++; }
                else if (ret > 0) d++;
                else { v++; d++; }                                  /* "unexpected case" branch of the reference (:509-512) */
            } else {
                if (ret == 0) { SET_PROFILE(v, aq < in->min_bq ? -2 : 1, in->digar_qi[d]); v++; }
                else { SET_PROFILE(v, -1, -1); v++; }
            }
        }
        for (; v < nv; ++v) {
            if (in->site_pos[v] > end) break;
            /* is_in_noisy_reg(pos, digar->noisy_regs): any interval [st, en) with st < pos + 1 and pos < en */
            int noisy = 0;
            for (int64_t k = ex->nreg_first[r]; k < ex->nreg_first[r] + ex->n_nreg[r]; ++k) if (ex->nreg_beg[k] < in->site_pos[v] + 1 && in->site_pos[v] < ex->nreg_end[k]) { noisy = 1; break; }
            if (noisy) continue;
            SET_PROFILE(v, 0, -1);
        }
#undef SET_PROFILE
        out->prof_start[r] = start; out->prof_end[r] = last;
        out->allele_off[r] = start >= 0 ? row0 + (start - v0) : row0;          /* so that alleles[allele_off[r] + (var - prof_start[r])] */
    }
    out->n_alleles = top;
    return 0;
}
