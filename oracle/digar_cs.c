/* oracle/digar_cs.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the cs-tag variant of the difference-list pass, collect_digar_from_cs_tag (reference src/bam_utils.c:844-1001): the
 * records come from the walk over the cs string (":n" / "=seq" matching runs, "*xy" mismatches, "+seq" insertions, "-seq" deletions, "~..."
 * introns -- which the reference skips WITHOUT advancing the reference position), the alt bases from the tag's own letters, and clips from the
 * first and the last CIGAR op only, with an event count that differs from the =/X variant's at contig ends (a long clip counts even when its
 * interval is not added).  Everything after the walk (pending window, skip test, cr_index order, chunk list) is the =/X variant's.
 * Groundwork for the next K1 variant on the GPU; pinned against the unmodified reference (oracle/_ref/libref_shim.so: ref_collect_digar_cs)
 * in tests/test_oracle_digar_cs.py.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

enum { CMATCH = 0, CINS = 1, CDEL = 2, CREF_SKIP = 3, CSOFT = 4, CHARD = 5, CEQUAL = 7, CDIFF = 8 };

typedef struct {
    int64_t *pos; int *len, *cnt; int front, rear; int64_t count;
    int64_t cur_start, cur_end; int q_start, q_end;
} Win;

typedef struct { lcd_digar_output_t *out; int64_t reg_first, n_reg, reg_cap; } Regs;
static int add_reg(Regs *g, int64_t st, int64_t en, int32_t label) {      /* cr_add, src/cgranges.c:145-160 */
    if (st < 0) st = 0;
    if (st > en) return 0;
    if (g->n_reg >= g->reg_cap) return -1;
    g->out->nreg_beg[g->reg_first + g->n_reg] = st; g->out->nreg_end[g->reg_first + g->n_reg] = en; g->out->nreg_label[g->reg_first + g->n_reg] = label;
    g->n_reg++;
    return 0;
}
/* flush the pending dense window as a noisy interval: label = max(sum of event sizes, window length) (:191-196, :776-781) */
static int flush_win(Win *q, Regs *g) {
    int64_t var_size = 0;
    for (int i = q->q_start; i <= q->q_end; ++i) var_size += q->cnt[i];
    if (var_size < q->cur_end - q->cur_start + 1) var_size = q->cur_end - q->cur_start + 1;
    return add_reg(g, q->cur_start - 1, q->cur_end, (int32_t)var_size);
}
/* push_xid_size_queue_win, src/bam_utils.c:161-205 */
static int push_win(Win *q, int64_t pos, int len, int count, int win, int max_s, Regs *g) {
    ++q->rear; q->pos[q->rear] = pos; q->len[q->rear] = len; q->cnt[q->rear] = count; q->count += count;
    while (q->pos[q->front] + q->len[q->front] - 1 <= pos - win) { q->count -= q->cnt[q->front]; q->front++; }
    if (count > 0 && q->count > max_s) {
        const int64_t ns = q->pos[q->front], ne = q->pos[q->rear] + q->len[q->rear];
        if (q->cur_start == -1) { q->cur_start = ns; q->cur_end = ne; q->q_start = q->front; q->q_end = q->rear; }
        else if (ns <= q->cur_end) { q->cur_end = ne; q->q_end = q->rear; }
        else {
            if (flush_win(q, g)) return -1;
            q->cur_start = ns; q->cur_end = ne; q->q_start = q->front; q->q_end = q->rear;
        }
    }
    return 0;
}

void lcd_oracle_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out);   /* phase.c: cgranges' cr_index order */

static int nt4(char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }

/* mode 1: the cs-tag variant; mode 2: the no-tag variant, collect_digar_from_ref_seq (src/bam_utils.c:1176-1290) -- every base of an M / = / X op
 * against the chunk's reference window ref_seq = positions ref_beg .. ref_end; a base outside the window is passed over without a record and
 * without closing the running match (:1209-1215), whose record is placed by counting back from where it ends (:1219, :1236); clips as in the
 * =/X variant (any op, :1264-1282). */
static int collect_tagged(int mode, const lcd_digar_input_t *in, const int64_t *cs_off, const char *cs_all, const char *ref_seq, int64_t ref_beg, int64_t ref_end,
                          lcd_digar_output_t *out) {
    int64_t dtop = 0, atop = 0, rtop = 0;
    out->n_cnreg = 0;
    memset(out->qual_counts, 0, sizeof(int64_t) * 256);
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        out->skip[r] = 0;
        if (in->is_skipped[r]) continue;
        const uint32_t *cigar = in->cigar + in->cigar_off[r]; const int nc = in->n_cigar[r];
        const uint8_t *qual = in->qual + in->qual_off[r];
        const int qlen = in->l_qseq[r];
        for (int k = 0; k < qlen; ++k) out->qual_counts[qual[k]]++;                      /* longcalld_copy_digar_read_buffers :90-103 */
        int64_t pos = in->read_pos0[r] + 1, rlen = 0; int qi = 0;
        for (int k = 0; k < nc; ++k) { const int op = cigar[k] & 15; if (op == CMATCH || op == CDEL || op == CREF_SKIP || op == CEQUAL || op == CDIFF) rlen += cigar[k] >> 4; }
        out->read_beg[r] = pos; out->read_end[r] = in->read_pos0[r] + (rlen ? rlen : 1);  /* bam_endpos: pos + rlen (rlen 0 -> pos + 1) */
        out->digar_first[r] = dtop; out->nreg_first[r] = rtop;
        Win q; q.pos = (int64_t*)malloc(sizeof(int64_t) * (rlen + 2)); q.len = (int*)malloc(sizeof(int) * (rlen + 2)); q.cnt = (int*)malloc(sizeof(int) * (rlen + 2));
        q.pos[0] = 0; q.len[0] = 0; q.cnt[0] = 0;
        q.front = 0; q.rear = -1; q.count = 0; q.cur_start = q.cur_end = -1; q.q_start = q.q_end = -1;
        Regs g = { out, rtop, 0, out->nreg_cap - rtop };
        const int left_pal = in->is_palindrome[r] && in->read_is_rev[r], right_pal = in->is_palindrome[r] && !in->read_is_rev[r];
        int n_cand = 0, rc = 0;
#define PUSH_DIGAR(p_, t_, l_, q_, low_) do { if (dtop >= out->digar_cap) { rc = -3; goto done; } out->digar_pos[dtop] = (p_); out->digar_type[dtop] = (int8_t)(t_); \
            out->digar_len[dtop] = (l_); out->digar_qi[dtop] = (q_); out->digar_low_qual[dtop] = (uint8_t)(low_); out->digar_alt_off[dtop] = atop; dtop++; } while (0)
        const char *cs = mode == 1 ? cs_all + cs_off[r] : "";
        if (mode == 2) {
            const uint8_t *bs = in->bseq + in->seq_off[r];
            for (int k = 0; k < nc; ++k) {
                const int op = cigar[k] & 15, len = (int)(cigar[k] >> 4);
                if (op == CMATCH || op == CDIFF || op == CEQUAL) {
                    int eq_len = 0;
                    for (int j = 0; j < len; ++j) {
                        if (pos < ref_beg || pos > ref_end) { pos++; qi++; continue; }
                        const int ref_base = nt4(ref_seq[pos - ref_beg]);
                        const int c = (bs[qi >> 1] >> ((~qi & 1) << 2)) & 15;
                        const int read_base = c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
                        if (ref_base != read_base) {
                            if (eq_len > 0) { PUSH_DIGAR(pos - eq_len, CEQUAL, eq_len, qi - eq_len, 0); eq_len = 0; }
                            const int low = !(qual[qi] >= in->min_bq);
                            if (!low && push_win(&q, pos, 1, 1, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                            PUSH_DIGAR(pos, CDIFF, 1, qi, low);
                            if (atop >= out->alt_cap) { rc = -3; goto done; }
                            out->digar_alt[atop++] = (uint8_t)read_base;
                            n_cand++;
                        } else eq_len++;
                        pos++; qi++;
                    }
                    if (eq_len > 0) PUSH_DIGAR(pos - eq_len, CEQUAL, eq_len, qi - eq_len, 0);
                } else if (op == CDEL) {
                    const int ok = (qi == 0 || qual[qi - 1] >= in->min_bq) && qual[qi] >= in->min_bq;
                    if (ok && push_win(&q, pos, len, len, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                    PUSH_DIGAR(pos, CDEL, len, qi, !ok);
                    n_cand++; pos += len;
                } else if (op == CINS) {
                    int low = 1;
                    for (int j = 0; j < len; ++j) if (qual[qi + j] >= in->min_bq) { low = 0; break; }
                    if (!low && push_win(&q, pos, 0, len, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                    PUSH_DIGAR(pos, CINS, len, qi, low);
                    if (atop + len > out->alt_cap) { rc = -3; goto done; }
                    for (int j = 0; j < len; ++j) { const int c = (bs[(qi + j) >> 1] >> ((~(qi + j) & 1) << 2)) & 15; out->digar_alt[atop++] = (uint8_t)(c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4); }
                    n_cand++; qi += len;
                } else if (op == CSOFT || op == CHARD) {
                    const int pal = (k == 0 && left_pal) || (k != 0 && right_pal);
                    PUSH_DIGAR(pos, pal ? CHARD : op, len, qi, 0);
                    if (((k == 0 && pos > 10) || (k != 0 && pos < in->whole_ref_len - 10)) && len > in->end_clip_reg) {
                        if (k == 0 && !left_pal) { if (pos > 1 && add_reg(&g, pos - 1, pos + in->end_clip_reg_flank_win, 0)) { rc = -4; goto done; } n_cand++; }
                        else if (k != 0 && !right_pal) { if (pos < in->whole_ref_len && add_reg(&g, pos - 1 - in->end_clip_reg_flank_win, pos, 0)) { rc = -4; goto done; } n_cand++; }
                    }
                    if (op == CSOFT) qi += len;
                } else if (op == CREF_SKIP) pos += len;
            }
        }
        if (mode == 1 && nc <= 0) { rc = -2; goto done; }
        if (mode == 1) {   /* left-end clipping: the first CIGAR op only (:876-889) */
            const int op = cigar[0] & 15, len = (int)(cigar[0] >> 4);
            if (op == CSOFT || op == CHARD) {
                PUSH_DIGAR(pos, left_pal ? CHARD : op, len, qi, 0);
                if (len > in->end_clip_reg && !left_pal) { if (pos > 10 && add_reg(&g, pos - 1, pos + in->end_clip_reg_flank_win, 0)) { rc = -4; goto done; } n_cand++; }
                if (op == CSOFT) qi += len;
            }
        }
        while (*cs) {
            if (*cs == ':') {
                char *end; const int len = (int)strtol(cs + 1, &end, 10); cs = end;
                PUSH_DIGAR(pos, CEQUAL, len, qi, 0); pos += len; qi += len;
            } else if (*cs == '=') {
                int len = 0; cs++;
                while (isalpha((unsigned char)*cs)) { len++; cs++; }
                PUSH_DIGAR(pos, CEQUAL, len, qi, 0); pos += len; qi += len;
            } else if (*cs == '*') {
                const int low = !(qual[qi] >= in->min_bq);
                if (!low && push_win(&q, pos, 1, 1, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                PUSH_DIGAR(pos, CDIFF, 1, qi, low);
                if (atop >= out->alt_cap) { rc = -3; goto done; }
                out->digar_alt[atop++] = (uint8_t)nt4(cs[2]);
                n_cand++; pos++; qi++; cs += 3;
            } else if (*cs == '+') {
                int len = 0, low = 1; cs++;
                while (isalpha((unsigned char)*cs)) { len++; cs++; }
                for (int j = 0; j < len; ++j) if (qual[qi + j] >= in->min_bq) { low = 0; break; }
                if (!low && push_win(&q, pos, 0, len, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                PUSH_DIGAR(pos, CINS, len, qi, low);
                if (atop + len > out->alt_cap) { rc = -3; goto done; }
                for (int j = 0; j < len; ++j) out->digar_alt[atop++] = (uint8_t)nt4(cs[j - len]);
                n_cand++; qi += len;
            } else if (*cs == '-') {
                int len = 0; cs++;
                while (isalpha((unsigned char)*cs)) { len++; cs++; }
                const int ok = (qi == 0 || qual[qi - 1] >= in->min_bq) && qual[qi] >= in->min_bq;
                if (ok && push_win(&q, pos, len, len, in->noisy_reg_slide_win, in->noisy_reg_max_xgaps, &g)) { rc = -4; goto done; }
                PUSH_DIGAR(pos, CDEL, len, qi, !ok);
                n_cand++; pos += len;
            } else if (*cs == '~') {                                                       /* intron: skipped, the position is NOT advanced (:945-947) */
                cs++;
                while (isalpha((unsigned char)*cs) || isdigit((unsigned char)*cs)) cs++;
            } else { rc = -2; goto done; }
        }
        if (mode == 1) {   /* right-end clipping: the last CIGAR op only (:953-966) */
            const int op = cigar[nc - 1] & 15, len = (int)(cigar[nc - 1] >> 4);
            if (op == CSOFT || op == CHARD) {
                PUSH_DIGAR(pos, right_pal ? CHARD : op, len, qi, 0);
                if (len > in->end_clip_reg && !right_pal) { if (pos < in->whole_ref_len - 10 && add_reg(&g, pos - 1 - in->end_clip_reg_flank_win, pos, 0)) { rc = -4; goto done; } n_cand++; }
                if (op == CSOFT) qi += len;
            }
        }
        if (q.cur_start != -1) {                                                           /* :776-782 (the pending window is closed with its own end) */
            int64_t var_size = 0;
            for (int x = q.q_start; x <= q.q_end; ++x) var_size += q.cnt[x];
            if (var_size < q.cur_end - q.cur_start + 1) var_size = q.cur_end - q.cur_start + 1;
            if (add_reg(&g, q.cur_start - 1, q.cur_end, (int32_t)var_size)) { rc = -4; goto done; }
        }
        {
            int64_t noisy_len = 0;                                                         /* collect_noisy_region_len :624-631 */
            for (int64_t x = 0; x < g.n_reg; ++x) noisy_len += out->nreg_end[rtop + x] - out->nreg_beg[rtop + x] + 1;
            const int64_t mapped = out->read_end[r] - out->read_beg[r] + 1;
            if ((int)noisy_len > mapped * in->max_noisy_frac_per_read || n_cand > mapped * in->max_var_ratio_per_read) out->skip[r] = 1;
        }
        if (g.n_reg > 1) {                                                                 /* cr_index(digar->noisy_regs) :784 */
            const int n = (int)g.n_reg;
            int32_t *st = (int32_t*)malloc(sizeof(int32_t) * n * 3), *id = st + n, *ord = st + 2 * n;
            int64_t *tb = (int64_t*)malloc(sizeof(int64_t) * n * 2), *te = tb + n; int32_t *tl = (int32_t*)malloc(sizeof(int32_t) * n);
            for (int x = 0; x < n; ++x) { st[x] = (int32_t)out->nreg_beg[rtop + x]; id[x] = x; tb[x] = out->nreg_beg[rtop + x]; te[x] = out->nreg_end[rtop + x]; tl[x] = out->nreg_label[rtop + x]; }
            lcd_oracle_cr_order(n, st, id, ord);
            for (int x = 0; x < n; ++x) { out->nreg_beg[rtop + x] = tb[ord[x]]; out->nreg_end[rtop + x] = te[ord[x]]; out->nreg_label[rtop + x] = tl[ord[x]]; }
            free(st); free(tb); free(tl);
        }
        if (!out->skip[r])                                                                 /* :819-832: is_overlap_reg(start + 1, end, reg_beg, reg_end) */
            for (int64_t x = 0; x < g.n_reg; ++x) {
                const int64_t b = out->nreg_beg[rtop + x], e = out->nreg_end[rtop + x];
                if (!(b + 1 > in->reg_end || e < in->reg_beg)) {
                    if (out->n_cnreg >= out->cnreg_cap) { rc = -4; goto done; }
                    out->cnreg_beg[out->n_cnreg] = b; out->cnreg_end[out->n_cnreg] = e; out->cnreg_label[out->n_cnreg] = out->nreg_label[rtop + x]; out->n_cnreg++;
                }
            }
done:
#undef PUSH_DIGAR
        free(q.pos); free(q.len); free(q.cnt);
        if (rc) return rc;
        out->n_digar[r] = (int32_t)(dtop - out->digar_first[r]); out->n_nreg[r] = (int32_t)g.n_reg;
        rtop += g.n_reg;
    }
    out->n_digar_total = dtop; out->n_alt_total = atop; out->n_nreg_total = rtop;
    return 0;
}

int lcd_oracle_collect_digar_cs(const lcd_digar_input_t *in, const int64_t *cs_off, const char *cs_all, lcd_digar_output_t *out) {
    return collect_tagged(1, in, cs_off, cs_all, NULL, 0, -1, out);
}
int lcd_oracle_collect_digar_refseq(const lcd_digar_input_t *in, const char *ref_seq, int64_t ref_beg, int64_t ref_end, lcd_digar_output_t *out) {
    return collect_tagged(2, in, NULL, NULL, ref_seq, ref_beg, ref_end, out);
}
