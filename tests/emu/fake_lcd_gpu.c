/* tests/emu/fake_lcd_gpu.c -- TEST INFRASTRUCTURE: a CPU test double of the C-ABI of include/lcd_gpu.h (the batch entry points the
 * drop-in binds), implemented with the oracle (oracle/liblcd_oracle.so).  It lets the reference-side binding
 * (longcalld_b200/dropin/lcd_dropin.c: marshalling, the coroutine region driver, the cross-thread combiner) be exercised end to end --
 * `longcallD call` -> VCF -- on a box without a GPU.  It is never built into, linked with or loaded by the product. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/lcd_gpu.h"
#define LCD_ORACLE_TYPES_FROM_GPU_HEADER 1

/* oracle entry points (oracle/lcd_oracle.h declares them on the same struct layouts) */
int lcd_oracle_wfa_align(const uint8_t *pattern, int plen, const uint8_t *text, int tlen, const lcd_wfa_params_t *par, char *ops, lcd_wfa_result_t *res);
int lcd_oracle_edlib_align(const uint8_t *query, int qlen, const uint8_t *target, int tlen, int mode, int want_path, uint8_t *aln, lcd_edlib_result_t *res);
int lcd_oracle_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
                   uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
int lcd_oracle_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out);
int lcd_oracle_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out);
int lcd_oracle_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out);
int lcd_oracle_collect_digar_eqx(const lcd_digar_input_t *in, lcd_digar_output_t *out);
int64_t lcd_oracle_md_to_eqx(int n_cigar, const uint32_t *cigar, const char *md, uint32_t *out, int64_t cap);
int lcd_oracle_collect_sites(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end, lcd_sites_output_t *out);

static unsigned long long n_batches;
const char *lcd_gpu_last_error(void) { return "fake_lcd_gpu (oracle-backed test double)"; }
uint64_t lcd_gpu_launch_count(void) { return n_batches; }
void *lcd_gpu_new_stream(void) { return (void*)1; }
void lcd_gpu_set_thread_stream(void *s) { (void)s; }
int lcd_gpu_init(int device, size_t pool_bytes) { (void)device; (void)pool_bytes; return 0; }
int lcd_gpu_split_pool(size_t lower_bytes) { (void)lower_bytes; return 0; }
int lcd_gpu_pool_windows(int a, int b, size_t lower_bytes) { (void)a; (void)b; (void)lower_bytes; return 0; }
int lcd_gpu_reserve_sms(int n) { (void)n; return 0; }

int lcd_digar_capacity(const lcd_digar_input_t *in, int64_t *digar_cap, int64_t *alt_cap, int64_t *nreg_cap) {
    long long nd = 0, na = 0, ni = 0;
    for (int r = 0; r < in->n_reads; ++r) {
        const uint32_t *cg = in->cigar + in->cigar_off[r];
        for (int k = 0; k < in->n_cigar[r]; ++k) {
            const int op = cg[k] & 15; const long long len = cg[k] >> 4;
            if (op == 8) { nd += len; na += len; ni += len; } else if (op == 1) { nd++; na += len; ni++; } else if (op == 2) { nd++; ni++; } else if (op == 7 || op == 4 || op == 5) nd++;
        }
        ni += 2;
    }
    *digar_cap = nd + 1; *alt_cap = na + 1; *nreg_cap = ni + 1;
    return 0;
}
int lcd_digar_batch(int n, const lcd_digar_input_t *in, lcd_digar_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_collect_digar_eqx(in + i, out + i)) return -1;
    return 0;
}
int lcd_digar_md_batch(int n, const lcd_digar_input_t *in, const lcd_md_tags_t *tags, lcd_digar_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) {              /* the MD walk per read, then the =/X pass on the converted CIGARs */
        lcd_digar_input_t x = in[i];
        size_t cap = 16;
        for (int r = 0; r < x.n_reads; ++r) cap += (size_t)x.l_qseq[r] + x.n_cigar[r] + 8;
        uint32_t *cig = (uint32_t*)malloc(cap * 4); int64_t *off = (int64_t*)calloc(x.n_reads + 1, 8); int32_t *nc = (int32_t*)calloc(x.n_reads + 1, 4);
        size_t top = 0;
        for (int r = 0; r < x.n_reads; ++r) {
            off[r] = (int64_t)top;
            const uint32_t *src = x.cigar + x.cigar_off[r];
            if (tags[i].md_off[r] < 0) { memcpy(cig + top, src, 4 * (size_t)x.n_cigar[r]); nc[r] = x.n_cigar[r]; }
            else { const int64_t m = lcd_oracle_md_to_eqx(x.n_cigar[r], src, tags[i].md + tags[i].md_off[r], cig + top, (int64_t)(cap - top)); if (m < 0) { free(cig); free(off); free(nc); return -1; } nc[r] = (int32_t)m; }
            top += (size_t)nc[r];
        }
        x.cigar = cig; x.cigar_off = off; x.n_cigar = nc;
        const int rc = lcd_oracle_collect_digar_eqx(&x, out + i);
        free(cig); free(off); free(nc);
        if (rc) return -1;
    }
    return 0;
}
int lcd_oracle_collect_digar_cs(const lcd_digar_input_t *in, const int64_t *cs_off, const char *cs, lcd_digar_output_t *out);
int lcd_oracle_collect_digar_refseq(const lcd_digar_input_t *in, const char *ref_seq, int64_t ref_beg, int64_t ref_end, lcd_digar_output_t *out);
int lcd_digar_tags_batch(int n, const lcd_digar_input_t *in, const lcd_read_tags_t *tags, lcd_digar_output_t *out) {
    /* the double answers chunks whose tagged reads are all of one variant (what the bundled data holds); =/X and MD reads may mix */
    for (int i = 0; i < n; ++i) {
        int n_cs = 0, n_ref = 0, n_other = 0;
        for (int k = 0; k < in[i].n_reads; ++k) { const int r = in[i].ordered_read_ids[k]; if (in[i].is_skipped[r]) continue; if (tags[i].kind[r] == LCD_TAG_CS) n_cs++; else if (tags[i].kind[r] == LCD_TAG_REFSEQ) n_ref++; else n_other++; }
        if (n_cs && !n_ref && !n_other) { __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED); if (lcd_oracle_collect_digar_cs(in + i, tags[i].off, tags[i].text, out + i)) return -1; }
        else if (n_ref && !n_cs && !n_other) { __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED); if (lcd_oracle_collect_digar_refseq(in + i, tags[i].ref_seq, tags[i].ref_beg, tags[i].ref_end, out + i)) return -1; }
        else if (!n_cs && !n_ref) {
            int64_t *mo = (int64_t*)malloc(sizeof(int64_t) * (in[i].n_reads + 1));
            for (int r = 0; r < in[i].n_reads; ++r) mo[r] = tags[i].kind[r] == LCD_TAG_MD ? tags[i].off[r] : -1;
            lcd_md_tags_t t = { mo, tags[i].text };
            const int rc = lcd_digar_md_batch(1, in + i, &t, out + i);
            free(mo);
            if (rc) return -1;
        } else return -1;
    }
    return 0;
}
int lcd_sites_batch(int n, const lcd_pileup_input_t *in, const lcd_sites_params_t *par, lcd_sites_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) { lcd_pileup_input_t x = in[i]; x.min_sv_len = par[i].min_sv_len; if (lcd_oracle_collect_sites(&x, par[i].reg_beg, par[i].reg_end, out + i)) return -1; }
    return 0;
}
int lcd_pileup_batch(int n, const lcd_pileup_input_t *in, lcd_pileup_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_collect_cand_vars(in + i, out + i)) return -1;
    return 0;
}
int64_t lcd_profile_capacity(const lcd_pileup_input_t *in) {        /* upper bound: every site whose span can touch the read */
    int64_t tot = 0; int mrl = 2;
    for (int s = 0; s < in->n_sites; ++s) if (in->site_ref_len[s] + 2 > mrl) mrl = in->site_ref_len[s] + 2;
    for (int r = 0; r < in->n_reads; ++r) {
        int lo = 0, hi = in->n_sites;
        while (lo < hi) { const int m = (lo + hi) / 2; if (in->site_pos[m] < in->read_beg[r] - mrl) lo = m + 1; else hi = m; }
        int a = lo; hi = in->n_sites;
        while (lo < hi) { const int m = (lo + hi) / 2; if (in->site_pos[m] <= in->read_end[r] + 2) lo = m + 1; else hi = m; }
        tot += lo - a + 2;
    }
    return tot;
}
int lcd_profile_batch(int n, const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_read_var_profile(in + i, ex + i, out + i)) return -1;
    return 0;
}
int lcd_phase_batch(int n, const lcd_phase_input_t *in, lcd_phase_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_assign_hap(in + i, out + i)) return -1;
    return 0;
}
int lcd_wfa_batch(int n, const uint8_t *seqs, size_t seqs_len, const int64_t *po, const int32_t *pl, const int64_t *to, const int32_t *tl,
                  const lcd_wfa_params_t *par, char *ops, const int64_t *oo, lcd_wfa_result_t *res) {
    (void)seqs_len; __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_wfa_align(seqs + po[i], pl[i], seqs + to[i], tl[i], par + i, ops + oo[i], res + i)) return -1;
    return 0;
}
int lcd_edlib_batch(int n, const uint8_t *seqs, size_t seqs_len, const int64_t *qo, const int32_t *ql, const int64_t *to, const int32_t *tl,
                    const int32_t *mode, const int32_t *want, uint8_t *aln, const int64_t *ao, lcd_edlib_result_t *res) {
    (void)seqs_len; __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_edlib_align(seqs + qo[i], ql[i], seqs + to[i], tl[i], mode[i], want[i], aln + ao[i], res + i)) return -1;
    return 0;
}
int lcd_oracle_poa_sub(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                       const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
int lcd_poa_sub_batch(int n, const uint8_t *seqs, size_t seqs_len, const int32_t *first, const int32_t *nr, const int64_t *off, const int32_t *len, int n_total,
                      const int32_t *sub_beg, const int32_t *sub_end,
                      const lcd_poa_params_t *par, uint8_t *cons, const int64_t *coff, uint8_t *msa, const int64_t *moff, const int64_t *mcap, lcd_poa_result_t *res) {
    (void)seqs_len; (void)n_total; __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        int32_t cl = 0, ml = 0;
        const int want = msa && moff && mcap && mcap[i] > 0;
        int64_t cap = 0; for (int k = 0; k < nr[i]; ++k) cap += 2 * (int64_t)len[first[i] + k] + 64;
        cap *= nr[i] + 1;
        uint8_t *tmp = (uint8_t*)malloc((size_t)cap + 16);
        const int rc = lcd_oracle_poa_sub(nr[i], seqs, off + first[i], len + first[i], sub_beg + first[i], sub_end + first[i], par + i, cons + coff[i], &cl, tmp, &ml,
                                          (int32_t)(cap > 0x7fffffff ? 0x7fffffff : cap));
        res[i].status = rc; res[i].cons_len = cl; res[i].msa_len = ml; res[i].n_nodes = 0;
        if (rc == 0 && want) { if ((int64_t)(nr[i] + 1) * ml > mcap[i]) res[i].status = LCD_POA_MSA_CAP; else memcpy(msa + moff[i], tmp, (size_t)(nr[i] + 1) * ml); }
        if (res[i].status) ++bad;
        free(tmp);
    }
    return bad ? -2 : 0;
}
int lcd_poa_batch(int n, const uint8_t *seqs, size_t seqs_len, const int32_t *first, const int32_t *nr, const int64_t *off, const int32_t *len, int n_total,
                  const lcd_poa_params_t *par, uint8_t *cons, const int64_t *coff, uint8_t *msa, const int64_t *moff, const int64_t *mcap, lcd_poa_result_t *res) {
    (void)seqs_len; (void)n_total; __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        int32_t cl = 0, ml = 0;
        const int want = msa && moff && mcap && mcap[i] > 0;
        int64_t cap = 0; for (int k = 0; k < nr[i]; ++k) cap += 2 * (int64_t)len[first[i] + k] + 64;
        cap *= nr[i] + 1;
        uint8_t *tmp = (uint8_t*)malloc((size_t)cap + 16);
        const int rc = lcd_oracle_poa(nr[i], seqs, off + first[i], len + first[i], par + i, cons + coff[i], &cl, tmp, &ml, (int32_t)(cap > 0x7fffffff ? 0x7fffffff : cap));
        res[i].status = rc; res[i].cons_len = cl; res[i].msa_len = ml; res[i].n_nodes = 0;
        if (rc == 0 && want) { if ((int64_t)(nr[i] + 1) * ml > mcap[i]) res[i].status = LCD_POA_MSA_CAP; else memcpy(msa + moff[i], tmp, (size_t)(nr[i] + 1) * ml); }
        if (res[i].status) ++bad;
        free(tmp);
    }
    return bad ? -2 : 0;
}

int lcd_oracle_poa_ncons(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const lcd_poa_params_t *p, double min_freq,
                         uint8_t *cons, int32_t *cons_len, int32_t *n_cons, uint8_t *read_clu, uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
int lcd_poa_ncons_batch(int n, const uint8_t *seqs, size_t seqs_len, const int32_t *first, const int32_t *nr, const int64_t *off, const int32_t *len, int n_total,
                        const lcd_poa_params_t *par, const double *min_freq, uint8_t *cons, const int64_t *coff, uint8_t *msa, const int64_t *moff, const int64_t *mcap,
                        lcd_poa_result_t *res, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster) {
    (void)seqs_len; (void)n_total; __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        int32_t cl[2] = {0, 0}, ml = 0, nc = 0;
        const int want = msa && moff && mcap && mcap[i] > 0;
        int64_t cap = 0; for (int k = 0; k < nr[i]; ++k) cap += 2 * (int64_t)len[first[i] + k] + 64;
        cap *= nr[i] + 2;
        uint8_t *tmp = (uint8_t*)malloc((size_t)cap + 16);
        const int rc = lcd_oracle_poa_ncons(nr[i], seqs, off + first[i], len + first[i], par + i, min_freq[i], cons + coff[i], cl, &nc, read_cluster + first[i], tmp, &ml,
                                            (int32_t)(cap > 0x7fffffff ? 0x7fffffff : cap));
        res[i].status = rc; res[i].cons_len = cl[0]; res[i].msa_len = ml; res[i].n_nodes = 0; n_cons[i] = nc; cons_len2[i] = cl[1];
        if (rc == 0 && want) { if ((int64_t)(nr[i] + nc) * ml > mcap[i]) res[i].status = LCD_POA_MSA_CAP; else memcpy(msa + moff[i], tmp, (size_t)(nr[i] + nc) * ml); }
        if (res[i].status) ++bad;
        free(tmp);
    }
    return bad ? -2 : 0;
}

int lcd_oracle_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate);
int lcd_classify_batch(int n, const lcd_classify_input_t *in, lcd_classify_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_classify_sites(&in[i], out[i].var_cate)) return -1;
    return 0;
}
int lcd_oracle_noisy_regs(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out);
int lcd_noisyreg_batch(int n, const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) {
    __atomic_fetch_add(&n_batches, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) if (lcd_oracle_noisy_regs(&in[i], &out[i])) return -1;
    return 0;
}
