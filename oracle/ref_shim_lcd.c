/* oracle/ref_shim_lcd.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat-C entry points over the UNMODIFIED longcallD sources (oracle/_ref/liblcdref.so) for the parts of the
 * per-region worker that operate on bam_chunk_t: a chunk is assembled around the caller's flat arrays (only the
 * fields the called function reads), the reference function is called, and its results are copied back out.
 * Compiled only where /root/reference exists (oracle/Makefile, target _ref/libref_shim.so). */
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"
#include "call_var_main.h"
#include "bam_utils.h"
#include "collect_var.h"
#include "assign_hap.h"
#include "cgranges.h"
#include "math_utils.h"

/* globals of the reference's main.c, which is not linked here */
int LONGCALLD_VERBOSE = 0;
const char PROG[20] = "longcallD";
char *CMD = (char *)"";

/* assign_hap_based_on_germline_het_vars_kmeans (src/assign_hap.c:473) on a synthetic chunk */
int ref_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out) {
    const int nr = in->n_reads, nv = in->n_vars;
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.is_ont = in->is_ont;
    bam_chunk_t chunk; memset(&chunk, 0, sizeof(chunk));
    chunk.n_reads = chunk.m_reads = nr;
    chunk.ordered_read_ids = (int*)malloc(sizeof(int) * (nr + 1));
    chunk.is_skipped = (uint8_t*)malloc(nr + 1);
    chunk.haps = (int*)calloc(nr + 1, sizeof(int)); chunk.phase_scores = (int*)calloc(nr + 1, sizeof(int));
    chunk.phase_sets = (hts_pos_t*)calloc(nr + 1, sizeof(hts_pos_t));
    chunk.n_clean_agree_snps = (int*)calloc(nr + 1, sizeof(int)); chunk.n_clean_conflict_snps = (int*)calloc(nr + 1, sizeof(int));
    for (int i = 0; i < nr; ++i) { chunk.ordered_read_ids[i] = in->ordered_read_ids[i]; chunk.is_skipped[i] = in->is_skipped[i]; }
    for (int r = 0; r < nr; ++r) {          /* the output arrays are in/out: what the reference does not touch keeps its value */
        chunk.haps[r] = out->haps[r]; chunk.phase_sets[r] = out->phase_sets[r];
        chunk.n_clean_agree_snps[r] = out->n_clean_agree_snps[r]; chunk.n_clean_conflict_snps[r] = out->n_clean_conflict_snps[r];
    }
    chunk.n_cand_vars = nv;
    chunk.cand_vars = (cand_var_t*)calloc(nv + 1, sizeof(cand_var_t));
    chunk.var_i_to_cate = (int*)malloc(sizeof(int) * (nv + 1));
    for (int v = 0; v < nv; ++v) {
        cand_var_t *c = chunk.cand_vars + v;
        c->pos = in->pos[v]; c->phase_set = -1; c->var_type = in->var_type[v]; c->is_homopolymer_indel = in->is_hp_indel[v];
        c->total_cov = in->total_cov[v]; c->n_uniq_alles = in->n_uniq_alles[v];
        c->alle_covs = (int*)malloc(sizeof(int) * 4);
        for (int i = 0; i < 4; ++i) c->alle_covs[i] = in->alle_covs[4 * v + i];
        c->ref_len = 1; c->alt_len = 1;
        chunk.var_i_to_cate[v] = in->var_cate[v];
    }
    /* read_var_profile + read_var_cr exactly as collect_read_var_profile leaves them (src/collect_var.c:1389-1431) */
    read_var_profile_t *p = (read_var_profile_t*)calloc(nr + 1, sizeof(read_var_profile_t));
    cgranges_t *cr = cr_init();
    for (int r = 0; r < nr; ++r) {
        p[r].read_id = r; p[r].start_var_idx = in->prof_start[r]; p[r].end_var_idx = in->prof_end[r];
        const int n = in->prof_end[r] - in->prof_start[r] + 1;
        p[r].alleles = (int*)malloc(sizeof(int) * (n > 0 ? n : 1)); p[r].alt_qi = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
        for (int i = 0; i < n; ++i) { p[r].alleles[i] = in->alleles[in->allele_off[r] + i]; p[r].alt_qi[i] = -1; }
    }
    for (int i = 0; i < nr; ++i) {
        const int r = in->ordered_read_ids[i];
        if (chunk.is_skipped[r]) continue;
        if (p[r].start_var_idx < 0 || p[r].end_var_idx < 0) continue;
        cr_add(cr, "cr", p[r].start_var_idx, p[r].end_var_idx + 1, r);
    }
    cr_index(cr);
    chunk.read_var_profile = p; chunk.read_var_cr = cr;
    assign_hap_based_on_germline_het_vars_kmeans(&opt, &chunk, in->target_var_cate);
    for (int r = 0; r < nr; ++r) {
        out->haps[r] = chunk.haps[r]; out->phase_sets[r] = chunk.phase_sets[r];
        out->n_clean_agree_snps[r] = chunk.n_clean_agree_snps[r]; out->n_clean_conflict_snps[r] = chunk.n_clean_conflict_snps[r];
    }
    for (int v = 0; v < nv; ++v) {
        cand_var_t *c = chunk.cand_vars + v;
        if (c->hap_to_cons_alle) {
            for (int h = 0; h < 3; ++h) {
                out->hap_to_cons_alle[3 * v + h] = c->hap_to_cons_alle[h];
                for (int a = 0; a < 4; ++a) out->hap_to_alle_profile[12 * v + 4 * h + a] = a < c->n_uniq_alles ? c->hap_to_alle_profile[h][a] : 0;
                free(c->hap_to_alle_profile[h]);
            }
            free(c->hap_to_alle_profile); free(c->hap_to_cons_alle);
            out->var_phase_set[v] = c->phase_set;
        }
        free(c->alle_covs);
    }
    for (int r = 0; r < nr; ++r) { free(p[r].alleles); free(p[r].alt_qi); }
    free(p); cr_destroy(cr);
    free(chunk.cand_vars); free(chunk.var_i_to_cate); free(chunk.ordered_read_ids); free(chunk.is_skipped); free(chunk.haps);
    free(chunk.phase_scores); free(chunk.phase_sets); free(chunk.n_clean_agree_snps); free(chunk.n_clean_conflict_snps);
    return 0;
}

/* cr_index + in-order listing: the order reads come back from cr_overlap */
void ref_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out) {
    cgranges_t *cr = cr_init();
    for (int i = 0; i < n; ++i) cr_add(cr, "cr", start[i], start[i] + 1, label[i]);
    cr_index(cr);
    for (int i = 0; i < n; ++i) order_out[i] = cr_label(cr, i);
    cr_destroy(cr);
}

int collect_cand_vars(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_var_sites, var_site_t *var_sites);   /* src/collect_var.c:238 (no prototype in the headers) */
/* collect_cand_vars (src/collect_var.c:238) on a synthetic chunk: digar_t / var_site_t records around the flat arrays */
int ref_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out) {
    const int nr = in->n_reads, ns = in->n_sites;
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.min_bq = in->min_bq; opt.min_sv_len = in->min_sv_len;
    bam_chunk_t chunk; memset(&chunk, 0, sizeof(chunk));
    chunk.n_reads = chunk.m_reads = nr; chunk.tid = 0;
    chunk.ordered_read_ids = (int*)malloc(sizeof(int) * (nr + 1));
    chunk.is_skipped = (uint8_t*)malloc(nr + 1);
    chunk.digars = (digar_t*)calloc(nr + 1, sizeof(digar_t));
    for (int r = 0; r < nr; ++r) {
        chunk.ordered_read_ids[r] = in->ordered_read_ids[r]; chunk.is_skipped[r] = in->is_skipped[r];
        digar_t *g = chunk.digars + r;
        g->beg = in->read_beg[r]; g->end = in->read_end[r]; g->is_rev = in->read_is_rev[r];
        g->n_digar = g->m_digar = in->n_digar[r];
        g->digars = (digar1_t*)calloc(g->n_digar + 1, sizeof(digar1_t));
        g->qual = (uint8_t*)(in->qual + in->qual_off[r]);
        for (int k = 0; k < g->n_digar; ++k) {
            const int64_t d = in->digar_first[r] + k;
            digar1_t *x = g->digars + k;
            x->pos = in->digar_pos[d]; x->type = in->digar_type[d]; x->len = in->digar_len[d]; x->qi = in->digar_qi[d];
            x->is_low_qual = in->digar_low_qual[d];
            x->alt_seq = (x->type == BAM_CDIFF || x->type == BAM_CINS) ? (uint8_t*)(in->digar_alt + in->digar_alt_off[d]) : NULL;
        }
    }
    var_site_t *sites = (var_site_t*)calloc(ns + 1, sizeof(var_site_t));
    for (int i = 0; i < ns; ++i) {
        sites[i].tid = 0; sites[i].pos = in->site_pos[i]; sites[i].var_type = in->site_type[i];
        sites[i].ref_len = in->site_ref_len[i]; sites[i].alt_len = in->site_alt_len[i];
        sites[i].alt_seq = (uint8_t*)(in->site_alt + in->site_alt_off[i]);
    }
    collect_cand_vars(&opt, &chunk, ns, sites);
    for (int i = 0; i < ns; ++i) {
        cand_var_t *c = chunk.cand_vars + i; int32_t *o = out->site_counts + 8 * i;
        o[0] = c->total_cov; o[1] = c->low_qual_cov; o[2] = c->alle_covs[0]; o[3] = c->alle_covs[1];
        for (int s = 0; s < 2; ++s) for (int a = 0; a < 2; ++a) o[4 + 2 * s + a] = c->strand_to_alle_covs[s][a];
        free(c->alle_covs); free(c->strand_to_alle_covs[0]); free(c->strand_to_alle_covs[1]); free(c->strand_to_alle_covs);
        if (c->alt_seq) free(c->alt_seq);
    }
    free(chunk.cand_vars); free(sites);
    for (int r = 0; r < nr; ++r) free(chunk.digars[r].digars);
    free(chunk.digars); free(chunk.ordered_read_ids); free(chunk.is_skipped);
    return 0;
}

read_var_profile_t *collect_read_var_profile(const call_var_opt_t *opt, bam_chunk_t *chunk);   /* src/collect_var.c:1389 (no prototype in the headers) */
/* collect_read_var_profile (src/collect_var.c:1389) on a synthetic chunk; the rows are returned in the CSR layout of
 * lcd_profile_output_t (row r = the reference's alleles[0 .. end_var_idx - start_var_idx]). */
int ref_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out) {
    const int nr = in->n_reads, nv = in->n_sites;
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.min_bq = in->min_bq; opt.min_sv_len = in->min_sv_len; opt.out_somatic = 0;
    bam_chunk_t chunk; memset(&chunk, 0, sizeof(chunk));
    chunk.n_reads = chunk.m_reads = nr; chunk.tid = 0;
    chunk.ordered_read_ids = (int*)malloc(sizeof(int) * (nr + 1));
    chunk.is_skipped = (uint8_t*)malloc(nr + 1);
    chunk.digars = (digar_t*)calloc(nr + 1, sizeof(digar_t));
    for (int r = 0; r < nr; ++r) {
        chunk.ordered_read_ids[r] = in->ordered_read_ids[r]; chunk.is_skipped[r] = in->is_skipped[r];
        digar_t *g = chunk.digars + r;
        g->beg = in->read_beg[r]; g->end = in->read_end[r]; g->is_rev = in->read_is_rev[r];
        g->n_digar = g->m_digar = in->n_digar[r];
        g->digars = (digar1_t*)calloc(g->n_digar + 1, sizeof(digar1_t));
        g->qual = (uint8_t*)(in->qual + in->qual_off[r]);
        for (int k = 0; k < g->n_digar; ++k) {
            const int64_t d = in->digar_first[r] + k;
            digar1_t *x = g->digars + k;
            x->pos = in->digar_pos[d]; x->type = in->digar_type[d]; x->len = in->digar_len[d]; x->qi = in->digar_qi[d];
            x->is_low_qual = in->digar_low_qual[d];
            x->alt_seq = (x->type == BAM_CDIFF || x->type == BAM_CINS) ? (uint8_t*)(in->digar_alt + in->digar_alt_off[d]) : NULL;
        }
        g->noisy_regs = cr_init();
        for (int64_t k = ex->nreg_first[r]; k < ex->nreg_first[r] + ex->n_nreg[r]; ++k) cr_add(g->noisy_regs, "cr", (int32_t)ex->nreg_beg[k], (int32_t)ex->nreg_end[k], 0);
        cr_index(g->noisy_regs);
    }
    chunk.n_cand_vars = nv;
    chunk.cand_vars = (cand_var_t*)calloc(nv + 1, sizeof(cand_var_t));
    chunk.var_i_to_cate = (int*)malloc(sizeof(int) * (nv + 1));
    for (int v = 0; v < nv; ++v) {
        cand_var_t *c = chunk.cand_vars + v;
        c->tid = 0; c->pos = in->site_pos[v]; c->var_type = in->site_type[v]; c->ref_len = in->site_ref_len[v]; c->alt_len = in->site_alt_len[v];
        c->alt_seq = (uint8_t*)(in->site_alt + in->site_alt_off[v]);
        chunk.var_i_to_cate[v] = ex->var_cate[v];
    }
    read_var_profile_t *p = collect_read_var_profile(&opt, &chunk);
    int64_t top = 0; int rc = 0;
    for (int r = 0; r < nr; ++r) {
        out->prof_start[r] = p[r].start_var_idx; out->prof_end[r] = p[r].end_var_idx; out->allele_off[r] = top;
        const int n = p[r].end_var_idx - p[r].start_var_idx + 1;
        if (p[r].start_var_idx < 0 || n <= 0) continue;
        if (top + n > out->alleles_cap) { rc = -3; break; }
        for (int k = 0; k < n; ++k) { out->alleles[top + k] = (int8_t)p[r].alleles[k]; out->alt_qi[top + k] = p[r].alt_qi[k]; }
        top += n;
    }
    out->n_alleles = top;
    free(p);                                      /* one allocation: init_read_var_profile_inner, src/bam_utils.c:13-36 */
    cr_destroy(chunk.read_var_cr);
    for (int r = 0; r < nr; ++r) { free(chunk.digars[r].digars); cr_destroy(chunk.digars[r].noisy_regs); }
    free(chunk.digars); free(chunk.cand_vars); free(chunk.var_i_to_cate); free(chunk.ordered_read_ids); free(chunk.is_skipped);
    return rc;
}

/* collect_digar_from_eqx_cigar (src/bam_utils.c:701-841) for every listed, not yet skipped read of a synthetic chunk, as
 * collect_digars_from_bam drives it (src/collect_var.c:1063-1082).  Three steps so that a benchmark can time the reference's own
 * work alone: prepare (bam1_t records built with htslib's bam_set1 from the flat CIGAR / packed SEQ / QUAL arrays; reads flagged
 * is_palindrome get an SA tag that passes the reference's 90 % overlap test when opt.is_ont is set), core (the reference calls),
 * finish (results copied out in the layout of lcd_digar_output_t, everything freed). */
#include "htslib/sam.h"
typedef struct { bam_chunk_t chunk; call_var_opt_t opt; uint8_t *ret; } ref_digar_job_t;

void *ref_digar_prepare(const lcd_digar_input_t *in) {
    const int nr = in->n_reads;
    ref_digar_job_t *job = (ref_digar_job_t*)calloc(1, sizeof(ref_digar_job_t));
    call_var_opt_t *opt = &job->opt; bam_chunk_t *chunk = &job->chunk;
    opt->min_bq = in->min_bq; opt->noisy_reg_max_xgaps = in->noisy_reg_max_xgaps; opt->noisy_reg_slide_win = in->noisy_reg_slide_win;
    opt->end_clip_reg = in->end_clip_reg; opt->end_clip_reg_flank_win = in->end_clip_reg_flank_win;
    opt->max_noisy_frac_per_read = in->max_noisy_frac_per_read; opt->max_var_ratio_per_read = in->max_var_ratio_per_read;
    opt->is_ont = 0;
    for (int r = 0; r < nr; ++r) if (in->is_palindrome[r]) opt->is_ont = 1;
    chunk->n_reads = chunk->m_reads = nr; chunk->tid = 0; chunk->tname = (char*)"chr";
    chunk->reg_beg = in->reg_beg; chunk->reg_end = in->reg_end; chunk->whole_ref_len = in->whole_ref_len;
    chunk->qual_counts = (int*)calloc(256, sizeof(int));
    chunk->is_ont_palindrome = (uint8_t*)calloc(nr + 1, 1);
    chunk->digars = (digar_t*)calloc(nr + 1, sizeof(digar_t));
    chunk->reads = (bam1_t**)calloc(nr + 1, sizeof(bam1_t*));
    chunk->chunk_noisy_regs = cr_init();
    job->ret = (uint8_t*)calloc(nr + 1, 1);
    for (int r = 0; r < nr; ++r) {
        const int L = in->l_qseq[r];
        char *seq = (char*)malloc(L + 1), *ql = (char*)malloc(L + 1);
        const uint8_t *bs = in->bseq + in->seq_off[r];
        for (int k = 0; k < L; ++k) { seq[k] = seq_nt16_str[(bs[k >> 1] >> ((~k & 1) << 2)) & 15]; ql[k] = (char)in->qual[in->qual_off[r] + k]; }
        seq[L] = 0;
        char name[32]; snprintf(name, sizeof(name), "r%d", r);
        bam1_t *b = bam_init1();
        if (bam_set1(b, strlen(name), name, in->read_is_rev[r] ? BAM_FREVERSE : 0, 0, in->read_pos0[r], 60, in->n_cigar[r], in->cigar + in->cigar_off[r],
                     -1, -1, 0, L, seq, ql, 64) < 0) return NULL;
        if (in->is_palindrome[r]) {
            char sa[64]; snprintf(sa, sizeof(sa), "chr,%lld,%c,%lldM,60,0;", (long long)in->read_pos0[r] + 1, in->read_is_rev[r] ? '+' : '-',
                                  (long long)(bam_endpos(b) - in->read_pos0[r]));
            bam_aux_append(b, "SA", 'Z', (int)strlen(sa) + 1, (uint8_t*)sa);
        }
        chunk->reads[r] = b; free(seq); free(ql);
    }
    return job;
}

void ref_digar_core(void *h, const lcd_digar_input_t *in) {          /* the reference's own work: the loop of collect_digars_from_bam */
    ref_digar_job_t *job = (ref_digar_job_t*)h;
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        job->ret[r] = collect_digar_from_eqx_cigar(&job->chunk, r, &job->opt, job->chunk.digars + r) < 0;
    }
}

int ref_digar_finish(void *h, const lcd_digar_input_t *in, lcd_digar_output_t *out) {
    ref_digar_job_t *job = (ref_digar_job_t*)h; bam_chunk_t chunk = job->chunk;
    const int nr = in->n_reads;
    int64_t dtop = 0, atop = 0, rtop = 0; int rc = 0;
    for (int i = 0; i < nr && !rc; ++i) {
        const int r = in->ordered_read_ids[i];
        out->skip[r] = 0;
        if (in->is_skipped[r]) continue;
        digar_t *g = chunk.digars + r;
        out->skip[r] = job->ret[r];
        out->read_beg[r] = g->beg; out->read_end[r] = g->end;
        out->digar_first[r] = dtop; out->n_digar[r] = g->n_digar; out->nreg_first[r] = rtop; out->n_nreg[r] = (int32_t)g->noisy_regs->n_r;
        if (chunk.is_ont_palindrome[r] != in->is_palindrome[r]) rc = -8;
        for (int k = 0; k < g->n_digar && !rc; ++k) {
            const digar1_t *x = g->digars + k;
            if (dtop >= out->digar_cap) { rc = -3; break; }
            out->digar_pos[dtop] = x->pos; out->digar_type[dtop] = (int8_t)x->type; out->digar_len[dtop] = x->len; out->digar_qi[dtop] = x->qi;
            out->digar_low_qual[dtop] = x->is_low_qual; out->digar_alt_off[dtop] = atop; dtop++;
            if (x->type == BAM_CDIFF || x->type == BAM_CINS) {
                if (atop + x->len > out->alt_cap) { rc = -3; break; }
                memcpy(out->digar_alt + atop, x->alt_seq, x->len); atop += x->len;
            }
        }
        for (int64_t k = 0; k < g->noisy_regs->n_r && !rc; ++k) {
            if (rtop >= out->nreg_cap) { rc = -4; break; }
            out->nreg_beg[rtop] = cr_start(g->noisy_regs, k); out->nreg_end[rtop] = cr_end(g->noisy_regs, k); out->nreg_label[rtop] = cr_label(g->noisy_regs, k); rtop++;
        }
    }
    out->n_digar_total = dtop; out->n_alt_total = atop; out->n_nreg_total = rtop;
    for (int k = 0; k < 256; ++k) out->qual_counts[k] = chunk.qual_counts[k];
    out->n_cnreg = 0;
    for (int64_t k = 0; k < chunk.chunk_noisy_regs->n_r && !rc; ++k) {
        if (out->n_cnreg >= out->cnreg_cap) { rc = -4; break; }
        /* not indexed yet: x = ctg << 32 | start, y = end (src/cgranges.h:38-42) */
        out->cnreg_beg[out->n_cnreg] = (int32_t)chunk.chunk_noisy_regs->r[k].x; out->cnreg_end[out->n_cnreg] = (int32_t)chunk.chunk_noisy_regs->r[k].y;
        out->cnreg_label[out->n_cnreg] = cr_label(chunk.chunk_noisy_regs, k); out->n_cnreg++;
    }
    for (int r = 0; r < nr; ++r) {
        digar_t *g = chunk.digars + r;
        if (g->digars) { for (int k = 0; k < g->n_digar; ++k) if (g->digars[k].alt_seq) free(g->digars[k].alt_seq); free(g->digars); }
        if (g->noisy_regs) cr_destroy(g->noisy_regs);
        free(g->bseq); free(g->qual);
        bam_destroy1(chunk.reads[r]);
    }
    cr_destroy(chunk.chunk_noisy_regs);
    free(chunk.reads); free(chunk.digars); free(chunk.is_ont_palindrome); free(chunk.qual_counts); free(job->ret); free(job);
    return rc;
}

int ref_collect_digar_eqx(const lcd_digar_input_t *in, lcd_digar_output_t *out) {
    void *job = ref_digar_prepare(in);
    if (!job) return -9;
    ref_digar_core(job, in);
    return ref_digar_finish(job, in, out);
}


int collect_all_cand_var_sites(const call_var_opt_t *opt, bam_chunk_t *chunk, var_site_t **var_sites);   /* src/collect_var.c:1209 (no prototype in the headers) */
/* collect_all_cand_var_sites (src/collect_var.c:1209) on a synthetic chunk: digar_t records around the flat arrays.
 * Three steps so that a batch can time the reference's own call alone: prepare (build the records), core, finish (copy out). */
typedef struct { call_var_opt_t opt; bam_chunk_t chunk; var_site_t *sites; int n; } ref_sites_job_t;
void *ref_sites_prepare(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end) {
    const int nr = in->n_reads;
    ref_sites_job_t *job = (ref_sites_job_t*)calloc(1, sizeof(ref_sites_job_t));
    job->opt.min_bq = in->min_bq; job->opt.min_sv_len = in->min_sv_len;
    bam_chunk_t *chunk = &job->chunk;
    chunk->n_reads = chunk->m_reads = nr; chunk->tid = 0; chunk->tname = (char*)"chr"; chunk->reg_beg = reg_beg; chunk->reg_end = reg_end;
    chunk->ordered_read_ids = (int*)malloc(sizeof(int) * (nr + 1));
    chunk->is_skipped = (uint8_t*)malloc(nr + 1);
    chunk->digars = (digar_t*)calloc(nr + 1, sizeof(digar_t));
    for (int r = 0; r < nr; ++r) {
        chunk->ordered_read_ids[r] = in->ordered_read_ids[r]; chunk->is_skipped[r] = in->is_skipped[r];
        digar_t *g = chunk->digars + r;
        g->n_digar = g->m_digar = in->n_digar[r];
        g->digars = (digar1_t*)calloc(g->n_digar + 1, sizeof(digar1_t));
        for (int k = 0; k < g->n_digar; ++k) {
            const int64_t d = in->digar_first[r] + k; digar1_t *x = g->digars + k;
            x->pos = in->digar_pos[d]; x->type = in->digar_type[d]; x->len = in->digar_len[d]; x->qi = in->digar_qi[d]; x->is_low_qual = in->digar_low_qual[d];
            x->alt_seq = (x->type == BAM_CDIFF || x->type == BAM_CINS) ? (uint8_t*)(in->digar_alt + in->digar_alt_off[d]) : NULL;
        }
    }
    return job;
}
void ref_sites_core(void *h) {                    /* the reference's own work */
    ref_sites_job_t *job = (ref_sites_job_t*)h;
    job->n = collect_all_cand_var_sites(&job->opt, &job->chunk, &job->sites);
}
int ref_sites_finish(void *h, const lcd_pileup_input_t *in, lcd_sites_output_t *out) {
    ref_sites_job_t *job = (ref_sites_job_t*)h; var_site_t *sites = job->sites; const int n = job->n;
    int rc = 0;
    out->n_sites = n;
    if (n > out->cap) rc = -3;
    for (int i = 0; i < n && !rc; ++i) {
        out->site_pos[i] = sites[i].pos; out->site_type[i] = sites[i].var_type; out->site_ref_len[i] = sites[i].ref_len; out->site_alt_len[i] = sites[i].alt_len;
        out->site_src[i] = sites[i].alt_seq ? (int64_t)(sites[i].alt_seq - in->digar_alt) : -1;     /* offset of the alt bases in digar_alt (not a record index) */
    }
    if (sites) free(sites);
    for (int r = 0; r < in->n_reads; ++r) free(job->chunk.digars[r].digars);
    free(job->chunk.digars); free(job->chunk.ordered_read_ids); free(job->chunk.is_skipped); free(job);
    return rc;
}
int ref_collect_sites(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end, lcd_sites_output_t *out) {
    void *job = ref_sites_prepare(in, reg_beg, reg_end);
    ref_sites_core(job);
    return ref_sites_finish(job, in, out);
}

int classify_var_cate(const call_var_opt_t *opt, char *ref_seq, hts_pos_t ref_beg, hts_pos_t ref_end, cand_var_t *var, int min_dp_thres, int min_alt_dp_thres,
                      double min_af_thres, double max_af_thres, double min_noisy_reg_ratio);          /* src/collect_var.c:413 (no prototype in the headers) */
/* classify_var_cate (src/collect_var.c:413) for every site, as the first loop of classify_cand_vars (:915-918) calls it: cand_var_t records around the flat arrays */
int ref_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate) {
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.noisy_reg_max_xgaps = in->max_xgaps; opt.is_ont = in->is_ont; opt.min_dp = in->min_dp; opt.min_alt_dp = in->min_alt_dp; opt.min_af = in->min_af; opt.max_af = in->max_af;
    if (in->is_ont) { opt.strand_bias_pval = LONGCALLD_STRAND_BIAS_PVAL_ONT; initialize_lgamma_cache(&opt); }      /* set_ont_opt, call_var_init_para */
    for (int i = 0; i < in->n_sites; ++i) {
        const int32_t *c = in->site_counts + 8 * (int64_t)i;
        cand_var_t var; memset(&var, 0, sizeof(var));
        int alle_covs[2] = { c[2], c[3] }, s0[2] = { c[4], c[5] }, s1[2] = { c[6], c[7] }; int *strand[2] = { s0, s1 };
        var.pos = in->site_pos[i]; var.var_type = in->site_type[i]; var.total_cov = c[0]; var.low_qual_cov = c[1]; var.n_uniq_alles = 2;
        var.alle_covs = alle_covs; var.strand_to_alle_covs = strand; var.ref_len = in->site_ref_len[i]; var.alt_len = in->site_alt_len[i];
        var.alt_seq = (var.var_type == BAM_CDIFF || var.var_type == BAM_CINS) ? (uint8_t*)(in->site_alt + in->site_alt_off[i]) : NULL;
        var_cate[i] = classify_var_cate(&opt, (char*)in->ref_seq, in->ref_beg, in->ref_end, &var, opt.min_dp, opt.min_alt_dp, opt.min_af, opt.max_af, opt.min_af);
    }
    return 0;
}

/* collect_digar_from_MD_tag (src/bam_utils.c:1003-1174) on the same synthetic chunk: the reads carry plain-M CIGARs and get an MD tag
 * (md + md_off[r], NUL terminated). */
int ref_collect_digar_md(const lcd_digar_input_t *in, const int64_t *md_off, const char *md, lcd_digar_output_t *out) {
    ref_digar_job_t *job = (ref_digar_job_t*)ref_digar_prepare(in);
    if (!job) return -9;
    for (int r = 0; r < in->n_reads; ++r) {
        const char *m = md + md_off[r];
        if (bam_aux_append(job->chunk.reads[r], "MD", 'Z', (int)strlen(m) + 1, (const uint8_t*)m) < 0) return -9;
    }
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        job->ret[r] = collect_digar_from_MD_tag(&job->chunk, r, &job->opt, job->chunk.digars + r) < 0;
    }
    return ref_digar_finish(job, in, out);
}

/* collect_digar_from_ref_seq (src/bam_utils.c:1176-1290) on the same synthetic chunk: plain-M reads without tags against the reference window */
int ref_collect_digar_refseq(const lcd_digar_input_t *in, const char *ref_seq, int64_t ref_beg, int64_t ref_end, lcd_digar_output_t *out) {
    ref_digar_job_t *job = (ref_digar_job_t*)ref_digar_prepare(in);
    if (!job) return -9;
    job->chunk.ref_seq = (char*)ref_seq; job->chunk.ref_beg = ref_beg; job->chunk.ref_end = ref_end;
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        job->ret[r] = collect_digar_from_ref_seq(&job->chunk, r, &job->opt, job->chunk.digars + r) < 0;
    }
    return ref_digar_finish(job, in, out);
}

/* collect_digar_from_cs_tag (src/bam_utils.c:844-1001) on the same synthetic chunk: every read gets a cs tag (cs + cs_off[r], NUL terminated) */
int ref_collect_digar_cs(const lcd_digar_input_t *in, const int64_t *cs_off, const char *cs, lcd_digar_output_t *out) {
    ref_digar_job_t *job = (ref_digar_job_t*)ref_digar_prepare(in);
    if (!job) return -9;
    for (int r = 0; r < in->n_reads; ++r) {
        const char *m = cs + cs_off[r];
        if (bam_aux_append(job->chunk.reads[r], "cs", 'Z', (int)strlen(m) + 1, (const uint8_t*)m) < 0) return -9;
    }
    for (int i = 0; i < in->n_reads; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r]) continue;
        job->ret[r] = collect_digar_from_cs_tag(&job->chunk, r, &job->opt, job->chunk.digars + r) < 0;
    }
    return ref_digar_finish(job, in, out);
}

/* abpoa_aln_msa_cons (src/align.c:872-953) itself: the de-novo POA of a region's fully covering reads with max_n_cons consensus sequences.
 * cons: the consensus sequences back to back; read_clu: cluster of each read; msa: (n_seq + n_cons) rows (reads in input order, then the consensus rows). */
int abpoa_aln_msa_cons(const call_var_opt_t *opt, int n_reads, int *read_ids, uint8_t **read_seqs, int *read_lens, int max_n_cons, int *cons_lens, uint8_t **cons_seqs,
                       int *clu_n_seqs, int **clu_read_ids, int *msa_seq_len, uint8_t ***msa_seq);      /* (no prototype in the headers) */
int ref_poa_ncons(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const lcd_poa_params_t *p, double min_freq,
                  uint8_t *cons, int32_t *cons_len, int32_t *n_cons, uint8_t *read_clu, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.min_af = min_freq; opt.match = p->match; opt.mismatch = p->mismatch;
    opt.gap_open1 = p->gap_open1; opt.gap_ext1 = p->gap_ext1; opt.gap_open2 = p->gap_open2; opt.gap_ext2 = p->gap_ext2;
    int *ids = (int*)malloc(sizeof(int) * (n_seq + 1)), *lens = (int*)malloc(sizeof(int) * (n_seq + 1));
    uint8_t **ptr = (uint8_t**)malloc(sizeof(uint8_t*) * (n_seq + 1));
    for (int i = 0; i < n_seq; ++i) { ids[i] = i; lens[i] = seq_len[i]; ptr[i] = (uint8_t*)seqs + seq_off[i]; }
    int cl[2] = {0, 0}, clu_n[2] = {0, 0}, ml[2] = {0, 0}; uint8_t *cs[2] = {NULL, NULL}; int *clu_ids[2] = {NULL, NULL};
    uint8_t ***ms = (uint8_t***)malloc(2 * sizeof(uint8_t**));
    for (int i = 0; i < 2; ++i) ms[i] = (uint8_t**)calloc(n_seq + 1, sizeof(uint8_t*));
    const int nc = abpoa_aln_msa_cons(&opt, n_seq, ids, ptr, lens, p->max_n_cons, cl, cs, clu_n, clu_ids, ml, ms);
    int rc = 0, at = 0;
    *n_cons = nc; cons_len[0] = cons_len[1] = 0; *msa_len = nc > 0 ? ml[0] : 0;
    memset(read_clu, 0, n_seq);
    if ((int64_t)(n_seq + nc) * (nc > 0 ? ml[0] : 0) > msa_cap) rc = -5;
    for (int c = 0; c < nc; ++c) {
        cons_len[c] = cl[c]; memcpy(cons + at, cs[c], cl[c]); at += cl[c];
        const int cn = nc == 2 ? clu_n[c] : n_seq;
        for (int j = 0; j < cn; ++j) {
            const int r = nc == 2 ? clu_ids[c][j] : clu_ids[0][j];
            read_clu[r] = (uint8_t)c;
            if (rc == 0) memcpy(msa + (size_t)r * ml[0], ms[c][j], ml[0]);
        }
        if (rc == 0) memcpy(msa + (size_t)(n_seq + c) * ml[0], ms[c][cn], ml[0]);
    }
    for (int c = 0; c < 2; ++c) { for (int j = 0; j < n_seq + 1; ++j) free(ms[c][j]); free(ms[c]); free(cs[c]); free(clu_ids[c]); }
    free(ms); free(ids); free(lens); free(ptr);
    return rc;
}

/* pre_process_noisy_regs (src/collect_var.c:557) + classify_cand_vars (:902) themselves, on a chunk assembled from the flat arrays: the candidate
 * sites with their counters (the reference runs its own classify_var_cate on them), the reads' difference lists and noisy intervals as K1 leaves
 * them, chunk_noisy_regs in cr_add order, the low-complexity intervals.  Out: the compacted chunk->cand_vars (kept_pos / type / ref_len / cate,
 * *n_kept) and chunk->chunk_noisy_regs. */
void pre_process_noisy_regs(bam_chunk_t *chunk, call_var_opt_t *opt);
int classify_cand_vars(bam_chunk_t *chunk, int n_var_sites, const call_var_opt_t *opt);
int ref_noisy_regs(const lcd_classify_input_t *ci, const lcd_noisyreg_input_t *in, int64_t *kept_pos, int32_t *kept_type, int32_t *kept_ref_len, int32_t *kept_cate, int32_t *n_kept,
                   lcd_noisyreg_output_t *out) {
    call_var_opt_t opt; memset(&opt, 0, sizeof(opt));
    opt.noisy_reg_max_xgaps = ci->max_xgaps; opt.is_ont = ci->is_ont; opt.min_dp = ci->min_dp; opt.min_alt_dp = ci->min_alt_dp; opt.min_af = ci->min_af; opt.max_af = ci->max_af;
    opt.noisy_reg_merge_dis = LONGCALLD_NOISY_REG_MERGE_DIS; opt.min_sv_len = LONGCALLD_MIN_SV_LEN; opt.noisy_reg_flank_len = in->noisy_reg_flank_len; opt.out_somatic = 0;
    if (ci->is_ont) { opt.strand_bias_pval = LONGCALLD_STRAND_BIAS_PVAL_ONT; initialize_lgamma_cache(&opt); }
    const int n = ci->n_sites, nr = in->n_reads;
    bam_chunk_t chunk; memset(&chunk, 0, sizeof(chunk));
    chunk.tname = (char*)"cr"; chunk.reg_beg = in->reg_beg; chunk.reg_end = in->reg_end;
    chunk.ref_seq = (char*)ci->ref_seq; chunk.ref_beg = ci->ref_beg; chunk.ref_end = ci->ref_end;
    chunk.n_reads = chunk.m_reads = nr;
    chunk.ordered_read_ids = (int*)malloc(sizeof(int) * (nr + 1)); chunk.is_skipped = (uint8_t*)malloc(nr + 1);
    chunk.digars = (digar_t*)calloc(nr + 1, sizeof(digar_t));
    for (int r = 0; r < nr; ++r) {
        chunk.ordered_read_ids[r] = r; chunk.is_skipped[r] = in->is_skipped[r];
        digar_t *d = chunk.digars + r;
        d->beg = in->read_beg[r]; d->end = in->read_end[r]; d->n_digar = d->m_digar = in->n_digar[r];
        d->digars = (digar1_t*)calloc(in->n_digar[r] + 1, sizeof(digar1_t));
        for (int j = 0; j < in->n_digar[r]; ++j) { const int64_t q = in->digar_first[r] + j; d->digars[j].pos = in->digar_pos[q]; d->digars[j].type = in->digar_type[q]; d->digars[j].len = in->digar_len[q]; }
        d->noisy_regs = cr_init();
        for (int x = 0; x < in->n_nreg[r]; ++x) { const int64_t q = in->nreg_first[r] + x; cr_add(d->noisy_regs, "cr", (int32_t)in->nreg_beg[q], (int32_t)in->nreg_end[q], 1); }
        cr_index(d->noisy_regs);
    }
    chunk.chunk_noisy_regs = cr_init();
    for (int64_t i = 0; i < in->n_cnreg; ++i) cr_add(chunk.chunk_noisy_regs, "cr", (int32_t)in->cnreg_beg[i], (int32_t)in->cnreg_end[i], in->cnreg_label[i]);
    if (in->n_low > 0) {
        chunk.low_comp_cr = cr_init();
        for (int64_t i = 0; i < in->n_low; ++i) cr_add(chunk.low_comp_cr, "cr", (int32_t)in->low_beg[i], (int32_t)in->low_end[i], 0);
        cr_index(chunk.low_comp_cr);
    }
    chunk.n_cand_vars = n;
    chunk.cand_vars = (cand_var_t*)calloc(n + 1, sizeof(cand_var_t));
    for (int i = 0; i < n; ++i) {
        const int32_t *c = ci->site_counts + 8 * (int64_t)i;
        cand_var_t *v = chunk.cand_vars + i;
        v->pos = ci->site_pos[i]; v->var_type = ci->site_type[i]; v->total_cov = c[0]; v->low_qual_cov = c[1]; v->n_uniq_alles = 2;
        v->alle_covs = (int*)malloc(2 * sizeof(int)); v->alle_covs[0] = c[2]; v->alle_covs[1] = c[3];
        v->strand_to_alle_covs = (int**)malloc(2 * sizeof(int*));
        for (int s = 0; s < 2; ++s) { v->strand_to_alle_covs[s] = (int*)malloc(2 * sizeof(int)); v->strand_to_alle_covs[s][0] = c[4 + 2 * s]; v->strand_to_alle_covs[s][1] = c[5 + 2 * s]; }
        v->ref_len = ci->site_ref_len[i]; v->alt_len = ci->site_alt_len[i];
        v->te_seq_i = i;          /* the site's index rides along (copy_var copies the field; nothing reads it when tsd_len == 0): tells which sites were kept */
        if (v->var_type == BAM_CDIFF || v->var_type == BAM_CINS) { v->alt_seq = (uint8_t*)malloc(v->alt_len > 0 ? v->alt_len : 1); memcpy(v->alt_seq, ci->site_alt + ci->site_alt_off[i], v->alt_len); }
    }
    pre_process_noisy_regs(&chunk, &opt);
    int nk = 0;
    if (n > 0) nk = classify_cand_vars(&chunk, n, &opt);
    *n_kept = nk;
    for (int i = 0; i < nk; ++i) { kept_pos[i] = chunk.cand_vars[i].pos; kept_type[i] = chunk.cand_vars[i].var_type; kept_ref_len[i] = chunk.cand_vars[i].ref_len; kept_cate[i] = chunk.var_i_to_cate[i]; }
    if (out->keep && out->var_cate) {          /* the same as a mask over the input sites */
        for (int i = 0; i < n; ++i) { out->keep[i] = 0; out->var_cate[i] = -1; }
        for (int i = 0; i < nk; ++i) { const int k = chunk.cand_vars[i].te_seq_i; out->keep[k] = 1; out->var_cate[k] = chunk.var_i_to_cate[i]; }
    }
    int rc = 0;
    out->n_regs = chunk.chunk_noisy_regs ? chunk.chunk_noisy_regs->n_r : 0;
    if (out->n_regs > out->reg_cap) rc = -5;
    else for (int64_t k = 0; k < out->n_regs; ++k) { out->reg_beg[k] = cr_start(chunk.chunk_noisy_regs, k); out->reg_end[k] = cr_end(chunk.chunk_noisy_regs, k); out->reg_label[k] = cr_label(chunk.chunk_noisy_regs, k); }
    for (int i = 0; i < nk; ++i) {          /* the kept entries (classify_cand_vars freed the dropped ones itself) */
        cand_var_t *v = chunk.cand_vars + i;
        free(v->alle_covs); for (int s = 0; s < 2; ++s) free(v->strand_to_alle_covs[s]); free(v->strand_to_alle_covs); free(v->alt_seq);
    }
    if (n > 0) free(chunk.var_i_to_cate);
    else for (int i = 0; i < n; ++i) { cand_var_t *v = chunk.cand_vars + i; free(v->alle_covs); free(v->alt_seq); }
    free(chunk.cand_vars);
    for (int r = 0; r < nr; ++r) { free(chunk.digars[r].digars); cr_destroy(chunk.digars[r].noisy_regs); }
    free(chunk.digars); free(chunk.ordered_read_ids); free(chunk.is_skipped);
    if (chunk.chunk_noisy_regs) cr_destroy(chunk.chunk_noisy_regs);
    if (chunk.low_comp_cr) cr_destroy(chunk.low_comp_cr);
    if (chunk.var_noisy_read_cov_cr) cr_destroy(chunk.var_noisy_read_cov_cr);
    if (chunk.var_noisy_read_err_cr) cr_destroy(chunk.var_noisy_read_err_cr);
    free(chunk.var_noisy_read_marks);
    return rc;
}

/* sdust() itself (src/sdust.c:184) on a sequence: the 0-based half-open intervals it returns */
#include "sdust.h"
int ref_sdust(const uint8_t *seq, int l_seq, int T, int W, int64_t *beg, int64_t *end, int64_t cap) {
    int n = 0;
    uint64_t *r = sdust(0, seq, l_seq, T, W, &n);
    if (n > cap) { free(r); return -1; }
    for (int i = 0; i < n; ++i) { beg[i] = (int64_t)(r[i] >> 32); end[i] = (int64_t)(uint32_t)r[i]; }
    free(r);
    return n;
}
