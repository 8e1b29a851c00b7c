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
