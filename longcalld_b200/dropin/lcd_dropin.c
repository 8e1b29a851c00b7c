/* lcd_dropin.c -- the reference-side binding of liblcd_gpu.so, as a link-time / LD_PRELOAD drop-in.
 *
 * Defines, with the reference's own signatures, the functions of `longcallD call`'s per-region worker that have a
 * B200 implementation, marshals the reference's structures (bam_chunk_t, digar_t, cand_var_t, read_var_profile_t:
 * reference src/bam_utils.h, src/collect_var.h) into the flat views of include/lcd_gpu.h, calls the library, and
 * writes the results back exactly where the reference leaves them:
 *     collect_digars_from_bam                        (src/collect_var.c:1063)  -> lcd_digar_batch / lcd_digar_md_batch (K1; reads with =/X CIGARs or
 *                                                     MD tags -- chunks with cs-tagged or untagged plain-M reads are forwarded to the reference)
 *     collect_all_cand_var_sites                     (src/collect_var.c:1209)  -> lcd_sites_batch    (K1b)
 *     collect_cand_vars                              (src/collect_var.c:238)   -> lcd_pileup_batch   (K2)
 *     collect_read_var_profile                       (src/collect_var.c:1389)  -> lcd_profile_batch  (K3)
 *     assign_hap_based_on_germline_het_vars_kmeans   (src/assign_hap.c:473)    -> lcd_phase_batch    (K4)
 *     edlib_edit_distance / edlib_xgaps / edlib_end2end_aln / edlib_infix_aln (src/align.c:210-275) -> lcd_edlib_batch (K7)
 *     wfa_end2end_aln                                (src/align.c:374)         -> lcd_wfa_batch      (K6)
 *     abpoa_partial_aln_msa_cons                     (src/align.c:762)         -> lcd_poa_batch      (K5; regions whose reads all
 *                                                     cover the region -- partial-cover / sampled regions, which need the sub-graph
 *                                                     alignment the GPU library does not have yet, are forwarded to the reference)
 *     collect_var_main                               (src/collect_var.c:2897)  -> the batched region driver below: the pending noisy regions
 *                                                     of a chunk run side by side (coroutines) and their POA / WFA / edlib problems go to the
 *                                                     library as ONE batch per engine, merged across the reference's worker threads
 * Everything else of the reference runs unchanged.  K1 - K4 are called once per chunk (a batch of one chunk each, on the worker thread's
 * own stream); K5 - K7 are batched over regions and threads.  With this file preloaded the reference writes the same VCF.
 * There is no CPU fallback: a failing library call aborts the run.  (Only the -s somatic profile path, which the GPU
 * library rejects, is forwarded to the reference's own implementation.)
 *
 * Built only where the reference headers exist (make -C longcalld_b200/dropin REF=/root/reference). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "call_var_main.h"
#include "bam_utils.h"
#include "collect_var.h"
#include "assign_hap.h"
#include "cgranges.h"
#include "utils.h"
#include "lcd_gpu.h"

static void die(const char *what) { fprintf(stderr, "[lcd_dropin] %s failed: %s\n", what, lcd_gpu_last_error()); exit(1); }
#include <time.h>
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static double t_batch[3], t_blocked, t_fwd_poa;       /* seconds inside the library per engine (leader threads), blocked in the combiner (all threads), in forwarded abPOA */
static pthread_mutex_t t_mu = PTHREAD_MUTEX_INITIALIZER;
static void t_add(double *acc, double dt) { pthread_mutex_lock(&t_mu); *acc += dt; pthread_mutex_unlock(&t_mu); }
/* LCD_DROPIN_TRACE=<file>: a timeline of the worker threads (chunk begin / end, waits in the combiner, engine batches) as TSV */
typedef struct { double t; unsigned long tid; const char *what; long a, b; } trace_t;
static trace_t *trace_buf; static size_t trace_n, trace_cap; static int trace_on = -1;
static void trace(const char *what, long a, long b) {
    if (trace_on < 0) trace_on = getenv("LCD_DROPIN_TRACE") != NULL;
    if (!trace_on) return;
    const double t = now_s();
    pthread_mutex_lock(&t_mu);
    if (trace_n == trace_cap) { trace_cap = trace_cap ? 2 * trace_cap : 1 << 16; trace_buf = (trace_t*)realloc(trace_buf, trace_cap * sizeof(trace_t)); }
    trace_t e = { t, (unsigned long)pthread_self(), what, a, b }; trace_buf[trace_n++] = e;
    pthread_mutex_unlock(&t_mu);
}
static unsigned long n_calls[16];      /* [12]: de-novo POA problems with two consensus sequences on the GPU, [13]: abpoa_aln_msa_cons calls forwarded, [14] / [15]: chunks whose noisy-region set (K2b + K2c) ran on the GPU / was forwarded */
#define COUNT(i) __atomic_fetch_add(&n_calls[i], 1, __ATOMIC_RELAXED)      /* the reference's worker threads call in concurrently */
__attribute__((destructor)) static void report(void) {
    if (trace_on > 0 && trace_n) { FILE *f = fopen(getenv("LCD_DROPIN_TRACE"), "w"); if (f) { for (size_t i = 0; i < trace_n; ++i) fprintf(f, "%.6f\t%lx\t%s\t%ld\t%ld\n", trace_buf[i].t - trace_buf[0].t, trace_buf[i].tid, trace_buf[i].what, trace_buf[i].a, trace_buf[i].b); fclose(f); } }
    if (getenv("LCD_DROPIN_VERBOSE")) fprintf(stderr, "[lcd_dropin] GPU calls: digar %lu (forwarded: %lu), sites %lu, pileup %lu, noisy-region set %lu (forwarded: %lu), profile %lu, phase %lu, edlib %lu, wfa %lu, poa %lu (with partially covering reads: %lu; de-novo with max_n_cons = 2: %lu; forwarded to abPOA: %lu + %lu) in %lu engine batches (library time: poa %.2f s, wfa %.2f s, edlib %.2f s; threads blocked %.2f s in total; forwarded abPOA %.2f s); kernel launches %llu\n",
                                              n_calls[8], n_calls[9], n_calls[10], n_calls[0], n_calls[14], n_calls[15], n_calls[1], n_calls[2], n_calls[3], n_calls[4], n_calls[5], n_calls[11], n_calls[12], n_calls[6], n_calls[13], n_calls[7], t_batch[0], t_batch[1], t_batch[2], t_blocked, t_fwd_poa, (unsigned long long)lcd_gpu_launch_count());
}

/* LCD_DROPIN_STAGES=engines keeps the pileup scan and the phasing (K1 - K4) on the reference's own host code and sends only the DP engines
 * (K5 - K7, batched) to the GPU: per chunk the reference's AoS structures have to be flattened for, and rebuilt from, every K1 - K4 call,
 * which costs the host about what those stages cost it in the first place (measured: DESIGN.md, whole-program numbers).  Default: all. */
static int pileup_on_gpu(void) { static int v = -1; if (v < 0) { const char *e = getenv("LCD_DROPIN_STAGES"); v = !(e && strcmp(e, "engines") == 0); } return v; }
#define FORWARD_UNLESS_PILEUP_ON_GPU(ret_t, name, proto, args) \
    if (!pileup_on_gpu()) { static ret_t (*orig_) proto = NULL; if (!orig_) orig_ = (ret_t (*) proto)dlsym(RTLD_NEXT, name); return orig_ args; }

/* ------------------------------------------------------------------------------------------ digars -> flat */
typedef struct {
    lcd_pileup_input_t in;
    int64_t *read_beg, *read_end, *digar_first, *qual_off, *digar_pos, *digar_alt_off, *site_pos, *site_alt_off;
    uint8_t *read_is_rev, *qual, *digar_low_qual, *digar_alt, *site_alt;
    int32_t *n_digar, *digar_len, *digar_qi, *site_type, *site_ref_len, *site_alt_len; int8_t *digar_type;
} flat_t;

static void flat_free(flat_t *f) {
    free(f->read_beg); free(f->read_end); free(f->digar_first); free(f->qual_off); free(f->digar_pos); free(f->digar_alt_off); free(f->site_pos);
    free(f->site_alt_off); free(f->read_is_rev); free(f->qual); free(f->digar_low_qual); free(f->digar_alt); free(f->site_alt); free(f->n_digar);
    free(f->digar_len); free(f->digar_qi); free(f->site_type); free(f->site_ref_len); free(f->site_alt_len); free(f->digar_type);
}

/* sites: either var_site_t[] (collect_cand_vars) or cand_var_t[] (collect_read_var_profile) */
static void flatten(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_sites, const var_site_t *sites, const cand_var_t *vars, flat_t *f) {
    const int nr = chunk->n_reads;
    memset(f, 0, sizeof(*f));
    size_t n_ev = 0, n_q = 0, n_alt = 0, n_salt = 0;
    for (int r = 0; r < nr; ++r) {
        const digar_t *g = chunk->digars + r;
        if (chunk->is_skipped[r]) continue;
        n_ev += g->n_digar; n_q += g->qlen;
        for (int k = 0; k < g->n_digar; ++k) if (g->digars[k].type == BAM_CDIFF || g->digars[k].type == BAM_CINS) n_alt += g->digars[k].len;
    }
    for (int i = 0; i < n_sites; ++i) n_salt += sites ? sites[i].alt_len : vars[i].alt_len;
#define A(field, type, n) f->field = (type*)calloc((n) + 1, sizeof(type))
    A(read_beg, int64_t, nr); A(read_end, int64_t, nr); A(digar_first, int64_t, nr); A(qual_off, int64_t, nr); A(read_is_rev, uint8_t, nr); A(n_digar, int32_t, nr);
    A(digar_pos, int64_t, n_ev); A(digar_alt_off, int64_t, n_ev); A(digar_low_qual, uint8_t, n_ev); A(digar_len, int32_t, n_ev); A(digar_qi, int32_t, n_ev);
    A(digar_type, int8_t, n_ev); A(qual, uint8_t, n_q); A(digar_alt, uint8_t, n_alt);
    A(site_pos, int64_t, n_sites); A(site_alt_off, int64_t, n_sites); A(site_type, int32_t, n_sites); A(site_ref_len, int32_t, n_sites);
    A(site_alt_len, int32_t, n_sites); A(site_alt, uint8_t, n_salt);
#undef A
    size_t ev = 0, q = 0, alt = 0;
    for (int r = 0; r < nr; ++r) {
        const digar_t *g = chunk->digars + r;
        f->digar_first[r] = (int64_t)ev; f->qual_off[r] = (int64_t)q;
        if (chunk->is_skipped[r]) continue;                 /* skipped reads carry no usable digars; the kernels never look at them */
        f->read_beg[r] = g->beg; f->read_end[r] = g->end; f->read_is_rev[r] = g->is_rev; f->n_digar[r] = g->n_digar;
        memcpy(f->qual + q, g->qual, g->qlen); q += g->qlen;
        for (int k = 0; k < g->n_digar; ++k, ++ev) {
            const digar1_t *x = g->digars + k;
            f->digar_pos[ev] = x->pos; f->digar_type[ev] = (int8_t)x->type; f->digar_len[ev] = x->len; f->digar_qi[ev] = x->qi;
            f->digar_low_qual[ev] = x->is_low_qual; f->digar_alt_off[ev] = (int64_t)alt;
            if (x->type == BAM_CDIFF || x->type == BAM_CINS) { memcpy(f->digar_alt + alt, x->alt_seq, x->len); alt += x->len; }
        }
    }
    size_t sa = 0;
    for (int i = 0; i < n_sites; ++i) {
        const int type = sites ? sites[i].var_type : vars[i].var_type, alt_len = sites ? sites[i].alt_len : vars[i].alt_len;
        const uint8_t *as = sites ? sites[i].alt_seq : vars[i].alt_seq;
        f->site_pos[i] = sites ? sites[i].pos : vars[i].pos; f->site_type[i] = type; f->site_ref_len[i] = sites ? sites[i].ref_len : vars[i].ref_len;
        f->site_alt_len[i] = alt_len; f->site_alt_off[i] = (int64_t)sa;
        if ((type == BAM_CDIFF || type == BAM_CINS) && as) { memcpy(f->site_alt + sa, as, alt_len); sa += alt_len; }
    }
    lcd_pileup_input_t *in = &f->in;
    in->n_reads = nr; in->n_sites = n_sites; in->min_bq = opt->min_bq; in->min_sv_len = opt->min_sv_len;
    in->ordered_read_ids = chunk->ordered_read_ids; in->is_skipped = chunk->is_skipped;
    in->read_beg = f->read_beg; in->read_end = f->read_end; in->read_is_rev = f->read_is_rev; in->digar_first = f->digar_first; in->n_digar = f->n_digar;
    in->qual_off = f->qual_off; in->qual = f->qual; in->digar_pos = f->digar_pos; in->digar_type = f->digar_type; in->digar_len = f->digar_len;
    in->digar_qi = f->digar_qi; in->digar_low_qual = f->digar_low_qual; in->digar_alt_off = f->digar_alt_off; in->digar_alt = f->digar_alt;
    in->site_pos = f->site_pos; in->site_type = f->site_type; in->site_ref_len = f->site_ref_len; in->site_alt_len = f->site_alt_len;
    in->site_alt_off = f->site_alt_off; in->site_alt = f->site_alt;
}

/* ------------------------------------------------------------------------------------------ K1 */
int is_ont_palindrome_clip(const call_var_opt_t *opt, bam1_t *read);                            /* src/bam_utils.c:662 */
extern int LONGCALLD_VERBOSE;                                                                    /* src/main.c */

void collect_digars_from_bam(bam_chunk_t *chunk, const struct call_var_pl_t *pl) {               /* src/collect_var.c:1063-1110 */
    FORWARD_UNLESS_PILEUP_ON_GPU(void, "collect_digars_from_bam", (bam_chunk_t *, const struct call_var_pl_t *), (chunk, pl))
    const call_var_opt_t *opt = pl->opt;
    const int nr = chunk->n_reads;
    /* the reference picks per read: =/X CIGAR, else cs tag, else MD tag, else the reference sequence (src/collect_var.c:1072-1080); the
       library's tag front end takes the same four variants (lcd_digar_tags_batch) */
    int n_tagged = 0;
    int8_t *kind = (int8_t*)calloc(nr + 1, 1);
    for (int r = 0; r < nr; ++r) {
        kind[r] = LCD_TAG_EQX;
        if (chunk->is_skipped[r] || has_equal_X_in_bam_cigar(chunk->reads[r])) continue;
        kind[r] = has_cs_in_bam(chunk->reads[r]) ? LCD_TAG_CS : has_MD_in_bam(chunk->reads[r]) ? LCD_TAG_MD : LCD_TAG_REFSEQ;
        n_tagged++;
    }
    const int n_md = n_tagged;
    chunk->chunk_noisy_regs = cr_init();
    size_t n_cig = 0, n_seq = 0, n_q = 0;
    for (int r = 0; r < nr; ++r) { const bam1_t *b = chunk->reads[r]; n_cig += b->core.n_cigar; n_seq += ((size_t)b->core.l_qseq + 1) / 2; n_q += b->core.l_qseq; }
    lcd_digar_input_t in; memset(&in, 0, sizeof(in));
    int64_t *pos0 = (int64_t*)calloc(nr + 1, sizeof(int64_t)), *coff = (int64_t*)calloc(nr + 1, sizeof(int64_t)), *soff = (int64_t*)calloc(nr + 1, sizeof(int64_t)), *qoff = (int64_t*)calloc(nr + 1, sizeof(int64_t));
    uint8_t *rev = (uint8_t*)calloc(nr + 1, 1), *pal = (uint8_t*)calloc(nr + 1, 1), *bseq = (uint8_t*)calloc(n_seq + 1, 1), *qual = (uint8_t*)calloc(n_q + 1, 1);
    int32_t *ncig = (int32_t*)calloc(nr + 1, sizeof(int32_t)), *lq = (int32_t*)calloc(nr + 1, sizeof(int32_t)); uint32_t *cig = (uint32_t*)calloc(n_cig + 1, sizeof(uint32_t));
    size_t c = 0, s = 0, q = 0;
    for (int r = 0; r < nr; ++r) {
        bam1_t *b = chunk->reads[r];
        pos0[r] = b->core.pos; rev[r] = bam_is_rev(b); ncig[r] = b->core.n_cigar; lq[r] = b->core.l_qseq; coff[r] = (int64_t)c; soff[r] = (int64_t)s; qoff[r] = (int64_t)q;
        if (!chunk->is_skipped[r]) { pal[r] = (uint8_t)is_ont_palindrome_clip(opt, b); if (pal[r]) chunk->is_ont_palindrome[r] = 1; }
        memcpy(cig + c, bam_get_cigar(b), sizeof(uint32_t) * b->core.n_cigar); c += b->core.n_cigar;
        memcpy(bseq + s, bam_get_seq(b), ((size_t)b->core.l_qseq + 1) / 2); s += ((size_t)b->core.l_qseq + 1) / 2;
        memcpy(qual + q, bam_get_qual(b), b->core.l_qseq); q += b->core.l_qseq;
    }
    in.n_reads = nr; in.min_bq = opt->min_bq; in.noisy_reg_max_xgaps = opt->noisy_reg_max_xgaps; in.noisy_reg_slide_win = opt->noisy_reg_slide_win;
    in.end_clip_reg = opt->end_clip_reg; in.end_clip_reg_flank_win = opt->end_clip_reg_flank_win;
    in.max_noisy_frac_per_read = opt->max_noisy_frac_per_read; in.max_var_ratio_per_read = opt->max_var_ratio_per_read;
    in.whole_ref_len = chunk->whole_ref_len; in.reg_beg = chunk->reg_beg; in.reg_end = chunk->reg_end;
    in.ordered_read_ids = chunk->ordered_read_ids; in.is_skipped = chunk->is_skipped; in.read_pos0 = pos0; in.read_is_rev = rev; in.is_palindrome = pal;
    in.n_cigar = ncig; in.cigar_off = coff; in.cigar = cig; in.l_qseq = lq; in.seq_off = soff; in.bseq = bseq; in.qual_off = qoff; in.qual = qual;
    int64_t dcap = 0, acap = 0, ncap = 0;
    if (lcd_digar_capacity(&in, &dcap, &acap, &ncap)) die("lcd_digar_capacity");
    if (n_md) { dcap += (int64_t)n_q; acap += (int64_t)n_q; ncap += (int64_t)n_q; }          /* an M op expands into up to its length in records */
    lcd_digar_output_t o; memset(&o, 0, sizeof(o));
#define A(field, type, n) o.field = (type*)calloc((size_t)(n) + 1, sizeof(type))
    A(skip, uint8_t, nr); A(read_beg, int64_t, nr); A(read_end, int64_t, nr); A(digar_first, int64_t, nr); A(n_digar, int32_t, nr);
    A(digar_pos, int64_t, dcap); A(digar_type, int8_t, dcap); A(digar_len, int32_t, dcap); A(digar_qi, int32_t, dcap); A(digar_low_qual, uint8_t, dcap);
    A(digar_alt_off, int64_t, dcap); A(digar_alt, uint8_t, acap); A(nreg_first, int64_t, nr); A(n_nreg, int32_t, nr);
    A(nreg_beg, int64_t, ncap); A(nreg_end, int64_t, ncap); A(nreg_label, int32_t, ncap); A(cnreg_beg, int64_t, ncap); A(cnreg_end, int64_t, ncap); A(cnreg_label, int32_t, ncap);
    A(qual_counts, int64_t, 256);
#undef A
    o.digar_cap = dcap; o.alt_cap = acap; o.nreg_cap = ncap; o.cnreg_cap = ncap;
    if (n_tagged == 0) { if (lcd_digar_batch(1, &in, &o)) die("lcd_digar_batch"); }
    else {                       /* tagged / untagged plain-M reads: tags and the reference window go along, the library walks them on the device */
        int64_t *t_off = (int64_t*)calloc(nr + 1, sizeof(int64_t)); size_t t_len = 1;
        for (int r = 0; r < nr; ++r) {
            t_off[r] = -1;
            if (kind[r] != LCD_TAG_CS && kind[r] != LCD_TAG_MD) continue;
            t_off[r] = (int64_t)t_len; t_len += strlen(bam_aux2Z(bam_aux_get(chunk->reads[r], kind[r] == LCD_TAG_CS ? "cs" : "MD"))) + 1;
        }
        char *text = (char*)calloc(t_len + 1, 1);
        for (int r = 0; r < nr; ++r) if (t_off[r] >= 0) strcpy(text + t_off[r], bam_aux2Z(bam_aux_get(chunk->reads[r], kind[r] == LCD_TAG_CS ? "cs" : "MD")));
        lcd_read_tags_t tags = { kind, t_off, text, chunk->ref_seq, chunk->ref_beg, chunk->ref_end };
        const int rc = lcd_digar_tags_batch(1, &in, &tags, &o);
        free(t_off); free(text);
        if (rc && strstr(lcd_gpu_last_error(), "differ from its SEQ")) {
            /* a cs tag that spells other bases than the read's SEQ: the reference takes its alt bases from the tag -- its own path for this chunk */
            static void (*orig)(bam_chunk_t *, const struct call_var_pl_t *) = NULL;
            if (!orig) orig = (void (*)(bam_chunk_t *, const struct call_var_pl_t *))dlsym(RTLD_NEXT, "collect_digars_from_bam");
            free(pos0); free(coff); free(soff); free(qoff); free(rev); free(pal); free(bseq); free(qual); free(ncig); free(lq); free(cig); free(kind);
            free(o.skip); free(o.read_beg); free(o.read_end); free(o.digar_first); free(o.n_digar); free(o.digar_pos); free(o.digar_type); free(o.digar_len); free(o.digar_qi);
            free(o.digar_low_qual); free(o.digar_alt_off); free(o.digar_alt); free(o.nreg_first); free(o.n_nreg); free(o.nreg_beg); free(o.nreg_end); free(o.nreg_label);
            free(o.cnreg_beg); free(o.cnreg_end); free(o.cnreg_label); free(o.qual_counts);
            cr_destroy(chunk->chunk_noisy_regs);
            memset(chunk->is_ont_palindrome, 0, nr);
            COUNT(9);
            orig(chunk, pl);
            return;
        }
        if (rc) die("lcd_digar_tags_batch");
    }
    free(kind);
    COUNT(8);
    for (int i = 0; i < nr; ++i) {
        const int r = chunk->ordered_read_ids[i];
        if (chunk->is_skipped[r]) continue;
        bam1_t *b = chunk->reads[r]; digar_t *g = chunk->digars + r;
        g->beg = o.read_beg[r]; g->end = o.read_end[r]; g->is_rev = rev[r]; g->qlen = b->core.l_qseq;
        const size_t nb = ((size_t)g->qlen + 1) / 2;
        g->bseq = nb ? (uint8_t*)malloc(nb) : NULL; if (nb) memcpy(g->bseq, bam_get_seq(b), nb);
        g->qual = g->qlen ? (uint8_t*)malloc(g->qlen) : NULL; if (g->qlen) memcpy(g->qual, bam_get_qual(b), g->qlen);
        g->n_digar = o.n_digar[r]; g->m_digar = g->n_digar > 0 ? g->n_digar : 1;
        g->digars = (digar1_t*)malloc(sizeof(digar1_t) * g->m_digar);
        for (int k = 0; k < g->n_digar; ++k) {
            const int64_t d = o.digar_first[r] + k; digar1_t *x = g->digars + k;
            x->pos = o.digar_pos[d]; x->type = o.digar_type[d]; x->len = o.digar_len[d]; x->qi = o.digar_qi[d]; x->is_low_qual = o.digar_low_qual[d]; x->alt_seq = NULL;
            if (x->type == BAM_CDIFF || x->type == BAM_CINS) { x->alt_seq = (uint8_t*)malloc(x->len > 0 ? x->len : 1); memcpy(x->alt_seq, o.digar_alt + o.digar_alt_off[d], x->len); }
        }
        g->noisy_regs = cr_init();      /* already in cr_index order: cr_index_prepare finds them sorted and leaves the order alone */
        for (int64_t k = o.nreg_first[r]; k < o.nreg_first[r] + o.n_nreg[r]; ++k) cr_add(g->noisy_regs, "cr", (int32_t)o.nreg_beg[k], (int32_t)o.nreg_end[k], o.nreg_label[k]);
        cr_index(g->noisy_regs);
        if (o.skip[r]) chunk->is_skipped[r] = BAM_RECORD_WRONG_MAP;
    }
    for (int64_t k = 0; k < o.n_cnreg; ++k) cr_add(chunk->chunk_noisy_regs, "cr", (int32_t)o.cnreg_beg[k], (int32_t)o.cnreg_end[k], o.cnreg_label[k]);
    for (int k = 0; k < 256; ++k) chunk->qual_counts[k] += (int)o.qual_counts[k];
    free(pos0); free(coff); free(soff); free(qoff); free(rev); free(pal); free(bseq); free(qual); free(ncig); free(lq); free(cig);
    free(o.skip); free(o.read_beg); free(o.read_end); free(o.digar_first); free(o.n_digar); free(o.digar_pos); free(o.digar_type); free(o.digar_len); free(o.digar_qi);
    free(o.digar_low_qual); free(o.digar_alt_off); free(o.digar_alt); free(o.nreg_first); free(o.n_nreg); free(o.nreg_beg); free(o.nreg_end); free(o.nreg_label);
    free(o.cnreg_beg); free(o.cnreg_end); free(o.cnreg_label); free(o.qual_counts);
    /* the tail of collect_digars_from_bam, verbatim in behaviour: base-quality quartiles of the chunk, then the records are released (:1083-1109) */
    int valid_quals[256], n_valid_quals = 0; int64_t n_total_counts = 0;
    for (int i = 0; i < 256; ++i) n_total_counts += chunk->qual_counts[i];
    for (int i = 0; i < 256; ++i) if (chunk->qual_counts[i] > 0 && chunk->qual_counts[i] >= 0.0001 * n_total_counts) valid_quals[n_valid_quals++] = i;
    if (n_valid_quals == 0) chunk->min_qual = chunk->first_quar_qual = chunk->median_qual = chunk->third_quar_qual = chunk->max_qual = 0;
    else {
        chunk->min_qual = valid_quals[0]; chunk->first_quar_qual = valid_quals[n_valid_quals / 4]; chunk->median_qual = valid_quals[n_valid_quals / 2];
        chunk->third_quar_qual = valid_quals[n_valid_quals * 3 / 4]; chunk->max_qual = valid_quals[n_valid_quals - 1];
    }
    if (LONGCALLD_VERBOSE < 2) {
        for (int i = 0; i < chunk->m_reads; ++i) bam_destroy1(chunk->reads[i]);
        free(chunk->reads);
    }
}

/* ------------------------------------------------------------------------------------------ K1b */
int collect_all_cand_var_sites(const call_var_opt_t *opt, bam_chunk_t *chunk, var_site_t **var_sites) {      /* src/collect_var.c:1209-1254 */
    FORWARD_UNLESS_PILEUP_ON_GPU(int, "collect_all_cand_var_sites", (const call_var_opt_t *, bam_chunk_t *, var_site_t **), (opt, chunk, var_sites))
    *var_sites = NULL;
    flat_t f; flatten(opt, chunk, 0, NULL, NULL, &f);
    size_t n_ev = 0, cap = 0;
    for (int r = 0; r < chunk->n_reads; ++r) if (!chunk->is_skipped[r]) n_ev += chunk->digars[r].n_digar;
    digar1_t **rec = (digar1_t**)malloc((n_ev + 1) * sizeof(digar1_t*));           /* flattened record index -> the reference's record */
    n_ev = 0;
    for (int r = 0; r < chunk->n_reads; ++r) {
        if (chunk->is_skipped[r]) continue;
        for (int k = 0; k < chunk->digars[r].n_digar; ++k) {
            digar1_t *x = chunk->digars[r].digars + k; rec[n_ev++] = x;
            if (x->type == BAM_CDIFF || x->type == BAM_CINS || x->type == BAM_CDEL) cap++;
        }
    }
    lcd_sites_params_t par = { chunk->reg_beg, chunk->reg_end, opt->min_sv_len, 0 };
    lcd_sites_output_t out; memset(&out, 0, sizeof(out));
    out.site_pos = (int64_t*)malloc((cap + 1) * sizeof(int64_t)); out.site_src = (int64_t*)malloc((cap + 1) * sizeof(int64_t));
    out.site_type = (int32_t*)malloc((cap + 1) * sizeof(int32_t)); out.site_ref_len = (int32_t*)malloc((cap + 1) * sizeof(int32_t));
    out.site_alt_len = (int32_t*)malloc((cap + 1) * sizeof(int32_t)); out.cap = (int64_t)cap;
    if (lcd_sites_batch(1, &f.in, &par, &out)) die("lcd_sites_batch");
    const int n = (int)out.n_sites;
    if (n > 0) {
        *var_sites = (var_site_t*)malloc((size_t)n * sizeof(var_site_t));
        for (int i = 0; i < n; ++i) {
            var_site_t v = { chunk->tid, out.site_pos[i], out.site_type[i], out.site_ref_len[i], out.site_alt_len[i], rec[out.site_src[i]]->alt_seq };
            (*var_sites)[i] = v;
        }
    }
    free(out.site_pos); free(out.site_src); free(out.site_type); free(out.site_ref_len); free(out.site_alt_len); free(rec); flat_free(&f);
    COUNT(10);
    return n;
}

/* ------------------------------------------------------------------------------------------ K2 */
cand_var_t *init_cand_vars_based_on_sites(int n_var_sites, var_site_t *var_sites);          /* src/collect_var.c:20 */

int collect_cand_vars(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_var_sites, var_site_t *var_sites) {
    FORWARD_UNLESS_PILEUP_ON_GPU(int, "collect_cand_vars", (const call_var_opt_t *, bam_chunk_t *, int, var_site_t *), (opt, chunk, n_var_sites, var_sites))
    chunk->cand_vars = init_cand_vars_based_on_sites(n_var_sites, var_sites);
    chunk->n_cand_vars = n_var_sites;
    flat_t f; flatten(opt, chunk, n_var_sites, var_sites, NULL, &f);
    int32_t *counts = (int32_t*)calloc(8 * (size_t)n_var_sites + 8, sizeof(int32_t));
    lcd_pileup_output_t out = { counts };
    if (lcd_pileup_batch(1, &f.in, &out)) die("lcd_pileup_batch");
    for (int i = 0; i < n_var_sites; ++i) {
        cand_var_t *c = chunk->cand_vars + i; const int32_t *o = counts + 8 * i;
        c->total_cov = o[0]; c->low_qual_cov = o[1]; c->alle_covs[0] = o[2]; c->alle_covs[1] = o[3];
        for (int s = 0; s < 2; ++s) for (int a = 0; a < 2; ++a) c->strand_to_alle_covs[s][a] = o[4 + 2 * s + a];
    }
    free(counts); flat_free(&f);
    COUNT(0);
    return 0;
}

/* ------------------------------------------------------------------------------------------ K3 */
read_var_profile_t *init_read_var_profile(int n_reads, int n_total_vars);                     /* src/bam_utils.c:38 */

read_var_profile_t *collect_read_var_profile(const call_var_opt_t *opt, bam_chunk_t *chunk) {
    if (opt->out_somatic || !pileup_on_gpu()) {      /* -s: candidate somatic variants take the reference's fuzzy path, which the GPU library rejects */
        static read_var_profile_t *(*orig)(const call_var_opt_t *, bam_chunk_t *) = NULL;
        if (!orig) orig = (read_var_profile_t *(*)(const call_var_opt_t *, bam_chunk_t *))dlsym(RTLD_NEXT, "collect_read_var_profile");
        return orig(opt, chunk);
    }
    const int nr = chunk->n_reads, nv = chunk->n_cand_vars;
    flat_t f; flatten(opt, chunk, nv, NULL, chunk->cand_vars, &f);
    size_t n_iv = 0;
    for (int r = 0; r < nr; ++r) if (!chunk->is_skipped[r] && chunk->digars[r].noisy_regs) n_iv += chunk->digars[r].noisy_regs->n_r;
    int64_t *nfirst = (int64_t*)calloc(nr + 1, sizeof(int64_t)), *nbeg = (int64_t*)calloc(n_iv + 1, sizeof(int64_t)), *nend = (int64_t*)calloc(n_iv + 1, sizeof(int64_t));
    int32_t *nn = (int32_t*)calloc(nr + 1, sizeof(int32_t));
    size_t iv = 0;
    for (int r = 0; r < nr; ++r) {
        nfirst[r] = (int64_t)iv;
        cgranges_t *cr = chunk->digars[r].noisy_regs;
        if (chunk->is_skipped[r] || cr == NULL) continue;
        nn[r] = (int32_t)cr->n_r;
        for (int64_t k = 0; k < cr->n_r; ++k, ++iv) { nbeg[iv] = cr_start(cr, k); nend[iv] = cr_end(cr, k); }
    }
    lcd_profile_extra_t ex = { chunk->var_i_to_cate, nfirst, nn, nbeg, nend };
    const int64_t cap = lcd_profile_capacity(&f.in);
    lcd_profile_output_t out; memset(&out, 0, sizeof(out));
    out.prof_start = (int32_t*)calloc(nr + 1, sizeof(int32_t)); out.prof_end = (int32_t*)calloc(nr + 1, sizeof(int32_t));
    out.allele_off = (int64_t*)calloc(nr + 1, sizeof(int64_t)); out.alleles = (int8_t*)calloc(cap + 1, 1); out.alt_qi = (int32_t*)calloc(cap + 1, sizeof(int32_t));
    out.alleles_cap = cap;
    if (lcd_profile_batch(1, &f.in, &ex, &out)) die("lcd_profile_batch");
    read_var_profile_t *p = init_read_var_profile(nr, nv);
    cgranges_t *read_var_cr = cr_init();
    for (int i = 0; i < nr; ++i) {                       /* same order of cr_add as the reference (src/collect_var.c:1407-1412) */
        const int r = chunk->ordered_read_ids[i];
        if (chunk->is_skipped[r]) continue;
        p[r].start_var_idx = out.prof_start[r]; p[r].end_var_idx = out.prof_end[r];
        for (int v = out.prof_start[r]; v <= out.prof_end[r] && out.prof_start[r] >= 0; ++v) {
            p[r].alleles[v - out.prof_start[r]] = out.alleles[out.allele_off[r] + (v - out.prof_start[r])];
            p[r].alt_qi[v - out.prof_start[r]] = out.alt_qi[out.allele_off[r] + (v - out.prof_start[r])];
        }
        if (p[r].start_var_idx < 0 || p[r].end_var_idx < 0) continue;
        cr_add(read_var_cr, "cr", p[r].start_var_idx, p[r].end_var_idx + 1, r);
    }
    cr_index(read_var_cr); chunk->read_var_cr = read_var_cr;
    free(out.prof_start); free(out.prof_end); free(out.allele_off); free(out.alleles); free(out.alt_qi);
    free(nfirst); free(nbeg); free(nend); free(nn); flat_free(&f);
    COUNT(1);
    return p;
}

/* ------------------------------------------------------------------------------------------ K4 */
int assign_hap_based_on_germline_het_vars_kmeans(const call_var_opt_t *opt, bam_chunk_t *chunk, int target_var_cate) {
    FORWARD_UNLESS_PILEUP_ON_GPU(int, "assign_hap_based_on_germline_het_vars_kmeans", (const call_var_opt_t *, bam_chunk_t *, int), (opt, chunk, target_var_cate))
    const int nr = chunk->n_reads, nv = chunk->n_cand_vars;
    read_var_profile_t *p = chunk->read_var_profile;
    int n_valid = 0;
    for (int v = 0; v < nv; ++v) if (chunk->var_i_to_cate[v] & target_var_cate) n_valid++;
    if (n_valid == 0) return 0;                                            /* src/assign_hap.c:483-486 */
    size_t n_al = 0;
    for (int r = 0; r < nr; ++r) if (p[r].start_var_idx >= 0 && p[r].end_var_idx >= p[r].start_var_idx) n_al += p[r].end_var_idx - p[r].start_var_idx + 1;
    int32_t *ps = (int32_t*)calloc(nr + 1, sizeof(int32_t)), *pe = (int32_t*)calloc(nr + 1, sizeof(int32_t));
    int64_t *ao = (int64_t*)calloc(nr + 1, sizeof(int64_t)); int8_t *al = (int8_t*)calloc(n_al + 1, 1);
    size_t top = 0;
    for (int r = 0; r < nr; ++r) {
        ps[r] = p[r].start_var_idx; pe[r] = p[r].end_var_idx; ao[r] = (int64_t)top;
        if (ps[r] < 0 || pe[r] < ps[r]) continue;
        for (int k = 0; k <= pe[r] - ps[r]; ++k) al[top++] = (int8_t)p[r].alleles[k];
    }
    int32_t *type = (int32_t*)calloc(nv + 1, 4), *hp = (int32_t*)calloc(nv + 1, 4), *nu = (int32_t*)calloc(nv + 1, 4), *covs = (int32_t*)calloc(4 * (size_t)nv + 4, 4),
            *tc = (int32_t*)calloc(nv + 1, 4), *cons = (int32_t*)calloc(3 * (size_t)nv + 3, 4), *prof = (int32_t*)calloc(12 * (size_t)nv + 12, 4);
    int64_t *pos = (int64_t*)calloc(nv + 1, 8), *vps = (int64_t*)calloc(nv + 1, 8);
    for (int v = 0; v < nv; ++v) {
        const cand_var_t *c = chunk->cand_vars + v;
        type[v] = c->var_type; hp[v] = c->is_homopolymer_indel; nu[v] = c->n_uniq_alles; tc[v] = c->total_cov; pos[v] = c->pos; vps[v] = c->phase_set;
        for (int a = 0; a < c->n_uniq_alles && a < 4; ++a) covs[4 * v + a] = c->alle_covs[a];
    }
    lcd_phase_input_t in = { nr, nv, target_var_cate, opt->is_ont, chunk->ordered_read_ids, chunk->is_skipped, ps, pe, ao, al,
                             chunk->var_i_to_cate, type, hp, nu, covs, tc, pos };
    lcd_phase_output_t out = { chunk->haps, (int64_t*)chunk->phase_sets, cons, prof, vps, chunk->n_clean_agree_snps, chunk->n_clean_conflict_snps };
    if (lcd_phase_batch(1, &in, &out)) die("lcd_phase_batch");
    for (int r = 0; r < nr; ++r) chunk->phase_scores[r] = 0;              /* read_init_hap_phase_set, src/assign_hap.c:16-20 */
    for (int v = 0; v < nv; ++v) {
        if ((chunk->var_i_to_cate[v] & target_var_cate) == 0) continue;
        cand_var_t *c = chunk->cand_vars + v;
        if (c->hap_to_alle_profile == NULL) {                             /* var_init_hap_profile_cons_allele, src/assign_hap.c:42-45 */
            c->hap_to_alle_profile = (int**)malloc((LONGCALLD_DEF_PLOID + 1) * sizeof(int*));
            for (int h = 0; h <= LONGCALLD_DEF_PLOID; ++h) c->hap_to_alle_profile[h] = (int*)calloc(c->n_uniq_alles, sizeof(int));
            c->hap_to_cons_alle = (int*)malloc((LONGCALLD_DEF_PLOID + 1) * sizeof(int));
        }
        for (int h = 0; h <= LONGCALLD_DEF_PLOID; ++h) {
            c->hap_to_cons_alle[h] = cons[3 * v + h];
            for (int a = 0; a < c->n_uniq_alles && a < 4; ++a) c->hap_to_alle_profile[h][a] = prof[12 * v + 4 * h + a];
        }
        c->phase_set = vps[v];
    }
    free(ps); free(pe); free(ao); free(al); free(type); free(hp); free(nu); free(covs); free(tc); free(cons); free(prof); free(pos); free(vps);
    COUNT(2);
    return 0;
}

/* ------------------------------------------------------------------------------------------ the batched region driver (a8)
 *
 * The reference aligns one (region, haplotype) at a time from inside collect_noisy_vars1 (src/collect_var.c:2648); a GPU needs thousands
 * of problems per launch.  The reference's host code is kept as it is and run as COROUTINES: every pending noisy region of a chunk gets
 * its own stack (ucontext) on which the unmodified collect_noisy_vars1 runs until it reaches an engine call (abpoa_partial_aln_msa_cons,
 * wfa_end2end_aln, edlib_*: defined below with the reference's signatures).  There the request is parked and the next region runs; when
 * every region of the chunk is parked, the requests of one kind go to the library as ONE lcd_poa_batch / lcd_wfa_batch / lcd_edlib_batch,
 * merged with what the other worker threads (kt_for) have parked meanwhile (leader / follower combiner).  Results are handed back and the
 * regions continue.  What must stay sequential does: make_vars_from_msa_cons_aln + merge_var_profile (they edit the chunk's variant list)
 * run in the reference's region order -- a region waits at the entry of make_vars_from_msa_cons_aln until all regions before it are done.
 * The alignment phase reads only what a pass leaves fixed (difference lists, haplotypes, phase sets: src/align.c:1377-1461), so running
 * the regions of one pass side by side computes exactly what the reference's loop computes. */
#include <pthread.h>
#include <ucontext.h>
#include <unistd.h>
#include "align.h"
#include "abpoa.h"

enum { RQ_POA = 0, RQ_WFA = 1, RQ_EDLIB = 2, RQ_KINDS = 3 };
typedef struct req_t {
    int kind; volatile int done; int rc;
    /* inputs (owned by the caller, alive until done) */
    const uint8_t *seqs; size_t seqs_len;                 /* POA: the reads back to back; WFA: pattern then text; edlib: query then target */
    int n_reads; const int64_t *read_off; const int32_t *read_len; lcd_poa_params_t ppar; int want_msa; int max_len;
    const int32_t *sub_beg, *sub_end;                     /* POA: partially covering reads (NULL: none), see lcd_poa_sub_batch */
    double min_freq; uint8_t *read_clu; int32_t n_cons, cons_len2;   /* POA with ppar.max_n_cons = 2 (de-novo clustering), see lcd_poa_ncons_batch */
    int plen, tlen; lcd_wfa_params_t wpar;
    int qlen, mode, want_path;
    /* outputs (buffers owned by the caller) */
    uint8_t *cons; uint8_t *msa; int64_t msa_cap; lcd_poa_result_t pres;
    char *ops; lcd_wfa_result_t wres;
    uint8_t *aln; lcd_edlib_result_t eres;
} req_t;

static pthread_once_t init_once = PTHREAD_ONCE_INIT;
#define MAX_INFLIGHT 8
static int max_inflight = 3;                               /* engine batches on the GPU at the same time, each in its own windows of the workspace pool */
static void dropin_init(void) {                            /* LCD_DROPIN_DEVICE / _POOL_GB / _INFLIGHT / _RESERVE_SMS: the rank's GPU, the workspace pool, CTA slots left to K1 - K4 */
    const char *d = getenv("LCD_DROPIN_DEVICE"), *g = getenv("LCD_DROPIN_POOL_GB"), *r = getenv("LCD_DROPIN_RESERVE_SMS"), *f = getenv("LCD_DROPIN_INFLIGHT");
    const size_t pool = (size_t)(g ? atof(g) : 48.0) << 30;
    trace("init_begin", 0, 0);
    if (f) max_inflight = atoi(f) < 1 ? 1 : atoi(f) > MAX_INFLIGHT ? MAX_INFLIGHT : atoi(f);
    if (lcd_gpu_init(d ? atoi(d) : 0, pool)) die("lcd_gpu_init");
    /* K5 below, K6 / K7 above: the engines of one batch run side by side, and so do max_inflight batches */
    if (lcd_gpu_pool_windows(max_inflight, max_inflight, pool / 8 * 5)) die("lcd_gpu_pool_windows");
    if (lcd_gpu_reserve_sms(r ? atoi(r) : 8)) die("lcd_gpu_reserve_sms");
    trace("init_end", 0, 0);
}
/* CUDA context creation and the pool allocation take 0.8 - 1.9 s (measured on B200: tools/init_probe.py): started when the library is
 * loaded, on a thread of its own, they overlap the reference's start-up and the first chunks' BAM decoding */
static void *init_thread(void *a) { (void)a; pthread_once(&init_once, dropin_init); return NULL; }
__attribute__((constructor)) static void start_init(void) {
    if (getenv("LCD_DROPIN_LAZY_INIT")) return;
    pthread_t t; if (pthread_create(&t, NULL, init_thread, NULL) == 0) pthread_detach(t);
}
static __thread void *tl_stream = NULL;
static void *new_thread_stream(void) {                    /* every host thread that calls the library gets a stream of its own */
    pthread_once(&init_once, dropin_init);
    if (!tl_stream) { tl_stream = lcd_gpu_new_stream(); if (!tl_stream) die("lcd_gpu_new_stream"); lcd_gpu_set_thread_stream(tl_stream); }
    return tl_stream;
}

/* one library call over n requests of one kind */
static void run_batch(int kind, req_t **r, int n) {
    const double t0_ = now_s();
    trace("batch_begin", kind, n);
    new_thread_stream();
    if (kind == RQ_POA) {
        int n2 = 0;
        for (int i = 0; i < n; ++i) n2 += r[i]->ppar.max_n_cons == 2;
        if (n2 > 0 && n2 < n) {                              /* de-novo problems (two consensus sequences) go in a library call of their own */
            req_t **a = (req_t**)malloc(sizeof(req_t*) * n); int na = 0, nb = 0;
            for (int i = 0; i < n; ++i) if (r[i]->ppar.max_n_cons == 2) a[na++] = r[i];
            for (int i = 0; i < n; ++i) if (r[i]->ppar.max_n_cons != 2) a[na + nb++] = r[i];
            run_batch(kind, a, na); run_batch(kind, a + na, nb);
            free(a);
            return;
        }
        size_t tot = 0, n_rd = 0, cons_tot = 0, msa_tot = 0;
        for (int i = 0; i < n; ++i) { tot += r[i]->seqs_len; n_rd += r[i]->n_reads; cons_tot += r[i]->seqs_len + 16; msa_tot += r[i]->want_msa ? (size_t)r[i]->msa_cap : 0; }
        uint8_t *seqs = (uint8_t*)malloc(tot + 1), *cons = (uint8_t*)malloc(cons_tot + 1), *msa = (uint8_t*)malloc(msa_tot + 1);
        int32_t *first = (int32_t*)malloc(sizeof(int32_t) * n), *nr = (int32_t*)malloc(sizeof(int32_t) * n), *len = (int32_t*)malloc(sizeof(int32_t) * (n_rd + 1));
        int64_t *off = (int64_t*)malloc(sizeof(int64_t) * (n_rd + 1)), *coff = (int64_t*)malloc(sizeof(int64_t) * n), *moff = (int64_t*)malloc(sizeof(int64_t) * n), *mcap = (int64_t*)malloc(sizeof(int64_t) * n);
        int32_t *sb = (int32_t*)calloc(n_rd + 1, sizeof(int32_t)), *se = (int32_t*)calloc(n_rd + 1, sizeof(int32_t)); int any_sub = 0;
        lcd_poa_params_t *par = (lcd_poa_params_t*)malloc(sizeof(lcd_poa_params_t) * n);
        lcd_poa_result_t *res = (lcd_poa_result_t*)calloc(n, sizeof(lcd_poa_result_t));
        size_t o = 0, rd = 0, co = 0, mo = 0;
        for (int i = 0; i < n; ++i) {
            memcpy(seqs + o, r[i]->seqs, r[i]->seqs_len);
            first[i] = (int32_t)rd; nr[i] = r[i]->n_reads; par[i] = r[i]->ppar; coff[i] = (int64_t)co; moff[i] = (int64_t)mo; mcap[i] = r[i]->want_msa ? r[i]->msa_cap : 0;
            for (int k = 0; k < r[i]->n_reads; ++k, ++rd) { off[rd] = (int64_t)o + r[i]->read_off[k]; len[rd] = r[i]->read_len[k]; if (r[i]->sub_beg) { sb[rd] = r[i]->sub_beg[k]; se[rd] = r[i]->sub_end[k]; any_sub = 1; } }
            o += r[i]->seqs_len; co += r[i]->seqs_len + 16; mo += (size_t)mcap[i];
        }
        trace("lib_begin", kind, n);
        double *mf = NULL; int32_t *ncs = NULL, *cl2 = NULL; uint8_t *rclu = NULL;
        if (n2) { mf = (double*)malloc(sizeof(double) * n); ncs = (int32_t*)calloc(n, sizeof(int32_t)); cl2 = (int32_t*)calloc(n, sizeof(int32_t)); rclu = (uint8_t*)calloc(n_rd + 1, 1); for (int i = 0; i < n; ++i) mf[i] = r[i]->min_freq; }
        const int rc = n2 ? lcd_poa_ncons_batch(n, seqs, tot, first, nr, off, len, (int)n_rd, par, mf, cons, coff, msa, moff, mcap, res, ncs, cl2, rclu)
                     : any_sub ? lcd_poa_sub_batch(n, seqs, tot, first, nr, off, len, (int)n_rd, sb, se, par, cons, coff, msa, moff, mcap, res)
                               : lcd_poa_batch(n, seqs, tot, first, nr, off, len, (int)n_rd, par, cons, coff, msa, moff, mcap, res);
        trace("lib_end", kind, n);
        if (rc == -1) die("lcd_poa_batch");                /* -2: some problems were refused on the device (their status says why) */
        for (int i = 0; i < n; ++i) {
            r[i]->pres = res[i]; r[i]->rc = rc;
            if (res[i].status == LCD_POA_OK) {
                const int nc = n2 ? ncs[i] : 1, l2 = n2 ? cl2[i] : 0;
                memcpy(r[i]->cons, cons + coff[i], (size_t)res[i].cons_len + l2);
                if (r[i]->want_msa) memcpy(r[i]->msa, msa + moff[i], (size_t)(r[i]->n_reads + nc) * res[i].msa_len);
                if (n2) { r[i]->n_cons = nc; r[i]->cons_len2 = l2; memcpy(r[i]->read_clu, rclu + first[i], r[i]->n_reads); }
            }
        }
        free(mf); free(ncs); free(cl2); free(rclu);
        free(seqs); free(cons); free(msa); free(first); free(nr); free(len); free(off); free(coff); free(moff); free(mcap); free(par); free(res); free(sb); free(se);
        __atomic_fetch_add(&n_calls[5], (unsigned long)n, __ATOMIC_RELAXED); COUNT(7);
    } else if (kind == RQ_WFA) {
        size_t tot = 0, ops_tot = 0;
        for (int i = 0; i < n; ++i) { tot += r[i]->seqs_len; ops_tot += 2 * r[i]->seqs_len + 16; }
        uint8_t *seqs = (uint8_t*)malloc(tot + 1); char *ops = (char*)malloc(ops_tot + 1);
        int64_t *po = (int64_t*)malloc(sizeof(int64_t) * n), *to = (int64_t*)malloc(sizeof(int64_t) * n), *oo = (int64_t*)malloc(sizeof(int64_t) * n);
        int32_t *pl = (int32_t*)malloc(sizeof(int32_t) * n), *tl = (int32_t*)malloc(sizeof(int32_t) * n);
        lcd_wfa_params_t *par = (lcd_wfa_params_t*)malloc(sizeof(lcd_wfa_params_t) * n);
        lcd_wfa_result_t *res = (lcd_wfa_result_t*)calloc(n, sizeof(lcd_wfa_result_t));
        size_t o = 0, op = 0;
        for (int i = 0; i < n; ++i) {
            memcpy(seqs + o, r[i]->seqs, r[i]->seqs_len);
            po[i] = (int64_t)o; pl[i] = r[i]->plen; to[i] = (int64_t)o + r[i]->plen; tl[i] = r[i]->tlen; oo[i] = (int64_t)op; par[i] = r[i]->wpar;
            o += r[i]->seqs_len; op += 2 * r[i]->seqs_len + 16;
        }
        trace("lib_begin", kind, n);
        if (lcd_wfa_batch(n, seqs, tot, po, pl, to, tl, par, ops, oo, res)) die("lcd_wfa_batch");
        trace("lib_end", kind, n);
        for (int i = 0; i < n; ++i) { r[i]->wres = res[i]; memcpy(r[i]->ops, ops + oo[i], res[i].n_ops > 0 ? (size_t)res[i].n_ops : 0); }
        free(seqs); free(ops); free(po); free(to); free(oo); free(pl); free(tl); free(par); free(res);
        __atomic_fetch_add(&n_calls[4], (unsigned long)n, __ATOMIC_RELAXED); COUNT(7);
    } else {
        size_t tot = 0;
        for (int i = 0; i < n; ++i) tot += r[i]->seqs_len + 2;
        uint8_t *seqs = (uint8_t*)malloc(tot + 1), *aln = (uint8_t*)malloc(tot + 1);
        int64_t *qo = (int64_t*)malloc(sizeof(int64_t) * n), *to = (int64_t*)malloc(sizeof(int64_t) * n), *ao = (int64_t*)malloc(sizeof(int64_t) * n);
        int32_t *ql = (int32_t*)malloc(sizeof(int32_t) * n), *tl = (int32_t*)malloc(sizeof(int32_t) * n), *md = (int32_t*)malloc(sizeof(int32_t) * n), *wp = (int32_t*)malloc(sizeof(int32_t) * n);
        lcd_edlib_result_t *res = (lcd_edlib_result_t*)calloc(n, sizeof(lcd_edlib_result_t));
        size_t o = 0;
        for (int i = 0; i < n; ++i) {
            memcpy(seqs + o, r[i]->seqs, r[i]->seqs_len);
            qo[i] = (int64_t)o; ql[i] = r[i]->qlen; to[i] = (int64_t)o + r[i]->qlen; tl[i] = r[i]->tlen; ao[i] = (int64_t)o; md[i] = r[i]->mode; wp[i] = r[i]->want_path;
            o += r[i]->seqs_len + 2;
        }
        if (lcd_edlib_batch(n, seqs, tot, qo, ql, to, tl, md, wp, aln, ao, res)) die("lcd_edlib_batch");
        for (int i = 0; i < n; ++i) { r[i]->eres = res[i]; if (r[i]->want_path && res[i].aln_len > 0) memcpy(r[i]->aln, aln + ao[i], (size_t)res[i].aln_len); }
        free(seqs); free(aln); free(qo); free(to); free(ao); free(ql); free(tl); free(md); free(wp); free(res);
        __atomic_fetch_add(&n_calls[3], (unsigned long)n, __ATOMIC_RELAXED); COUNT(7);
    }
    for (int i = 0; i < n; ++i) __atomic_store_n(&r[i]->done, 1, __ATOMIC_RELEASE);
    trace("batch_end", kind, n);
    t_add(&t_batch[kind], now_s() - t0_);
}

/* Group-commit combiner over ALL worker threads (kt_for) and all three engines.  Whoever has parked requests adds them to the queue; up to
 * max_inflight RUNNER threads take whatever the queue holds and make one library call per engine (the three engines of a batch side by
 * side: K5 in a window of the lower part of the workspace pool, K6 / K7 in a window of the upper part -- lcd_gpu_pool_windows).  While every
 * runner is on the GPU the queue keeps growing, so the batch size follows the GPU's round-trip time: the slower a round, the wider the
 * next one.  A waiting worker does not sleep if it can help it: kt_for (below) keeps several CHUNKS in flight per worker thread and
 * switches to another one. */
typedef struct { int n; cand_var_t *vars; int *cate; read_var_profile_t *p; int *map; } pend_merge_t;
typedef struct { pend_merge_t *v; int n, cap, on; } merge_tl_t;
static __thread merge_tl_t tl_merge;                       /* a13: the pass's queued merge_var_profile calls (further down) */
static struct { pthread_mutex_t mu; pthread_cond_t cv_work, cv_done; req_t **pend; int n, cap; unsigned long done_gen; } cq =
    { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0 };
static int linger_us(void) { static int v = -1; if (v < 0) { const char *e = getenv("LCD_DROPIN_LINGER_US"); v = e ? atoi(e) : 300; } return v; }

typedef struct { int kind, n, slot; req_t **r; } kind_job_t;
static struct { void *st[RQ_KINDS]; } slots[MAX_INFLIGHT];          /* a batch in flight: one stream per engine */
static void run_kind(kind_job_t *j) {
    void *mine = tl_stream;
    if (!slots[j->slot].st[j->kind]) { slots[j->slot].st[j->kind] = lcd_gpu_new_stream(); if (!slots[j->slot].st[j->kind]) die("lcd_gpu_new_stream"); }
    tl_stream = slots[j->slot].st[j->kind]; lcd_gpu_set_thread_stream(tl_stream);
    run_batch(j->kind, j->r, j->n);
    tl_stream = mine; lcd_gpu_set_thread_stream(mine);
}
static void *kind_thread(void *a) { pthread_once(&init_once, dropin_init); run_kind((kind_job_t*)a); return NULL; }

static void *runner_main(void *a) {
    const int slot = (int)(long)a;
    pthread_once(&init_once, dropin_init);
    for (;;) {
        pthread_mutex_lock(&cq.mu);
        while (cq.n == 0) pthread_cond_wait(&cq.cv_work, &cq.mu);
        if (linger_us() > 0) { pthread_mutex_unlock(&cq.mu); usleep(linger_us()); pthread_mutex_lock(&cq.mu); }       /* a burst arrives over a few hundred microseconds */
        req_t **take = cq.pend; const int nt = cq.n;
        cq.pend = NULL; cq.n = cq.cap = 0;
        pthread_mutex_unlock(&cq.mu);
        if (nt == 0) continue;                              /* another runner took the burst */
        req_t **byk = (req_t**)malloc(sizeof(req_t*) * nt);
        kind_job_t job[RQ_KINDS]; pthread_t th[RQ_KINDS]; int started[RQ_KINDS], m = 0, n_kinds = 0;
        for (int k = 0; k < RQ_KINDS; ++k) { job[k].kind = k; job[k].slot = slot; job[k].r = byk + m; job[k].n = 0; for (int i = 0; i < nt; ++i) if (take[i]->kind == k) { byk[m++] = take[i]; job[k].n++; } if (job[k].n) n_kinds++; }
        for (int k = RQ_KINDS - 1; k >= 1; --k) started[k] = job[k].n > 0 && n_kinds > 1 && pthread_create(&th[k], NULL, kind_thread, &job[k]) == 0;
        if (job[0].n) run_kind(&job[0]);
        for (int k = 1; k < RQ_KINDS; ++k) { if (started[k]) pthread_join(th[k], NULL); else if (job[k].n) run_kind(&job[k]); }
        free(byk); free(take);
        pthread_mutex_lock(&cq.mu);
        cq.done_gen++;
        pthread_cond_broadcast(&cq.cv_done);
        pthread_mutex_unlock(&cq.mu);
    }
    return NULL;
}
static pthread_once_t runners_once = PTHREAD_ONCE_INIT;
static void start_runners(void) {
    pthread_once(&init_once, dropin_init);
    for (long k = 0; k < max_inflight; ++k) { pthread_t t; if (pthread_create(&t, NULL, runner_main, (void*)k) != 0) die("pthread_create"); pthread_detach(t); }
}

/* ---- chunks in flight: kt_for (src/kthread.c:48) with several chunks per worker thread.
 * The reference gives every worker thread one chunk at a time (call_var_main.c:773); its thread would sleep through every engine batch.
 * This kt_for runs each call of the worker function on a stack of its own (ucontext) and keeps up to LCD_DROPIN_CHUNKS_PER_THREAD of them
 * going per thread: a chunk that waits for its requests yields, and the thread loads / scans / post-processes another chunk meanwhile.
 * A thread's chunks share its tid (the reference's per-thread BAM handle): they only ever switch inside combine(), never inside the loader. */
enum { CK_FREE = 0, CK_READY, CK_WAITING, CK_DONE };
struct sched_t; struct co_t;
typedef struct chunk_co_t { ucontext_t ctx; void *stack; int state; long i; req_t **wait; int n_wait; struct sched_t *sv_sched; struct co_t *sv_co; merge_tl_t sv_merge; } chunk_co_t;
typedef struct { ucontext_t main; chunk_co_t *cur; void (*func)(void*, long, int); void *data; int tid; long n; long *next; } kworker_t;
static __thread kworker_t *tl_kw = NULL;
#define CHUNK_STACK ((size_t)8 << 20)
static int chunks_per_thread(void) { static int v = -1; if (v < 0) { const char *e = getenv("LCD_DROPIN_CHUNKS_PER_THREAD"); v = e ? atoi(e) : 2; if (v < 1) v = 1; if (v > 16) v = 16; } return v; }
static int all_done(req_t **r, int n) { for (int i = 0; i < n; ++i) if (!__atomic_load_n(&r[i]->done, __ATOMIC_ACQUIRE)) return 0; return 1; }

static void combine(req_t **r, int n) {
    if (n == 0) return;
    const double t0_ = now_s();
    trace("wait_begin", n, 0);
    pthread_once(&runners_once, start_runners);
    pthread_mutex_lock(&cq.mu);
    if (cq.n + n > cq.cap) { cq.cap = 2 * (cq.n + n); cq.pend = (req_t**)realloc(cq.pend, sizeof(req_t*) * cq.cap); }
    memcpy(cq.pend + cq.n, r, sizeof(req_t*) * n); cq.n += n;
    pthread_cond_signal(&cq.cv_work);
    if (tl_kw && tl_kw->cur) {                              /* a chunk of kt_for: let the thread work on another chunk meanwhile */
        pthread_mutex_unlock(&cq.mu);
        chunk_co_t *c = tl_kw->cur; c->wait = r; c->n_wait = n;
        while (!all_done(r, n)) { c->state = CK_WAITING; swapcontext(&c->ctx, &tl_kw->main); }
    } else {
        while (!all_done(r, n)) pthread_cond_wait(&cq.cv_done, &cq.mu);
        pthread_mutex_unlock(&cq.mu);
    }
    trace("wait_end", n, 0);
    t_add(&t_blocked, now_s() - t0_);
}

/* ---- coroutines: one per pending noisy region of the chunk a worker thread is on */
enum { CO_READY, CO_PARKED, CO_TURN, CO_DONE };
typedef struct co_t { ucontext_t ctx; void *stack; int state, idx, reg_i, ret; req_t **reqs; int n_reqs; bam_chunk_t *chunk; const call_var_opt_t *opt; } co_t;
typedef struct sched_t { ucontext_t main; co_t *cos; int n, next_turn; } sched_t;
static __thread sched_t *tl_sched = NULL;
static __thread co_t *tl_co = NULL;
#define CO_STACK ((size_t)1 << 20)

int collect_noisy_vars1(bam_chunk_t *chunk, const call_var_opt_t *opt, int noisy_reg_i);                                   /* src/collect_var.c:2648 */
static void co_entry(void) {
    co_t *c = tl_co;
    c->ret = collect_noisy_vars1(c->chunk, c->opt, c->reg_i);
    c->state = CO_DONE;
    swapcontext(&c->ctx, &tl_sched->main);
}

/* an engine call: parked when made from a region's coroutine, a batch of one (merged with other threads' requests) otherwise */
static void gpu_call_many(req_t **rs, int n) {             /* independent requests of one region: they travel in the same batch */
    if (n <= 0) return;
    for (int i = 0; i < n; ++i) rs[i]->done = 0;
    if (tl_co) { tl_co->reqs = rs; tl_co->n_reqs = n; tl_co->state = CO_PARKED; swapcontext(&tl_co->ctx, &tl_sched->main); }
    else combine(rs, n);
}
static void gpu_call(req_t *r) { req_t *one = r; gpu_call_many(&one, 1); }

/* coroutine stacks come from a free list of the OS thread: the chunks a thread keeps in flight (kt_for) run their regions at the same time */
static __thread void **free_stacks = NULL; static __thread int n_free_stacks = 0, cap_free_stacks = 0;
static void *co_stack_get(void) { return n_free_stacks > 0 ? free_stacks[--n_free_stacks] : malloc(CO_STACK); }
static void co_stack_put(void *st) {
    if (n_free_stacks == cap_free_stacks) { cap_free_stacks = 2 * cap_free_stacks + 64; free_stacks = (void**)realloc(free_stacks, sizeof(void*) * cap_free_stacks); }
    free_stacks[n_free_stacks++] = st;
}

/* fork / join inside a region: independent pieces of one region's work (the two haplotypes' POA, then their two WFA alignments) run on child
 * stacks until each has parked its engine requests; the region then parks with the union, so the pieces cost one round trip, not one each. */
typedef struct { void (*fn)(void *); void *arg; } task_t;
static __thread task_t *tl_tasks = NULL;
static void task_entry(void) {
    co_t *c = tl_co;
    tl_tasks[c->idx].fn(tl_tasks[c->idx].arg);
    c->state = CO_DONE;
    swapcontext(&c->ctx, &tl_sched->main);
}
static void fork_join(task_t *tasks, int n) {
    if (!tl_co || n == 1) { for (int i = 0; i < n; ++i) tasks[i].fn(tasks[i].arg); return; }       /* not inside a region coroutine: one after the other */
    sched_t *outer_sched = tl_sched; co_t *outer_co = tl_co; task_t *outer_tasks = tl_tasks;
    sched_t sc; memset(&sc, 0, sizeof(sc));
    sc.cos = (co_t*)calloc(n, sizeof(co_t)); sc.n = n;
    for (int i = 0; i < n; ++i) {
        co_t *c = sc.cos + i;
        c->stack = co_stack_get(); c->state = CO_READY; c->idx = i;
        getcontext(&c->ctx); c->ctx.uc_stack.ss_sp = c->stack; c->ctx.uc_stack.ss_size = CO_STACK; c->ctx.uc_link = NULL;
        makecontext(&c->ctx, task_entry, 0);
    }
    req_t **parked = NULL; int cap = 0;
    for (;;) {
        for (int i = 0; i < n; ++i) if (sc.cos[i].state == CO_READY) { tl_sched = &sc; tl_tasks = tasks; tl_co = sc.cos + i; swapcontext(&sc.main, &sc.cos[i].ctx); }
        int np = 0;
        for (int i = 0; i < n; ++i) if (sc.cos[i].state == CO_PARKED) {
            if (np + sc.cos[i].n_reqs > cap) { cap = 2 * (np + sc.cos[i].n_reqs) + 16; parked = (req_t**)realloc(parked, sizeof(req_t*) * cap); }
            for (int k = 0; k < sc.cos[i].n_reqs; ++k) parked[np++] = sc.cos[i].reqs[k];
        }
        if (np == 0) break;
        tl_sched = outer_sched; tl_co = outer_co; tl_tasks = outer_tasks;
        gpu_call_many(parked, np);                          /* the region parks with all the pieces' requests */
        for (int i = 0; i < n; ++i) if (sc.cos[i].state == CO_PARKED) sc.cos[i].state = CO_READY;
    }
    tl_sched = outer_sched; tl_co = outer_co; tl_tasks = outer_tasks;
    for (int i = 0; i < n; ++i) co_stack_put(sc.cos[i].stack);
    free(parked); free(sc.cos);
}

/* all pending regions of one pass side by side; ret[k] = collect_noisy_vars1's return value for regs[k] */
static void run_regions(bam_chunk_t *chunk, const call_var_opt_t *opt, int n, const int *regs, int *ret) {
    void **stack_pool = (void**)malloc(sizeof(void*) * n);
    for (int i = 0; i < n; ++i) stack_pool[i] = co_stack_get();
    sched_t sc; memset(&sc, 0, sizeof(sc));
    sc.cos = (co_t*)calloc(n, sizeof(co_t)); sc.n = n; sc.next_turn = 0;
    for (int i = 0; i < n; ++i) {
        co_t *c = sc.cos + i;
        c->stack = stack_pool[i]; c->state = CO_READY; c->idx = i; c->reg_i = regs[i]; c->chunk = chunk; c->opt = opt;
        getcontext(&c->ctx); c->ctx.uc_stack.ss_sp = c->stack; c->ctx.uc_stack.ss_size = CO_STACK; c->ctx.uc_link = NULL;
        makecontext(&c->ctx, co_entry, 0);
    }
    req_t **parked = NULL; int parked_cap = 0;
    tl_sched = &sc;
    for (;;) {
        int ran = 0;
        for (int i = 0; i < n; ++i) {
            co_t *c = sc.cos + i;
            if (c->state == CO_READY || (c->state == CO_TURN && i == sc.next_turn)) {
                tl_co = c; swapcontext(&sc.main, &c->ctx); tl_co = NULL; ran = 1;
                while (sc.next_turn < n && sc.cos[sc.next_turn].state == CO_DONE) sc.next_turn++;
            }
        }
        int np = 0;
        for (int i = 0; i < n; ++i) if (sc.cos[i].state == CO_PARKED) {
            if (np + sc.cos[i].n_reqs > parked_cap) { parked_cap = 2 * (np + sc.cos[i].n_reqs) + 64; parked = (req_t**)realloc(parked, sizeof(req_t*) * parked_cap); }
            for (int k = 0; k < sc.cos[i].n_reqs; ++k) parked[np++] = sc.cos[i].reqs[k];
        }
        if (np) {
            combine(parked, np);
            for (int i = 0; i < n; ++i) if (sc.cos[i].state == CO_PARKED) sc.cos[i].state = CO_READY;
            continue;
        }
        if (sc.next_turn >= n) break;
        if (!ran) { fprintf(stderr, "[lcd_dropin] region scheduler stalled\n"); exit(1); }
    }
    tl_sched = NULL;
    for (int i = 0; i < n; ++i) ret[i] = sc.cos[i].ret;
    for (int i = 0; i < n; ++i) co_stack_put(stack_pool[i]);
    free(stack_pool);
    free(parked); free(sc.cos);
}

/* kt_for with several chunks in flight per worker thread (see above) */
static void chunk_entry(void) {
    kworker_t *kw = tl_kw; chunk_co_t *c = kw->cur;
    kw->func(kw->data, c->i, kw->tid);
    c->state = CK_DONE;
    swapcontext(&c->ctx, &kw->main);
}
static void *kworker_main(void *a) {
    kworker_t *kw = (kworker_t*)a;
    tl_kw = kw;
    const int K = chunks_per_thread();
    chunk_co_t *sl = (chunk_co_t*)calloc(K, sizeof(chunk_co_t));
    int exhausted = 0;
    for (;;) {
        pthread_mutex_lock(&cq.mu); const unsigned long gen = cq.done_gen; pthread_mutex_unlock(&cq.mu);
        int progressed = 0, n_live = 0;
        for (int k = 0; k < K; ++k) {
            chunk_co_t *c = sl + k;
            if (c->state == CK_WAITING && all_done(c->wait, c->n_wait)) c->state = CK_READY;
            if (c->state == CK_FREE && !exhausted) {
                const long i = __atomic_fetch_add(kw->next, 1, __ATOMIC_RELAXED);
                if (i >= kw->n) exhausted = 1;
                else {
                    if (!c->stack) c->stack = malloc(CHUNK_STACK);
                    getcontext(&c->ctx); c->ctx.uc_stack.ss_sp = c->stack; c->ctx.uc_stack.ss_size = CHUNK_STACK; c->ctx.uc_link = NULL;
                    makecontext(&c->ctx, chunk_entry, 0);
                    c->i = i; c->state = CK_READY; c->sv_sched = NULL; c->sv_co = NULL; memset(&c->sv_merge, 0, sizeof(c->sv_merge));
                }
            }
            if (c->state == CK_READY) {
                tl_sched = c->sv_sched; tl_co = c->sv_co; tl_merge = c->sv_merge;          /* the chunk's thread-local state travels with it */
                kw->cur = c; swapcontext(&kw->main, &c->ctx); kw->cur = NULL;
                c->sv_sched = tl_sched; c->sv_co = tl_co; c->sv_merge = tl_merge;
                tl_sched = NULL; tl_co = NULL; memset(&tl_merge, 0, sizeof(tl_merge));
                progressed = 1;
                if (c->state == CK_DONE) { free(c->sv_merge.v); c->state = CK_FREE; }
            }
            if (c->state != CK_FREE) n_live++;
        }
        if (n_live == 0 && exhausted) break;
        if (!progressed) {                                   /* every chunk of this thread waits for the GPU */
            pthread_mutex_lock(&cq.mu);
            if (cq.done_gen == gen) pthread_cond_wait(&cq.cv_done, &cq.mu);
            pthread_mutex_unlock(&cq.mu);
        }
    }
    for (int k = 0; k < K; ++k) free(sl[k].stack);
    free(sl);
    tl_kw = NULL;
    return NULL;
}
void kt_for(int n_threads, void (*func)(void*, long, int), void *data, long n) {                      /* src/kthread.c:48-66 */
    if (n_threads < 1) n_threads = 1;
    long next = 0;
    kworker_t *kw = (kworker_t*)calloc(n_threads, sizeof(kworker_t)); pthread_t *th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) { kw[t].func = func; kw[t].data = data; kw[t].tid = t; kw[t].n = n; kw[t].next = &next; }
    if (n_threads == 1) kworker_main(&kw[0]);
    else {
        for (int t = 0; t < n_threads; ++t) if (pthread_create(&th[t], NULL, kworker_main, &kw[t]) != 0) die("pthread_create");
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    free(kw); free(th);
}

/* make_vars_from_msa_cons_aln (src/collect_var.c:2279) starts the part of a region that edits the chunk's variant list: regions take it
 * in the reference's order */
int make_vars_from_msa_cons_aln(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_reads, int *read_ids, hts_pos_t noisy_reg_beg, int n_cons, int *clu_n_seqs,
                                int **clu_read_ids, aln_str_t **aln_strs, cand_var_t **noisy_vars, int **noisy_var_cate, read_var_profile_t **p) {
    typedef int (*fn_t)(const call_var_opt_t *, bam_chunk_t *, int, int *, hts_pos_t, int, int *, int **, aln_str_t **, cand_var_t **, int **, read_var_profile_t **);
    static fn_t orig = NULL;
    if (!orig) orig = (fn_t)dlsym(RTLD_NEXT, "make_vars_from_msa_cons_aln");
    if (tl_co) while (tl_sched->next_turn != tl_co->idx) { tl_co->state = CO_TURN; swapcontext(&tl_co->ctx, &tl_sched->main); }
    return orig(opt, chunk, n_reads, read_ids, noisy_reg_beg, n_cons, clu_n_seqs, clu_read_ids, aln_strs, noisy_vars, noisy_var_cate, p);
}

/* ------------------------------------------------------------------------------------------ a13: merge_var_profile, once per pass
 * The reference folds every region's new variants into the chunk's list as soon as the region is done (merge_var_profile,
 * src/collect_var.c:1298-1385) and rebuilds the chunk's whole dense read x variant matrix each time (init_read_var_profile: n_reads x
 * n_vars x 8 bytes allocated and filled per REGION -- a quarter of the reference's CPU time on a 30x HiFi BAM).  Nothing between two merges
 * of a pass reads the merged state (make_vars_from_msa_cons_aln and the alignment phase work on the region's own arrays), so the regions'
 * results are queued here and folded in once, after the last region of the pass: the variant lists by the same sequence of two-way merges
 * (same comparator, same "the earlier one wins" rule for equal variants), the profile by one pass per read over its sources.  The chunk
 * ends the pass with the same variants, categories, profile values and spans, and read_var_cr as after the reference's region-by-region merges. */
int exact_comp_cand_var(const call_var_opt_t *opt, cand_var_t *var1, cand_var_t *var2);     /* src/collect_var.c:1255 */
void free_cand_vars1(cand_var_t *cand_vars);                                                /* :54 */
void free_read_var_profile(read_var_profile_t *p, int n_reads);                             /* src/bam_utils.c:46 */

int merge_var_profile(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_new_vars, cand_var_t *new_vars, int *new_var_cate, read_var_profile_t *new_p) {
    if (!tl_merge.on) {
        static int (*orig)(const call_var_opt_t *, bam_chunk_t *, int, cand_var_t *, int *, read_var_profile_t *) = NULL;
        if (!orig) orig = (int (*)(const call_var_opt_t *, bam_chunk_t *, int, cand_var_t *, int *, read_var_profile_t *))dlsym(RTLD_NEXT, "merge_var_profile");
        return orig(opt, chunk, n_new_vars, new_vars, new_var_cate, new_p);
    }
    if (n_new_vars <= 0) return 0;
    if (tl_merge.n == tl_merge.cap) { tl_merge.cap = tl_merge.cap ? 2 * tl_merge.cap : 64; tl_merge.v = (pend_merge_t*)realloc(tl_merge.v, sizeof(pend_merge_t) * tl_merge.cap); }
    pend_merge_t m = { n_new_vars, new_vars, new_var_cate, new_p, NULL };
    tl_merge.v[tl_merge.n++] = m;
    return n_new_vars;
}

static void flush_merges(const call_var_opt_t *opt, bam_chunk_t *chunk) {
    const int np = tl_merge.n;
    if (np == 0) return;
    const int n_old = chunk->n_cand_vars;
    int total = n_old;
    for (int j = 0; j < np; ++j) total += tl_merge.v[j].n;
    /* the variant list: np two-way merges; (src, idx) of every entry is kept so that the final position of every source variant is known */
    cand_var_t *cur = (cand_var_t*)malloc(sizeof(cand_var_t) * (total ? total : 1)), *nxt = (cand_var_t*)malloc(sizeof(cand_var_t) * (total ? total : 1));
    int *ccate = (int*)malloc(sizeof(int) * (total ? total : 1)), *ncate = (int*)malloc(sizeof(int) * (total ? total : 1));
    int *csrc = (int*)malloc(sizeof(int) * 2 * (total ? total : 1)), *nsrc = (int*)malloc(sizeof(int) * 2 * (total ? total : 1));
    int nc = n_old;
    for (int i = 0; i < n_old; ++i) { cur[i] = chunk->cand_vars[i]; ccate[i] = chunk->var_i_to_cate[i]; csrc[2 * i] = -1; csrc[2 * i + 1] = i; }
    for (int j = 0; j < np; ++j) {
        pend_merge_t *pm = tl_merge.v + j;
        int a = 0, b = 0, o = 0;
#define TAKE_CUR() do { nxt[o] = cur[a]; ncate[o] = ccate[a]; nsrc[2 * o] = csrc[2 * a]; nsrc[2 * o + 1] = csrc[2 * a + 1]; ++o; ++a; } while (0)
#define TAKE_NEW() do { nxt[o] = pm->vars[b]; ncate[o] = pm->cate[b]; nsrc[2 * o] = j; nsrc[2 * o + 1] = b; ++o; ++b; } while (0)
        while (a < nc && b < pm->n) {
            const int ret = exact_comp_cand_var(opt, cur + a, pm->vars + b);
            if (ret < 0) TAKE_CUR();
            else if (ret > 0) TAKE_NEW();
            else { TAKE_CUR(); free_cand_vars1(pm->vars + b); ++b; }          /* equal: the one already in the list stays */
        }
        while (a < nc) TAKE_CUR();
        while (b < pm->n) TAKE_NEW();
#undef TAKE_CUR
#undef TAKE_NEW
        { cand_var_t *t = cur; cur = nxt; nxt = t; int *u = ccate; ccate = ncate; ncate = u; u = csrc; csrc = nsrc; nsrc = u; }
        nc = o;
    }
    int *map_old = (int*)malloc(sizeof(int) * (n_old ? n_old : 1));
    for (int j = 0; j < np; ++j) { tl_merge.v[j].map = (int*)malloc(sizeof(int) * tl_merge.v[j].n); for (int i = 0; i < tl_merge.v[j].n; ++i) tl_merge.v[j].map[i] = -1; }
    for (int f = 0; f < nc; ++f) { if (csrc[2 * f] < 0) map_old[csrc[2 * f + 1]] = f; else tl_merge.v[csrc[2 * f]].map[csrc[2 * f + 1]] = f; }
    /* the profile: per read the hull of its sources' spans in final positions, then every source entry at its final position */
    read_var_profile_t *old_p = chunk->read_var_profile;
    read_var_profile_t *mp = init_read_var_profile(chunk->n_reads, total);
    cgranges_t *cr = cr_init();
    for (int i = 0; i < chunk->n_reads; ++i) {
        const int rd = chunk->ordered_read_ids[i];
        if (chunk->is_skipped[rd]) continue;
        int lo = INT32_MAX, hi = -1;
        for (int pass = 0; pass < 2; ++pass) {
            for (int j = -1; j < np; ++j) {
                const read_var_profile_t *sp = j < 0 ? (old_p ? old_p + rd : NULL) : tl_merge.v[j].p + rd;
                const int *map = j < 0 ? map_old : tl_merge.v[j].map;
                if (sp == NULL || sp->start_var_idx < 0 || sp->end_var_idx < sp->start_var_idx) continue;
                for (int v = sp->start_var_idx; v <= sp->end_var_idx; ++v) {
                    const int f = map[v];
                    if (f < 0) continue;
                    if (pass == 0) { if (f < lo) lo = f; if (f > hi) hi = f; }
                    else { mp[rd].alleles[f - lo] = sp->alleles[v - sp->start_var_idx]; mp[rd].alt_qi[f - lo] = sp->alt_qi[v - sp->start_var_idx]; }
                }
            }
            if (hi < 0) break;
            if (pass == 0) { mp[rd].start_var_idx = lo; mp[rd].end_var_idx = hi; }
        }
        if (mp[rd].start_var_idx < 0 || mp[rd].end_var_idx < 0) continue;
        cr_add(cr, "cr", mp[rd].start_var_idx, mp[rd].end_var_idx + 1, rd);
    }
    cr_index(cr);
    if (n_old > 0) free_read_var_profile(old_p, chunk->n_reads);
    for (int j = 0; j < np; ++j) { free_read_var_profile(tl_merge.v[j].p, chunk->n_reads); free(tl_merge.v[j].vars); free(tl_merge.v[j].cate); free(tl_merge.v[j].map); }
    free(chunk->var_i_to_cate); free(chunk->cand_vars); cr_destroy(chunk->read_var_cr);
    chunk->read_var_profile = mp; chunk->cand_vars = cur; chunk->n_cand_vars = nc; chunk->var_i_to_cate = ccate; chunk->read_var_cr = cr;
    free(nxt); free(ncate); free(csrc); free(nsrc); free(map_old);
    tl_merge.n = 0;
}

/* collect_var_main (src/collect_var.c:2897-2981): the same sequence of steps; step 4's inner loop over the pending regions runs them side by side */
void pre_process_noisy_regs(bam_chunk_t *chunk, call_var_opt_t *opt);                                                       /* src/collect_var.c:557 */
int classify_cand_vars(bam_chunk_t *chunk, int n_var_sites, const call_var_opt_t *opt);                                     /* :902 */
int *sort_noisy_regs(bam_chunk_t *chunk);                                                                                   /* :2745 */
void collect_somatic_var(bam_chunk_t *chunk, const call_var_opt_t *opt);                                                    /* :2857 */


/* ------------------------------------------------------------------------------------------ K2b + K2c
 * pre_process_noisy_regs (src/collect_var.c:557) + classify_cand_vars (:902) of a chunk on the GPU (out_somatic = 0): the candidate sites with their
 * counters go through lcd_classify_batch (the category loop), then sites + categories, the reads' spans / records / noisy intervals, the chunk's noisy
 * list and the low-complexity intervals through lcd_noisyreg_batch; chunk->chunk_noisy_regs, the compacted chunk->cand_vars and chunk->var_i_to_cate
 * are rebuilt from the answer exactly as classify_cand_vars leaves them (copy_var :436, the frees of :1028-1034).  Returns 0 when the chunk was
 * handed to the reference's own functions instead (a site too close to the reference window's ends for K2b). */
void copy_var(cand_var_t *to_var, cand_var_t *from_var);                                                                      /* src/collect_var.c:436 */
static int classify_on_gpu(bam_chunk_t *chunk, int n_var_sites, call_var_opt_t *opt) {
    const int n = n_var_sites, nr = chunk->n_reads;
    cand_var_t *cv = chunk->cand_vars;
    int rc_ok = 1;
    /* K2b */
    int32_t *cate = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    int64_t *spos = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t)), *saoff = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    int32_t *stype = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t)), *sref = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t)), *salt = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *counts = (int32_t*)calloc(8 * (size_t)n + 8, sizeof(int32_t));
    size_t na = 0;
    for (int i = 0; i < n; ++i) na += (size_t)cv[i].alt_len;
    uint8_t *alt = (uint8_t*)calloc(na + 1, 1); na = 0;
    for (int i = 0; i < n; ++i) {
        spos[i] = cv[i].pos; stype[i] = cv[i].var_type; sref[i] = cv[i].ref_len; salt[i] = cv[i].alt_len; saoff[i] = (int64_t)na;
        if ((cv[i].var_type == BAM_CDIFF || cv[i].var_type == BAM_CINS) && cv[i].alt_seq) { memcpy(alt + na, cv[i].alt_seq, cv[i].alt_len); na += cv[i].alt_len; }
        int32_t *o = counts + 8 * (size_t)i;
        o[0] = cv[i].total_cov; o[1] = cv[i].low_qual_cov; o[2] = cv[i].alle_covs[0]; o[3] = cv[i].alle_covs[1];
        for (int st = 0; st < 2; ++st) for (int a = 0; a < 2; ++a) o[4 + 2 * st + a] = cv[i].strand_to_alle_covs[st][a];
    }
    if (n > 0) {
        lcd_classify_input_t ci; memset(&ci, 0, sizeof(ci));
        ci.n_sites = n; ci.min_dp = opt->min_dp; ci.min_alt_dp = opt->min_alt_dp; ci.max_xgaps = opt->noisy_reg_max_xgaps; ci.is_ont = opt->is_ont; ci.min_af = opt->min_af; ci.max_af = opt->max_af;
        ci.ref_beg = chunk->ref_beg; ci.ref_end = chunk->ref_end; ci.ref_seq = chunk->ref_seq;
        ci.site_pos = spos; ci.site_type = stype; ci.site_ref_len = sref; ci.site_alt_len = salt; ci.site_alt_off = saoff; ci.site_alt = alt; ci.site_counts = counts;
        lcd_classify_output_t co = { cate };
        if (lcd_classify_batch(1, &ci, &co)) rc_ok = 0;           /* (a small indel within the margin of the window's ends: the reference reads it unchecked) */
    }
    if (rc_ok) {
        /* K2c inputs */
        size_t nd = 0, nn = 0;
        for (int r = 0; r < nr; ++r) if (!chunk->is_skipped[r]) { nd += chunk->digars[r].n_digar; nn += chunk->digars[r].noisy_regs ? chunk->digars[r].noisy_regs->n_r : 0; }
        int64_t *rb = (int64_t*)calloc((size_t)nr + 1, 8), *re = (int64_t*)calloc((size_t)nr + 1, 8), *df = (int64_t*)calloc((size_t)nr + 1, 8), *nf = (int64_t*)calloc((size_t)nr + 1, 8);
        int32_t *ndg = (int32_t*)calloc((size_t)nr + 1, 4), *nng = (int32_t*)calloc((size_t)nr + 1, 4);
        int64_t *dpos = (int64_t*)calloc(nd + 1, 8), *nb = (int64_t*)calloc(nn + 1, 8), *ne = (int64_t*)calloc(nn + 1, 8);
        int8_t *dtype = (int8_t*)calloc(nd + 1, 1); int32_t *dlen = (int32_t*)calloc(nd + 1, 4);
        size_t d = 0, q = 0;
        for (int r = 0; r < nr; ++r) {
            df[r] = (int64_t)d; nf[r] = (int64_t)q;
            if (chunk->is_skipped[r]) continue;
            const digar_t *g = chunk->digars + r;
            rb[r] = g->beg; re[r] = g->end; ndg[r] = g->n_digar;
            for (int k = 0; k < g->n_digar; ++k, ++d) { dpos[d] = g->digars[k].pos; dtype[d] = (int8_t)g->digars[k].type; dlen[d] = g->digars[k].len; }
            if (g->noisy_regs) { nng[r] = (int32_t)g->noisy_regs->n_r; for (int64_t k = 0; k < g->noisy_regs->n_r; ++k, ++q) { nb[q] = cr_start(g->noisy_regs, k); ne[q] = cr_end(g->noisy_regs, k); } }
        }
        cgranges_t *cn = chunk->chunk_noisy_regs, *lc = chunk->low_comp_cr;
        const int64_t ncn = cn ? cn->n_r : 0, nl = lc ? lc->n_r : 0;
        if (ncn > 0) cr_index(cn);                  /* (as pre_process_noisy_regs does first: before cr_index the entries hold (contig, start) / end, not start / end) */
        int64_t *cb = (int64_t*)calloc((size_t)ncn + 1, 8), *ce = (int64_t*)calloc((size_t)ncn + 1, 8), *lb = (int64_t*)calloc((size_t)nl + 1, 8), *le = (int64_t*)calloc((size_t)nl + 1, 8);
        int32_t *cl = (int32_t*)calloc((size_t)ncn + 1, 4);
        for (int64_t k = 0; k < ncn; ++k) { cb[k] = cr_start(cn, k); ce[k] = cr_end(cn, k); cl[k] = cr_label(cn, k); }
        for (int64_t k = 0; k < nl; ++k) { lb[k] = cr_start(lc, k); le[k] = cr_end(lc, k); }          /* cr_index'ed by the loader: ascending starts */
        lcd_noisyreg_input_t in; memset(&in, 0, sizeof(in));
        in.reg_beg = chunk->reg_beg; in.reg_end = chunk->reg_end; in.min_alt_dp = opt->min_alt_dp; in.noisy_reg_flank_len = opt->noisy_reg_flank_len; in.is_ont = opt->is_ont; in.min_af = opt->min_af;
        in.n_sites = n; in.n_reads = nr; in.site_pos = spos; in.site_type = stype; in.site_ref_len = sref; in.var_cate = cate;
        in.n_cnreg = ncn; in.cnreg_beg = cb; in.cnreg_end = ce; in.cnreg_label = cl; in.n_low = nl; in.low_beg = lb; in.low_end = le;
        in.is_skipped = chunk->is_skipped; in.read_beg = rb; in.read_end = re; in.digar_first = df; in.n_digar = ndg; in.digar_pos = dpos; in.digar_type = dtype; in.digar_len = dlen;
        in.nreg_first = nf; in.n_nreg = nng; in.nreg_beg = nb; in.nreg_end = ne;
        const int64_t cap = ncn + n + 8;
        lcd_noisyreg_output_t out; memset(&out, 0, sizeof(out));
        out.var_cate = (int32_t*)calloc((size_t)n + 1, 4); out.keep = (uint8_t*)calloc((size_t)n + 1, 1);
        out.reg_beg = (int64_t*)calloc((size_t)cap, 8); out.reg_end = (int64_t*)calloc((size_t)cap, 8); out.reg_label = (int32_t*)calloc((size_t)cap, 4); out.reg_cap = cap;
        if (lcd_noisyreg_batch(1, &in, &out)) die("lcd_noisyreg_batch");
        if (getenv("LCD_DROPIN_CHECK_K2C")) {          /* self-check: the reference's own functions on the same chunk, compared with the library's answer (which is then dropped) */
            if (ncn > 0) { cgranges_t *raw = cr_init(); for (int64_t k = 0; k < ncn; ++k) cr_add(raw, "cr", (int32_t)cb[k], (int32_t)ce[k], cl[k]); cr_destroy(cn); cn = NULL; chunk->chunk_noisy_regs = raw; }
            pre_process_noisy_regs(chunk, opt);
            {   /* stage 1 alone: the library with no sites returns the list after pre_process_noisy_regs */
                lcd_noisyreg_input_t in0 = in; in0.n_sites = 0;
                lcd_noisyreg_output_t o0 = out; o0.reg_beg = (int64_t*)calloc((size_t)cap, 8); o0.reg_end = (int64_t*)calloc((size_t)cap, 8); o0.reg_label = (int32_t*)calloc((size_t)cap, 4);
                if (lcd_noisyreg_batch(1, &in0, &o0)) die("lcd_noisyreg_batch");
                const int64_t n1 = chunk->chunk_noisy_regs ? chunk->chunk_noisy_regs->n_r : 0;
                int b1 = n1 != o0.n_regs; int64_t first = -1;
                for (int64_t k = 0; k < n1 && k < o0.n_regs; ++k) if (cr_start(chunk->chunk_noisy_regs, k) != o0.reg_beg[k] || cr_end(chunk->chunk_noisy_regs, k) != o0.reg_end[k] || cr_label(chunk->chunk_noisy_regs, k) != o0.reg_label[k]) { b1 = 1; if (first < 0) first = k; }
                fprintf(stderr, "[k2c check] after pre_process_noisy_regs: %ld regions (reference %ld): %s", (long)o0.n_regs, (long)n1, b1 ? "DIFFERENT" : "identical");
                if (first >= 0) fprintf(stderr, " first at %ld: (%ld, %ld, %d) vs reference (%d, %d, %d)", (long)first, (long)o0.reg_beg[first], (long)o0.reg_end[first], o0.reg_label[first], cr_start(chunk->chunk_noisy_regs, first), cr_end(chunk->chunk_noisy_regs, first), cr_label(chunk->chunk_noisy_regs, first));
                fprintf(stderr, "\n");
                free(o0.reg_beg); free(o0.reg_end); free(o0.reg_label);
            }
            const int nk_ref = n > 0 ? classify_cand_vars(chunk, n, opt) : 0;
            int bad = 0, w = 0;
            for (int i = 0; i < n; ++i) if (out.keep[i]) {
                if (w >= nk_ref || cv[w].pos != spos[i] || cv[w].var_type != stype[i] || cv[w].ref_len != sref[i] || chunk->var_i_to_cate[w] != out.var_cate[i]) { if (!bad) fprintf(stderr, "[k2c check] kept site %d differs (pos %ld)\n", w, (long)spos[i]); bad = 1; }
                ++w;
            }
            if (w != nk_ref) bad = 1;
            const int64_t nr_ref = chunk->chunk_noisy_regs ? chunk->chunk_noisy_regs->n_r : 0;
            if (nr_ref != out.n_regs) bad = 1;
            else for (int64_t k = 0; k < nr_ref; ++k) if (cr_start(chunk->chunk_noisy_regs, k) != out.reg_beg[k] || cr_end(chunk->chunk_noisy_regs, k) != out.reg_end[k] || cr_label(chunk->chunk_noisy_regs, k) != out.reg_label[k]) bad = 1;
            fprintf(stderr, "[k2c check] chunk %ld: %d sites, kept %d (reference %d), regions %ld (reference %ld): %s\n", (long)chunk->reg_beg, n, w, nk_ref, (long)out.n_regs, (long)nr_ref, bad ? "DIFFERENT" : "identical");
            free(out.var_cate); free(out.keep); free(out.reg_beg); free(out.reg_end); free(out.reg_label);
            free(rb); free(re); free(df); free(nf); free(ndg); free(nng); free(dpos); free(nb); free(ne); free(dtype); free(dlen); free(cb); free(ce); free(lb); free(le); free(cl);
            free(cate); free(spos); free(saoff); free(stype); free(sref); free(salt); free(counts); free(alt);
            return 1;
        }
        /* chunk->chunk_noisy_regs */
        cgranges_t *R = cr_init();
        for (int64_t k = 0; k < out.n_regs; ++k) cr_add(R, "cr", (int32_t)out.reg_beg[k], (int32_t)out.reg_end[k], out.reg_label[k]);
        cr_index(R);
        if (cn) cr_destroy(cn);
        chunk->chunk_noisy_regs = R;
        /* the compacted cand_vars (src/collect_var.c:1012-1034) */
        if (n > 0) {
            chunk->var_i_to_cate = (int*)malloc((size_t)n * sizeof(int));
            int w = 0;
            for (int i = 0; i < n; ++i) {
                if (!out.keep[i]) continue;
                if (i != w) copy_var(cv + w, cv + i);
                chunk->var_i_to_cate[w++] = out.var_cate[i];
            }
            for (int i = w; i < n; ++i) {
                free(cv[i].alle_covs);
                for (int j = 0; j < cv[i].n_uniq_alles; ++j) free(cv[i].strand_to_alle_covs[j]);
                free(cv[i].strand_to_alle_covs);
                if (cv[i].alt_seq != NULL) free(cv[i].alt_seq);
                if (cv[i].tsd_len > 0) free(cv[i].tsd_seq);
            }
            chunk->n_cand_vars = w;
        }
        free(out.var_cate); free(out.keep); free(out.reg_beg); free(out.reg_end); free(out.reg_label);
        free(rb); free(re); free(df); free(nf); free(ndg); free(nng); free(dpos); free(nb); free(ne); free(dtype); free(dlen); free(cb); free(ce); free(lb); free(le); free(cl);
    }
    free(cate); free(spos); free(saoff); free(stype); free(sref); free(salt); free(counts); free(alt);
    return rc_ok;
}

void collect_aln_beg_end(uint32_t *cigar_buf, int cigar_len, int ext_direction, int ref_len, int *ref_beg, int *ref_end, int read_len, int *read_beg, int *read_end);   /* src/align.c:633 */
void collect_var_main(const call_var_pl_t *pl, bam_chunk_t *chunk) {
    call_var_opt_t *opt = pl->opt;
    new_thread_stream();
    trace("chunk_begin", (long)chunk->reg_beg, chunk->n_reads);
    collect_digars_from_bam(chunk, pl);                                                                                      /* 1.1 */
    var_site_t *var_sites = NULL;
    const int n_var_sites = collect_all_cand_var_sites(opt, chunk, &var_sites);                                              /* 1.2 */
    if (n_var_sites > 0) collect_cand_vars(opt, chunk, n_var_sites, var_sites);                                              /* 1.3 */
    free(var_sites);
    if (pileup_on_gpu() && !opt->out_somatic && classify_on_gpu(chunk, n_var_sites, opt)) COUNT(14);                          /* 2.1 - 2.4 on the GPU (K2b + K2c) */
    else {
        if (pileup_on_gpu() && !opt->out_somatic) COUNT(15);
        pre_process_noisy_regs(chunk, opt);                                                                                  /* 2.1 */
        if (n_var_sites > 0) classify_cand_vars(chunk, n_var_sites, opt);                                                    /* 2.2 - 2.4 */
    }
    if (chunk->n_cand_vars == 0 && (chunk->chunk_noisy_regs == NULL || chunk->chunk_noisy_regs->n_r == 0)) return;
    if (chunk->n_cand_vars > 0) {
        chunk->read_var_profile = collect_read_var_profile(opt, chunk);                                                      /* 3.1 */
        assign_hap_based_on_germline_het_vars_kmeans(opt, chunk, LONGCALLD_CLEAN_HET_SNP | LONGCALLD_CLEAN_HET_INDEL | LONGCALLD_CLEAN_HOM_VAR);   /* 3.2 */
    }
    if (chunk->chunk_noisy_regs != NULL && chunk->chunk_noisy_regs->n_r > 0) {                                               /* 4 */
        const int n_regs = (int)chunk->chunk_noisy_regs->n_r;
        int *sorted = sort_noisy_regs(chunk), *is_done = (int*)calloc(n_regs, sizeof(int));
        int *pend = (int*)malloc(sizeof(int) * n_regs), *ret = (int*)malloc(sizeof(int) * n_regs);
        /* -s / --refine-aln rewrite the reads' difference lists region by region (update_digars_from_aln_str, src/align.c:1796): one at a time */
        const int one_by_one = opt->out_somatic || (opt->refine_bam && opt->out_aln_fp != NULL) || getenv("LCD_DROPIN_SERIAL") != NULL;
        trace("regions_begin", (long)chunk->reg_beg, n_regs);
        for (;;) {
            int new_region_is_done = 0, new_var = 0, np = 0;
            for (int i = 0; i < n_regs; ++i) if (!is_done[sorted[i]]) pend[np++] = sorted[i];
            if (one_by_one) for (int k = 0; k < np; ++k) ret[k] = collect_noisy_vars1(chunk, opt, pend[k]);
            else if (np > 0) { tl_merge.on = 1; run_regions(chunk, opt, np, pend, ret); tl_merge.on = 0; flush_merges(opt, chunk); }
            for (int k = 0; k < np; ++k) if (ret[k] >= 0) { is_done[pend[k]] = 1; new_region_is_done = 1; if (ret[k] > 0) new_var = 1; }
            if (new_var) assign_hap_based_on_germline_het_vars_kmeans(opt, chunk, LONGCALLD_CAND_GERMLINE_VAR_CATE);
            if (new_region_is_done == 0) break;
        }
        trace("regions_end", (long)chunk->reg_beg, n_regs);
        free(sorted); free(is_done); free(pend); free(ret);
    }
    trace("chunk_end", (long)chunk->reg_beg, chunk->n_cand_vars);
    if (opt->out_somatic == 1) collect_somatic_var(chunk, opt);                                                              /* 5 */
}

/* ------------------------------------------------------------------------------------------ K7 */
static void edlib_req_init(req_t *r, uint8_t *target, int tlen, uint8_t *query, int qlen, int mode, int want_path) {
    uint8_t *seqs = (uint8_t*)malloc((size_t)qlen + tlen + 1);
    memcpy(seqs, query, qlen); memcpy(seqs + qlen, target, tlen);
    memset(r, 0, sizeof(*r));
    r->kind = RQ_EDLIB; r->seqs = seqs; r->seqs_len = (size_t)qlen + tlen; r->qlen = qlen; r->tlen = tlen; r->mode = mode; r->want_path = want_path;
    r->aln = (uint8_t*)malloc((size_t)qlen + tlen + 2);
}
static int edlib_req_xgaps(req_t *r) {                    /* edlibAlignmentToXGAPS, src/align.c:189-208; releases the request's buffers */
    int n_gaps = 0, n_mis = 0;
    for (int i = 0; i < r->eres.aln_len; ++i) {
        if (r->aln[i] == 3) n_mis++;
        else if ((r->aln[i] == 1 || r->aln[i] == 2) && (i == 0 || r->aln[i - 1] != r->aln[i])) n_gaps++;
    }
    free((void*)r->seqs); free(r->aln);
    return n_mis + n_gaps;
}
static int edlib1(uint8_t *target, int tlen, uint8_t *query, int qlen, int mode, int want_path, uint8_t **aln, lcd_edlib_result_t *res) {
    req_t r; edlib_req_init(&r, target, tlen, query, qlen, mode, want_path);
    gpu_call(&r);
    *res = r.eres; *aln = r.aln;
    free((void*)r.seqs);
    return 0;
}
int edlib_edit_distance(uint8_t *target, int tlen, uint8_t *query, int qlen) {                 /* src/align.c:210 */
    uint8_t *aln; lcd_edlib_result_t r; edlib1(target, tlen, query, qlen, LCD_EDLIB_MODE_NW, 0, &aln, &r); free(aln);
    return r.edit_distance;
}
int edlib_xgaps(uint8_t *target, int tlen, uint8_t *query, int qlen) {                         /* src/align.c:222 + edlibAlignmentToXGAPS :189 */
    uint8_t *aln; lcd_edlib_result_t r; edlib1(target, tlen, query, qlen, LCD_EDLIB_MODE_NW, 1, &aln, &r);
    int n_gaps = 0, n_mis = 0;
    for (int i = 0; i < r.aln_len; ++i) {
        if (aln[i] == 3) n_mis++;
        else if ((aln[i] == 1 || aln[i] == 2) && (i == 0 || aln[i - 1] != aln[i])) n_gaps++;
    }
    free(aln);
    return n_mis + n_gaps;
}
static int edlib_path_counts(uint8_t *target, int tlen, uint8_t *query, int qlen, int mode, int *n_eq, int *n_xid) {
    uint8_t *aln; lcd_edlib_result_t r; edlib1(target, tlen, query, qlen, mode, 1, &aln, &r);
    if (n_eq != NULL && n_xid != NULL) {                                                        /* edlibAlignmentToXID, src/align.c:164-187 */
        int eq = 0, x = 0;
        for (int i = 0; i < r.aln_len; ++i) { if (aln[i] == 0) eq++; else x++; }
        *n_eq = eq; *n_xid = x;
    }
    free(aln);
    return r.edit_distance;
}
int edlib_end2end_aln(uint8_t *target, int tlen, uint8_t *query, int qlen, int *n_eq, int *n_xid) { return edlib_path_counts(target, tlen, query, qlen, LCD_EDLIB_MODE_NW, n_eq, n_xid); }   /* src/align.c:234 */
int edlib_infix_aln(uint8_t *target, int tlen, uint8_t *query, int qlen, int *n_eq, int *n_xid) { return edlib_path_counts(target, tlen, query, qlen, LCD_EDLIB_MODE_HW, n_eq, n_xid); }     /* src/align.c:256 */

/* ------------------------------------------------------------------------------------------ K6 */
typedef struct { req_t r; uint8_t *seqs, *p, *t; char *ops; int plen, tlen, left; } wfa_job_t;
static void wfa_job_init(wfa_job_t *jb, uint8_t *pattern, int plen, uint8_t *text, int tlen, int gap_aln, int b, int q, int e, int q2, int e2, int heuristic, int affine_gap) {
    lcd_wfa_params_t par; memset(&par, 0, sizeof(par));
    par.mismatch = b; par.gap_open1 = q; par.gap_ext1 = e; par.gap_open2 = q2; par.gap_ext2 = e2;
    par.affine2p = affine_gap == LONGCALLD_WFA_AFFINE_2P;
    par.min_wavefront_length = 10; par.max_distance_threshold = 50; par.steps_between_cutoffs = 1;     /* wavefront_aligner_attr_default */
    if (heuristic == LONGCALLD_WFA_ADAPTIVE) par.heuristic = LCD_WFA_HEUR_ADAPTIVE;
    else if (heuristic == LONGCALLD_WFA_ZDROP) {
        par.heuristic = LCD_WFA_HEUR_ZDROP;
        const int mn = plen < tlen ? plen : tlen; const int z = (int)(mn * 0.1);
        par.zdrop = z < 500 ? z : 500; par.steps_between_cutoffs = 100;
    } else par.heuristic = LCD_WFA_HEUR_NONE;
    const int left = gap_aln == LONGCALLD_GAP_LEFT_ALN;
    uint8_t *seqs = (uint8_t*)malloc((size_t)plen + tlen + 1), *p = seqs, *t = seqs + plen;
    if (left) {                                              /* gaps at the left-most position: align the reversed sequences (:410-414) */
        for (int i = 0; i < plen; ++i) p[i] = pattern[plen - i - 1];
        for (int i = 0; i < tlen; ++i) t[i] = text[tlen - i - 1];
    } else { memcpy(p, pattern, plen); memcpy(t, text, tlen); }
    char *ops = (char*)malloc(2 * ((size_t)plen + tlen) + 16);
    memset(&jb->r, 0, sizeof(jb->r));
    jb->r.kind = RQ_WFA; jb->r.seqs = seqs; jb->r.seqs_len = (size_t)plen + tlen; jb->r.plen = plen; jb->r.tlen = tlen; jb->r.wpar = par; jb->r.ops = ops;
    jb->seqs = seqs; jb->p = p; jb->t = t; jb->ops = ops; jb->plen = plen; jb->tlen = tlen; jb->left = left;
}
static void wfa_job_finish(wfa_job_t *jb, uint32_t **cigar_buf, int *cigar_length, uint8_t **pattern_alg, uint8_t **text_alg, int *alg_length) {
    const lcd_wfa_result_t res = jb->r.wres;
    uint8_t *seqs = jb->seqs, *p = jb->p, *t = jb->t; char *ops = jb->ops; const int plen = jb->plen, tlen = jb->tlen, left = jb->left;
    const int n = res.n_ops;
    if (cigar_buf != NULL && cigar_length != NULL) {         /* cigar_get_CIGAR(cigar, true, ...) (WFA2-lib/alignment/cigar.c:181-240), reversed for left alignment */
        uint32_t *tmp = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t)); int m = 0;
        for (int i = 0; i < n;) {
            int j = i; while (j < n && ops[j] == ops[i]) ++j;
            const uint32_t op = ops[i] == 'M' ? BAM_CEQUAL : ops[i] == 'X' ? BAM_CDIFF : ops[i] == 'I' ? BAM_CINS : BAM_CDEL;
            tmp[m++] = ((uint32_t)(j - i) << 4) | op;
            i = j;
        }
        *cigar_length = m;
        *cigar_buf = (uint32_t*)malloc(((size_t)m + 1) * sizeof(uint32_t));
        for (int i = 0; i < m; ++i) (*cigar_buf)[i] = left ? tmp[m - i - 1] : tmp[i];
        free(tmp);
    }
    if (pattern_alg != NULL && text_alg != NULL) {           /* wfa_collect_pretty_alignment (:277-329), reversed for left alignment (:440-452) */
        const int max_len = tlen + plen + 1;
        uint8_t *mem = (uint8_t*)calloc(2 * (size_t)max_len, 1), *pa = mem, *ta = mem + max_len;
        int k = 0, pp = 0, tp = 0;
        for (int i = 0; i < n; ++i) {
            switch (ops[i]) {
                case 'M': case 'X': pa[k] = p[pp++]; ta[k++] = t[tp++]; break;
                case 'I': pa[k] = 5; ta[k++] = t[tp++]; break;
                case 'D': pa[k] = p[pp++]; ta[k++] = 5; break;
                default: break;
            }
        }
        if (left) for (int i = 0; i < k / 2; ++i) { uint8_t x = pa[i]; pa[i] = pa[k - i - 1]; pa[k - i - 1] = x; x = ta[i]; ta[i] = ta[k - i - 1]; ta[k - i - 1] = x; }
        *pattern_alg = pa; *text_alg = ta; *alg_length = k;
    }
    free(ops); free(seqs);
}
int wfa_end2end_aln(uint8_t *pattern, int plen, uint8_t *text, int tlen, int gap_aln, int b, int q, int e, int q2, int e2, int heuristic, int affine_gap,
                    uint32_t **cigar_buf, int *cigar_length, uint8_t **pattern_alg, uint8_t **text_alg, int *alg_length) {      /* src/align.c:374-460 */
    wfa_job_t jb; wfa_job_init(&jb, pattern, plen, text, tlen, gap_aln, b, q, e, q2, e2, heuristic, affine_gap);
    gpu_call(&jb.r);
    wfa_job_finish(&jb, cigar_buf, cigar_length, pattern_alg, text_alg, alg_length);
    return 0;
}

/* ------------------------------------------------------------------------------------------ K5 */
int abpoa_partial_aln_msa_cons(const call_var_opt_t *opt, abpoa_t *ab, int sampling_reads, int n_reads, int *read_ids, uint8_t **read_seqs, uint8_t **read_quals,
                               int *read_lens, int *read_full_cover, char **names, int max_n_cons, int *cons_lens, uint8_t **cons_seqs, int *clu_n_seqs,
                               int **clu_read_ids, int *msa_seq_lens, uint8_t **msa_seqs) {                                        /* src/align.c:762-857 */
    typedef int (*fn_t)(const call_var_opt_t *, abpoa_t *, int, int, int *, uint8_t **, uint8_t **, int *, int *, char **, int, int *, uint8_t **, int *, int **, int *, uint8_t **);
    static fn_t orig = NULL;
    if (!orig) orig = (fn_t)dlsym(RTLD_NEXT, "abpoa_partial_aln_msa_cons");
    /* Per read i > 0 the reference decides (collect_partial_aln_beg_end, :709-745) whether the read goes against the whole graph, against the
     * sub-graph between two bases of the first read (it covers the region only on one side: an extension alignment to the first read finds where
     * it ends, cal_wfa_partial_aln_beg_end :672-707), or is left out (sampled regions: more than 10 % mismatches + gap openings to the first read;
     * one-sided reads: the same test on the overlapping end).  Here the same decisions are taken for all reads of the region at once: one batch
     * of edlib problems (K7), one batch of WFA extension alignments (K6), then ONE POA problem with the reads' anchors (K5, lcd_poa_sub_batch). */
    int ok = ab == NULL && max_n_cons == 1 && n_reads >= 1 && cons_lens && cons_seqs && LONGCALLD_NOISY_IS_BOTH_COVER(read_full_cover[0]) && read_lens[0] > 0;
    for (int i = 0; ok && i < n_reads; ++i) if (read_lens[i] <= 0) ok = 0;
    if (ok) {
        enum { RD_FULL = 0, RD_FULL_SAMPLED, RD_L2R, RD_R2L, RD_SKIP };
        const int tlen0 = read_lens[0];
        int *cls = (int*)calloc(n_reads, sizeof(int)), *sub_beg = (int*)calloc(n_reads, sizeof(int)), *sub_end = (int*)calloc(n_reads, sizeof(int));
        int *cut_beg = (int*)calloc(n_reads, sizeof(int)), *cut_end = (int*)calloc(n_reads, sizeof(int));
        typedef struct { uint8_t *target, *query; int tlen, qlen, gap_aln; } ext_t;            /* the (trimmed) pair of cal_wfa_partial_aln_beg_end */
        ext_t *ext = (ext_t*)calloc(n_reads, sizeof(ext_t));
        req_t *er = (req_t*)calloc(n_reads, sizeof(req_t)); req_t **list = (req_t**)malloc(sizeof(req_t*) * n_reads); int nl = 0, any_sub = 0;
        for (int i = 1; i < n_reads; ++i) {
            const int c = read_full_cover[i], qlen0 = read_lens[i];
            if (LONGCALLD_NOISY_IS_BOTH_COVER(c) || (LONGCALLD_NOISY_IS_LEFT_COVER(c) && LONGCALLD_NOISY_IS_RIGHT_GAP(c)) || (LONGCALLD_NOISY_IS_RIGHT_COVER(c) && LONGCALLD_NOISY_IS_LEFT_GAP(c))) {
                cls[i] = sampling_reads ? RD_FULL_SAMPLED : RD_FULL;
                if (sampling_reads) { edlib_req_init(er + i, read_seqs[0], tlen0, read_seqs[i], qlen0, LCD_EDLIB_MODE_NW, 1); list[nl++] = er + i; }
            } else if (LONGCALLD_NOISY_IS_LEFT_COVER(c) || LONGCALLD_NOISY_IS_RIGHT_COVER(c)) {
                const int l2r = LONGCALLD_NOISY_IS_LEFT_COVER(c) != 0;
                const double ratio = opt->partial_aln_ratio;
                ext_t *x = ext + i; x->target = read_seqs[0]; x->query = read_seqs[i]; x->tlen = tlen0; x->qlen = qlen0;
                if (l2r) { if (tlen0 > qlen0 * ratio) x->tlen = (int)(qlen0 * ratio); else if (qlen0 > tlen0 * ratio) x->qlen = (int)(tlen0 * ratio); }
                else {
                    if (tlen0 > qlen0 * ratio) { x->target = read_seqs[0] + tlen0 - (int)(qlen0 * ratio); x->tlen = (int)(qlen0 * ratio); }
                    else if (qlen0 > tlen0 * ratio) { x->query = read_seqs[i] + qlen0 - (int)(tlen0 * ratio); x->qlen = (int)(tlen0 * ratio); }
                }
                x->gap_aln = opt->gap_aln;
                if (l2r) x->gap_aln = (opt->gap_aln == LONGCALLD_GAP_RIGHT_ALN) ? LONGCALLD_GAP_LEFT_ALN : LONGCALLD_GAP_RIGHT_ALN;
                cls[i] = l2r ? RD_L2R : RD_R2L;
                const int mn = x->tlen < x->qlen ? x->tlen : x->qlen;
                if (l2r) edlib_req_init(er + i, x->target, mn, x->query, mn, LCD_EDLIB_MODE_NW, 1);
                else edlib_req_init(er + i, x->target + x->tlen - mn, mn, x->query + x->qlen - mn, mn, LCD_EDLIB_MODE_NW, 1);
                list[nl++] = er + i;
            } else cls[i] = RD_FULL;                       /* covers neither end: the reference keeps the full ranges (:711, ret = 1) */
        }
        gpu_call_many(list, nl);                            /* K7: every read's filter in one batch */
        wfa_job_t *wj = (wfa_job_t*)calloc(n_reads, sizeof(wfa_job_t)); nl = 0;
        for (int i = 1; i < n_reads; ++i) {
            if (cls[i] == RD_FULL_SAMPLED) {
                const int mn = tlen0 < read_lens[i] ? tlen0 : read_lens[i];
                if (edlib_req_xgaps(er + i) > mn * 0.10) cls[i] = RD_SKIP;
            } else if (cls[i] == RD_L2R || cls[i] == RD_R2L) {
                ext_t *x = ext + i; const int mn = x->tlen < x->qlen ? x->tlen : x->qlen;
                if (edlib_req_xgaps(er + i) > mn * 0.10) { cls[i] = RD_SKIP; continue; }
                wfa_job_init(wj + i, x->target, x->tlen, x->query, x->qlen, x->gap_aln, opt->mismatch, opt->gap_open1, opt->gap_ext1, opt->gap_open2, opt->gap_ext2,
                             LONGCALLD_WFA_NO_HEURISTIC, LONGCALLD_WFA_AFFINE_2P);
                list[nl++] = &wj[i].r;
            }
        }
        gpu_call_many(list, nl);                            /* K6: every one-sided read's extension alignment in one batch */
        size_t tot = 0;
        for (int i = 0; i < n_reads; ++i) {
            if (cls[i] == RD_L2R || cls[i] == RD_R2L) {
                uint32_t *cg = NULL; int ncg = 0;
                wfa_job_finish(wj + i, &cg, &ncg, NULL, NULL, NULL);
                if (ncg == 0) cls[i] = RD_SKIP;
                else {
                    int rb, re, qb, qe;
                    collect_aln_beg_end(cg, ncg, cls[i] == RD_L2R ? LONGCALLD_EXT_ALN_LEFT_TO_RIGHT : LONGCALLD_EXT_ALN_RIGHT_TO_LEFT, tlen0, &rb, &re, read_lens[i], &qb, &qe);
                    sub_beg[i] = rb + 1; sub_end[i] = re + 1; cut_beg[i] = qb - 1; cut_end[i] = read_lens[i] - qe; any_sub = 1;
                    if (rb == 1 && re == tlen0) { sub_beg[i] = 0; sub_end[i] = 0; }            /* the whole first read: abpoa_subgraph_nodes yields the whole graph */
                }
                if (cg) free(cg);
            }
            if (cls[i] == RD_SKIP) { sub_beg[i] = sub_end[i] = -1; any_sub = 1; }
            if (read_lens[i] - cut_beg[i] - cut_end[i] <= 0 && cls[i] != RD_SKIP) ok = 0;      /* (an empty piece: the reference's abPOA handles it its own way) */
            tot += (size_t)read_lens[i];
        }
        free(wj); free(er); free(list); free(ext);
        if (ok) {
            uint8_t *seqs = (uint8_t*)malloc(tot + 1), *cons = (uint8_t*)malloc(tot + 17);
            int64_t *off = (int64_t*)malloc(sizeof(int64_t) * n_reads); int32_t *len = (int32_t*)malloc(sizeof(int32_t) * n_reads);
            size_t o = 0; int max_len = 0;
            for (int i = 0; i < n_reads; ++i) {
                const int l = cls[i] == RD_SKIP ? read_lens[i] : read_lens[i] - cut_beg[i] - cut_end[i];
                off[i] = (int64_t)o; len[i] = l; memcpy(seqs + o, read_seqs[i] + (cls[i] == RD_SKIP ? 0 : cut_beg[i]), l); o += l; if (l > max_len) max_len = l;
            }
            const lcd_poa_params_t par = { opt->match, opt->mismatch, opt->gap_open1, opt->gap_ext1, opt->gap_open2, opt->gap_ext2, 10, 0.01f, 1, 1 };   /* abpoa_init_para defaults wb / wf */
            const int64_t msa_cap = (int64_t)(n_reads + 1) * (2 * (int64_t)max_len + 64);
            uint8_t *msa = (msa_seq_lens && msa_seqs) ? (uint8_t*)malloc((size_t)msa_cap) : NULL;
            req_t r; memset(&r, 0, sizeof(r));
            r.kind = RQ_POA; r.seqs = seqs; r.seqs_len = o; r.n_reads = n_reads; r.read_off = off; r.read_len = len; r.ppar = par; r.want_msa = msa != NULL; r.max_len = max_len;
            r.cons = cons; r.msa = msa; r.msa_cap = msa_cap;
            if (any_sub) { r.sub_beg = sub_beg; r.sub_end = sub_end; }
            gpu_call(&r);
            const lcd_poa_result_t res = r.pres;
            if (res.status == LCD_POA_OK && res.cons_len > 0) {
                cons_lens[0] = res.cons_len; cons_seqs[0] = (uint8_t*)malloc(res.cons_len); memcpy(cons_seqs[0], cons, res.cons_len);
                if (clu_n_seqs != NULL && clu_read_ids != NULL) { *clu_n_seqs = n_reads; *clu_read_ids = (int*)malloc(n_reads * sizeof(int)); for (int i = 0; i < n_reads; ++i) (*clu_read_ids)[i] = read_ids[i]; }
                if (msa) {
                    *msa_seq_lens = res.msa_len;
                    for (int i = 0; i < n_reads + 1; ++i) { msa_seqs[i] = (uint8_t*)malloc(res.msa_len); memcpy(msa_seqs[i], msa + (size_t)i * res.msa_len, res.msa_len); }
                }
                free(seqs); free(cons); free(off); free(len); free(msa);
                free(cls); free(sub_beg); free(sub_end); free(cut_beg); free(cut_end);
                if (any_sub) COUNT(11);
                return 1;
            }
            /* outside the kernel's envelope (e.g. LCD_POA_NEEDS_INT32, MSA wider than the estimate): let abPOA handle this region */
            free(seqs); free(cons); free(off); free(len); free(msa);
        }
        free(cls); free(sub_beg); free(sub_end); free(cut_beg); free(cut_end);
    }
    COUNT(6);
    const double t0_ = now_s();
    const int rc_ = orig(opt, ab, sampling_reads, n_reads, read_ids, read_seqs, read_quals, read_lens, read_full_cover, names, max_n_cons, cons_lens, cons_seqs,
                clu_n_seqs, clu_read_ids, msa_seq_lens, msa_seqs);
    t_add(&t_fwd_poa, now_s() - t0_);
    return rc_;
}

/* abpoa_aln_msa_cons (src/align.c:872-953): the de-novo POA of a region without a usable phase set -- all fully covering reads, unbanded, up to
 * two consensus sequences from abPOA's read clustering -- as ONE problem of lcd_poa_ncons_batch (the clustering runs on the device). */
int abpoa_aln_msa_cons(const call_var_opt_t *opt, int n_reads, int *read_ids, uint8_t **read_seqs, int *read_lens, int max_n_cons, int *cons_lens, uint8_t **cons_seqs,
                       int *clu_n_seqs, int **clu_read_ids, int *msa_seq_len, uint8_t ***msa_seq) {
    typedef int (*fn_t)(const call_var_opt_t *, int, int *, uint8_t **, int *, int, int *, uint8_t **, int *, int **, int *, uint8_t ***);
    static fn_t orig = NULL;
    if (!orig) orig = (fn_t)dlsym(RTLD_NEXT, "abpoa_aln_msa_cons");
    int ok = n_reads >= 1 && (max_n_cons == 1 || max_n_cons == 2) && cons_lens && cons_seqs && clu_n_seqs && clu_read_ids;
    size_t tot = 0; int max_len = 0;
    for (int i = 0; ok && i < n_reads; ++i) { if (read_lens[i] <= 0) ok = 0; else { tot += (size_t)read_lens[i]; if (read_lens[i] > max_len) max_len = read_lens[i]; } }
    if (ok) {
        uint8_t *seqs = (uint8_t*)malloc(tot + 1), *cons = (uint8_t*)malloc(tot + 17), *rclu = (uint8_t*)calloc(n_reads + 1, 1);
        int64_t *off = (int64_t*)malloc(sizeof(int64_t) * n_reads); int32_t *len = (int32_t*)malloc(sizeof(int32_t) * n_reads);
        size_t o = 0;
        for (int i = 0; i < n_reads; ++i) { off[i] = (int64_t)o; len[i] = read_lens[i]; memcpy(seqs + o, read_seqs[i], read_lens[i]); o += read_lens[i]; }
        const lcd_poa_params_t par = { opt->match, opt->mismatch, opt->gap_open1, opt->gap_ext1, opt->gap_open2, opt->gap_ext2, -1, 0.01f, 0, max_n_cons };
        const int want_msa = msa_seq != NULL && msa_seq_len != NULL;
        const int64_t msa_cap = (int64_t)(n_reads + 2) * (2 * (int64_t)max_len + 64);
        uint8_t *msa = want_msa ? (uint8_t*)malloc((size_t)msa_cap) : NULL;
        req_t r; memset(&r, 0, sizeof(r));
        r.kind = RQ_POA; r.seqs = seqs; r.seqs_len = o; r.n_reads = n_reads; r.read_off = off; r.read_len = len; r.ppar = par; r.want_msa = want_msa; r.max_len = max_len;
        r.cons = cons; r.msa = msa; r.msa_cap = msa_cap; r.min_freq = opt->min_af; r.read_clu = rclu; r.n_cons = 1;
        gpu_call(&r);
        const lcd_poa_result_t res = r.pres;
        int done = 0;
        if (res.status == LCD_POA_OK && res.cons_len > 0 && (max_n_cons == 1 || r.n_cons >= 1)) {
            const int nc = max_n_cons == 2 ? r.n_cons : 1;
            const int cl[2] = { res.cons_len, r.cons_len2 };
            int at = 0;
            for (int c = 0; c < nc; ++c) { cons_lens[c] = cl[c]; cons_seqs[c] = (uint8_t*)malloc(cl[c] > 0 ? cl[c] : 1); memcpy(cons_seqs[c], cons + at, cl[c]); at += cl[c]; }
            int cn[2] = { n_reads, 0 };
            if (nc == 2) {
                cn[0] = cn[1] = 0;
                for (int c = 0; c < 2; ++c) clu_read_ids[c] = (int*)malloc((n_reads + 1) * sizeof(int));
                for (int i = 0; i < n_reads; ++i) { const int c = rclu[i] ? 1 : 0; clu_read_ids[c][cn[c]++] = read_ids[i]; }
                clu_n_seqs[0] = cn[0]; clu_n_seqs[1] = cn[1];
            } else {
                *clu_n_seqs = n_reads; *clu_read_ids = (int*)malloc((n_reads + 1) * sizeof(int));
                for (int i = 0; i < n_reads; ++i) (*clu_read_ids)[i] = read_ids[i];
            }
            if (want_msa) {
                const size_t ml = (size_t)res.msa_len;
                for (int c = 0; c < nc; ++c) {
                    msa_seq_len[c] = res.msa_len;
                    int j = 0;
                    for (int i = 0; i < n_reads; ++i) if (nc == 1 || (rclu[i] ? 1 : 0) == c) { msa_seq[c][j] = (uint8_t*)malloc(ml > 0 ? ml : 1); memcpy(msa_seq[c][j], msa + (size_t)i * ml, ml); ++j; }
                    msa_seq[c][cn[c]] = (uint8_t*)malloc(ml > 0 ? ml : 1); memcpy(msa_seq[c][cn[c]], msa + (size_t)(n_reads + c) * ml, ml);
                }
            }
            if (nc == 2) COUNT(12);
            done = nc;
        }
        free(seqs); free(cons); free(rclu); free(off); free(len); free(msa);
        if (done) return done;
        /* outside the kernel's envelope (e.g. LCD_POA_NEEDS_INT32): let abPOA handle this region */
    }
    COUNT(13);
    const double t0_ = now_s();
    const int rc_ = orig(opt, n_reads, read_ids, read_seqs, read_lens, max_n_cons, cons_lens, cons_seqs, clu_n_seqs, clu_read_ids, msa_seq_len, msa_seq);
    t_add(&t_fwd_poa, now_s() - t0_);
    return rc_;
}

/* ------------------------------------------------------------------------------------------ a8: the two haplotypes of a region side by side
 * wfa_collect_noisy_aln_str_with_ps_hap (src/align.c:1286-1375): the same steps -- the reads of each haplotype of the phase set, one POA per
 * haplotype, ref-vs-consensus alignment strings, consensus-vs-read strings from the MSA rows -- with the two POA problems issued together and
 * then the two WFA problems issued together (fork_join), so a region costs two engine round trips here instead of four. */
int is_homopolymer(uint8_t *seq, int seq_len, int flank_len, int *hp_start, int *hp_end, int *hp_len);                                  /* src/align.c:1000 */
int make_cons_read_aln_str(const call_var_opt_t *opt, uint8_t *cons_str, uint8_t *read_str, int msa_len, int full_cover, aln_str_t *cons_read_aln_str);   /* :1029 */
int make_ref_read_aln_str(const call_var_opt_t *opt, aln_str_t *ref_cons_aln_str, aln_str_t *cons_read_aln_str, aln_str_t *ref_read_aln_str);          /* :1056 */
int wfa_collect_aln_str(const call_var_opt_t *opt, uint8_t *target, int tlen, uint8_t *query, int qlen, int full_cover, int heuristic, int affine_gap, aln_str_t *aln_str);   /* :565 */
typedef struct {
    const call_var_opt_t *opt; int sampling_reads, n; int *ids; uint8_t **seqs, **quals; int *lens, *covers; char **names;
    int *cons_len; uint8_t **cons_seq; int *clu_n; int **clu_ids; int *msa_len; uint8_t **msa; int ret;
} hap_poa_t;
static void hap_poa_task(void *a) {
    hap_poa_t *h = (hap_poa_t*)a;
    h->ret = abpoa_partial_aln_msa_cons(h->opt, NULL, h->sampling_reads, h->n, h->ids, h->seqs, h->quals, h->lens, h->covers, h->names, 1, h->cons_len, h->cons_seq, h->clu_n, h->clu_ids, h->msa_len, h->msa);
}
typedef struct { const call_var_opt_t *opt; uint8_t *ref; int ref_len; uint8_t *cons; int cons_len; aln_str_t *out; } hap_wfa_t;
static void hap_wfa_task(void *a) {
    hap_wfa_t *h = (hap_wfa_t*)a;
    wfa_collect_aln_str(h->opt, h->ref, h->ref_len, h->cons, h->cons_len, LONGCALLD_NOISY_BOTH_COVER, LONGCALLD_WFA_NO_HEURISTIC, LONGCALLD_WFA_AFFINE_2P, h->out);
}
int wfa_collect_noisy_aln_str_with_ps_hap(const call_var_opt_t *opt, int sampling_reads, int n_reads, int *noisy_read_ids, int *lens, uint8_t **seqs, uint8_t *strands, uint8_t **quals, char **names,
                                          int *haps, hts_pos_t *phase_sets, int *fully_covers, hts_pos_t ps, int min_hap_full_reads, int min_hap_all_reads, uint8_t *ref_seq, int ref_seq_len,
                                          int *clu_n_seqs, int **clu_read_ids, aln_str_t **aln_strs, int collect_ref_read_aln_str) {
    (void)strands; (void)min_hap_full_reads; (void)min_hap_all_reads;
    const int total = n_reads + 2;
    int n_cons = 0;
    int cons_lens[2] = {0, 0}; uint8_t *cons_seqs[2] = {NULL, NULL}; int msa_seq_lens[2] = {0, 0}; uint8_t **msa_seqs[2];
    int *h_ids[2], *h_lens[2], *h_cov[2]; uint8_t **h_seqs[2], **h_quals[2]; char **h_names[2];
    for (int i = 0; i < 2; ++i) {
        msa_seqs[i] = (uint8_t**)calloc(n_reads + 1, sizeof(uint8_t*));
        h_ids[i] = (int*)malloc(total * sizeof(int)); h_lens[i] = (int*)malloc(total * sizeof(int)); h_cov[i] = (int*)calloc(total, sizeof(int));
        h_seqs[i] = (uint8_t**)malloc(total * sizeof(uint8_t*)); h_quals[i] = (uint8_t**)malloc(total * sizeof(uint8_t*)); h_names[i] = (char**)malloc(total * sizeof(char*));
    }
    int hp_s, hp_e, hp_l, use_non_full = 1;
    if (is_homopolymer(ref_seq, ref_seq_len, opt->noisy_reg_flank_len, &hp_s, &hp_e, &hp_l)) use_non_full = 0;
    /* the reads of each haplotype (:1312-1326).  The reference gives up at the first haplotype whose first read is as long as max_noisy_reg_len
     * (:1327-1330; for a haplotype without reads it looks at the slot the haplotype before left, which has passed the test) and skips a
     * haplotype without reads (:1332); one consensus alone counts as none (:1338). */
    hap_poa_t hp[2]; task_t tasks[2]; int n_tasks = 0;
    for (int hap = 1; hap <= 2; ++hap) {
        const int x = hap - 1; int m = 0;
        for (int i = 0; i < n_reads; ++i) {
            if (lens[i] <= 0 || phase_sets[i] != ps || haps[i] != hap) continue;
            if (use_non_full == 0 && LONGCALLD_NOISY_IS_BOTH_COVER(fully_covers[i]) == 0) continue;
            h_ids[x][m] = noisy_read_ids[i]; h_lens[x][m] = lens[i]; h_seqs[x][m] = seqs[i]; h_quals[x][m] = quals[i]; h_cov[x][m] = fully_covers[i]; h_names[x][m] = names[i]; ++m;
        }

        if (m > 0 && h_lens[x][0] >= opt->max_noisy_reg_len) break;
        if (m == 0) continue;
        hap_poa_t h = { opt, sampling_reads, m, h_ids[x], h_seqs[x], h_quals[x], h_lens[x], h_cov[x], h_names[x], cons_lens + x, cons_seqs + x, clu_n_seqs + x, clu_read_ids + x, msa_seq_lens + x, msa_seqs[x], 0 };
        hp[n_tasks] = h; tasks[n_tasks].fn = hap_poa_task; tasks[n_tasks].arg = &hp[n_tasks]; ++n_tasks;
    }
    fork_join(tasks, n_tasks);                               /* K5 (and the filters / extension alignments of partially covering reads) of both haplotypes */
    for (int k = 0; k < n_tasks; ++k) n_cons += hp[k].ret;
    if (n_cons != 2) n_cons = 0;
    else {
        hap_wfa_t hw[2];
        for (int x = 0; x < 2; ++x) { hap_wfa_t h = { opt, ref_seq, ref_seq_len, cons_seqs[x], cons_lens[x], LONGCALLD_REF_CONS_ALN_STR(aln_strs[x]) }; hw[x] = h; tasks[x].fn = hap_wfa_task; tasks[x].arg = &hw[x]; }
        fork_join(tasks, 2);                                 /* K6: ref vs consensus, both haplotypes */
        for (int hap = 1; hap <= 2; ++hap) {
            aln_str_t *clu_aln_str = aln_strs[hap - 1];
            int m = 0;
            for (int i = 0; i < n_reads; ++i) {
                if (lens[i] <= 0 || phase_sets[i] != ps || haps[i] != hap) continue;
                if (use_non_full == 0 && LONGCALLD_NOISY_IS_BOTH_COVER(fully_covers[i]) == 0) continue;
                make_cons_read_aln_str(opt, msa_seqs[hap - 1][clu_n_seqs[hap - 1]], msa_seqs[hap - 1][m], msa_seq_lens[hap - 1], fully_covers[i], LONGCALLD_CONS_READ_ALN_STR(clu_aln_str, m));
                if (collect_ref_read_aln_str)
                    make_ref_read_aln_str(opt, LONGCALLD_REF_CONS_ALN_STR(clu_aln_str), LONGCALLD_CONS_READ_ALN_STR(clu_aln_str, m), LONGCALLD_REF_READ_ALN_STR(clu_aln_str, m));
                ++m;
            }
        }
    }
    for (int i = 0; i < 2; ++i) {
        free(cons_seqs[i]);
        for (int j = 0; j < n_reads + 1; ++j) free(msa_seqs[i][j]);
        free(msa_seqs[i]); free(h_ids[i]); free(h_lens[i]); free(h_cov[i]); free(h_seqs[i]); free(h_quals[i]); free(h_names[i]);
    }
    return n_cons;
}
