/* include/lcd_gpu.h -- C ABI of liblcd_gpu.so, the B200 (sm_100a) implementation of the
 * re-alignment stack on the hot path of `longcallD call` (yangao07/longcallD @ 491f055).
 *
 * Plain pointers and sizes only; no torch / C++ types.  Every entry point names the reference
 * interface it replaces.  The reference calls its engines one problem at a time from inside
 * collect_noisy_reg_aln_strs() (src/align.c:1760); a GPU needs many problems per launch, so each
 * engine is exposed
 *   (1) as a *batch* call over host buffers  -- lcd_<engine>_batch()      (drop-in, e2e)
 *   (2) as a *plan* whose inputs stay resident in HBM -- lcd_<engine>_plan_*() (re-runnable;
 *       what bench.py times for the device-resident number and what ncu profiles).
 * Results are bit-identical to the reference's for the same inputs (tests/ -m gpu).
 *
 * Threading: one context per process (lcd_gpu_init).  Plans may be created and run from several
 * host threads (each is bound to the library's device on entry); runs on the same CUDA stream
 * serialise, plans that own their buffers (K1, K1b, K2, K2b, K3, K4) run unserialised next to the
 * DP engines (lcd_gpu_reserve_sms, lcd_gpu_split_pool).  There is NO CPU fallback: every call
 * fails with a non-zero code (and lcd_gpu_last_error()) when the GPU or the kernels are missing.
 */
#ifndef LCD_GPU_H
#define LCD_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCD_GPU_ABI_VERSION 1

/* ---------------------------------------------------------------- context */
/* Select the device, create the library stream and size the wavefront/DP workspace pool.
 * pool_bytes == 0 picks a default (a fraction of free HBM).  Returns 0 on success. */
int  lcd_gpu_init(int device, size_t pool_bytes);
void lcd_gpu_shutdown(void);
const char *lcd_gpu_last_error(void);
int  lcd_gpu_abi_version(void);
/* Number of kernel launches issued by this library since init (bench.py's gpu_launches). */
uint64_t lcd_gpu_launch_count(void);
/* The library's CUDA stream (a cudaStream_t cast to void*) -- callers record their timing events
 * on it; NULL before lcd_gpu_init. */
void *lcd_gpu_stream(void);
/* A second stream of the library, and the per-host-thread default: plans created / run / fetched by a thread that passes no stream use
 * the stream it set here (NULL: back to the library stream).  One host thread can so stage one stage's buffers (K1's H2D copies) while
 * another thread's kernels (the POA launch) occupy the library stream.  The POA and WFA plans share the workspace pool and must all
 * run on one stream. */
void *lcd_gpu_aux_stream(void);
void *lcd_gpu_new_stream(void);        /* a new non-blocking stream on the library's device (owned by the library, released by lcd_gpu_shutdown); NULL on error */
void lcd_gpu_set_thread_stream(void *stream);
/* The DP engines (K5, K6) run persistent grids that fill every SM until their queues drain; kernels of other plans launched meanwhile
 * (K1 - K4 from a second host thread / stream: their plans own their buffers and are not serialised with the DP engines) would wait
 * for the DP launch to end.  Reserving n SMs' worth of CTA slots shrinks the persistent grids by that much so that such kernels run
 * alongside.  The DP launches are bounded by their longest problem, not by the SM count, so a small reservation costs them nothing. */
int lcd_gpu_reserve_sms(int n_sms);
/* Splits the workspace pool in two windows: the lower `lower_bytes` for the POA plans (K5), the rest for the WFA / edlib plans (K6, K7),
 * serialised independently, so that the WFA launch of one region batch and the POA launch of the next may be in flight together (run
 * them on different streams).  Call it right after lcd_gpu_init, before any plan exists; 0 restores the single window. */
int lcd_gpu_split_pool(size_t lower_bytes);
/* The general form: n_poa equal windows below `lower_bytes` for the POA plans and n_aln equal windows above it for the WFA / edlib plans.
 * A plan takes an idle window of its class for the time of a run (else the next one in turn, ordered behind that window's last user),
 * so up to n batches of one engine -- created and run by different host threads on different streams -- are on the GPU together: a
 * batch whose launch has shrunk to its last long problems no longer keeps the next one waiting.  lcd_gpu_split_pool(b) is (1, 1, b). */
int lcd_gpu_pool_windows(int n_poa, int n_aln, size_t lower_bytes);
/* Plan buffers come from the device's stream-ordered memory pool (kept between calls).  A host that creates plans from several threads while persistent DP
 * grids are resident should map a reserve once (after lcd_gpu_init): a block freed on one stream is not reused on another before that stream synchronises, and
 * a pool that has to grow under a resident grid stalls the allocating threads until the grid retires.  Costs its mapping time once (~0.1 s per GiB): a short-lived
 * process (one `longcallD call` on a small input) is better off without. */
int lcd_gpu_reserve_plan_memory(size_t bytes);

/* ---------------------------------------------------------------- K6: WFA gap-affine(-2p)
 * Replaces wavefront_aligner_new + wavefront_align + reading wf_aligner->cigar, as called by
 * wfa_end2end_aln (src/align.c:374-460) and is_diff_between_ref_hap_aln (src/assign_hap.c:972).
 * The left-alignment reversal and the aligned-string construction (src/align.c:410-452,277-329)
 * are host glue: see wfa_job_init / wfa_job_finish in longcalld_b200/dropin/lcd_dropin.c. */
enum { LCD_WFA_HEUR_NONE = 0, LCD_WFA_HEUR_ADAPTIVE = 1, LCD_WFA_HEUR_ZDROP = 2 };
enum { LCD_WFA_STATUS_COMPLETED = 0, LCD_WFA_STATUS_PARTIAL = 1, LCD_WFA_STATUS_ERROR = -1,
       LCD_WFA_STATUS_OOM = -2 };

typedef struct {
    int32_t mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2; /* match == 0 (src/align.c:383) */
    int32_t affine2p;              /* 1: gap_affine_2p, 0: gap_affine (o1,e1) */
    int32_t heuristic;             /* LCD_WFA_HEUR_* (src/align.c:398-406) */
    int32_t min_wavefront_length, max_distance_threshold;       /* wf-adaptive (10,50) */
    int32_t zdrop;                 /* z-drop */
    int32_t steps_between_cutoffs; /* 1 (adaptive) / 100 (z-drop) */
} lcd_wfa_params_t;

typedef struct {
    int32_t status;                /* LCD_WFA_STATUS_* */
    int32_t score;                 /* cigar->score */
    int32_t n_ops;                 /* edit operations in ops[] ('M','X','I','D') */
    int32_t end_v, end_h;          /* cigar->end_v / end_h */
} lcd_wfa_result_t;

/* Problem i aligns pattern = seqs[pat_off[i] .. +plen[i]) against text = seqs[txt_off[i] .. +tlen[i])
 * (base codes 0..4 as in the reference).  Its operations are written to ops[ops_off[i] ..], which
 * must have room for 2*(plen+tlen)+8 bytes.  All pointers are HOST memory. */
int lcd_wfa_batch(int n, const uint8_t *seqs, size_t seqs_len,
                  const int64_t *pat_off, const int32_t *plen,
                  const int64_t *txt_off, const int32_t *tlen,
                  const lcd_wfa_params_t *params,       /* n entries */
                  char *ops, const int64_t *ops_off, lcd_wfa_result_t *results);

typedef struct lcd_plan lcd_plan_t;
/* Upload once (H2D), run many times.  `stream` is a cudaStream_t cast to void* (NULL: library stream). */
lcd_plan_t *lcd_wfa_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                const int64_t *pat_off, const int32_t *plen,
                                const int64_t *txt_off, const int32_t *tlen,
                                const lcd_wfa_params_t *params);
int  lcd_plan_run(lcd_plan_t *plan, void *stream);           /* async launch(es) on stream */
int  lcd_plan_sync(lcd_plan_t *plan, void *stream);          /* wait + check device status; a POA plan looks at its statuses here (or in its
                                                                * fetch) and re-runs, with the worst-case workspace, any problem that outgrew its budget */
/* Copy results back (D2H).  ops may be NULL.  Layout as in the matching *_batch call. */
int  lcd_wfa_plan_fetch(lcd_plan_t *plan, void *stream, char *ops, const int64_t *ops_off,
                        lcd_wfa_result_t *results);
void lcd_plan_destroy(lcd_plan_t *plan);
/* Algorithmic work of one run of the plan (units of SURVEY.md section 8d: WFA wavefront cells,
 * edlib block-columns, POA banded cells) -- filled by the kernels themselves. */
int  lcd_plan_work_units(lcd_plan_t *plan, void *stream, uint64_t *units);

/* ---------------------------------------------------------------- K7: edlib NW / HW with path
 * Replaces edlibAlign(query, qlen, target, tlen, edlibNewAlignConfig(-1, mode, task, NULL, 0)) as called by
 * edlib_edit_distance / edlib_xgaps / edlib_end2end_aln / edlib_infix_aln (src/align.c:210-275) and reading
 * result.editDistance, startLocations[0], endLocations[0], alignment, alignmentLength.  The reductions of the
 * path that the reference applies next (edlibAlignmentToXGAPS / XID, src/align.c:164-208) are host glue. */
enum { LCD_EDLIB_MODE_NW = 0, LCD_EDLIB_MODE_SHW = 1, LCD_EDLIB_MODE_HW = 2 };          /* EdlibAlignMode */
enum { LCD_EDLIB_STATUS_OK = 0, LCD_EDLIB_STATUS_SYMBOL = -1 /* a base code > 5 */, LCD_EDLIB_STATUS_NO_SPLIT = -2,
       LCD_EDLIB_STATUS_OOM = -3, LCD_EDLIB_STATUS_NO_SOLUTION = -4 };
typedef struct {
    int32_t status;                /* LCD_EDLIB_STATUS_* */
    int32_t edit_distance;         /* result.editDistance */
    int32_t start_loc, end_loc;    /* startLocations[0] (-1 when want_path == 0: EDLIB_TASK_DISTANCE) / endLocations[0] */
    int32_t aln_len;               /* result.alignmentLength (0 when want_path == 0) */
} lcd_edlib_result_t;

/* Problem i aligns query = seqs[query_off[i] .. +qlen[i]) to target = seqs[target_off[i] .. +tlen[i]) (base codes 0..5).
 * want_path[i] != 0 (EDLIB_TASK_PATH): the path (EDLIB_EDOP_* codes 0 '=', 1 insert, 2 delete, 3 'X') goes to
 * aln[aln_off[i] ..], which needs qlen + tlen bytes.  All pointers are HOST memory. */
int lcd_edlib_batch(int n, const uint8_t *seqs, size_t seqs_len,
                    const int64_t *query_off, const int32_t *qlen,
                    const int64_t *target_off, const int32_t *tlen,
                    const int32_t *mode, const int32_t *want_path,
                    uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results);
lcd_plan_t *lcd_edlib_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                  const int64_t *query_off, const int32_t *qlen,
                                  const int64_t *target_off, const int32_t *tlen,
                                  const int32_t *mode, const int32_t *want_path);
int  lcd_edlib_plan_fetch(lcd_plan_t *plan, void *stream, uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results);

/* ---------------------------------------------------------------- K1: pileup scan, difference lists from =/X CIGARs
 * Replaces void collect_digars_from_bam(bam_chunk_t *chunk, const struct call_var_pl_t *pl) (src/collect_var.c:1063-1110) for reads
 * whose CIGAR uses =/X: per read int collect_digar_from_eqx_cigar(bam_chunk_t *chunk, int read_i, const call_var_opt_t *opt,
 * digar_t *digar) (src/bam_utils.c:701-841) with push_xid_size_queue_win (src/bam_utils.c:161-205), plus the chunk's base-quality
 * histogram (longcalld_copy_digar_read_buffers, src/bam_utils.c:90-103).  Inputs are the BAM record fields the reference reads
 * (SURVEY 8b): core.pos, reverse flag, CIGAR words, 4-bit packed SEQ, QUAL; is_palindrome is is_ont_palindrome_clip's SA-tag
 * test (src/bam_utils.c:659-698: host string parsing, 0 unless --ont).  Outputs are, per read, what the reference leaves in
 * digar_t (beg, end, digars[] in the flat layout lcd_pileup_input_t consumes, noisy_regs after cr_index), its return value
 * (skip: the caller sets chunk->is_skipped[r] = BAM_RECORD_WRONG_MAP), and per chunk qual_counts[256] and the intervals the kept
 * reads add to chunk->chunk_noisy_regs, in cr_add order.  Reads with an 'M' op (cs / MD / reference-compare paths) are rejected. */
typedef struct {
    int32_t n_reads;
    int32_t min_bq, noisy_reg_max_xgaps, noisy_reg_slide_win, end_clip_reg, end_clip_reg_flank_win;   /* call_var_opt_t */
    double max_noisy_frac_per_read, max_var_ratio_per_read;
    int64_t whole_ref_len;             /* chunk->whole_ref_len */
    int64_t reg_beg, reg_end;          /* chunk->reg_beg / reg_end */
    const int32_t *ordered_read_ids;   /* chunk->ordered_read_ids [n_reads] */
    const uint8_t *is_skipped;         /* chunk->is_skipped (reads already skipped by the loader) */
    const int64_t *read_pos0;          /* bam1_t core.pos (0-based) */
    const uint8_t *read_is_rev;        /* bam_is_rev */
    const uint8_t *is_palindrome;      /* is_ont_palindrome_clip */
    const int32_t *n_cigar; const int64_t *cigar_off; const uint32_t *cigar;      /* BAM CIGAR words: cigar[cigar_off[r] .. +n_cigar[r]) */
    const int32_t *l_qseq; const int64_t *seq_off; const uint8_t *bseq;           /* 4-bit packed SEQ, (l_qseq + 1) / 2 bytes per read */
    const int64_t *qual_off; const uint8_t *qual;                                   /* QUAL, l_qseq bytes per read */
} lcd_digar_input_t;
typedef struct {
    uint8_t *skip;                     /* [n_reads] 1: collect_digar_from_eqx_cigar returns -1 */
    int64_t *read_beg, *read_end;      /* digar_t.beg / end */
    int64_t *digar_first; int32_t *n_digar;                                        /* read r's records: digar_*[digar_first[r] .. +n_digar[r]) */
    int64_t *digar_pos; int8_t *digar_type; int32_t *digar_len, *digar_qi; uint8_t *digar_low_qual; int64_t *digar_alt_off; uint8_t *digar_alt;
    int64_t digar_cap, alt_cap;        /* capacities of digar_*[] / digar_alt[] (lcd_digar_capacity or lcd_digar_plan_sizes) */
    int64_t *nreg_first; int32_t *n_nreg; int64_t *nreg_beg, *nreg_end; int32_t *nreg_label;   /* digar_t.noisy_regs, cr_index order */
    int64_t nreg_cap;
    int64_t *cnreg_beg, *cnreg_end; int32_t *cnreg_label; int64_t cnreg_cap, n_cnreg;          /* chunk->chunk_noisy_regs additions, cr_add order */
    int64_t *qual_counts;              /* chunk->qual_counts [256] */
    int64_t n_digar_total, n_alt_total, n_nreg_total;                                /* entries used */
} lcd_digar_output_t;
/* Capacities that always suffice for one chunk (one host pass over its CIGAR words). */
int lcd_digar_capacity(const lcd_digar_input_t *in, int64_t *digar_cap, int64_t *alt_cap, int64_t *nreg_cap);
int lcd_digar_batch(int n_chunks, const lcd_digar_input_t *in, lcd_digar_output_t *out);
/* Reads with plain-M CIGARs and an MD tag (collect_digar_from_MD_tag, src/bam_utils.c:1003-1174; the reference takes this path for reads
 * without =/X ops and without a cs tag, src/collect_var.c:1072-1080): the MD strings go to the device, where the reference's walk over
 * (CIGAR, MD) (:1037-1094) turns them into the =/X CIGARs the kernels consume.  md_off[r] < 0: read r's CIGAR is =/X already.
 * Capacities: lcd_digar_capacity counts an M op as one record; size the outputs for l_qseq more records / alt bases per read instead
 * (or ask lcd_digar_plan_sizes after lcd_plan_run). */
typedef struct { const int64_t *md_off; const char *md; } lcd_md_tags_t;      /* per chunk: NUL-terminated tag of read r at md + md_off[r] */
int lcd_digar_md_batch(int n_chunks, const lcd_digar_input_t *in, const lcd_md_tags_t *tags, lcd_digar_output_t *out);
lcd_plan_t *lcd_digar_md_plan_create(int n_chunks, const lcd_digar_input_t *in, const lcd_md_tags_t *tags);
/* The general form: every read names the variant the reference's driver would pick for it (src/collect_var.c:1072-1080) --
 *   LCD_TAG_EQX     its CIGAR has =/X ops                       (collect_digar_from_eqx_cigar, src/bam_utils.c:701)
 *   LCD_TAG_CS      plain-M CIGAR + cs tag at text + off[r]     (collect_digar_from_cs_tag,   :844; clips from the first / last CIGAR op only,
 *                   introns not advanced over; the tag's letters must spell the read's own SEQ bases -- else the call fails loudly)
 *   LCD_TAG_MD      plain-M CIGAR + MD tag at text + off[r]     (collect_digar_from_MD_tag,   :1003)
 *   LCD_TAG_REFSEQ  plain-M CIGAR, no tag: bases are compared with the chunk's reference window ref_seq = positions ref_beg .. ref_end,
 *                   1-based inclusive (collect_digar_from_ref_seq, :1176; bases outside the window are passed over as the reference does)
 * -- and a front end on the device rewrites (CIGAR, tag / reference) into the op stream the difference-list kernels consume. */
enum { LCD_TAG_EQX = -1, LCD_TAG_MD = 0, LCD_TAG_CS = 1, LCD_TAG_REFSEQ = 2 };
typedef struct {
    const int8_t *kind;            /* [n_reads] LCD_TAG_* */
    const int64_t *off;            /* [n_reads] offset of the read's NUL-terminated tag value in text (LCD_TAG_MD / LCD_TAG_CS), else ignored */
    const char *text;
    const char *ref_seq;           /* the chunk's reference window (ASCII), needed when a read is LCD_TAG_REFSEQ */
    int64_t ref_beg, ref_end;
} lcd_read_tags_t;
int lcd_digar_tags_batch(int n_chunks, const lcd_digar_input_t *in, const lcd_read_tags_t *tags, lcd_digar_output_t *out);
lcd_plan_t *lcd_digar_tags_plan_create(int n_chunks, const lcd_digar_input_t *in, const lcd_read_tags_t *tags);
lcd_plan_t *lcd_digar_plan_create(int n_chunks, const lcd_digar_input_t *in);
/* After lcd_plan_run: exact sizes of chunk i's outputs (records, alt bases, per-read intervals; the chunk list needs <= the last). */
int  lcd_digar_plan_sizes(lcd_plan_t *plan, void *stream, int chunk, int64_t *n_digar, int64_t *n_alt, int64_t *n_nreg);
int  lcd_digar_plan_fetch(lcd_plan_t *plan, void *stream, lcd_digar_output_t *out);

/* ---------------------------------------------------------------- K2: pileup scan, per-site coverage
 * Replaces int collect_cand_vars(const call_var_opt_t *opt, bam_chunk_t *chunk, int n_var_sites, var_site_t *var_sites)
 * (src/collect_var.c:238-249: update_cand_vars_from_digar, src/bam_utils.c:287-329, for every kept read), called from
 * collect_var_main step 1.3 (src/collect_var.c:2913).  Inputs are the flattened digar_t / digar1_t records of the chunk's
 * reads (src/bam_utils.h:27-43) and its sorted var_site_t list (src/collect_var.h:43-47); the output is, per site, what the
 * reference leaves in cand_var_t: total_cov, low_qual_cov, alle_covs[0..1], strand_to_alle_covs[0..1][0..1]. */
typedef struct {
    int32_t n_reads, n_sites;
    int32_t min_bq, min_sv_len;        /* opt->min_bq, opt->min_sv_len */
    const int32_t *ordered_read_ids;   /* chunk->ordered_read_ids [n_reads] */
    const uint8_t *is_skipped;         /* chunk->is_skipped [n_reads] */
    const int64_t *read_beg, *read_end;/* digar_t.beg / end (1-based, inclusive) */
    const uint8_t *read_is_rev;        /* digar_t.is_rev */
    const int64_t *digar_first;        /* first digar1_t of read r in the digar_* arrays */
    const int32_t *n_digar;            /* digar_t.n_digar */
    const int64_t *qual_off;           /* read r's base qualities: qual[qual_off[r] + qi] */
    const uint8_t *qual;
    const int64_t *digar_pos;          /* digar1_t.pos */
    const int8_t  *digar_type;         /* BAM_CEQUAL 7 / BAM_CDIFF 8 / BAM_CINS 1 / BAM_CDEL 2 / clips 4,5 */
    const int32_t *digar_len, *digar_qi;
    const uint8_t *digar_low_qual;
    const int64_t *digar_alt_off;      /* X / I: alt bases digar_alt[digar_alt_off[d] .. +len) (digar1_t.alt_seq) */
    const uint8_t *digar_alt;
    const int64_t *site_pos;           /* var_site_t.pos, sorted as collect_all_cand_var_sites leaves them */
    const int32_t *site_type, *site_ref_len, *site_alt_len;
    const int64_t *site_alt_off;       /* var_site_t.alt_seq */
    const uint8_t *site_alt;
} lcd_pileup_input_t;
typedef struct {
    int32_t *site_counts;              /* [n_sites][8]: total_cov, low_qual_cov, alle_covs[0..1], strand_to_alle_covs[0..1][0..1] */
} lcd_pileup_output_t;
int lcd_pileup_batch(int n_chunks, const lcd_pileup_input_t *in, lcd_pileup_output_t *out);
lcd_plan_t *lcd_pileup_plan_create(int n_chunks, const lcd_pileup_input_t *in);
int  lcd_pileup_plan_fetch(lcd_plan_t *plan, void *stream, lcd_pileup_output_t *out);

/* ---------------------------------------------------------------- K2b: pileup scan, category of every candidate site
 * Replaces the first loop of int classify_cand_vars(bam_chunk_t *chunk, int n_var_sites, const call_var_opt_t *opt)
 * (src/collect_var.c:902-925): int classify_var_cate(opt, ref_seq, ref_beg, ref_end, var, min_dp, min_alt_dp, min_af, max_af, ...)
 * (:413-432) for every site, with var_is_homopolymer (:306-358) and var_is_repeat_region (:361-400): depth and allele-fraction
 * thresholds, then for small indels the reference context (1-6 bp unit repeated 3x on either side; the indel itself repeated 3x).
 * The sites and counters are K2's (lcd_pileup_input_t / lcd_pileup_output_t); the result is chunk->var_i_to_cate before the
 * noisy-region pass.  The reference window must reach 24 bases beyond every site on both sides (the reference reads it unchecked;
 * chunks carry +-50 kb).  ONT's strand-bias Fisher test (var_is_strand_bias, :270) is floating point and stays on the host:
 * chunks with is_ont set add var_is_strand_bias (src/collect_var.c:270: two-tailed Fisher exact test, p < 0.01 -> LONGCALLD_STRAND_BIAS_VAR). */
typedef struct {
    int32_t n_sites;
    int32_t min_dp, min_alt_dp;        /* opt->min_dp, opt->min_alt_dp */
    int32_t max_xgaps;                 /* opt->noisy_reg_max_xgaps: longer indels skip the homopolymer / repeat tests */
    int32_t is_ont;                    /* opt->is_ont: adds the strand-bias test of ONT data */
    int32_t pad;
    double min_af, max_af;             /* opt->min_af, opt->max_af */
    int64_t ref_beg, ref_end;          /* chunk->ref_beg / ref_end: ref_seq[0] is base ref_beg */
    const char *ref_seq;               /* chunk->ref_seq (ASCII) */
    const int64_t *site_pos;           /* cand_var_t.pos / var_type / ref_len / alt_len / alt_seq, as in lcd_pileup_input_t */
    const int32_t *site_type, *site_ref_len, *site_alt_len;
    const int64_t *site_alt_off;
    const uint8_t *site_alt;
    const int32_t *site_counts;        /* [n_sites][8]: lcd_pileup_output_t (total_cov, low_qual_cov, alle_covs[0..1], strand x allele) */
} lcd_classify_input_t;
typedef struct { int32_t *var_cate; } lcd_classify_output_t;      /* [n_sites]: LONGCALLD_* category (src/collect_var.h:11-24) */
int lcd_classify_batch(int n_chunks, const lcd_classify_input_t *in, lcd_classify_output_t *out);
lcd_plan_t *lcd_classify_plan_create(int n_chunks, const lcd_classify_input_t *in);
int  lcd_classify_plan_fetch(lcd_plan_t *plan, void *stream, lcd_classify_output_t *out);
/* K2 -> K2b in place: the categories of the sites a pileup plan (any of lcd_pileup_plan_create*, run) holds in HBM, from the counters
 * it left there; only the thresholds and the chunk's reference window are uploaded.  The window must reach 24 bases beyond every site
 * (a chunk's own ref_beg / ref_end, +-50 kb around its region, always does).  The pileup plan must outlive the plan. */
typedef struct {
    int32_t min_dp, min_alt_dp, max_xgaps, is_ont;   /* as in lcd_classify_input_t */
    double min_af, max_af;
    int64_t ref_beg, ref_end;
    const char *ref_seq;
} lcd_classify_params_t;
lcd_plan_t *lcd_classify_plan_create_on_pileup(lcd_plan_t *pileup_plan, int n_chunks, const lcd_classify_params_t *params);

/* ---------------------------------------------------------------- K0: low-complexity intervals of the reference
 * Replaces uint64_t *sdust(void *km, const uint8_t *seq, int l_seq, int T, int W, int *n) (src/sdust.c:184, symmetric DUST) as the chunk loader calls it to
 * fill chunk->low_comp_cr (src/bam_utils.c:1574-1583: seq = the chunk's region of the reference, T = LONGCALLD_SDUST_T = 5, W = LONGCALLD_SDUST_W = 20,
 * every interval added as (reg_beg + start - 1, reg_beg + end - 1): pass base = reg_beg - 1).  Output: the intervals in ascending order, base added to
 * the 0-based half-open (start, end) pairs sdust() returns; l_seq / 4 + 16 entries always suffice.  W up to 24 is supported. */
typedef struct { const char *seq; int32_t l_seq; int32_t T, W; int32_t pad; int64_t base; } lcd_sdust_input_t;
typedef struct { int64_t *beg, *end; int64_t cap, n; } lcd_sdust_output_t;
int lcd_sdust_batch(int n_chunks, const lcd_sdust_input_t *in, lcd_sdust_output_t *out);
lcd_plan_t *lcd_sdust_plan_create(int n_chunks, const lcd_sdust_input_t *in);
int  lcd_sdust_plan_fetch(lcd_plan_t *plan, void *stream, lcd_sdust_output_t *out);

/* ---------------------------------------------------------------- K2c: pileup scan, the noisy-region set (second half)
 * Replaces void pre_process_noisy_regs(bam_chunk_t *chunk, call_var_opt_t *opt) (src/collect_var.c:557-643) followed by
 * int classify_cand_vars(bam_chunk_t *chunk, int n_var_sites, const call_var_opt_t *opt) (:902-1033) after its classify_var_cate loop (K2b),
 * for out_somatic = 0: the reads' noisy intervals are widened by the low-complexity intervals they touch (:538-553), merged with the
 * reference's own cr_merge (src/cgranges.c:225-301: an interval absorbs later ones within min(label, label') of its growing end), kept when
 * at least min_alt_dp reads and min_af of the reads spanning them are noisy there; sites inside a region are dropped, repeat-region sites and
 * sites overlapping another site (when enough of their reads carry differences there: var_noisy_reads_ratio :718) open regions of their own
 * (cr_add_var_cr :754, cr_merge2), every region is widened by noisy_reg_flank_len past the candidate sites next to it (:482-535) and merged
 * again; what is left of the sites is chunk->cand_vars.  Interval lists are cgranges (st, en, label) triples as cr_add takes them.
 * Inputs: the candidate sites in collect_all_cand_var_sites' order (ascending anchors: exact_comp_var_site) with K2b's categories, K1's per-read outputs
 * (spans, records, noisy intervals; is_skipped = the loader's flag OR K1's skip), chunk->chunk_noisy_regs in cr_add order (K1's cnreg_*), and
 * chunk->low_comp_cr with ascending starts (cr_index order; sdust on the host).  All pointers are HOST memory. */
typedef struct {
    int64_t reg_beg, reg_end;          /* chunk->reg_beg / reg_end */
    int32_t min_alt_dp, noisy_reg_flank_len, is_ont, pad;       /* call_var_opt_t (noisy_reg_merge_dis / min_sv_len reach cr_merge but are unused there) */
    double min_af;
    int32_t n_sites, n_reads;
    const int64_t *site_pos; const int32_t *site_type, *site_ref_len, *var_cate;
    int64_t n_cnreg; const int64_t *cnreg_beg, *cnreg_end; const int32_t *cnreg_label;
    int64_t n_low; const int64_t *low_beg, *low_end;
    const uint8_t *is_skipped;
    const int64_t *read_beg, *read_end;
    const int64_t *digar_first; const int32_t *n_digar; const int64_t *digar_pos; const int8_t *digar_type; const int32_t *digar_len;
    const int64_t *nreg_first; const int32_t *n_nreg; const int64_t *nreg_beg, *nreg_end;
} lcd_noisyreg_input_t;
typedef struct {
    int32_t *var_cate;                 /* [n_sites] the working var_i_to_cate after classify_cand_vars */
    uint8_t *keep;                     /* [n_sites] 1: the site stays in chunk->cand_vars (chunk->var_i_to_cate = var_cate of the kept sites, in order) */
    int64_t *reg_beg, *reg_end; int32_t *reg_label; int64_t reg_cap, n_regs;      /* chunk->chunk_noisy_regs at the end (cr_index order); n_cnreg + n_sites entries always suffice */
} lcd_noisyreg_output_t;
int lcd_noisyreg_batch(int n_chunks, const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out);
lcd_plan_t *lcd_noisyreg_plan_create(int n_chunks, const lcd_noisyreg_input_t *in);
int  lcd_noisyreg_plan_fetch(lcd_plan_t *plan, void *stream, lcd_noisyreg_output_t *out);
/* K1 -> ... -> K2b -> K2c in place: the reads' spans, records and noisy intervals are read where a digar plan (run) left them in HBM -- the chunk's
 * own noisy list is gathered from them on the device (the intervals of kept reads that touch the region, src/bam_utils.c:819-832) -- and the sites
 * with their categories where a classify plan holds them (lcd_classify_plan_create_on_pileup; it must have run, on the same stream or one this
 * plan's stream waits for, before this plan runs).  Only the options and the low-complexity intervals come from the host.  The region is the
 * digar plan's.  Fetch as above (reg_cap: the chunk's reads' noisy intervals + n_sites entries always suffice). */
typedef struct {
    int32_t min_alt_dp, noisy_reg_flank_len, is_ont, pad;
    double min_af;
    int64_t n_low; const int64_t *low_beg, *low_end;
} lcd_noisyreg_params_t;
lcd_plan_t *lcd_noisyreg_plan_create_on_classify(lcd_plan_t *digar_plan, lcd_plan_t *classify_plan, int n_chunks, const lcd_noisyreg_params_t *params);
/* The same with the low-complexity intervals read where a K0 plan (lcd_sdust_plan_create, one window per chunk, `base` chosen so that its intervals are in
 * the chunk's coordinates: what src/bam_utils.c:1574-1583 adds to chunk->low_comp_cr) leaves them -- their number too is read on the device, so the K0 plan
 * only has to have run before this one in stream order.  params[i].n_low must be 0.  A chunk whose K0 window failed fails here as well (fetch reports it). */
lcd_plan_t *lcd_noisyreg_plan_create_on_sdust(lcd_plan_t *digar_plan, lcd_plan_t *classify_plan, lcd_plan_t *sdust_plan, int n_chunks, const lcd_noisyreg_params_t *params);

/* ---------------------------------------------------------------- K3: pileup scan, read x variant profile
 * Replaces read_var_profile_t *collect_read_var_profile(const call_var_opt_t *opt, bam_chunk_t *chunk)
 * (src/collect_var.c:1389-1431: update_read_vs_all_var_profile_from_digar, src/bam_utils.c:446-552, for every kept read;
 * germline categories -- a chunk with LONGCALLD_CAND_SOMATIC_VAR candidates (-s) is rejected), called from collect_var_main
 * step 3.1 (src/collect_var.c:2933).  The sites of lcd_pileup_input_t are the chunk's cand_vars.  The output is exactly what
 * K4 (lcd_phase_input_t) consumes: prof_start / prof_end / allele_off / alleles, plus alt_qi. */
typedef struct {
    const int32_t *var_cate;           /* chunk->var_i_to_cate [n_sites] */
    const int64_t *nreg_first;         /* read r's noisy intervals (digar_t.noisy_regs): nreg_*[nreg_first[r] .. +n_nreg[r]) */
    const int32_t *n_nreg;
    const int64_t *nreg_beg, *nreg_end;/* cgranges coordinates: [beg, end) */
} lcd_profile_extra_t;
typedef struct {
    int32_t *prof_start, *prof_end;    /* read_var_profile_t.start_var_idx / end_var_idx [n_reads]; (-1, -2): none */
    int64_t *allele_off;               /* read r's row: alleles[allele_off[r] + (var - prof_start[r])] */
    int8_t  *alleles;                  /* read_var_profile_t.alleles: 0 ref, 1 alt, -1 other / not set, -2 low-quality alt */
    int32_t *alt_qi;                   /* read_var_profile_t.alt_qi */
    int64_t alleles_cap;               /* capacity of alleles[] / alt_qi[] (>= lcd_profile_capacity(in)) */
    int64_t n_alleles;                 /* out: entries used */
} lcd_profile_output_t;
int64_t lcd_profile_capacity(const lcd_pileup_input_t *in);
int lcd_profile_batch(int n_chunks, const lcd_pileup_input_t *in, const lcd_profile_extra_t *extra, lcd_profile_output_t *out);
lcd_plan_t *lcd_profile_plan_create(int n_chunks, const lcd_pileup_input_t *in, const lcd_profile_extra_t *extra);
int  lcd_profile_plan_fetch(lcd_plan_t *plan, void *stream, lcd_profile_output_t *out);

/* ---------------------------------------------------------------- K1 -> K2 / K3 in place
 * The difference lists a digar plan (lcd_digar_plan_create + lcd_plan_run) left in HBM are consumed where they lie: only the
 * chunk's candidate sites (collect_all_cand_var_sites, src/collect_var.c:1209) -- for K3 the classified candidate variants with
 * their categories (chunk->var_i_to_cate) -- are uploaded.  Reads K1 dropped (skip) are left out, as the reference leaves out
 * reads with chunk->is_skipped set.  The digar plan must outlive the plans created on it.  Results come back through
 * lcd_pileup_plan_fetch / lcd_profile_plan_fetch (row capacity of chunk i: lcd_profile_plan_capacity). */
typedef struct {
    int32_t n_sites, min_sv_len;       /* opt->min_sv_len */
    const int64_t *site_pos;           /* as in lcd_pileup_input_t */
    const int32_t *site_type, *site_ref_len, *site_alt_len;
    const int64_t *site_alt_off;
    const uint8_t *site_alt;
    const int32_t *var_cate;           /* K3 only: chunk->var_i_to_cate [n_sites] */
} lcd_site_list_t;
lcd_plan_t *lcd_pileup_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_site_list_t *sites);
lcd_plan_t *lcd_profile_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_site_list_t *sites);
int64_t lcd_profile_plan_capacity(lcd_plan_t *plan, int chunk);

/* ---------------------------------------------------------------- K1b: pileup scan, candidate-site list
 * Replaces int collect_all_cand_var_sites(const call_var_opt_t *opt, bam_chunk_t *chunk, var_site_t **var_sites)
 * (src/collect_var.c:1209-1254; comparators exact_comp_var_site / exact_comp_var_site_ins :1878-1935; filter
 * is_collectible_var_digar :1153-1160), called from collect_var_main step 1.2 (src/collect_var.c:2905): the sorted list of the
 * distinct X / I / D records (not low quality, starting inside [reg_beg, reg_end]; -1 = open) of the kept reads, large insertions
 * (>= min_sv_len) of similar length (shorter >= 0.8 x longer) at one anchor merged into the first of them.  The reads and difference
 * lists are those of lcd_pileup_input_t (its site fields are not read; n_sites may be 0).
 * Site k of a chunk stands for the record site_src[k] (index into the chunk's digar_* arrays; the lowest index among identical
 * records): its alt bases are digar_alt[digar_alt_off[site_src[k]] .. + site_alt_len[k]). */
typedef struct { int64_t reg_beg, reg_end; int32_t min_sv_len, pad; } lcd_sites_params_t;
typedef struct {
    int64_t *site_pos;                 /* var_site_t.pos / var_type / ref_len / alt_len, in collect_all_cand_var_sites' order */
    int32_t *site_type, *site_ref_len, *site_alt_len;
    int64_t *site_src;
    int64_t cap;                       /* capacity of the arrays (lcd_sites_plan_sizes, or the chunk's X / I / D record count) */
    int64_t n_sites;                   /* out */
} lcd_sites_output_t;
int lcd_sites_batch(int n_chunks, const lcd_pileup_input_t *in, const lcd_sites_params_t *params, lcd_sites_output_t *out);
lcd_plan_t *lcd_sites_plan_create(int n_chunks, const lcd_pileup_input_t *in, const lcd_sites_params_t *params);
/* on the difference lists a digar plan left in HBM (reads K1 dropped are left out); the digar plan must outlive the plan */
lcd_plan_t *lcd_sites_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_sites_params_t *params);
int  lcd_sites_plan_sizes(lcd_plan_t *plan, void *stream, int chunk, int64_t *n_sites);   /* after lcd_plan_run */
int  lcd_sites_plan_fetch(lcd_plan_t *plan, void *stream, lcd_sites_output_t *out);
/* K1 -> K1b -> K2 in place: the coverage pass on the site lists a sites plan (created on the same digar plan, and run) left in HBM;
 * nothing is uploaded.  Results through lcd_pileup_plan_fetch, chunk i's counters in the order of lcd_sites_plan_fetch. */
lcd_plan_t *lcd_pileup_plan_create_on_sites(lcd_plan_t *digar_plan, lcd_plan_t *sites_plan);

/* ---------------------------------------------------------------- K4: read -> haplotype assignment and phasing
 * Replaces int assign_hap_based_on_germline_het_vars_kmeans(const call_var_opt_t *opt, bam_chunk_t *chunk,
 * int target_var_cate) (src/assign_hap.h:12, src/assign_hap.c:473-547), called from collect_var_main
 * (src/collect_var.c:2942,2975).  One lcd_phase_input_t is the flattened view of what the reference reads from a
 * bam_chunk_t; lcd_phase_output_t is what it writes (in/out: entries the reference leaves alone keep their values,
 * e.g. variants outside target_var_cate).  Chunks of a batch are independent (one CTA each). */
typedef struct {
    int32_t n_reads, n_vars;
    int32_t target_var_cate;           /* LONGCALLD_* category mask (src/collect_var.h:11-28) */
    int32_t is_ont;                    /* opt->is_ont */
    const int32_t *ordered_read_ids;   /* chunk->ordered_read_ids [n_reads] */
    const uint8_t *is_skipped;         /* chunk->is_skipped [n_reads] */
    const int32_t *prof_start, *prof_end;  /* read_var_profile_t.start_var_idx / end_var_idx; (-1, -2) = no variant */
    const int64_t *allele_off;         /* read r's alleles: alleles[allele_off[r] + (var - prof_start[r])] */
    const int8_t *alleles;             /* read_var_profile_t.alleles: 0 ref, 1 alt, -1 other, -2 low-quality alt */
    const int32_t *var_cate;           /* chunk->var_i_to_cate [n_vars] */
    const int32_t *var_type;           /* cand_var_t.var_type: BAM_CDIFF 8 / BAM_CINS 1 / BAM_CDEL 2 */
    const int32_t *is_hp_indel;        /* cand_var_t.is_homopolymer_indel */
    const int32_t *n_uniq_alles;       /* cand_var_t.n_uniq_alles (<= 4) */
    const int32_t *alle_covs;          /* cand_var_t.alle_covs, [n_vars][4] */
    const int32_t *total_cov;          /* cand_var_t.total_cov */
    const int64_t *pos;                /* cand_var_t.pos */
} lcd_phase_input_t;
typedef struct {
    int32_t *haps;                     /* chunk->haps [n_reads] */
    int64_t *phase_sets;               /* chunk->phase_sets [n_reads] */
    int32_t *hap_to_cons_alle;         /* cand_var_t.hap_to_cons_alle, [n_vars][3] */
    int32_t *hap_to_alle_profile;      /* cand_var_t.hap_to_alle_profile, [n_vars][3][4] */
    int64_t *var_phase_set;            /* cand_var_t.phase_set [n_vars] */
    int32_t *n_clean_agree_snps, *n_clean_conflict_snps;   /* chunk->n_clean_*_snps [n_reads] */
} lcd_phase_output_t;
int lcd_phase_batch(int n_chunks, const lcd_phase_input_t *in, lcd_phase_output_t *out);
lcd_plan_t *lcd_phase_plan_create(int n_chunks, const lcd_phase_input_t *in, const lcd_phase_output_t *state);
int  lcd_phase_plan_fetch(lcd_plan_t *plan, void *stream, lcd_phase_output_t *out);

/* ---------------------------------------------------------------- K5: abPOA consensus + MSA
 * One *problem* is the progressive partial-order alignment of the reads of one (noisy region,
 * haplotype): it replaces the abPOA call sequence of abpoa_partial_aln_msa_cons (src/align.c:762-870;
 * sub_aln = 1) and abpoa_aln_msa_cons (src/align.c:872-953; sub_aln = 0, wb = -1) for full-cover reads:
 * abpoa_init / abpoa_init_para / abpoa_post_set_para / abpoa_align_sequence_to_subgraph /
 * abpoa_add_subgraph_alignment (or abpoa_msa) / abpoa_output, then reading ab->abc
 * (cons_base[0], cons_len[0], msa_base, msa_len).  Results are those of abPOA's AVX-512BW build. */
typedef struct {
    int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2;  /* 2,6,6,2,24,1: src/align.h:21-26 */
    int32_t wb; float wf;          /* adaptive band (abpoa.h:17-18: 10, 0.01); wb = -1: unbanded */
    int32_t sub_aln;               /* 1: phased set-up (inc_both_ends = 0, span-read consensus rule); 0: abpoa_msa */
    int32_t max_n_cons;            /* 1; 2 (de-novo read clustering, sub_aln = 0) through lcd_poa_ncons_* only */
} lcd_poa_params_t;

enum { LCD_POA_OK = 0, LCD_POA_NEEDS_INT32 = -1, LCD_POA_BAND = -2, LCD_POA_BACKTRACK = -3,
       LCD_POA_NO_BASE = -4, LCD_POA_OOM = -5, LCD_POA_MSA_CAP = -6, LCD_POA_BAD_ANCHORS = -7, LCD_POA_SUB_UNSUPPORTED = -8 };
typedef struct {
    int32_t status;                /* LCD_POA_* */
    int32_t cons_len;              /* abc->cons_len[0] */
    int32_t msa_len;               /* abc->msa_len */
    int32_t n_nodes;               /* abg->node_n */
} lcd_poa_result_t;

/* Problem i owns reads first_read[i] .. first_read[i]+n_reads[i]-1; read r is seqs[read_off[r] .. +read_len[r])
 * (base codes 0..4), aligned in that order.  cons[cons_off[i] ..] needs the sum of the problem's read
 * lengths.  msa may be NULL; otherwise problem i's (n_reads+1) x msa_len row-major matrix (gap = 5, last
 * row = consensus) goes to msa[msa_off[i] ..] if it fits in msa_cap[i] bytes (else LCD_POA_MSA_CAP). */
int lcd_poa_batch(int n, const uint8_t *seqs, size_t seqs_len,
                  const int32_t *first_read, const int32_t *n_reads,
                  const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                  const lcd_poa_params_t *params,
                  uint8_t *cons, const int64_t *cons_off,
                  uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                  lcd_poa_result_t *results);
lcd_plan_t *lcd_poa_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                const int32_t *first_read, const int32_t *n_reads,
                                const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                const lcd_poa_params_t *params);
int  lcd_poa_plan_fetch(lcd_plan_t *plan, void *stream, uint8_t *cons, const int64_t *cons_off,
                        uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                        lcd_poa_result_t *results);

/* Partially covering reads (abpoa_partial_aln_msa_cons, src/align.c:790-812): read r > 0 of a problem is aligned against the SUB-GRAPH
 * abpoa_subgraph_nodes (abPOA/src/abpoa_graph.c:666) finds between two nodes of the first read -- sub_beg[r] = ref_beg + 1 and
 * sub_end[r] = ref_end + 1, with ref_beg / ref_end the 1-based positions in the first read that collect_partial_aln_beg_end
 * (src/align.c:709) returns, and the read passed here already cut to [read_beg, read_end] -- or against the whole graph (both 0), or is
 * left out of the graph (sub_beg[r] < 0: its MSA row stays empty).  sub_beg / sub_end are indexed like read_off / read_len; both NULL is
 * lcd_poa_batch.  The device keeps abPOA's BFS node index for such problems (the sub-graph and the nodes a read spans are index ranges). */
int lcd_poa_sub_batch(int n, const uint8_t *seqs, size_t seqs_len,
                      const int32_t *first_read, const int32_t *n_reads,
                      const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                      const int32_t *sub_beg, const int32_t *sub_end,
                      const lcd_poa_params_t *params,
                      uint8_t *cons, const int64_t *cons_off,
                      uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                      lcd_poa_result_t *results);
lcd_plan_t *lcd_poa_sub_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                    const int32_t *first_read, const int32_t *n_reads,
                                    const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                    const int32_t *sub_beg, const int32_t *sub_end,
                                    const lcd_poa_params_t *params);

/* Two consensus sequences (abpoa_aln_msa_cons with max_n_cons = 2, src/align.c:872-953, called by wfa_collect_noisy_aln_str_no_ps_hap :1180 for a
 * region without a usable phase set): after the progressive POA of all reads the device runs abPOA's read clustering on the row-column MSA
 * (abpoa_multip_read_clu_kmedoids, abPOA/src/abpoa_output.c:1135-1180: candidate het columns, merged read partitions, k-medoids) and, when it
 * finds two clusters, one most-frequent consensus per cluster (abpoa_most_frequent :549-586).  min_freq[i] = abpt->min_freq = opt->min_af (a
 * double, as the reference keeps it: min_w = MAX(2, ceil(n_reads * min_freq))).  Per problem: n_cons (abc->n_cons: 1 or 2), cons_len2 (length
 * of the second consensus, stored right after the first at cons[cons_off[i] + results[i].cons_len]); per read (indexed like read_off):
 * read_cluster (abc->clu_read_ids as a 0 / 1 label).  The MSA has n_reads + n_cons rows (msa_cap[i] must allow n_reads + 2).  Problems
 * with max_n_cons = 1 may share the batch. */
int lcd_poa_ncons_batch(int n, const uint8_t *seqs, size_t seqs_len,
                        const int32_t *first_read, const int32_t *n_reads,
                        const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                        const lcd_poa_params_t *params, const double *min_freq,
                        uint8_t *cons, const int64_t *cons_off,
                        uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                        lcd_poa_result_t *results, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster);
lcd_plan_t *lcd_poa_ncons_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                      const int32_t *first_read, const int32_t *n_reads,
                                      const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                      const lcd_poa_params_t *params, const double *min_freq);
/* after lcd_poa_plan_fetch; any of the three outputs may be NULL */
int lcd_poa_plan_fetch_clusters(lcd_plan_t *plan, void *stream, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster);

#ifdef __cplusplus
}
#endif
#endif
