/* oracle/lcd_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, single-threaded restatements of the reference algorithms on the hot path of
 * yangao07/longcallD @ 491f055 (per-region worker).  Each function cites the reference
 * file:line it follows.  The restatement is pinned against the unmodified reference compiled
 * into oracle/_ref/liblcdref.so (tests/test_oracle_vs_ref.py) and against WFA2-lib's own
 * golden vectors (WFA2-lib/tests/wfa.utest.check, committed subset under tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#ifndef LCD_ORACLE_H
#define LCD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- WFA (WFA2-lib/wavefront) ------------------------------------------------------- */
enum { LCD_WFA_HEUR_NONE = 0, LCD_WFA_HEUR_ADAPTIVE = 1, LCD_WFA_HEUR_ZDROP = 2 };
enum { LCD_WFA_STATUS_COMPLETED = 0, LCD_WFA_STATUS_PARTIAL = 1, LCD_WFA_STATUS_ERROR = -1 };

typedef struct {
    int32_t mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2; /* match is always 0 */
    int32_t affine2p;          /* 1: gap-affine-2p, 0: gap-affine (o1,e1 only) */
    int32_t heuristic;         /* LCD_WFA_HEUR_* */
    int32_t min_wavefront_length, max_distance_threshold; /* wf-adaptive */
    int32_t zdrop;             /* z-drop */
    int32_t steps_between_cutoffs;
} lcd_wfa_params_t;

typedef struct {
    int32_t status;            /* LCD_WFA_STATUS_* */
    int32_t score;             /* cigar->score after termination (WFA sign conventions) */
    int32_t n_ops;             /* number of edit operations written to ops[] */
    int32_t end_v, end_h;      /* cigar->end_v / end_h */
} lcd_wfa_result_t;

/* ops must hold 2*(plen+tlen)+2 bytes; filled with 'M','X','I','D' (begin..end of the cigar). */
int lcd_oracle_wfa_align(const uint8_t *pattern, int plen, const uint8_t *text, int tlen,
                         const lcd_wfa_params_t *par, char *ops, lcd_wfa_result_t *res);

/* ---- edlib (edlib/src/edlib.cpp) ---------------------------------------------------- */
enum { LCD_EDLIB_MODE_NW = 0, LCD_EDLIB_MODE_SHW = 1, LCD_EDLIB_MODE_HW = 2 };
typedef struct {
    int32_t status;            /* 0 ok */
    int32_t edit_distance;
    int32_t start_loc, end_loc;/* first start / first end location (target coords), -1 if none */
    int32_t aln_len;           /* number of entries written to aln[] (0 when want_path == 0) */
} lcd_edlib_result_t;
/* aln must hold qlen+tlen bytes; codes 0 '=', 1 insert (query base), 2 delete (target base), 3 'X'. */
int lcd_oracle_edlib_align(const uint8_t *query, int qlen, const uint8_t *target, int tlen,
                           int mode, int want_path, uint8_t *aln, lcd_edlib_result_t *res);

/* ---- abPOA (abPOA/src) --------------------------------------------------------------- */
typedef struct {
    int32_t match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2;  /* 2,6,6,2,24,1 (src/align.h:21-26) */
    int32_t wb; float wf;      /* adaptive band: 10, 0.01 (phased POA); wb = -1: unbanded (de-novo POA) */
    int32_t sub_aln;           /* 1: abpoa_partial_aln_msa_cons (inc_both_ends = 0, span-read consensus rule);
                                  0: abpoa_aln_msa_cons / abpoa_msa (inc_both_ends = 1) */
    int32_t max_n_cons;        /* 1 (2 = de-novo clustering: not restated yet) */
} lcd_poa_params_t;
/* One progressive POA over n_seq full-cover sequences (base codes 0..4): consensus (most frequent) and
 * the row-column MSA, (n_seq + 1) rows of msa_len columns (gap = 5), row n_seq = consensus.
 * cons must hold sum(seq_len) bytes.  Returns 0, or <0 when the case is outside the restated path. */
int lcd_oracle_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                   const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
                   uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
int lcd_oracle_poa_sub(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                       const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap);

/* ---- pileup scan, step 2 (rest): the chunk's noisy-region set and the sites that stay clean-region candidates ---------------------
 * pre_process_noisy_regs (src/collect_var.c:557-643) then classify_cand_vars after its first loop (:925-1033, out_somatic = 0) with
 * cr_extend_noisy_regs_with_low_comp (:538-553), cr_add_var_cr (:754-778), var_noisy_reads_ratio (:657-751), post_process_noisy_regs +
 * collect_noisy_reg_start_end (:482-535,646-655) and the reference's own cr_merge / cr_merge2 / cr_is_contained (src/cgranges.c:225-335,512-530).
 * Interval lists are cgranges (st, en, label) triples as cr_add takes them. */
typedef struct {
    int64_t reg_beg, reg_end;          /* chunk->reg_beg / reg_end */
    int32_t min_alt_dp, noisy_reg_flank_len, is_ont, pad;       /* call_var_opt_t (noisy_reg_merge_dis / min_sv_len reach cr_merge but are unused there) */
    double min_af;
    int32_t n_sites;                   /* candidate sites in collect_all_cand_var_sites' order, with classify_var_cate's category (K2b) */
    int32_t n_reads;
    const int64_t *site_pos; const int32_t *site_type, *site_ref_len, *var_cate;
    int64_t n_cnreg; const int64_t *cnreg_beg, *cnreg_end; const int32_t *cnreg_label;       /* chunk->chunk_noisy_regs as K1 leaves it (cr_add order) */
    int64_t n_low; const int64_t *low_beg, *low_end;                                          /* chunk->low_comp_cr (sdust; may be empty) */
    const uint8_t *is_skipped;         /* chunk->is_skipped after K1 */
    const int64_t *read_beg, *read_end;                                                        /* digar_t.beg / end */
    const int64_t *digar_first; const int32_t *n_digar; const int64_t *digar_pos; const int8_t *digar_type; const int32_t *digar_len;
    const int64_t *nreg_first; const int32_t *n_nreg; const int64_t *nreg_beg, *nreg_end;     /* digar_t.noisy_regs */
} lcd_noisyreg_input_t;
typedef struct {
    int32_t *var_cate;                 /* [n_sites] the working var_i_to_cate after classify_cand_vars */
    uint8_t *keep;                     /* [n_sites] 1: the site stays in chunk->cand_vars (chunk->var_i_to_cate gets var_cate of the kept sites, in order) */
    int64_t *reg_beg, *reg_end; int32_t *reg_label; int64_t reg_cap, n_regs;                   /* chunk->chunk_noisy_regs at the end (cr_index order) */
} lcd_noisyreg_output_t;
int lcd_oracle_noisy_regs(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out);

/* ---- symmetric DUST over a reference window (src/sdust.c as the loader calls it, src/bam_utils.c:1574-1583) -> chunk->low_comp_cr ---- */
int lcd_oracle_sdust(const uint8_t *seq, int l_seq, int T, int W, int64_t *beg, int64_t *end, int64_t cap);

/* ---- read -> haplotype assignment and phasing (src/assign_hap.c:16-547) ----------------------------- */
typedef struct {
    int32_t n_reads, n_vars;
    int32_t target_var_cate;           /* LONGCALLD_* category mask (src/collect_var.h:11-28) */
    int32_t is_ont;                    /* opt->is_ont */
    const int32_t *ordered_read_ids;   /* chunk->ordered_read_ids [n_reads] */
    const uint8_t *is_skipped;         /* chunk->is_skipped [n_reads] */
    const int32_t *prof_start, *prof_end;  /* read_var_profile_t.start_var_idx / end_var_idx; (-1, -2) = no variant */
    const int64_t *allele_off;         /* read r's alleles: alleles[allele_off[r] + (var - prof_start[r])] */
    const int8_t *alleles;             /* 0 ref, 1 alt, -1 other, -2 low-quality alt */
    const int32_t *var_cate;           /* chunk->var_i_to_cate [n_vars] */
    const int32_t *var_type;           /* cand_var_t.var_type: BAM_CDIFF 8 / BAM_CINS 1 / BAM_CDEL 2 */
    const int32_t *is_hp_indel;        /* cand_var_t.is_homopolymer_indel */
    const int32_t *n_uniq_alles;       /* <= 4 */
    const int32_t *alle_covs;          /* [n_vars][4] */
    const int32_t *total_cov;
    const int64_t *pos;                /* cand_var_t.pos */
} lcd_phase_input_t;
typedef struct {
    int32_t *haps;                     /* chunk->haps [n_reads] */
    int64_t *phase_sets;               /* chunk->phase_sets [n_reads] */
    int32_t *hap_to_cons_alle;         /* [n_vars][3] (only variants in the target mask are written) */
    int32_t *hap_to_alle_profile;      /* [n_vars][3][4] */
    int64_t *var_phase_set;            /* [n_vars] */
    int32_t *n_clean_agree_snps, *n_clean_conflict_snps;   /* [n_reads] */
} lcd_phase_output_t;
int lcd_oracle_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out);
/* The order cgranges returns intervals in after cr_index (in-place MSD radix sort by start, not stable). */
void lcd_oracle_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out);

/* ---- pileup scan: per-site coverage (src/collect_var.c:238-249, src/bam_utils.c:287-329) ------------- */
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
    const int64_t *digar_alt_off;      /* X / I: alt bases digar_alt[digar_alt_off[d] .. +len) */
    const uint8_t *digar_alt;
    const int64_t *site_pos;           /* var_site_t.pos, sorted as collect_all_cand_var_sites leaves them */
    const int32_t *site_type, *site_ref_len, *site_alt_len;
    const int64_t *site_alt_off;
    const uint8_t *site_alt;
} lcd_pileup_input_t;
typedef struct {
    int32_t *site_counts;              /* [n_sites][8]: total_cov, low_qual_cov, alle_covs[0..1], strand_to_alle_covs[0..1][0..1] */
} lcd_pileup_output_t;
int lcd_oracle_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out);

/* ---- read x variant profile (src/collect_var.c:1389-1431, src/bam_utils.c:446-552), germline categories ---- */
typedef struct {
    const int32_t *var_cate;           /* chunk->var_i_to_cate [n_sites]; the sites of lcd_pileup_input_t are the cand_vars */
    const int64_t *nreg_first;         /* read r's noisy intervals (digar_t.noisy_regs after cr_index): nreg_*[nreg_first[r] .. +n_nreg[r]) */
    const int32_t *n_nreg;
    const int64_t *nreg_beg, *nreg_end;/* cgranges coordinates: [beg, end) */
} lcd_profile_extra_t;
typedef struct {
    int32_t *prof_start, *prof_end;    /* read_var_profile_t.start_var_idx / end_var_idx [n_reads]; (-1, -2): none */
    int64_t *allele_off;               /* read r's row: alleles[allele_off[r] + (var - prof_start[r])] (what K4 consumes) */
    int8_t  *alleles;                  /* 0 ref, 1 alt, -1 other / not set, -2 low-quality alt */
    int32_t *alt_qi;                   /* read offset of the alt allele, -1 otherwise */
    int64_t alleles_cap;               /* capacity of alleles[] / alt_qi[]: sum over reads of (variants with pos in the read's span + 2) suffices */
    int64_t n_alleles;                 /* entries used */
} lcd_profile_output_t;
int lcd_oracle_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out);

/* ---- pileup scan, step 1: difference lists from =/X CIGARs (src/bam_utils.c:701-841, :161-205) ------------ */
typedef struct {
    int32_t n_reads;
    int32_t min_bq, noisy_reg_max_xgaps, noisy_reg_slide_win, end_clip_reg, end_clip_reg_flank_win;   /* call_var_opt_t */
    double max_noisy_frac_per_read, max_var_ratio_per_read;
    int64_t whole_ref_len;             /* chunk->whole_ref_len */
    int64_t reg_beg, reg_end;          /* chunk->reg_beg / reg_end: a kept read's intervals overlapping it go to chunk_noisy_regs */
    const int32_t *ordered_read_ids;   /* chunk->ordered_read_ids [n_reads] */
    const uint8_t *is_skipped;         /* chunk->is_skipped (reads already skipped by the loader) */
    const int64_t *read_pos0;          /* bam1_t core.pos (0-based) */
    const uint8_t *read_is_rev;        /* bam_is_rev */
    const uint8_t *is_palindrome;      /* is_ont_palindrome_clip (SA-tag test, host side; 0 for HiFi) */
    const int32_t *n_cigar; const int64_t *cigar_off; const uint32_t *cigar;      /* BAM CIGAR words */
    const int32_t *l_qseq; const int64_t *seq_off; const uint8_t *bseq;           /* 4-bit packed SEQ, (l_qseq+1)/2 bytes per read */
    const int64_t *qual_off; const uint8_t *qual;                                   /* QUAL, l_qseq bytes per read */
} lcd_digar_input_t;
typedef struct {
    uint8_t *skip;                     /* 1: the reference returns -1 (read dropped as BAM_RECORD_WRONG_MAP) */
    int64_t *read_beg, *read_end;      /* digar_t.beg / end */
    int64_t *digar_first; int32_t *n_digar;                                        /* per read: its digar1_t records */
    int64_t *digar_pos; int8_t *digar_type; int32_t *digar_len, *digar_qi; uint8_t *digar_low_qual; int64_t *digar_alt_off; uint8_t *digar_alt;
    int64_t digar_cap, alt_cap;        /* capacities of the digar_* arrays / digar_alt */
    int64_t *nreg_first; int32_t *n_nreg; int64_t *nreg_beg, *nreg_end; int32_t *nreg_label;   /* digar_t.noisy_regs in the order cr_index leaves them */
    int64_t nreg_cap;
    int64_t *cnreg_beg, *cnreg_end; int32_t *cnreg_label; int64_t cnreg_cap, n_cnreg;          /* chunk->chunk_noisy_regs in cr_add order (:819-832) */
    int64_t *qual_counts;              /* chunk->qual_counts [256] */
    int64_t n_digar_total, n_alt_total, n_nreg_total;
} lcd_digar_output_t;
int lcd_oracle_collect_digar_eqx(const lcd_digar_input_t *in, lcd_digar_output_t *out);
/* (CIGAR with M, MD tag) -> the equivalent =/X CIGAR, as collect_digar_from_MD_tag walks them (src/bam_utils.c:1037-1094); oracle/md.c */
/* the cs-tag variant (src/bam_utils.c:844-1001): read r's tag is the NUL-terminated string at cs + cs_off[r]; oracle/digar_cs.c */
int lcd_oracle_collect_digar_cs(const lcd_digar_input_t *in, const int64_t *cs_off, const char *cs, lcd_digar_output_t *out);
int lcd_oracle_collect_digar_refseq(const lcd_digar_input_t *in, const char *ref_seq, int64_t ref_beg, int64_t ref_end, lcd_digar_output_t *out);
int64_t lcd_oracle_md_to_eqx(int n_cigar, const uint32_t *cigar, const char *md, uint32_t *out, int64_t cap);

/* ---- pileup scan, step 1.2: sorted unique candidate sites (src/collect_var.c:1209-1254) ---------------------------- */
typedef struct {
    int64_t *site_pos; int32_t *site_type, *site_ref_len, *site_alt_len;   /* var_site_t, in collect_all_cand_var_sites' order */
    int64_t *site_src;                 /* a record (index into the digar_* arrays) whose alt bases are the site's alt_seq */
    int64_t cap, n_sites;
} lcd_sites_output_t;
int lcd_oracle_collect_sites(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end, lcd_sites_output_t *out);

/* ---- pileup scan, step 2 (first loop): category of every candidate site (src/collect_var.c:413-432 with :306-400) ---------- */
typedef struct {
    int32_t n_sites;
    int32_t min_dp, min_alt_dp;        /* opt->min_dp, opt->min_alt_dp */
    int32_t max_xgaps;                 /* opt->noisy_reg_max_xgaps: longer indels skip the homopolymer / repeat tests */
    int32_t is_ont;                    /* opt->is_ont (the strand-bias Fisher test of ONT data is not restated: must be 0) */
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
int lcd_oracle_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate);

#ifdef __cplusplus
}
#endif
#endif
