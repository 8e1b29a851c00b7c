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

#ifdef __cplusplus
}
#endif
#endif
