// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin flat-C entry points over the UNMODIFIED reference libraries (WFA2-lib, edlib, abPOA as
// vendored by yangao07/longcallD @ 491f055) so that tests and bench.py's cpu_baseline leg can call
// them through ctypes with the same argument structs as our oracle port and our CUDA C-ABI.
// Compiled (only where /root/reference exists) into oracle/_ref/libref_shim.so by oracle/Makefile;
// it configures the libraries exactly as the reference's src/align.c does (cited per function).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"
#include <thread>
#include <atomic>
#include <vector>
extern "C" {
#include "wavefront/wavefront_align.h"
}
#include "edlib.h"

extern "C" {

// Mirrors wfa_end2end_aln's aligner set-up (src/align.c:379-408) minus the sequence reversal,
// which is host glue above the kernel.
int ref_wfa_align(const uint8_t *pattern, int plen, const uint8_t *text, int tlen,
                  const lcd_wfa_params_t *par, char *ops, lcd_wfa_result_t *res) {
    wavefront_aligner_attr_t attributes = wavefront_aligner_attr_default;
    if (par->affine2p) {
        attributes.distance_metric = gap_affine_2p;
        attributes.affine2p_penalties.match = 0;
        attributes.affine2p_penalties.mismatch = par->mismatch;
        attributes.affine2p_penalties.gap_opening1 = par->gap_open1;
        attributes.affine2p_penalties.gap_extension1 = par->gap_ext1;
        attributes.affine2p_penalties.gap_opening2 = par->gap_open2;
        attributes.affine2p_penalties.gap_extension2 = par->gap_ext2;
    } else {
        attributes.distance_metric = gap_affine;
        attributes.affine_penalties.match = 0;
        attributes.affine_penalties.mismatch = par->mismatch;
        attributes.affine_penalties.gap_opening = par->gap_open1;
        attributes.affine_penalties.gap_extension = par->gap_ext1;
    }
    attributes.alignment_scope = compute_alignment;
    attributes.alignment_form.span = alignment_end2end;
    if (par->heuristic == LCD_WFA_HEUR_NONE) attributes.heuristic.strategy = wf_heuristic_none;
    else if (par->heuristic == LCD_WFA_HEUR_ADAPTIVE) {
        attributes.heuristic.strategy = wf_heuristic_wfadaptive;
        attributes.heuristic.min_wavefront_length = par->min_wavefront_length;
        attributes.heuristic.max_distance_threshold = par->max_distance_threshold;
        attributes.heuristic.steps_between_cutoffs = par->steps_between_cutoffs;
    } else {
        attributes.heuristic.strategy = wf_heuristic_zdrop;
        attributes.heuristic.zdrop = par->zdrop;
        attributes.heuristic.steps_between_cutoffs = par->steps_between_cutoffs;
    }
    wavefront_aligner_t *const wf = wavefront_aligner_new(&attributes);
    wavefront_align(wf, (const char*)pattern, plen, (const char*)text, tlen);
    cigar_t *c = wf->cigar;
    int n = c->end_offset - c->begin_offset; if (n < 0) n = 0;
    if (ops) { memcpy(ops, c->operations + c->begin_offset, n); ops[n] = '\0'; }
    res->status = (wf->align_status.status == WF_STATUS_ALG_COMPLETED) ? LCD_WFA_STATUS_COMPLETED
                : (wf->align_status.status == WF_STATUS_ALG_PARTIAL) ? LCD_WFA_STATUS_PARTIAL : LCD_WFA_STATUS_ERROR;
    res->score = c->score; res->n_ops = n; res->end_v = c->end_v; res->end_h = c->end_h;
    wavefront_aligner_delete(wf);
    return 0;
}

// Mirrors edlib_xgaps / edlib_end2end_aln / edlib_infix_aln (src/align.c:222-275): k = -1, no
// custom equalities, TASK_PATH or TASK_DISTANCE.
int ref_edlib_align(const uint8_t *query, int qlen, const uint8_t *target, int tlen,
                    int mode, int want_path, uint8_t *aln, lcd_edlib_result_t *res) {
    EdlibAlignMode m = mode == LCD_EDLIB_MODE_NW ? EDLIB_MODE_NW : mode == LCD_EDLIB_MODE_SHW ? EDLIB_MODE_SHW : EDLIB_MODE_HW;
    EdlibAlignResult r = edlibAlign((const char*)query, qlen, (const char*)target, tlen,
                                    edlibNewAlignConfig(-1, m, want_path ? EDLIB_TASK_PATH : EDLIB_TASK_DISTANCE, NULL, 0));
    res->status = r.status; res->edit_distance = r.editDistance;
    res->start_loc = (r.startLocations && r.numLocations > 0) ? r.startLocations[0] : -1;
    res->end_loc = (r.endLocations && r.numLocations > 0) ? r.endLocations[0] : -1;
    res->aln_len = 0;
    if (want_path && r.alignment) { res->aln_len = r.alignmentLength; if (aln) memcpy(aln, r.alignment, r.alignmentLength); }
    edlibFreeAlignResult(r);
    return 0;
}


// Batch driver for the cpu_baseline / --impl reference legs of bench.py: the reference's own
// parallelism is "one problem per kt_for worker" (src/kthread.c), restated here with std::thread
// pulling problems from an atomic counter.  Same buffer layout as lcd_wfa_batch (include/lcd_gpu.h).
int ref_wfa_batch(int n, const uint8_t *seqs, const int64_t *pat_off, const int32_t *plen,
                  const int64_t *txt_off, const int32_t *tlen, const lcd_wfa_params_t *params,
                  char *ops, const int64_t *ops_off, lcd_wfa_result_t *results, int n_threads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            ref_wfa_align(seqs + pat_off[i], plen[i], seqs + txt_off[i], tlen[i], &params[i],
                          ops ? ops + ops_off[i] : NULL, &results[i]);
        }
    };
    if (n_threads <= 1) { work(); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}

}
