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
#include <chrono>
#include <atomic>
#include <vector>
extern "C" {
#include "wavefront/wavefront_align.h"
}
#include "edlib.h"
extern "C" {
#include "abpoa.h"
}

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


// Mirrors abpoa_partial_aln_msa_cons (src/align.c:762-870) for full-cover reads when sub_aln == 1
// and abpoa_aln_msa_cons (src/align.c:872-953) when sub_aln == 0, with max_n_cons == 1.
int ref_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
            const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
            uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    abpoa_t *ab = abpoa_init();
    abpoa_para_t *abpt = abpoa_init_para();
    abpt->cons_algrm = ABPOA_MF;
    abpt->inc_path_score = 1;
    abpt->out_cons = 1; abpt->out_msa = 1;
    abpt->max_n_cons = p->max_n_cons; abpt->min_freq = 0.20;
    abpt->match = p->match; abpt->mismatch = p->mismatch;
    abpt->gap_open1 = p->gap_open1; abpt->gap_ext1 = p->gap_ext1;
    abpt->gap_open2 = p->gap_open2; abpt->gap_ext2 = p->gap_ext2;
    abpt->wb = p->wb; abpt->wf = p->wf;
    *cons_len = 0; *msa_len = 0;
    if (p->sub_aln) {
        abpt->sub_aln = 1;
        abpoa_post_set_para(abpt);
        ab->abs->n_seq = n_seq;
        for (int i = 0; i < n_seq; ++i) {
            abpoa_res_t res; res.graph_cigar = 0; res.n_cigar = 0;
            uint8_t *q = (uint8_t*)seqs + seq_off[i];
            abpoa_align_sequence_to_subgraph(ab, abpt, 0, 1, q, seq_len[i], &res);
            abpoa_add_subgraph_alignment(ab, abpt, 0, 1, q, NULL, seq_len[i], NULL, res, i, n_seq, 0);
            if (res.n_cigar) free(res.graph_cigar);
        }
        abpoa_output(ab, abpt, NULL);
    } else {
        abpoa_post_set_para(abpt);
        std::vector<uint8_t*> ptr(n_seq); std::vector<int> len(n_seq);
        for (int i = 0; i < n_seq; ++i) { ptr[i] = (uint8_t*)seqs + seq_off[i]; len[i] = seq_len[i]; }
        abpoa_msa(ab, abpt, n_seq, NULL, len.data(), ptr.data(), NULL, NULL);
    }
    abpoa_cons_t *abc = ab->abc;
    int rc = 0;
    if (abc->n_cons > 0) { *cons_len = abc->cons_len[0]; memcpy(cons, abc->cons_base[0], abc->cons_len[0]); }
    if (msa) {
        if ((int64_t)(abc->n_seq + abc->n_cons) * abc->msa_len > msa_cap || abc->n_cons != 1) rc = -5;
        else { *msa_len = abc->msa_len; for (int i = 0; i < abc->n_seq + 1; ++i) memcpy(msa + (size_t)i * abc->msa_len, abc->msa_base[i], abc->msa_len); }
    }
    abpoa_free_para(abpt); abpoa_free(ab);
    return rc;
}


// abpoa_partial_aln_msa_cons (src/align.c:790-812) with partially covering reads: read i > 0 goes against the sub-graph abpoa_subgraph_nodes
// finds between the nodes sub_beg[i] and sub_end[i] (both 0: the whole graph), or is left out (sub_beg[i] < 0).  The unmodified abPOA does the work.
int ref_poa_sub(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    abpoa_t *ab = abpoa_init();
    abpoa_para_t *abpt = abpoa_init_para();
    abpt->cons_algrm = ABPOA_MF; abpt->sub_aln = 1; abpt->inc_path_score = 1;
    abpt->out_cons = 1; abpt->out_msa = 1;
    abpt->max_n_cons = p->max_n_cons; abpt->min_freq = 0.20;
    abpt->match = p->match; abpt->mismatch = p->mismatch;
    abpt->gap_open1 = p->gap_open1; abpt->gap_ext1 = p->gap_ext1; abpt->gap_open2 = p->gap_open2; abpt->gap_ext2 = p->gap_ext2;
    abpt->wb = p->wb; abpt->wf = p->wf;
    *cons_len = 0; *msa_len = 0;
    abpoa_post_set_para(abpt);
    ab->abs->n_seq = n_seq;
    for (int i = 0; i < n_seq; ++i) {
        abpoa_res_t res; res.graph_cigar = 0; res.n_cigar = 0;
        int exc_beg = 0, exc_end = 1;
        if (i != 0 && sub_beg[i] < 0) continue;
        if (i != 0 && sub_beg[i] > 0) abpoa_subgraph_nodes(ab, abpt, sub_beg[i], sub_end[i], &exc_beg, &exc_end);
        uint8_t *q = (uint8_t*)seqs + seq_off[i];
        abpoa_align_sequence_to_subgraph(ab, abpt, exc_beg, exc_end, q, seq_len[i], &res);
        abpoa_add_subgraph_alignment(ab, abpt, exc_beg, exc_end, q, NULL, seq_len[i], NULL, res, i, n_seq, 0);
        if (res.n_cigar) free(res.graph_cigar);
    }
    abpoa_output(ab, abpt, NULL);
    abpoa_cons_t *abc = ab->abc;
    int rc = 0;
    if (abc->n_cons > 0) { *cons_len = abc->cons_len[0]; memcpy(cons, abc->cons_base[0], abc->cons_len[0]); }
    if (msa) {
        if ((int64_t)(abc->n_seq + abc->n_cons) * abc->msa_len > msa_cap || abc->n_cons != 1) rc = -5;
        else { *msa_len = abc->msa_len; for (int i = 0; i < abc->n_seq + 1; ++i) memcpy(msa + (size_t)i * abc->msa_len, abc->msa_base[i], abc->msa_len); }
    }
    abpoa_free_para(abpt); abpoa_free(ab);
    return rc;
}


// Batch driver for bench.py's cpu_baseline / --impl reference legs (one problem per worker thread, like kt_for).
int ref_poa_batch(int n, const uint8_t *seqs, const int32_t *first_read, const int32_t *n_reads,
                  const int64_t *read_off, const int32_t *read_len, const lcd_poa_params_t *params,
                  uint8_t *cons, const int64_t *cons_off, int32_t *cons_len, int n_threads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            int32_t ml = 0;
            ref_poa(n_reads[i], seqs, read_off + first_read[i], read_len + first_read[i], &params[i],
                    cons + cons_off[i], &cons_len[i], NULL, &ml, 0);
        }
    };
    if (n_threads <= 1) { work(); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}


// Batch drivers for the K7 / K4 legs of bench.py's reference arm (one problem / one chunk per worker thread, like kt_for).
int ref_edlib_batch(int n, const uint8_t *seqs, const int64_t *q_off, const int32_t *qlen, const int64_t *t_off, const int32_t *tlen,
                    const int32_t *mode, const int32_t *want_path, uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results, int n_threads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            ref_edlib_align(seqs + q_off[i], qlen[i], seqs + t_off[i], tlen[i], mode[i], want_path[i], aln ? aln + aln_off[i] : NULL, &results[i]);
        }
    };
    if (n_threads <= 1) { work(); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}

int ref_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out);      // ref_shim_lcd.c
int ref_phase_batch(int n, const lcd_phase_input_t *in, lcd_phase_output_t *out, int n_threads) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n) break;
            ref_assign_hap(&in[i], &out[i]);
        }
    };
    if (n_threads <= 1) { work(); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}

// K1 / K2 / K3 over many chunks on n_threads host threads (the reference's kt_for runs one chunk per thread as well)
int ref_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out);
int ref_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out);
}
template <typename F> static int run_chunks(int n, int n_threads, F f) {
    std::atomic<int> next(0), bad(0);
    auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; if (f(i)) bad = 1; } };
    if (n_threads <= 1) { work(); return bad; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return bad;
}
extern "C" {
// bam1_t records are built first and results copied out last; *core_wall_s is the wall time of the reference's own calls
void *ref_digar_prepare(const lcd_digar_input_t *in);
void ref_digar_core(void *job, const lcd_digar_input_t *in);
int ref_digar_finish(void *job, const lcd_digar_input_t *in, lcd_digar_output_t *out);
int ref_digar_batch(int n, const lcd_digar_input_t *in, lcd_digar_output_t *out, int n_threads, double *core_wall_s) {
    std::vector<void *> jobs(n, nullptr);
    if (run_chunks(n, n_threads, [&](int i) { jobs[i] = ref_digar_prepare(&in[i]); return jobs[i] ? 0 : 1; })) return -9;
    const auto t0 = std::chrono::steady_clock::now();
    run_chunks(n, n_threads, [&](int i) { ref_digar_core(jobs[i], &in[i]); return 0; });
    if (core_wall_s) *core_wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return run_chunks(n, n_threads, [&](int i) { return ref_digar_finish(jobs[i], &in[i], &out[i]); });
}
int ref_pileup_batch(int n, const lcd_pileup_input_t *in, lcd_pileup_output_t *out, int n_threads) {
    return run_chunks(n, n_threads, [&](int i) { return ref_collect_cand_vars(&in[i], &out[i]); });
}
// candidate-site list: records built first, results copied out last; *core_wall_s is the wall time of collect_all_cand_var_sites alone
void *ref_sites_prepare(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end);
void ref_sites_core(void *job);
int ref_sites_finish(void *job, const lcd_pileup_input_t *in, lcd_sites_output_t *out);
int ref_sites_batch(int n, const lcd_pileup_input_t *in, const int64_t *reg, lcd_sites_output_t *out, int n_threads, double *core_wall_s) {
    std::vector<void *> jobs(n, nullptr);
    run_chunks(n, n_threads, [&](int i) { jobs[i] = ref_sites_prepare(&in[i], reg[2 * i], reg[2 * i + 1]); return 0; });
    const auto t0 = std::chrono::steady_clock::now();
    run_chunks(n, n_threads, [&](int i) { ref_sites_core(jobs[i]); return 0; });
    if (core_wall_s) *core_wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return run_chunks(n, n_threads, [&](int i) { return ref_sites_finish(jobs[i], &in[i], &out[i]); });
}
// per-site categories of many chunks (classify_var_cate per site, as the first loop of classify_cand_vars)
int ref_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate);
int ref_classify_batch(int n, const lcd_classify_input_t *in, int32_t **var_cate, int n_threads) {
    return run_chunks(n, n_threads, [&](int i) { return ref_classify_sites(&in[i], var_cate[i]); });
}
// pre_process_noisy_regs + classify_cand_vars of many chunks (the reference runs its own classify_var_cate loop inside)
int ref_noisy_regs(const lcd_classify_input_t *ci, const lcd_noisyreg_input_t *in, int64_t *kept_pos, int32_t *kept_type, int32_t *kept_ref_len, int32_t *kept_cate, int32_t *n_kept,
                   lcd_noisyreg_output_t *out);
int ref_noisyreg_batch(int n, const lcd_classify_input_t *ci, const lcd_noisyreg_input_t *in, int64_t **kept_pos, int32_t **kept_type, int32_t **kept_ref_len, int32_t **kept_cate,
                       int32_t *n_kept, lcd_noisyreg_output_t *out, int n_threads) {
    return run_chunks(n, n_threads, [&](int i) { return ref_noisy_regs(&ci[i], &in[i], kept_pos[i], kept_type[i], kept_ref_len[i], kept_cate[i], &n_kept[i], &out[i]); });
}
int ref_profile_batch(int n, const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out, int n_threads) {
    return run_chunks(n, n_threads, [&](int i) { return ref_read_var_profile(&in[i], &ex[i], &out[i]); });
}

}
