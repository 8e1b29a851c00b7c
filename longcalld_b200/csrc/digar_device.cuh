// digar_device.cuh -- device-side logic of K1: the first step of the pileup scan for reads with an =/X CIGAR, replacing
// collect_digar_from_eqx_cigar (reference src/bam_utils.c:701-841) with its sliding-window detector of dense difference
// regions push_xid_size_queue_win (src/bam_utils.c:161-205), as collect_digars_from_bam drives it over a chunk's reads
// (src/collect_var.c:1063-1082), and the per-chunk base-quality histogram of longcalld_copy_digar_read_buffers
// (src/bam_utils.c:90-103).
//
// B200 design.  Input is what the BAM records hold, flat and concatenated over all chunks of a batch: CIGAR words (u32),
// 4-bit packed SEQ, QUAL.  Three passes:
//   count_read : one thread per read walks its CIGAR words and sizes its outputs (digar1_t records, alt bases, interval
//                capacity); an exclusive scan turns the sizes into the read's slices of the flat output arrays.
//   fill_read  : one thread per read walks the CIGAR again and writes its difference list in the SoA layout K2 / K3 consume
//                (pos, type, len, qi, low_qual, alt offset + alt bases).  The reference's sliding-window queue is not
//                materialised: its entries are exactly the read's non-low-quality X / I / D records already written, so the
//                queue front is a second cursor over the thread's own output and the "sum of counts between two queue
//                slots" is the difference of a running total.  Noisy intervals, the noisy-length / event-ratio skip test
//                (two double comparisons, as the reference) and the read span come out of the same walk.
//   hist_reads : the only per-base work -- a warp streams a read's QUAL bytes with 16-byte loads, run-length aggregates in
//                registers, adds into a per-warp shared-memory histogram and flushes to the chunk's 256 counters when its
//                reads move on to another chunk.  This is the HBM-bound part (1 B per read base).
// The file compiles for the host as well (tests/emu).
#pragma once
#include <stdint.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace digar {

enum { CMATCH = 0, CINS = 1, CDEL = 2, CREF_SKIP = 3, CSOFT = 4, CHARD = 5, CPAD = 6, CEQUAL = 7, CDIFF = 8,
       // pseudo-ops the tag front end (md_device.cuh) emits for the cs-tag and the no-tag variants of the reference's pass
       CSKIP = 9,               // bases outside the chunk's reference window: position and query index advance, no record (src/bam_utils.c:1209-1215)
       CCS_SOFT = 10, CCS_HARD = 11 };   // a clip of a cs-tagged read: a long one counts as a candidate wherever it lies (src/bam_utils.c:876-889, 953-966)
enum { ST_BAD_OP = 1, ST_OVERFLOW = 2 };

struct __align__(8) Chunk {
    int32_t min_bq, max_xgaps, win, end_clip_reg, flank_win, pad;
    double max_noisy_frac, max_var_ratio;
    long long whole_ref_len;
    long long read0;              // index of the chunk's first read in the concatenated per-read arrays
};

struct KernelArgs {
    const Chunk *chunks; long long n_reads_total;
    // per read (concatenated over chunks)
    const int32_t *read_chunk; const uint8_t *read_active;      // active = listed in ordered_read_ids and not skipped by the loader
    const long long *read_pos0; const uint8_t *read_is_rev, *is_palindrome;
    const int32_t *n_cigar; const long long *cigar_off; const uint32_t *cigar;
    const long long *rlen;                                       // nullptr, or the reference length of every read's own CIGAR (tag front end: the op stream may differ)
    const int32_t *l_qseq; const long long *seq_off; const uint8_t *bseq; const long long *qual_off; const uint8_t *qual;
    // sizes (count pass) and their exclusive scans
    long long *cnt;                                              // [3][n_reads_total + 1]: records, alt bases, interval capacity
    const long long *first;                                      // [3][n_reads_total + 1]: exclusive scans of cnt
    long long stride;                                            // n_reads_total + 1
    // outputs
    uint8_t *skip; long long *read_beg, *read_end; int32_t *n_digar;
    long long *digar_pos; int8_t *digar_type; int32_t *digar_len, *digar_qi; uint8_t *digar_low_qual; long long *digar_alt_off; uint8_t *digar_alt;
    int32_t *n_nreg; long long *nreg_beg, *nreg_end; int32_t *nreg_label;
    unsigned long long *qual_counts;                             // [n_chunks][256]
    int32_t *status;
};

__device__ __forceinline__ int base_code(const uint8_t *bseq, long long qi) {       // seq_nt16_int[bam_seqi(bseq, qi)]
    const int c = (bseq[qi >> 1] >> ((~qi & 1) << 2)) & 15;
    return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
}

// pass 1: sizes of read g's outputs from its CIGAR words alone
__device__ void count_read(const KernelArgs &a, long long g) {
    long long nd = 0, na = 0, ncap = 0;
    if (a.read_active[g]) {
        const uint32_t *cg = a.cigar + a.cigar_off[g]; const int nc = a.n_cigar[g];
        const int max_s = a.chunks[a.read_chunk[g]].max_xgaps;
        long long n_x = 0, n_gap = 0;
        for (int k = 0; k < nc; ++k) {
            const uint32_t w = cg[k]; const int op = w & 15; const long long len = w >> 4;
            if (op == CDIFF) { nd += len; na += len; n_x += len; }
            else if (op == CINS) { nd++; na += len; n_gap++; }
            else if (op == CDEL) { nd++; n_gap++; }
            else if (op == CEQUAL || op == CSOFT || op == CHARD || op == CCS_SOFT || op == CCS_HARD) nd++;
        }
        // a dense window holds an indel or more than max_s X bases of its own; plus the two clip intervals
        ncap = n_gap + n_x / (max_s > 0 ? max_s + 1 : 1) + 2;
    }
    a.cnt[g] = nd; a.cnt[a.stride + g] = na; a.cnt[2 * a.stride + g] = ncap; a.n_digar[g] = (int32_t)nd;
}

struct Win {                      // the reference's xid_queue_t + the pending interval (cr_cur_start / cr_cur_end / cr_q_start / cr_q_end)
    long long front;              // index of the queue-front record in the digar arrays
    long long q_count;            // q->count
    long long cum;                // counts pushed so far
    long long cur_start, cur_end; // pending dense window, -1: none
    long long cum_before, cum_end;// running totals bracketing the pending window's queue slots
};

struct Regs { const KernelArgs *a; long long first, cap; int n; long long noisy_len; };

// cr_add (src/cgranges.c:145-160): negative starts are clamped, inverted intervals dropped
__device__ __forceinline__ bool add_reg(Regs &r, long long st, long long en, int label) {
    if (st < 0) st = 0;
    if (st > en) return true;
    if (r.n >= r.cap) return false;
    const long long i = r.first + r.n++;
    r.a->nreg_beg[i] = st; r.a->nreg_end[i] = en; r.a->nreg_label[i] = label;
    r.noisy_len += en - st + 1;                                  // collect_noisy_region_len, src/bam_utils.c:624-631
    return true;
}

__device__ __forceinline__ bool flush_win(const Win &q, Regs &r) {   // src/bam_utils.c:191-196, :776-781
    long long var_size = q.cum_end - q.cum_before;
    if (var_size < q.cur_end - q.cur_start + 1) var_size = q.cur_end - q.cur_start + 1;
    return add_reg(r, q.cur_start - 1, q.cur_end, (int)var_size);
}

__device__ __forceinline__ bool is_queue_entry(const KernelArgs &a, long long d) {
    const int t = a.digar_type[d];
    return (t == CDIFF || t == CINS || t == CDEL) && !a.digar_low_qual[d];
}

// push_xid_size_queue_win (src/bam_utils.c:161-205) for the record just written at index d: (pos, len, count)
__device__ __forceinline__ bool push_win(const KernelArgs &a, Win &q, Regs &r, long long d, long long pos, int len, int count, int win, int max_s) {
    if (q.front < 0) q.front = d;
    q.cum += count; q.q_count += count;
    for (;;) {
        const int ft = a.digar_type[q.front]; const int fl = a.digar_len[q.front];
        const int f_len = ft == CDIFF ? 1 : (ft == CDEL ? fl : 0), f_cnt = ft == CDIFF ? 1 : fl;
        if (a.digar_pos[q.front] + f_len - 1 > pos - win) break;
        q.q_count -= f_cnt;
        do { ++q.front; } while (q.front < d && !is_queue_entry(a, q.front));
    }
    if (count > 0 && q.q_count > max_s) {
        const long long ns = a.digar_pos[q.front], ne = pos + len;
        if (q.cur_start == -1) { q.cur_start = ns; q.cur_end = ne; q.cum_before = q.cum - q.q_count; q.cum_end = q.cum; }
        else if (ns <= q.cur_end) { q.cur_end = ne; q.cum_end = q.cum; }
        else {
            if (!flush_win(q, r)) return false;
            q.cur_start = ns; q.cur_end = ne; q.cum_before = q.cum - q.q_count; q.cum_end = q.cum;
        }
    }
    return true;
}

// pass 2: collect_digar_from_eqx_cigar (src/bam_utils.c:701-841) for read g
__device__ void fill_read(const KernelArgs &a, long long g) {
    a.skip[g] = 0; a.n_nreg[g] = 0;
    if (!a.read_active[g]) return;
    const Chunk ch = a.chunks[a.read_chunk[g]];
    const uint32_t *cg = a.cigar + a.cigar_off[g]; const int nc = a.n_cigar[g];
    const uint8_t *bseq = a.bseq + a.seq_off[g], *qual = a.qual + a.qual_off[g];
    long long pos = a.read_pos0[g] + 1, rlen = 0, qi = 0;
    if (a.rlen) rlen = a.rlen[g];
    else for (int k = 0; k < nc; ++k) { const int op = cg[k] & 15; if (op == CMATCH || op == CDEL || op == CREF_SKIP || op == CEQUAL || op == CDIFF) rlen += cg[k] >> 4; }
    const long long beg = pos, end = a.read_pos0[g] + (rlen ? rlen : 1);              // bam_endpos
    a.read_beg[g] = beg; a.read_end[g] = end;
    long long d = a.first[g], at = a.first[a.stride + g];
    const long long d_end = a.first[g + 1], at_end = a.first[a.stride + g + 1];
    const long long alt_base = a.first[a.stride + ch.read0];                          // alt offsets are relative to the chunk's first alt base
    Regs r; r.a = &a; r.first = a.first[2 * a.stride + g]; r.cap = a.first[2 * a.stride + g + 1] - r.first; r.n = 0; r.noisy_len = 0;
    Win q; q.front = -1; q.q_count = 0; q.cum = 0; q.cur_start = q.cur_end = -1; q.cum_before = q.cum_end = 0;
    const bool left_pal = a.is_palindrome[g] && a.read_is_rev[g], right_pal = a.is_palindrome[g] && !a.read_is_rev[g];
    long long n_cand = 0; int err = 0;
#define LCD_PUT(p_, t_, l_, low_) { if (d >= d_end) { err = ST_OVERFLOW; break; } a.digar_pos[d] = (p_); a.digar_type[d] = (int8_t)(t_); a.digar_len[d] = (int)(l_); \
        a.digar_qi[d] = (int)qi; a.digar_low_qual[d] = (uint8_t)(low_); a.digar_alt_off[d] = at - alt_base; }
    for (int k = 0; k < nc && !err; ++k) {
        const int op = cg[k] & 15; const long long len = cg[k] >> 4;
        if (op == CDIFF) {
            for (long long j = 0; j < len; ++j) {
                const int low = !(qual[qi] >= ch.min_bq);
                LCD_PUT(pos, op, 1, low);
                if (at >= at_end) { err = ST_OVERFLOW; break; }
                a.digar_alt[at++] = (uint8_t)base_code(bseq, qi);
                if (!low && !push_win(a, q, r, d, pos, 1, 1, ch.win, ch.max_xgaps)) { err = ST_OVERFLOW; break; }
                ++d; ++n_cand; ++pos; ++qi;
            }
        } else if (op == CEQUAL) {
            LCD_PUT(pos, op, len, 0);
            ++d; pos += len; qi += len;
        } else if (op == CDEL) {
            const int ok = (qi == 0 || qual[qi - 1] >= ch.min_bq) && qual[qi] >= ch.min_bq;
            LCD_PUT(pos, op, len, !ok);
            if (ok && !push_win(a, q, r, d, pos, (int)len, (int)len, ch.win, ch.max_xgaps)) { err = ST_OVERFLOW; break; }
            ++d; ++n_cand; pos += len;
        } else if (op == CINS) {
            int low = 1;
            for (long long j = 0; j < len; ++j) if (qual[qi + j] >= ch.min_bq) { low = 0; break; }
            LCD_PUT(pos, op, len, low);
            if (at + len > at_end) { err = ST_OVERFLOW; break; }
            for (long long j = 0; j < len; ++j) a.digar_alt[at++] = (uint8_t)base_code(bseq, qi + j);
            if (!low && !push_win(a, q, r, d, pos, 0, (int)len, ch.win, ch.max_xgaps)) { err = ST_OVERFLOW; break; }
            ++d; ++n_cand; qi += len;
        } else if (op == CSOFT || op == CHARD) {
            const bool pal = (k == 0 && left_pal) || (k != 0 && right_pal);
            LCD_PUT(pos, pal ? CHARD : op, len, 0);
            ++d;
            if (((k == 0 && pos > 10) || (k != 0 && pos < ch.whole_ref_len - 10)) && len > ch.end_clip_reg) {
                if (k == 0 && !left_pal) { if (pos > 1 && !add_reg(r, pos - 1, pos + ch.flank_win, 0)) { err = ST_OVERFLOW; break; } ++n_cand; }
                else if (k != 0 && !right_pal) { if (pos < ch.whole_ref_len && !add_reg(r, pos - 1 - ch.flank_win, pos, 0)) { err = ST_OVERFLOW; break; } ++n_cand; }
            }
            if (op == CSOFT) qi += len;
        } else if (op == CCS_SOFT || op == CCS_HARD) {                            // cs-tag variant: first / last CIGAR op, src/bam_utils.c:876-889, 953-966
            const int cop = op == CCS_SOFT ? CSOFT : CHARD;
            const bool pal = (k == 0 && left_pal) || (k != 0 && right_pal);
            LCD_PUT(pos, pal ? CHARD : cop, len, 0);
            ++d;
            if (len > ch.end_clip_reg && !pal) {
                if (k == 0) { if (pos > 10 && !add_reg(r, pos - 1, pos + ch.flank_win, 0)) { err = ST_OVERFLOW; break; } }
                else if (pos < ch.whole_ref_len - 10 && !add_reg(r, pos - 1 - ch.flank_win, pos, 0)) { err = ST_OVERFLOW; break; }
                ++n_cand;
            }
            if (cop == CSOFT) qi += len;
        } else if (op == CREF_SKIP) pos += len;
        else if (op == CSKIP) { pos += len; qi += len; }
        else if (op == CMATCH) err = ST_BAD_OP;                                   // 'M' is not expected in an =/X CIGAR (:766-768)
    }
#undef LCD_PUT
    if (!err && q.cur_start != -1 && !flush_win(q, r)) err = ST_OVERFLOW;
    if (err) { atomicMax(a.status, err); return; }
    a.n_nreg[g] = r.n;
    const int mapped_len = (int)(end - beg + 1);
    if ((int)r.noisy_len > mapped_len * ch.max_noisy_frac || (int)n_cand > mapped_len * ch.max_var_ratio) a.skip[g] = 1;   // :787-791
    // cr_index order for up to 64 intervals = stable insertion sort by start (src/cgranges.c:13-64); larger sets are ordered by the host plan
    if (r.n > 1 && r.n <= 64) {
        for (int i = 1; i < r.n; ++i) {
            const long long b = a.nreg_beg[r.first + i], e = a.nreg_end[r.first + i]; const int l = a.nreg_label[r.first + i];
            int j = i;
            while (j > 0 && a.nreg_beg[r.first + j - 1] > b) {
                a.nreg_beg[r.first + j] = a.nreg_beg[r.first + j - 1]; a.nreg_end[r.first + j] = a.nreg_end[r.first + j - 1]; a.nreg_label[r.first + j] = a.nreg_label[r.first + j - 1];
                --j;
            }
            if (j != i) { a.nreg_beg[r.first + j] = b; a.nreg_end[r.first + j] = e; a.nreg_label[r.first + j] = l; }
        }
    }
}

// Base-quality histogram of read g (longcalld_copy_digar_read_buffers, src/bam_utils.c:90-103), for lane `lane` of `lanes`:
// 16-byte aligned loads over the read's QUAL bytes; equal neighbours are aggregated in registers (a whole word equal to the
// running value costs one compare) before the shared-memory add.  Only the first and last 16-byte chunk need byte masks.
__device__ __forceinline__ void hist_word(unsigned wd, int &cur, unsigned &run, unsigned *hist) {
    if (wd == (unsigned)cur * 0x01010101u) { run += 4; return; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = (wd >> (8 * i)) & 255;
        if (b != cur) { if (run) atomicAdd(hist + cur, run); cur = b; run = 0; }
        ++run;
    }
}
__device__ __forceinline__ void hist_read(const KernelArgs &a, long long g, int lane, int lanes, unsigned *hist) {
    const uint8_t *q = a.qual + a.qual_off[g]; const long long n = a.l_qseq[g];
    if (n <= 0) return;
    const uintptr_t p0 = (uintptr_t)q, p1 = p0 + (uintptr_t)n;
    const uintptr_t base = p0 & ~(uintptr_t)15, last = (p1 - 1) & ~(uintptr_t)15;
    int cur = 0; unsigned run = 0;
    for (uintptr_t p = base + 16 * (uintptr_t)lane; p <= last; p += 16 * (uintptr_t)lanes) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        if (p != base && p != last) {
            hist_word(v.x, cur, run, hist); hist_word(v.y, cur, run, hist); hist_word(v.z, cur, run, hist); hist_word(v.w, cur, run, hist);
        } else {
            const unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uintptr_t addr = p + i;
                if (addr < p0 || addr >= p1) continue;
                const int b = (w[i >> 2] >> ((i & 3) * 8)) & 255;
                if (b != cur) { if (run) atomicAdd(hist + cur, run); cur = b; run = 0; }
                ++run;
            }
        }
    }
    if (run) atomicAdd(hist + cur, run);
}

} // namespace digar
} // namespace lcd
