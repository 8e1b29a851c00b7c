// sdust_device.cuh -- K0: the low-complexity intervals of a chunk's reference window (chunk->low_comp_cr): per-position decisions, independent segments, one replay thread each.
//
// What it replaces: sdust(0, seq, l_seq, T, W, &n) (src/sdust.c:184, H. Li's symmetric DUST) as the chunk loader calls it over the chunk's region of the
// reference (src/bam_utils.c:1574-1583: T = 5, W = 20), whose intervals K2c (noisyreg_device.cuh) extends the noisy regions and the sites' spans with.
//
// The reference is one sequential pass whose state is (a) the window: the last <= W - 2 words (triplets), with their counts, the number of equal pairs rw, and
// the longest suffix L whose words all occur at most 2T / 10 times (with its own counts and pair number rv), and (b) the list P of "perfect intervals" of the
// window, which are written out (merged with the previous output interval when they touch it) once the window has moved past their start.  (a) is a function
// of the last <= W - 2 words alone -- words, not bases: the reference does NOT clear the window at an N, only its run length, so the window reaches back over
// N runs --; (b) is empty whenever no interval has been inserted for W + 20 positions (an interval's start lies at most 17 past the window's start, which
// passes it at most W + 17 positions after its insertion; an N flushes the list).  So:
//   1. every position decides by itself, from the last <= W - 2 words, whether the reference would look for perfect intervals there (rw * 10 > L * T);
//   2. the positions where it would, with less than W + 20 quiet positions between them, form independent SEGMENTS; one thread replays the reference's
//      pass over one segment, starting from the reconstructed window and an empty P, into a staging range of its own; the intervals are then packed
//      (the segments' outputs cannot touch each other: the next segment's first interval starts past anything this one can end with).
#pragma once
#include <stdint.h>

namespace lcd {
namespace sdust {

constexpr int WLEN = 3, WTOT = 64, WMSK = 63, MAXW = 32;      // MAXW: capacity of the window ring (W - 2 <= MAXW: W <= MAX_W)
constexpr int MAX_W = 24;                                      // the reference runs W = 20 (src/call_var_main.h:83); the per-thread state is sized for windows up to this
constexpr int MAXP = 768;                                      // perfect intervals alive at once (~300 seen at W = 20, ~400 at W = 24; overflow is reported, not ignored)
enum { ST_OK = 0, ST_CAP = -5, ST_PLIST = -6 };

struct Chunk {
    const char *seq; int n;              // the region of the reference (ASCII) and its length
    int T, W;
    long long base;                      // added to the 0-based interval starts / ends on output (the loader adds reg_beg - 1)
    // outputs
    long long *out_beg, *out_end; long long cap; long long *n_out; int *status;
    // scratch
    int *prevvalid;                      // [n] last position <= i at which a word (three A/C/G/T bases in a row) ends
    unsigned char *trig;                 // [n] the reference calls find_perfect at this position
    int *seg_start, *seg_cnt, *seg_off;  // [seg_cap] segments: first position, number of output intervals, offset of the first one
    int seg_cap; int *ctr;               // ctr[0]: number of segments, ctr[1]: the next one to replay
    long long *stage_beg, *stage_end; long long stage_cap;      // the segments' intervals before they are packed: segment s writes from stage_at(c, s) on
};

__device__ __forceinline__ int nt4(unsigned char c) {
    if (c < 4) return c;
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}
__device__ __forceinline__ int word_at(const Chunk &c, int i) { return (nt4((unsigned char)c.seq[i - 2]) << 4 | nt4((unsigned char)c.seq[i - 1]) << 2 | nt4((unsigned char)c.seq[i])) & WMSK; }
__device__ __forceinline__ bool valid_at(const Chunk &c, int i) { return i >= 2 && c.prevvalid[i] == i; }
// Where segment s stages its intervals.  They are disjoint, do not touch and hold a word each, so their starts are >= 4 apart, and they lie between W before
// the segment's first position and the next segment's: at most (seg_start[s + 1] - seg_start[s] + W) / 4 + 1 of them, which the spacing below leaves room for.
__device__ __forceinline__ long long stage_at(const Chunk &c, int s) { return (long long)(c.seg_start[s] / 4) + (long long)s * (c.W / 4 + 3); }
// run length of A/C/G/T bases ending at i, as far as it matters (capped at lim)
__device__ __forceinline__ int run_len(const Chunk &c, int i, int lim) { int l = 0; while (l < lim && i - l >= 0 && nt4((unsigned char)c.seq[i - l]) < 4) ++l; return l; }

// the window after position i's word has been pushed: its words, oldest first; returns their number
__device__ __forceinline__ int window_at(const Chunk &c, int i, int *w) {
    const int maxn = c.W - WLEN + 1;
    int tmp[MAXW], k = 0;
    for (int j = i; j >= 2 && k < maxn;) {
        if (c.prevvalid[j] != j) { j = c.prevvalid[j]; if (j < 2) break; continue; }
        tmp[k++] = word_at(c, j); --j;
    }
    for (int x = 0; x < k; ++x) w[x] = tmp[k - 1 - x];
    return k;
}
// (rw, L, rv) of a window: pairs of equal words, the longest suffix whose words all occur at most 2T / 10 times, pairs within it
__device__ __forceinline__ void window_stats(const int *w, int cnt, int T, int &rw, int &L, int &rv) {
    rw = 0;
    for (int a = 0; a < cnt; ++a) for (int b = a + 1; b < cnt; ++b) rw += w[a] == w[b];
    L = 0; rv = 0;
    for (int s = cnt - 1; s >= 0; --s) {
        int occ = 0;
        for (int b = s + 1; b < cnt; ++b) occ += w[b] == w[s];
        if ((occ + 1) * 10 > (T << 1)) break;
        rv += occ; ++L;
    }
}

// One segment: the reference's pass from its first position, starting with the window as it is before that position's word is pushed and an empty list of
// perfect intervals, until W + 20 positions in a row have not looked for perfect intervals (or the sequence ends); its intervals go to its staging range,
// their number (or a negative status) to seg_cnt[].
//
// The state a step touches all the time -- the window's words and the two count tables -- is 160 bytes of bytes (a count is at most W - 2), which the kernel
// keeps in shared memory: as int arrays in the thread's stack frame it was 1.2 TB/s of DRAM traffic (ncu: L1 / L2 hit rates of 30 % / 26 %) with every SM
// full of threads, each on its own path.
struct Hot { unsigned char ring[MAXW], cw[WTOT], cv[WTOT], pad[4]; };       // (164 bytes = 41 words: neighbouring threads' copies start in different banks)
// fetch() hands the thread its next segment (or -1): the GPU's threads take them from a counter of the chunk, so that a thread whose segment was short goes
// on with another one, and the loop is one body per position -- a new segment is set up inside it -- so that the threads of a warp meet again at every position.
template <class Fetch> __device__ inline void replay_segments(const Chunk &c, Hot &hot, Fetch fetch) {
    const int T = c.T, W = c.W, maxn = W - WLEN + 1, QUIET = W + 20, ns = c.ctr[0];
    unsigned char *ring = hot.ring, *cv = hot.cv, *cw = hot.cw;
    // a perfect interval in 8 bytes, so that the list's shifts move one word pair per entry: finish - start <= W, r <= (W - 2)(W - 3) / 2, l <= W - 2
    struct __align__(8) Perf { int start; unsigned m; __device__ int finish() const { return start + (int)(m >> 24); } __device__ int r() const { return (int)(m & 0xffff); } __device__ int l() const { return (int)(m >> 16 & 0xff); } } P[MAXP]; int np = 0;
    int front = 0, cnt = 0, rv = 0, rw = 0, L = 0, l = 0, quiet = 0, err = 0, i = 0;
    long long n_res = 0, last_beg = 0, last_end = 0, lim = 0; bool have_last = false, over = false;
    long long *out_beg = nullptr, *out_end = nullptr;
    auto save = [&](int start) {                                 // save_masked_regions (src/sdust.c:87-103)
        if (np == 0 || P[np - 1].start >= start) return;
        const Perf p = P[np - 1];
        bool saved = false;
        if (have_last && p.start <= last_end) { saved = true; if (p.finish() > last_end) { last_end = p.finish(); if (!over) out_end[n_res - 1] = c.base + last_end; } }
        if (!saved) { last_beg = p.start; last_end = p.finish(); have_last = true; if (n_res < lim) { out_beg[n_res] = c.base + last_beg; out_end[n_res] = c.base + last_end; } else over = true; ++n_res; }
        int x = np - 1; while (x >= 0 && P[x].start < start) --x;
        np = x + 1;
    };
    int s = fetch();
    bool fresh = true;
    while (s >= 0) {
        if (fresh) {       // the window before the segment's first position: the words that end at the last valid positions before it; no perfect intervals
            fresh = false;
            const int a = c.seg_start[s];
            const long long at = stage_at(c, s);
            out_beg = c.stage_beg + at; out_end = c.stage_end + at; lim = (s + 1 < ns ? stage_at(c, s + 1) : c.stage_cap) - at;
            for (int x = 0; x < WTOT; ++x) { cv[x] = 0; cw[x] = 0; }
            int w0[MAXW];
            front = 0; cnt = a >= 1 ? window_at(c, a - 1, w0) : 0;
            for (int x = 0; x < cnt; ++x) { ring[x] = (unsigned char)w0[x]; cw[w0[x]]++; }
            window_stats(w0, cnt, T, rw, L, rv);
            for (int x = cnt - L; x < cnt; ++x) cv[w0[x]]++;
            l = a >= 1 ? run_len(c, a - 1, W + 2) : 0;           // (only min(l, W) and l >= WLEN matter below)
            np = 0; n_res = 0; have_last = false; over = false; quiet = 0; err = 0; i = a;
        }
        const int b = i < c.n ? nt4((unsigned char)c.seq[i]) : 4;
        if (b < 4) {
            if (l < W + 2) ++l;
            if (l >= WLEN) {
                const int t = word_at(c, i);
                const int start = (l - W > 0 ? l - W : 0) + (i + 1 - l);       // l is exact while l <= W; beyond, i + 1 - W either way
                const int start_ = l > W ? i + 1 - W : start;
                save(start_);
                if (cnt >= maxn) {                                // shift_window (src/sdust.c:66-85)
                    const int o = ring[front]; front = (front + 1) % MAXW; --cnt;
                    rw -= --cw[o];
                    if (L > cnt) { --L; rv -= --cv[o]; }
                }
                ring[(front + cnt) % MAXW] = t; ++cnt;
                ++L;
                rw += cw[t]++;
                rv += cv[t]++;
                if (cv[t] * 10 > (T << 1)) {
                    int o;
                    do { o = ring[(front + cnt - L) % MAXW]; rv -= --cv[o]; --L; } while (o != t);
                }
                if (rw * 10 > L * T) {                            // find_perfect (src/sdust.c:105-131)
                    quiet = 0;
                    // (the reference counts on in a copy of cv; here cv itself, and the loop after this one takes the words out again)
                    // The reference scans P from its head for every k: the intervals that start at or after k + start_, a prefix that only grows as k
                    // falls, for the best r / l among them -- a running maximum (an interval is inserted only when it is at least as good), so the scan
                    // goes on from where the previous k's stopped: O(np + W) per position instead of O(np * W), the same answers.
                    // The intervals found at this position (at most W - 2, by falling start) are kept aside with the place each belongs at and merged into P
                    // in one pass from its tail afterwards -- the scan never needs them (each is the running maximum itself when it is found) --: one
                    // move per entry behind the first place instead of one per entry and interval.
                    int r = rv, max_r = 0, max_l = 0, j = 0, m = 0;
                    Perf ins[MAXW]; int ins_at[MAXW];
                    for (int k = cnt - L - 1; k >= 0; --k) {
                        const int tt = ring[(front + k) % MAXW];
                        r += cv[tt]++;
                        const int new_r = r, new_l = cnt - k - 1;
                        if (new_r * 10 > T * new_l) {
                            for (; j < np && P[j].start >= k + start_; ++j)
                                { const int pr = P[j].r(), pl = P[j].l(); if (max_r == 0 || pr * max_l > max_r * pl) { max_r = pr; max_l = pl; } }
                            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                                max_r = new_r; max_l = new_l;
                                ins[m].start = k + start_; ins[m].m = (unsigned)(cnt + (WLEN - 1) - k) << 24 | (unsigned)new_l << 16 | (unsigned)new_r;
                                ins_at[m++] = j;
                            }
                        }
                    }
                    if (np + m > MAXP) err = ST_PLIST;                      // (reported with the segment's count; no early exit)
                    else if (m) {
                        int i = np - 1, w = np + m - 1;
                        for (int t = m - 1; t >= 0; --t) {
                            for (; i >= ins_at[t]; --i) P[w--] = P[i];
                            P[w--] = ins[t];
                        }
                        np += m;
                    }
                    for (int k = cnt - L - 1; k >= 0; --k) cv[ring[(front + k) % MAXW]]--;
                } else ++quiet;
            } else ++quiet;
        } else {
            int start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            if (l > W) start = i + 2 - W;
            while (np) { save(start); ++start; }
            l = 0; ++quiet;
        }
        ++i;
        if (quiet >= QUIET || i > c.n) {                          // the segment is over
            if (quiet >= QUIET && np) err = ST_PLIST;             // (the bound the segmentation rests on: checked, not assumed)
            c.seg_cnt[s] = err ? err : over ? ST_CAP : (int)n_res;
            s = fetch(); fresh = true;
        }
    }
}

// prevvalid[]: every thread a contiguous block; a block's carry-in is the last valid position before it
template <class SyncF> __device__ void scan_valid(const Chunk &c, int tid, int nt, SyncF SYNC) {
    const int n = c.n;
    if (tid == 0) { c.ctr[0] = 0; c.ctr[1] = 0; *c.status = ST_OK; *c.n_out = 0; }
    const int B = (n + nt - 1) / nt, lo = tid * B, hi = lo + B < n ? lo + B : n;
    int *carry = c.seg_cnt;                                      // (free until the segments are known; needs nt <= seg_cap)
    {
        int last = -1, l = lo > 0 ? run_len(c, lo - 1, 2) : 0;
        for (int i = lo; i < hi; ++i) { if (nt4((unsigned char)c.seq[i]) < 4) { if (l < WLEN) ++l; } else l = 0; if (l >= WLEN) last = i; }
        carry[tid] = last;
    }
    SYNC();
    if (tid == 0) { int m = -1; for (int t = 0; t < nt; ++t) { const int v = carry[t]; carry[t] = m; if (v > m) m = v; } }
    SYNC();
    {
        int last = carry[tid], l = lo > 0 ? run_len(c, lo - 1, 2) : 0;
        for (int i = lo; i < hi; ++i) { if (nt4((unsigned char)c.seq[i]) < 4) { if (l < WLEN) ++l; } else l = 0; if (l >= WLEN) last = i; c.prevvalid[i] = last; }
    }
    SYNC();
}
// 1. does the reference look for perfect intervals at position i?  For a run of positions [i0, i1): the window before i0 reconstructed once (O(W^2)), then
// the reference's own window update per position (a position by itself cost ~3 500 instructions: 8 ms of a 28 ms chain for 50 Mb).
constexpr int DECIDE_RUN = 64;
__device__ inline void decide_run(const Chunk &c, Hot &hot, int i0, int i1) {
    const int T = c.T, W = c.W, maxn = W - WLEN + 1;
    unsigned char *ring = hot.ring, *cv = hot.cv, *cw = hot.cw;
    for (int x = 0; x < WTOT; ++x) { cv[x] = 0; cw[x] = 0; }
    int w0[MAXW], rw, L, rv, front = 0;
    int cnt = i0 >= 1 ? window_at(c, i0 - 1, w0) : 0;
    for (int x = 0; x < cnt; ++x) { ring[x] = (unsigned char)w0[x]; cw[w0[x]]++; }
    window_stats(w0, cnt, T, rw, L, rv);
    for (int x = cnt - L; x < cnt; ++x) cv[w0[x]]++;
    int l = i0 >= 1 ? run_len(c, i0 - 1, WLEN) : 0;               // (only l >= WLEN matters here)
    for (int i = i0; i < i1; ++i) {
        unsigned char tr = 0;
        if (nt4((unsigned char)c.seq[i]) < 4) {
            if (l < WLEN) ++l;
            if (l >= WLEN) {
                const int t = word_at(c, i);
                if (cnt >= maxn) {                                // shift_window (src/sdust.c:66-85)
                    const int o = ring[front]; front = (front + 1) % MAXW; --cnt;
                    rw -= --cw[o];
                    if (L > cnt) { --L; rv -= --cv[o]; }
                }
                ring[(front + cnt) % MAXW] = t; ++cnt;
                ++L;
                rw += cw[t]++;
                rv += cv[t]++;
                if (cv[t] * 10 > (T << 1)) {
                    int o;
                    do { o = ring[(front + cnt - L) % MAXW]; rv -= --cv[o]; --L; } while (o != t);
                }
                tr = rw * 10 > L * T;
            }
        } else l = 0;
        c.trig[i] = tr;
    }
}
// 2. segments: a position that looks, after W + 20 that did not; their first positions in position order in seg_start[0 .. ctr[0])
template <class SyncF> __device__ void collect_segments(const Chunk &c, int tid, int nt, SyncF SYNC) {
    const int n = c.n, QUIET = c.W + 20;
    for (int i = tid; i < n; i += nt) {
        if (!c.trig[i]) continue;
        bool first = true;
        for (int j = i - 1; j >= 0 && j > i - 1 - QUIET && first; --j) if (c.trig[j]) first = false;
        if (first) { const int at = atomicAdd(&c.ctr[0], 1); if (at < c.seg_cap) c.seg_start[at] = i; }
    }
    SYNC();
    const int ns = c.ctr[0];
    if (ns > c.seg_cap) { if (tid == 0) *c.status = ST_CAP; return; }
    // position order (the appends raced): rank sort through seg_off
    for (int s = tid; s < ns; s += nt) { int r = 0; for (int t = 0; t < ns; ++t) r += c.seg_start[t] < c.seg_start[s]; c.seg_off[r] = c.seg_start[s]; }
    SYNC();
    for (int s = tid; s < ns; s += nt) c.seg_start[s] = c.seg_off[s];
    SYNC();
}
template <class SyncF> __device__ void find_segments(const Chunk &c, int tid, int nt, SyncF SYNC) {
    scan_valid(c, tid, nt, SYNC);
    { Hot hot; for (int i0 = tid * DECIDE_RUN; i0 < c.n; i0 += nt * DECIDE_RUN) decide_run(c, hot, i0, i0 + DECIDE_RUN < c.n ? i0 + DECIDE_RUN : c.n); }
    SYNC();
    collect_segments(c, tid, nt, SYNC);
}

// after the count pass: offsets of the segments' outputs, their total, the status (one thread per chunk)
__device__ inline void finish_counts(const Chunk &c) {
    if (*c.status != ST_OK) return;
    const int ns = c.ctr[0];
    long long tot = 0; int bad = 0;
    for (int s = 0; s < ns; ++s) { if (c.seg_cnt[s] < 0) { bad = c.seg_cnt[s]; break; } c.seg_off[s] = (int)tot; tot += c.seg_cnt[s]; }
    if (bad) *c.status = bad; else if (tot > c.cap) *c.status = ST_CAP;
    *c.n_out = tot;
}

// the whole chunk on one group of threads (host emulation; the GPU runs the segments of all chunks side by side: sdust_kernel.cu)
// pack_segment moves a segment's staged intervals to their place in the chunk's output
__device__ inline void pack_segment(const Chunk &c, int s) {
    const long long at = stage_at(c, s); const int o = c.seg_off[s];
    for (int x = 0; x < c.seg_cnt[s]; ++x) { c.out_beg[o + x] = c.stage_beg[at + x]; c.out_end[o + x] = c.stage_end[at + x]; }
}

// the whole chunk on one group of threads (host emulation; the GPU runs the segments of all chunks side by side: sdust_kernel.cu)
template <class SyncF> __device__ void run_chunk(Chunk c, int tid, int nt, SyncF SYNC) {
    find_segments(c, tid, nt, SYNC);
    SYNC();
    if (*c.status != ST_OK) return;
    const int ns = c.ctr[0];
    { Hot hot; int nxt = tid; replay_segments(c, hot, [&]() { const int s = nxt; nxt += nt; return s < ns ? s : -1; }); }
    SYNC();
    if (tid == 0) finish_counts(c);
    SYNC();
    if (*c.status != ST_OK) return;
    for (int s = tid; s < ns; s += nt) pack_segment(c, s);
}

} // namespace sdust
} // namespace lcd
