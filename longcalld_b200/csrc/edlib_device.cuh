// edlib_device.cuh -- device-side logic of K7: edlib's block bit-vector (Myers/Hyyro) global (NW) and
// infix (HW) alignment with path, as longcallD calls it (reference src/align.c:210-275: edlib_xgaps,
// edlib_end2end_aln, edlib_infix_aln, edlib_edit_distance -> edlibAlign, edlib/src/edlib.cpp:146).
//
// B200 design: ONE THREAD PER PROBLEM.  With k = the edit distance the Ukkonen band of a column is 1-3
// 64-cell blocks wide whatever the query length, and the blocks of a column form a carry chain, so
// there is no parallelism inside a problem worth a warp; the parallelism is the thousands of
// read-vs-consensus pairs of a batch.  Problems are sorted by size so that the 32 threads of a warp run
// columns of similar length; every thread owns a private slice of the workspace pool (query profile,
// column state, the per-column (P, M, score) store the traceback reads).
//
// Bit-exactness: the PATH depends on which blocks are inside the band when the traceback looks left,
// on the traceback preference (up, left, diagonal), on the Hirschberg split row and on the HW start
// location rule, so the band bookkeeping of the reference is restated literally (cited per function).
// The file compiles for the host as well (tests/emu) so the logic is diffed against the oracle on CPU.
#pragma once
#include <stdint.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace edlib {

typedef unsigned long long Word;
constexpr int WS = 64;
constexpr Word HIGH_BIT = (Word)1 << 63;
constexpr int NSYM = 6;                    // base codes 0..5 (A C G T N gap)
constexpr int STACK = 48;                  // Hirschberg recursion: the target halves at every level
constexpr long long TRACEBACK_LIMIT = 1024 * 1024;   // edlib.cpp:1195: column data below 1 MiB -> traceback

enum { OP_MATCH = 0, OP_INSERT = 1, OP_DELETE = 2, OP_MISMATCH = 3 };
enum { ST_OK = 0, ST_SYMBOL = -1, ST_NO_SPLIT = -2, ST_OOM = -3, ST_NO_SOLUTION = -4 };

struct __align__(16) Problem {
    uint64_t q_off, t_off;     // byte offsets in the packed sequence buffer
    uint64_t aln_off;          // byte offset in the path output buffer (capacity qlen + tlen + 2)
    uint64_t ws_off;           // 64-bit word offset of this problem's workspace slice
    uint64_t ws_words;
    int32_t qlen, tlen;
    int32_t mode, want_path;
};

struct __align__(16) DevResult {
    int32_t status, edit_distance, start_loc, end_loc, aln_len;
    uint32_t units_lo, units_hi;          // block x column advances (SURVEY 8d unit)
    int32_t pad;
};

struct KernelArgs {
    const Problem *problems;
    const int32_t *order;
    int32_t n;
    uint32_t *queue;
    const uint8_t *seqs;
    uint8_t *aln;
    DevResult *results;
    Word *pool;
};

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int ceil_div(int x, int y) { return (x + y - 1) / y; }

// a sequence slice read forwards or backwards (edlib's createReverseCopy without the copy)
struct Seq {
    const uint8_t *p; int n; int rev;
    __device__ __forceinline__ int at(int i) const { return rev ? p[n - 1 - i] : p[i]; }
    __device__ __forceinline__ Seq sub(int off, int len) const {           // slice [off, off+len) in THIS orientation
        Seq s; s.n = len; s.rev = rev; s.p = rev ? p + (n - off - len) : p + off; return s;
    }
    __device__ __forceinline__ Seq reversed() const { Seq s = *this; s.rev = !rev; return s; }
};

// calculateBlock, edlib.cpp:395-435
__device__ __forceinline__ int advance_block(Word Pv, Word Mv, Word Eq, int hin, Word &PvOut, Word &MvOut) {
    const Word hin_neg = hin < 0 ? 1 : 0;
    const Word Xv = Eq | Mv;
    Eq |= hin_neg;
    const Word Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    Word Ph = Mv | ~(Xh | Pv);
    Word Mh = Pv & Xh;
    const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
    Ph <<= 1; Mh <<= 1;
    Mh |= hin_neg;
    Ph |= hin > 0 ? 1 : 0;
    PvOut = Mh | ~(Xv | Ph);
    MvOut = Ph & Xv;
    return hout;
}

// value of the cell `i` positions above the bottom cell of a block (getBlockCellValues, edlib.cpp:458-470):
// the bottom cell is `score`; going up, a set P bit means the cell below was +1, a set M bit -1
__device__ __forceinline__ int cell_above(Word P, Word M, int score, int i) {
    if (i == 0) return score;
    const Word top = ~(Word)0 << (WS - i);           // the i highest bits
    return score - __popcll(P & top) + __popcll(M & top);
}

struct State {               // column state of the blocks + scratch, all in this problem's workspace slice
    Word *P, *M; int *S;
};

// buildPeq, edlib.cpp:359-384 (the padding rows of the last block match everything)
__device__ void build_peq(Word *peq, const Seq &q, int NB) {
    for (int b = 0; b < NB; ++b) {
        Word w[NSYM];
#pragma unroll
        for (int s = 0; s < NSYM; ++s) w[s] = 0;
        const int r0 = b * WS;
        for (int i = 0; i < WS; ++i) {
            const int r = r0 + i;
            if (r >= q.n) {
#pragma unroll
                for (int s = 0; s < NSYM; ++s) w[s] |= (Word)1 << i;
            } else {
                const int c = q.at(r);
#pragma unroll
                for (int s = 0; s < NSYM; ++s) if (c == s) w[s] |= (Word)1 << i;
            }
        }
#pragma unroll
        for (int s = 0; s < NSYM; ++s) peq[(size_t)s * NB + b] = w[s];
    }
}

__device__ __forceinline__ bool all_cells_larger(Word P, Word M, int score, int k) {        // edlib.cpp:510-516
    for (int i = 0; i < WS; ++i) if (cell_above(P, M, score, i) <= k) return false;
    return true;
}

// myersCalcEditDistanceSemiGlobal, edlib.cpp:537-705.  Returns the best score (-1: none <= k); first_pos / last_pos
// are the first and the last entry of the reference's `positions` vector.
__device__ int semi_global(const Word *peq, int W, int NB, int qlen, const Seq &t, int k, int mode, const State &st,
                           int &first_pos, int &last_pos, unsigned long long &units) {
    Word *P = st.P, *M = st.M; int *S = st.S;
    int first = 0, last = imin(ceil_div(k + 1, WS), NB) - 1;
    if (mode == LCD_EDLIB_MODE_HW) k = imin(qlen, k);
    for (int b = 0; b <= last; ++b) { S[b] = (b + 1) * WS; P[b] = ~(Word)0; M[b] = 0; }
    int best = -1;
    const int start_hout = mode == LCD_EDLIB_MODE_HW ? 0 : 1;
    const int tlen = t.n;
    for (int c = 0; c < tlen; ++c) {
        const Word *pc = peq + (size_t)t.at(c) * NB;
        int hout = start_hout;
        for (int b = first; b <= last; ++b) { hout = advance_block(P[b], M[b], pc[b], hout, P[b], M[b]); S[b] += hout; }
        units += (unsigned long long)(last - first + 1);
        if (last < NB - 1 && S[last] - hout <= k && ((pc[last + 1] & 1) || hout < 0)) {          // :600-608
            last++;
            P[last] = ~(Word)0; M[last] = 0;
            const int nh = advance_block(P[last], M[last], pc[last], hout, P[last], M[last]);
            S[last] = S[last - 1] - hout + WS + nh;
            units++;
        } else {
            while (last >= first && S[last] >= k + WS) last--;                                   // :609-613
        }
        if (c % 2048 == 0) while (last >= 0 && last >= first && all_cells_larger(P[last], M[last], S[last], k)) last--;
        if (mode == LCD_EDLIB_MODE_HW && last == -1) last++;                                     // :629-631
        if (mode != LCD_EDLIB_MODE_HW) {                                                         // :634-643
            while (first <= last && S[first] >= k + WS) first++;
            if (c % 2048 == 0) while (first <= last && all_cells_larger(P[first], M[first], S[first], k)) first++;
        }
        if (last < first) return best;                                                           // :646-655
        if (last == NB - 1) {                                                                    // :659-676
            const int cs = S[last];
            if (cs <= k && (best == -1 || cs <= best)) {
                if (cs != best) { first_pos = c - W; best = cs; k = best; }
                last_pos = c - W;
            }
        }
    }
    if (last == NB - 1) {                                                                        // :683-696
        for (int i = 0; i < W; ++i) {
            const int cs = cell_above(P[last], M[last], S[last], i + 1);
            if (cs <= k && (best == -1 || cs <= best)) {
                if (cs != best) { first_pos = tlen - W + i; k = best = cs; }
                last_pos = tlen - W + i;
            }
        }
    }
    return best;
}

// per-column store read by the traceback (AlignmentData, edlib.cpp:22-47)
struct AlignData { Word *Ps, *Ms; int *scores, *first, *last; };

// myersCalcEditDistanceNW, edlib.cpp:730-925.  ad != nullptr && stop < 0: keep every column; stop >= 0: compute up to
// column `stop` and keep it as the only column.  Returns the score (-1: none within k / stopped).  ok = false when
// the band died (no column stored).
__device__ int nw(const Word *peq, int W, int NB, int qlen, const Seq &t, int k, const State &st,
                  const AlignData *ad, int stop, bool &ok, unsigned long long &units) {
    Word *P = st.P, *M = st.M; int *S = st.S;
    const int tlen = t.n;
    ok = false;
    { const int d = tlen - qlen; if (k < (d < 0 ? -d : d)) return -1; }
    k = imin(k, imax(qlen, tlen));
    int first = 0;
    int last = imin(NB, ceil_div(imin(k, (k + qlen - tlen) / 2) + 1, WS)) - 1;
    for (int b = 0; b <= last; ++b) { S[b] = (b + 1) * WS; P[b] = ~(Word)0; M[b] = 0; }
    for (int c = 0; c < tlen; ++c) {
        const Word *pc = peq + (size_t)t.at(c) * NB;
        int hout = 1;
        for (int b = first; b <= last; ++b) { hout = advance_block(P[b], M[b], pc[b], hout, P[b], M[b]); S[b] += hout; }
        units += (unsigned long long)(last - first + 1);
        k = imin(k, S[last] + imax(tlen - c - 1, qlen - ((1 + last) * WS - 1) - 1) + (last == NB - 1 ? W : 0));   // :789-791
        if (last + 1 < NB && !((last + 1) * WS - 1 > k - S[last] + 2 * WS - 2 - tlen + c + qlen)) {             // :796-806
            last++;
            P[last] = ~(Word)0; M[last] = 0;
            const int nh = advance_block(P[last], M[last], pc[last], hout, P[last], M[last]);
            S[last] = S[last - 1] - hout + WS + nh;
            hout = nh;
            units++;
        }
        while (last >= first && (S[last] >= k + WS                                                                 // :811-818
                                 || ((last + 1) * WS - 1 > k - S[last] + 2 * WS - 2 - tlen + c + qlen + 1))) last--;
        while (first <= last && (S[first] >= k + WS                                                                // :823-829
                                 || ((first + 1) * WS - 1 < S[first] - k - tlen + qlen + c))) first++;
        if (c % 2048 == 0) {                                                                                       // :834-870
            while (last >= first) {
                const int ncell = last == NB - 1 ? WS - W : WS;
                int r = last * WS + ncell - 1; bool reduce = true;
                for (int i = WS - ncell; i < WS; ++i) {
                    const int v = cell_above(P[last], M[last], S[last], i);
                    if (v <= k && r <= k - v - tlen + c + qlen + 1) { reduce = false; break; }
                    r--;
                }
                if (!reduce) break;
                last--;
            }
            while (first <= last) {
                const int ncell = first == NB - 1 ? WS - W : WS;
                int r = first * WS + ncell - 1; bool reduce = true;
                for (int i = WS - ncell; i < WS; ++i) {
                    const int v = cell_above(P[first], M[first], S[first], i);
                    if (v <= k && r >= v - k - tlen + c + qlen) { reduce = false; break; }
                    r--;
                }
                if (!reduce) break;
                first++;
            }
        }
        if (last < first) return -1;                                                                               // :874-878
        if (ad && stop < 0) {                                                                                      // :883-893
            const size_t base = (size_t)NB * c;
            for (int b = first; b <= last; ++b) { ad->Ps[base + b] = P[b]; ad->Ms[base + b] = M[b]; ad->scores[base + b] = S[b]; }
            ad->first[c] = first; ad->last[c] = last;
        }
        if (c == stop) {                                                                                           // :896-908
            for (int b = first; b <= last; ++b) { ad->Ps[b] = P[b]; ad->Ms[b] = M[b]; ad->scores[b] = S[b]; }
            ad->first[0] = first; ad->last[0] = last;
            ok = true;
            return -1;
        }
    }
    ok = true;
    if (last == NB - 1) {                                                                                          // :913-921
        const int v = cell_above(P[last], M[last], S[last], W);
        if (v <= k) return v;
    }
    return -1;
}

// obtainAlignmentTraceback, edlib.cpp:940-1148.  Appends the path (forward order) at aln; returns its length.
__device__ int traceback(int qlen, int tlen, int best, const AlignData &ad, uint8_t *aln) {
    const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
    int n = 0, c = tlen - 1, b = NB - 1;
    int cur = best, ls = -1, us = -1, uls = -1;
    Word curP = ad.Ps[(size_t)c * NB + b], curM = ad.Ms[(size_t)c * NB + b];
    bool left_blk = c > 0 && b >= ad.first[c - 1] && b <= ad.last[c - 1];
    Word lP = 0, lM = 0;
    if (left_blk) { lP = ad.Ps[(size_t)(c - 1) * NB + b]; lM = ad.Ms[(size_t)(c - 1) * NB + b]; }
    curP <<= W; curM <<= W;
    int pos = WS - W - 1;
    for (;;) {
        if (c == 0) { left_blk = true; ls = b * WS + pos + 1; uls = ls - 1; }
        if (ls == -1 && left_blk) {
            // score of the left cell: walk up from the bottom of the left block (WS - pos - 1 cells)
            const int up = WS - pos - 1;
            ls = cell_above(lP, lM, ad.scores[(size_t)(c - 1) * NB + b], up);
            lP = up >= WS ? 0 : lP << up; lM = up >= WS ? 0 : lM << up;
        }
        if (uls == -1) {
            if (ls != -1) uls = ls - (int)(lP >> 63) + (int)(lM >> 63);
            else if (c > 0 && b - 1 >= ad.first[c - 1] && b - 1 <= ad.last[c - 1]) uls = ad.scores[(size_t)(c - 1) * NB + b - 1];
        }
        if (us == -1) {
            us = cur - (int)(curP >> 63) + (int)(curM >> 63);
            curP <<= 1; curM <<= 1;
        }
        if (us != -1 && us + 1 == cur) {                          // up
            cur = us; ls = uls; us = uls = -1;
            if (pos == 0) {
                if (b == 0) {
                    aln[n++] = OP_INSERT;
                    for (int i = 0; i < c + 1; ++i) aln[n++] = OP_DELETE;
                    break;
                }
                pos = WS - 1; b--;
                curP = ad.Ps[(size_t)c * NB + b]; curM = ad.Ms[(size_t)c * NB + b];
                if (c > 0 && b >= ad.first[c - 1] && b <= ad.last[c - 1]) {
                    left_blk = true; lP = ad.Ps[(size_t)(c - 1) * NB + b]; lM = ad.Ms[(size_t)(c - 1) * NB + b];
                } else left_blk = false;
            } else { pos--; lP <<= 1; lM <<= 1; }
            aln[n++] = OP_INSERT;
        } else if (ls != -1 && ls + 1 == cur) {                   // left
            cur = ls; us = uls; ls = uls = -1;
            c--;
            if (c == -1) {
                aln[n++] = OP_DELETE;
                const int up = b * WS + pos + 1;
                for (int i = 0; i < up; ++i) aln[n++] = OP_INSERT;
                break;
            }
            curP = lP; curM = lM;
            if (c > 0 && b >= ad.first[c - 1] && b <= ad.last[c - 1]) {
                left_blk = true; lP = ad.Ps[(size_t)(c - 1) * NB + b]; lM = ad.Ms[(size_t)(c - 1) * NB + b];
            } else if (c == 0) { left_blk = true; ls = b * WS + pos + 1; uls = ls - 1; }
            else left_blk = false;
            aln[n++] = OP_DELETE;
        } else if (uls != -1) {                                   // diagonal
            const uint8_t code = uls == cur ? OP_MATCH : OP_MISMATCH;
            cur = uls; us = ls = uls = -1;
            c--;
            if (c == -1) {
                aln[n++] = code;
                const int up = b * WS + pos;
                for (int i = 0; i < up; ++i) aln[n++] = OP_INSERT;
                break;
            }
            if (pos == 0) {
                if (b == 0) {
                    aln[n++] = code;
                    for (int i = 0; i < c + 1; ++i) aln[n++] = OP_DELETE;
                    break;
                }
                pos = WS - 1; b--;
                curP = ad.Ps[(size_t)c * NB + b]; curM = ad.Ms[(size_t)c * NB + b];
            } else { pos--; curP = lP << 1; curM = lM << 1; }
            if (c > 0 && b >= ad.first[c - 1] && b <= ad.last[c - 1]) {
                left_blk = true; lP = ad.Ps[(size_t)(c - 1) * NB + b]; lM = ad.Ms[(size_t)(c - 1) * NB + b];
            } else if (c == 0) { left_blk = true; ls = b * WS + pos + 1; uls = ls - 1; }
            else left_blk = false;
            aln[n++] = code;
        } else break;
    }
    for (int i = 0, j = n - 1; i < j; ++i, --j) { const uint8_t x = aln[i]; aln[i] = aln[j]; aln[j] = x; }
    return n;
}

// workspace words a problem needs (host and device agree through this one function)
__host__ __device__ inline uint64_t workspace_words(int qlen, int tlen, int want_path) {
    const uint64_t NB = (uint64_t)(qlen > 0 ? (qlen + WS - 1) / WS : 1);
    uint64_t w = 2 * NSYM * NB          // peq + reverse peq
               + 2 * NB + (NB + 1) / 2  // P, M, S
               + 8;
    if (want_path) {
        long long bytes = 20ll * (long long)NB * tlen + 8ll * tlen;
        if (bytes > TRACEBACK_LIMIT) bytes = TRACEBACK_LIMIT;
        w += (uint64_t)bytes / 8 + 16;                       // column store of one traceback leaf
        w += 2 * (2 * NB + (NB + 1) / 2 + 2);                // the two Hirschberg stop columns
        w += 2 * (NB * WS / 2 + 1);                          // unwrapped left / right scores
        w += (uint64_t)STACK * 3;                            // recursion stack (5 ints per entry)
    }
    return (w + 1) & ~1ull;
}

struct Frame { int qo, ql, to, tl, best; };

struct Aligner {
    Word *ws; uint64_t ws_words; uint64_t top;
    unsigned long long units;
    __device__ Word *take(uint64_t n) { Word *p = ws + top; top += n; return p; }

    // obtainAlignment / obtainAlignmentHirschberg (edlib.cpp:1169-1393) with an explicit stack: the upper-left
    // sub-problem is solved before the lower-right one and their paths are concatenated in that order.
    __device__ int path(const Seq &q, const Seq &t, int best, Word *peq, Word *rpeq, const State &st, uint8_t *aln, int &status) {
        const uint64_t mark = top;
        const int NBmax = q.n > 0 ? ceil_div(q.n, WS) : 1;
        long long cap_bytes = 20ll * NBmax * t.n + 8ll * t.n;
        if (cap_bytes > TRACEBACK_LIMIT) cap_bytes = TRACEBACK_LIMIT;
        Word *store = take((uint64_t)cap_bytes / 8 + 16);
        AlignData colL, colR;
        colL.Ps = take(NBmax); colL.Ms = take(NBmax); colL.scores = (int *)take((NBmax + 1) / 2); colL.first = (int *)take(1); colL.last = colL.first + 1;
        colR.Ps = take(NBmax); colR.Ms = take(NBmax); colR.scores = (int *)take((NBmax + 1) / 2); colR.first = (int *)take(1); colR.last = colR.first + 1;
        int *sl = (int *)take((uint64_t)NBmax * WS / 2 + 1), *sr0 = (int *)take((uint64_t)NBmax * WS / 2 + 1);
        Frame *stack = (Frame *)take((uint64_t)STACK * 3);
        if (top > ws_words) { status = ST_OOM; top = mark; return 0; }
        int sp = 0, n = 0;
        stack[sp++] = Frame{0, q.n, 0, t.n, best};
        while (sp > 0) {
            const Frame f = stack[--sp];
            if (f.ql == 0 || f.tl == 0) {                                       // :1176-1183
                for (int i = 0; i < f.ql + f.tl; ++i) aln[n++] = f.ql == 0 ? OP_DELETE : OP_INSERT;
                continue;
            }
            const Seq qs = q.sub(f.qo, f.ql), ts = t.sub(f.to, f.tl);
            const int NB = ceil_div(f.ql, WS), W = NB * WS - f.ql;
            const long long bytes = 20ll * NB * f.tl + 8ll * f.tl;
            bool ok;
            if (bytes < TRACEBACK_LIMIT) {                                      // :1195-1209
                AlignData ad;
                ad.Ps = store; ad.Ms = ad.Ps + (size_t)NB * f.tl; ad.scores = (int *)(ad.Ms + (size_t)NB * f.tl);
                ad.first = ad.scores + (size_t)NB * f.tl; ad.last = ad.first + f.tl;
                build_peq(peq, qs, NB);
                nw(peq, W, NB, f.ql, ts, f.best, st, &ad, -1, ok, units);
                if (!ok) { status = ST_NO_SOLUTION; break; }
                n += traceback(f.ql, f.tl, f.best, ad, aln + n);
                continue;
            }
            // Hirschberg split (:1236-1393)
            build_peq(peq, qs, NB);
            build_peq(rpeq, qs.reversed(), NB);
            const int lw = f.tl / 2, rw = f.tl - lw;
            nw(peq, W, NB, f.ql, ts, f.best, st, &colL, lw - 1, ok, units);
            if (!ok) { status = ST_NO_SOLUTION; break; }
            nw(rpeq, W, NB, f.ql, ts.reversed(), f.best, st, &colR, rw - 1, ok, units);
            if (!ok) { status = ST_NO_SOLUTION; break; }
            const int fl = colL.first[0], ll = colL.last[0];
            int nl = (ll - fl + 1) * WS;
            for (int b = fl; b <= ll; ++b)                                       // readBlock: top cell first
                for (int i = 0; i < WS; ++i) sl[(b - fl) * WS + i] = cell_above(colL.Ps[b], colL.Ms[b], colL.scores[b], WS - 1 - i);
            const int sl_start = fl * WS;
            if (ll == NB - 1) nl -= W;
            const int fr = colR.first[0], lr = colR.last[0];
            int nr = (lr - fr + 1) * WS;
            int *sr = sr0;
            for (int b = fr; b <= lr; ++b)                                       // readBlockReverse: bottom cell first
                for (int i = 0; i < WS; ++i) sr[(lr - b) * WS + i] = cell_above(colR.Ps[b], colR.Ms[b], colR.scores[b], i);
            int sr_start = f.ql - (lr + 1) * WS;
            if (sr_start < 0) { sr += W; sr_start += W; nr -= W; }
            const int q_s = imax(sl_start, sr_start - 1), q_e = imin(sl_start + nl - 1, sr_start + nr - 2);
            int ls = -1, rs = -1, row = -1; bool found = false;
            for (int i = q_s; i <= q_e; ++i) {
                ls = sl[i - sl_start]; rs = sr[i + 1 - sr_start];
                if (ls + rs == f.best) { row = i; found = true; break; }
            }
            if (!found && sl_start == 0 && sr_start == 0) {
                ls = lw; rs = sr[0];
                if (ls + rs == f.best) { row = -1; found = true; }
            }
            if (!found && sl_start + nl == f.ql && sr_start + nr == f.ql) {
                ls = sl[nl - 1]; rs = rw;
                if (ls + rs == f.best) { row = f.ql - 1; found = true; }
            }
            if (!found) { status = ST_NO_SPLIT; break; }
            const int ulh = row + 1, lrh = f.ql - ulh;
            if (sp + 2 > STACK) { status = ST_OOM; break; }
            stack[sp++] = Frame{f.qo + ulh, lrh, f.to + lw, rw, rs};              // popped second
            stack[sp++] = Frame{f.qo, ulh, f.to, lw, ls};                         // popped first
        }
        top = mark;
        return n;
    }

    // edlibAlign, edlib.cpp:146-305
    __device__ void align(const Problem &pb, const uint8_t *seqs, uint8_t *aln_out, Word *ws_, DevResult *res) {
        ws = ws_; ws_words = pb.ws_words; top = 0; units = 0;
        DevResult r; r.status = ST_OK; r.edit_distance = -1; r.start_loc = r.end_loc = -1; r.aln_len = 0; r.pad = 0;
        const int qlen = pb.qlen, tlen = pb.tlen, mode = pb.mode;
        Seq q; q.p = seqs + pb.q_off; q.n = qlen; q.rev = 0;
        Seq t; t.p = seqs + pb.t_off; t.n = tlen; t.rev = 0;
        bool bad = false;
        for (int i = 0; i < qlen; ++i) bad |= q.p[i] >= NSYM;
        for (int i = 0; i < tlen; ++i) bad |= t.p[i] >= NSYM;
        if (bad) r.status = ST_SYMBOL;
        else if (qlen == 0 || tlen == 0) {                                        // :165-183
            if (mode == LCD_EDLIB_MODE_NW) { r.edit_distance = imax(qlen, tlen); r.end_loc = tlen - 1; }
            else { r.edit_distance = qlen; r.end_loc = -1; }
        } else {
            const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
            Word *peq = take((uint64_t)NSYM * NB), *rpeq = take((uint64_t)NSYM * NB);
            State st; st.P = take(NB); st.M = take(NB); st.S = (int *)take((NB + 1) / 2);
            if (top > ws_words) r.status = ST_OOM;
            else {
                build_peq(peq, q, NB);
                int k = WS, best = -1, first_pos = -1, last_pos = -1;
                bool ok;
                do {                                                              // :199-217
                    if (mode == LCD_EDLIB_MODE_NW) best = nw(peq, W, NB, qlen, t, k, st, nullptr, -1, ok, units);
                    else best = semi_global(peq, W, NB, qlen, t, k, mode, st, first_pos, last_pos, units);
                    k *= 2;
                } while (best == -1 && k > 0);
                if (best < 0) r.status = ST_NO_SOLUTION;
                else {
                    r.edit_distance = best;
                    int end_loc = mode == LCD_EDLIB_MODE_NW ? tlen - 1 : first_pos, start_loc = 0;
                    if (!pb.want_path) start_loc = -1;
                    else if (mode == LCD_EDLIB_MODE_HW && end_loc != -1) {        // :229-262
                        build_peq(rpeq, q.reversed(), NB);
                        int f2 = -1, l2 = -1;
                        semi_global(rpeq, W, NB, qlen, t.sub(0, end_loc + 1).reversed(), best, LCD_EDLIB_MODE_SHW, st, f2, l2, units);
                        start_loc = end_loc - l2;
                    }
                    r.start_loc = start_loc; r.end_loc = end_loc;
                    if (pb.want_path) {                                           // :272-287
                        int status = ST_OK;
                        r.aln_len = path(q, t.sub(start_loc, end_loc - start_loc + 1), best, peq, rpeq, st, aln_out, status);
                        if (status != ST_OK) { r.status = status; r.aln_len = 0; }
                    }
                }
            }
        }
        r.units_lo = (uint32_t)units; r.units_hi = (uint32_t)(units >> 32);
        *res = r;
    }
};

} // namespace edlib
} // namespace lcd
