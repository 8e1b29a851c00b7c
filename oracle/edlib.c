/* oracle/edlib.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of edlib's Myers/Hyyro block bit-vector alignment as longcallD uses it
 * (reference src/align.c:210-275: k = -1, no extra equalities, EDLIB_MODE_NW or EDLIB_MODE_HW,
 * EDLIB_TASK_PATH or EDLIB_TASK_DISTANCE).  Every function names the reference lines it follows
 * (edlib/src/edlib.cpp).  What has to be reproduced bit for bit is the PATH: the Ukkonen block band
 * (which cells exist for the traceback), the traceback preference up -> left -> diagonal, the
 * Hirschberg split (first row whose left + right scores reach the optimum) and the HW start-location
 * rule (last position of the reverse SHW search).
 *
 * Pinned against the unmodified edlib compiled into oracle/_ref/libref_shim.so (tests/test_oracle_edlib.py)
 * and against committed outputs of the reference (tests/golden/edlib_lcd.json.gz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lcd_oracle.h"

typedef uint64_t Word;
#define WS 64
#define HIGH_BIT ((Word)1 << 63)

typedef struct { Word P, M; int score; } Block;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int ceil_div(int x, int y) { return (x + y - 1) / y; }

/* calculateBlock, edlib.cpp:395-435: one 64-cell block of a column; hin/hout in {-1,0,1} */
static inline int advance_block(Word Pv, Word Mv, Word Eq, int hin, Word *PvOut, Word *MvOut) {
    const Word hin_neg = hin < 0 ? 1 : 0;
    const Word Xv = Eq | Mv;
    Eq |= hin_neg;
    const Word Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    Word Ph = Mv | ~(Xh | Pv);
    Word Mh = Pv & Xh;
    int hout = 0;
    if (Ph & HIGH_BIT) hout = 1;
    if (Mh & HIGH_BIT) hout -= 1;      /* both bits are never set together */
    Ph <<= 1; Mh <<= 1;
    Mh |= hin_neg;
    if (hin > 0) Ph |= 1;
    *PvOut = Mh | ~(Xv | Ph);
    *MvOut = Ph & Xv;
    return hout;
}

/* getBlockCellValues, edlib.cpp:458-470: v[0] = bottom cell ... v[63] = top cell */
static void cell_values(const Block *b, int *v) {
    int score = b->score; Word mask = HIGH_BIT;
    for (int i = 0; i < WS - 1; ++i) {
        v[i] = score;
        if (b->P & mask) score--;
        if (b->M & mask) score++;
        mask >>= 1;
    }
    v[WS - 1] = score;
}
/* allBlockCellsLarger, edlib.cpp:510-516 */
static int all_cells_larger(const Block *b, int k) {
    int v[WS]; cell_values(b, v);
    for (int i = 0; i < WS; ++i) if (v[i] <= k) return 0;
    return 1;
}

/* buildPeq, edlib.cpp:359-384: (alphabet+1) x NB words; the W padding rows of the last block match everything */
static Word *build_peq(int alen, const uint8_t *q, int qlen) {
    const int NB = ceil_div(qlen, WS);
    Word *peq = (Word*)malloc(sizeof(Word) * (size_t)(alen + 1) * NB);
    for (int s = 0; s <= alen; ++s)
        for (int b = 0; b < NB; ++b) {
            Word w = 0;
            if (s == alen) w = ~(Word)0;
            else for (int r = (b + 1) * WS - 1; r >= b * WS; --r) { w <<= 1; if (r >= qlen || q[r] == s) w |= 1; }
            peq[(size_t)s * NB + b] = w;
        }
    return peq;
}

typedef struct { int *v; int n, cap; } IntVec;
static void iv_push(IntVec *a, int x) {
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 16; a->v = (int*)realloc(a->v, sizeof(int) * a->cap); }
    a->v[a->n++] = x;
}

/* myersCalcEditDistanceSemiGlobal, edlib.cpp:537-705 (HW: free gaps before and after the query in the
 * target; SHW: free gap after).  positions: every column where the best score is attained. */
static void semi_global(const Word *Peq, int W, int NB, int qlen, const uint8_t *target, int tlen, int k, int mode,
                        int *best_out, IntVec *pos) {
    pos->n = 0;
    int first = 0, last = imin(ceil_div(k + 1, WS), NB) - 1;
    Block *bl = (Block*)malloc(sizeof(Block) * NB);
    if (mode == LCD_EDLIB_MODE_HW) k = imin(qlen, k);
    for (int b = 0; b <= last; ++b) { bl[b].score = (b + 1) * WS; bl[b].P = ~(Word)0; bl[b].M = 0; }
    int best = -1;
    const int start_hout = mode == LCD_EDLIB_MODE_HW ? 0 : 1;
    for (int c = 0; c < tlen; ++c) {
        const Word *pc = Peq + (size_t)target[c] * NB;
        int hout = start_hout;
        for (int b = first; b <= last; ++b) { hout = advance_block(bl[b].P, bl[b].M, pc[b], hout, &bl[b].P, &bl[b].M); bl[b].score += hout; }
        /* band: one more block below, or drop blocks that cannot hold a cell <= k (:600-613) */
        if (last < NB - 1 && bl[last].score - hout <= k && ((pc[last + 1] & 1) || hout < 0)) {
            last++;
            bl[last].P = ~(Word)0; bl[last].M = 0;
            const int nh = advance_block(bl[last].P, bl[last].M, pc[last], hout, &bl[last].P, &bl[last].M);
            bl[last].score = bl[last - 1].score - hout + WS + nh;
        } else {
            while (last >= first && bl[last].score >= k + WS) last--;
        }
        if (c % 2048 == 0) while (last >= 0 && last >= first && all_cells_larger(&bl[last], k)) last--;   /* :620-624 */
        if (mode == LCD_EDLIB_MODE_HW && last == -1) last++;                                                  /* :629-631 */
        if (mode != LCD_EDLIB_MODE_HW) {                                                                      /* :634-643 */
            while (first <= last && bl[first].score >= k + WS) first++;
            if (c % 2048 == 0) while (first <= last && all_cells_larger(&bl[first], k)) first++;
        }
        if (last < first) { *best_out = best; free(bl); return; }                                             /* :646-655 */
        if (last == NB - 1) {                                                                                 /* :659-676 */
            const int cs = bl[last].score;
            if (cs <= k && (best == -1 || cs <= best)) {
                if (cs != best) { pos->n = 0; best = cs; k = best; }
                iv_push(pos, c - W);
            }
        }
    }
    if (last == NB - 1) {                                                                                     /* :683-696 */
        int v[WS]; cell_values(&bl[last], v);
        for (int i = 0; i < W; ++i) {
            const int cs = v[i + 1];
            if (cs <= k && (best == -1 || cs <= best)) {
                if (cs != best) { pos->n = 0; k = best = cs; }
                iv_push(pos, tlen - W + i);
            }
        }
    }
    *best_out = best;
    free(bl);
}

/* AlignmentData, edlib.cpp:22-47: every block of every column (band marked per column) */
typedef struct { Word *Ps, *Ms; int *scores, *first, *last; } AlignData;
static AlignData *ad_new(int NB, int ncol) {
    AlignData *a = (AlignData*)malloc(sizeof(AlignData));
    a->Ps = (Word*)malloc(sizeof(Word) * (size_t)NB * ncol); a->Ms = (Word*)malloc(sizeof(Word) * (size_t)NB * ncol);
    a->scores = (int*)malloc(sizeof(int) * (size_t)NB * ncol);
    a->first = (int*)malloc(sizeof(int) * ncol); a->last = (int*)malloc(sizeof(int) * ncol);
    return a;
}
static void ad_free(AlignData *a) { if (!a) return; free(a->Ps); free(a->Ms); free(a->scores); free(a->first); free(a->last); free(a); }

/* myersCalcEditDistanceNW, edlib.cpp:730-925.  find_aln: keep every column; stop >= 0: compute up to that column
 * (inclusive) and return it as the only column of *ad_out. */
static void nw(const Word *Peq, int W, int NB, int qlen, const uint8_t *target, int tlen, int k,
               int *best_out, int find_aln, AlignData **ad_out, int stop) {
    *ad_out = NULL;
    if (k < abs(tlen - qlen)) { *best_out = -1; return; }
    k = imin(k, imax(qlen, tlen));
    int first = 0;
    int last = imin(NB, ceil_div(imin(k, (k + qlen - tlen) / 2) + 1, WS)) - 1;
    Block *bl = (Block*)malloc(sizeof(Block) * NB);
    for (int b = 0; b <= last; ++b) { bl[b].score = (b + 1) * WS; bl[b].P = ~(Word)0; bl[b].M = 0; }
    AlignData *ad = NULL;
    if (find_aln) ad = ad_new(NB, tlen); else if (stop > -1) ad = ad_new(NB, 1);
    *ad_out = ad;
    for (int c = 0; c < tlen; ++c) {
        const Word *pc = Peq + (size_t)target[c] * NB;
        int hout = 1;
        for (int b = first; b <= last; ++b) { hout = advance_block(bl[b].P, bl[b].M, pc[b], hout, &bl[b].P, &bl[b].M); bl[b].score += hout; }
        /* tighten k from the last block of the column (:789-791) */
        k = imin(k, bl[last].score + imax(tlen - c - 1, qlen - ((1 + last) * WS - 1) - 1) + (last == NB - 1 ? W : 0));
        /* one more block if the next one may still be inside the band (:796-806) */
        if (last + 1 < NB && !((last + 1) * WS - 1 > k - bl[last].score + 2 * WS - 2 - tlen + c + qlen)) {
            last++;
            bl[last].P = ~(Word)0; bl[last].M = 0;
            const int nh = advance_block(bl[last].P, bl[last].M, pc[last], hout, &bl[last].P, &bl[last].M);
            bl[last].score = bl[last - 1].score - hout + WS + nh;
            hout = nh;
        }
        /* drop blocks below / above the band (:811-829) */
        while (last >= first && (bl[last].score >= k + WS
                                 || ((last + 1) * WS - 1 > k - bl[last].score + 2 * WS - 2 - tlen + c + qlen + 1))) last--;
        while (first <= last && (bl[first].score >= k + WS
                                 || ((first + 1) * WS - 1 < bl[first].score - k - tlen + qlen + c))) first++;
        if (c % 2048 == 0) {                                                                                  /* :834-870 */
            int v[WS];
            while (last >= first) {
                cell_values(&bl[last], v);
                const int ncell = last == NB - 1 ? WS - W : WS;
                int r = last * WS + ncell - 1, reduce = 1;
                for (int i = WS - ncell; i < WS; ++i) {
                    if (v[i] <= k && r <= k - v[i] - tlen + c + qlen + 1) { reduce = 0; break; }
                    r--;
                }
                if (!reduce) break;
                last--;
            }
            while (first <= last) {
                cell_values(&bl[first], v);
                const int ncell = first == NB - 1 ? WS - W : WS;
                int r = first * WS + ncell - 1, reduce = 1;
                for (int i = WS - ncell; i < WS; ++i) {
                    if (v[i] <= k && r >= v[i] - k - tlen + c + qlen) { reduce = 0; break; }
                    r--;
                }
                if (!reduce) break;
                first++;
            }
        }
        if (last < first) { *best_out = -1; free(bl); return; }                                               /* :874-878 */
        if (find_aln) {                                                                                       /* :883-893 */
            for (int b = first; b <= last; ++b) {
                ad->Ps[(size_t)NB * c + b] = bl[b].P; ad->Ms[(size_t)NB * c + b] = bl[b].M; ad->scores[(size_t)NB * c + b] = bl[b].score;
            }
            ad->first[c] = first; ad->last[c] = last;
        }
        if (c == stop) {                                                                                      /* :896-908 */
            for (int b = first; b <= last; ++b) { ad->Ps[b] = bl[b].P; ad->Ms[b] = bl[b].M; ad->scores[b] = bl[b].score; }
            ad->first[0] = first; ad->last[0] = last;
            *best_out = -1; free(bl); return;
        }
    }
    *best_out = -1;
    if (last == NB - 1) {                                                                                     /* :913-921 */
        int v[WS]; cell_values(&bl[last], v);
        if (v[W] <= k) *best_out = v[W];
    }
    free(bl);
}

enum { OP_MATCH = 0, OP_INSERT = 1, OP_DELETE = 2, OP_MISMATCH = 3 };

/* obtainAlignmentTraceback, edlib.cpp:940-1148.  Writes the path (forward order) to aln; returns its length. */
static int traceback(int qlen, int tlen, int best, const AlignData *ad, uint8_t *aln) {
    const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
    int n = 0, c = tlen - 1, b = NB - 1;
    int cur = best, ls = -1, us = -1, uls = -1;
    Word curP = ad->Ps[(size_t)c * NB + b], curM = ad->Ms[(size_t)c * NB + b];
    int left_blk = c > 0 && b >= ad->first[c - 1] && b <= ad->last[c - 1];
    Word lP = 0, lM = 0;
    if (left_blk) { lP = ad->Ps[(size_t)(c - 1) * NB + b]; lM = ad->Ms[(size_t)(c - 1) * NB + b]; }
    curP <<= W; curM <<= W;
    int pos = WS - W - 1;
    for (;;) {
        if (c == 0) { left_blk = 1; ls = b * WS + pos + 1; uls = ls - 1; }
        if (ls == -1 && left_blk) {
            ls = ad->scores[(size_t)(c - 1) * NB + b];
            for (int i = 0; i < WS - pos - 1; ++i) {
                if (lP & HIGH_BIT) ls--;
                if (lM & HIGH_BIT) ls++;
                lP <<= 1; lM <<= 1;
            }
        }
        if (uls == -1) {
            if (ls != -1) {
                uls = ls;
                if (lP & HIGH_BIT) uls--;
                if (lM & HIGH_BIT) uls++;
            } else if (c > 0 && b - 1 >= ad->first[c - 1] && b - 1 <= ad->last[c - 1]) {
                uls = ad->scores[(size_t)(c - 1) * NB + b - 1];
            }
        }
        if (us == -1) {
            us = cur;
            if (curP & HIGH_BIT) us--;
            if (curM & HIGH_BIT) us++;
            curP <<= 1; curM <<= 1;
        }
        if (us != -1 && us + 1 == cur) {                          /* up: query base against a gap */
            cur = us; ls = uls; us = uls = -1;
            if (pos == 0) {
                if (b == 0) {
                    aln[n++] = OP_INSERT;
                    for (int i = 0; i < c + 1; ++i) aln[n++] = OP_DELETE;
                    break;
                }
                pos = WS - 1; b--;
                curP = ad->Ps[(size_t)c * NB + b]; curM = ad->Ms[(size_t)c * NB + b];
                if (c > 0 && b >= ad->first[c - 1] && b <= ad->last[c - 1]) {
                    left_blk = 1; lP = ad->Ps[(size_t)(c - 1) * NB + b]; lM = ad->Ms[(size_t)(c - 1) * NB + b];
                } else left_blk = 0;
            } else { pos--; lP <<= 1; lM <<= 1; }
            aln[n++] = OP_INSERT;
        } else if (ls != -1 && ls + 1 == cur) {                   /* left: target base against a gap */
            cur = ls; us = uls; ls = uls = -1;
            c--;
            if (c == -1) {
                aln[n++] = OP_DELETE;
                const int up = b * WS + pos + 1;
                for (int i = 0; i < up; ++i) aln[n++] = OP_INSERT;
                break;
            }
            curP = lP; curM = lM;
            if (c > 0 && b >= ad->first[c - 1] && b <= ad->last[c - 1]) {
                left_blk = 1; lP = ad->Ps[(size_t)(c - 1) * NB + b]; lM = ad->Ms[(size_t)(c - 1) * NB + b];
            } else if (c == 0) { left_blk = 1; ls = b * WS + pos + 1; uls = ls - 1; }
            else left_blk = 0;
            aln[n++] = OP_DELETE;
        } else if (uls != -1) {                                   /* diagonal */
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
                curP = ad->Ps[(size_t)c * NB + b]; curM = ad->Ms[(size_t)c * NB + b];
            } else { pos--; curP = lP << 1; curM = lM << 1; }
            if (c > 0 && b >= ad->first[c - 1] && b <= ad->last[c - 1]) {
                left_blk = 1; lP = ad->Ps[(size_t)(c - 1) * NB + b]; lM = ad->Ms[(size_t)(c - 1) * NB + b];
            } else if (c == 0) { left_blk = 1; ls = b * WS + pos + 1; uls = ls - 1; }
            else left_blk = 0;
            aln[n++] = code;
        } else break;
    }
    for (int i = 0, j = n - 1; i < j; ++i, --j) { const uint8_t t = aln[i]; aln[i] = aln[j]; aln[j] = t; }
    return n;
}

/* readBlock / readBlockReverse, edlib.cpp:478-503 */
static void read_block(const Block *b, int *dest) {          /* dest[0] = top cell */
    int score = b->score; Word mask = HIGH_BIT;
    for (int i = 0; i < WS - 1; ++i) { dest[WS - 1 - i] = score; if (b->P & mask) score--; if (b->M & mask) score++; mask >>= 1; }
    dest[0] = score;
}
static void read_block_rev(const Block *b, int *dest) {      /* dest[0] = bottom cell */
    int score = b->score; Word mask = HIGH_BIT;
    for (int i = 0; i < WS - 1; ++i) { dest[i] = score; if (b->P & mask) score--; if (b->M & mask) score++; mask >>= 1; }
    dest[WS - 1] = score;
}

static int obtain_alignment(const uint8_t *q, const uint8_t *rq, int qlen, const uint8_t *t, const uint8_t *rt, int tlen,
                            int alen, int best, uint8_t *aln, int *n_out);

/* obtainAlignmentHirschberg, edlib.cpp:1236-1393 */
static int hirschberg(const uint8_t *q, const uint8_t *rq, int qlen, const uint8_t *t, const uint8_t *rt, int tlen,
                      int alen, int best, uint8_t *aln, int *n_out) {
    const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
    Word *peq = build_peq(alen, q, qlen), *rpeq = build_peq(alen, rq, qlen);
    const int lw = tlen / 2, rw = tlen - lw;
    int sc;
    AlignData *L = NULL, *R = NULL;
    nw(peq, W, NB, qlen, t, tlen, best, &sc, 0, &L, lw - 1);
    nw(rpeq, W, NB, qlen, rt, tlen, best, &sc, 0, &R, rw - 1);
    free(peq); free(rpeq);
    if (!L || !R) { ad_free(L); ad_free(R); return -1; }
    const int fl = L->first[0], ll = L->last[0];
    int nl = (ll - fl + 1) * WS;
    int *sl = (int*)malloc(sizeof(int) * nl);
    for (int b = fl; b <= ll; ++b) { Block x = { L->Ps[b], L->Ms[b], L->scores[b] }; read_block(&x, sl + (b - fl) * WS); }
    const int sl_start = fl * WS;
    if (ll == NB - 1) nl -= W;
    const int fr = R->first[0], lr = R->last[0];
    int nr = (lr - fr + 1) * WS;
    int *sr0 = (int*)malloc(sizeof(int) * nr), *sr = sr0;
    for (int b = fr; b <= lr; ++b) { Block x = { R->Ps[b], R->Ms[b], R->scores[b] }; read_block_rev(&x, sr + (lr - b) * WS); }
    int sr_start = qlen - (lr + 1) * WS;
    if (sr_start < 0) { sr += W; sr_start += W; nr -= W; }
    ad_free(L); ad_free(R);
    /* first row of the left column that, with its lower-right neighbour in the right column, reaches best (:1310-1350) */
    const int qs = imax(sl_start, sr_start - 1), qe = imin(sl_start + nl - 1, sr_start + nr - 2);
    int ls = -1, rs = -1, row = -1, found = 0;
    for (int i = qs; i <= qe; ++i) {
        ls = sl[i - sl_start]; rs = sr[i + 1 - sr_start];
        if (ls + rs == best) { row = i; found = 1; break; }
    }
    if (!found && sl_start == 0 && sr_start == 0) {
        ls = lw; rs = sr[0];
        if (ls + rs == best) { row = -1; found = 1; }
    }
    if (!found && sl_start + nl == qlen && sr_start + nr == qlen) {
        ls = sl[nl - 1]; rs = rw;
        if (ls + rs == best) { row = qlen - 1; found = 1; }
    }
    free(sl); free(sr0);
    if (!found) return -1;
    const int ulh = row + 1, lrh = qlen - ulh;
    int n1 = 0, n2 = 0;
    if (obtain_alignment(q, rq + lrh, ulh, t, rt + rw, lw, alen, ls, aln, &n1)) return -1;
    if (obtain_alignment(q + ulh, rq, lrh, t + lw, rt, rw, alen, rs, aln + n1, &n2)) return -1;
    *n_out = n1 + n2;
    return 0;
}

/* obtainAlignment, edlib.cpp:1169-1217: traceback below 1 MiB of column data, Hirschberg above */
static int obtain_alignment(const uint8_t *q, const uint8_t *rq, int qlen, const uint8_t *t, const uint8_t *rt, int tlen,
                            int alen, int best, uint8_t *aln, int *n_out) {
    if (qlen == 0 || tlen == 0) {
        for (int i = 0; i < qlen + tlen; ++i) aln[i] = qlen == 0 ? OP_DELETE : OP_INSERT;
        *n_out = qlen + tlen;
        return 0;
    }
    const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
    const long long bytes = (2ll * sizeof(Word) + sizeof(int)) * NB * tlen + 2ll * sizeof(int) * tlen;
    if (bytes < 1024 * 1024) {
        int sc; AlignData *ad = NULL;
        Word *peq = build_peq(alen, q, qlen);
        nw(peq, W, NB, qlen, t, tlen, best, &sc, 1, &ad, -1);
        free(peq);
        if (!ad) return -1;
        *n_out = traceback(qlen, tlen, best, ad, aln);
        ad_free(ad);
        return 0;
    }
    return hirschberg(q, rq, qlen, t, rt, tlen, alen, best, aln, n_out);
}

static uint8_t *reverse_copy(const uint8_t *s, int n) {
    uint8_t *r = (uint8_t*)malloc(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) r[i] = s[n - 1 - i];
    return r;
}

/* edlibAlign, edlib.cpp:146-305 */
int lcd_oracle_edlib_align(const uint8_t *query, int qlen, const uint8_t *target, int tlen,
                           int mode, int want_path, uint8_t *aln, lcd_edlib_result_t *res) {
    res->status = 0; res->edit_distance = -1; res->start_loc = res->end_loc = -1; res->aln_len = 0;
    /* transformSequences :1417-1463: symbols renumbered by first appearance (query first) */
    uint8_t *q = (uint8_t*)malloc(qlen > 0 ? qlen : 1), *t = (uint8_t*)malloc(tlen > 0 ? tlen : 1);
    int idx[256], alen = 0;
    for (int i = 0; i < 256; ++i) idx[i] = -1;
    for (int i = 0; i < qlen; ++i) { if (idx[query[i]] < 0) idx[query[i]] = alen++; q[i] = (uint8_t)idx[query[i]]; }
    for (int i = 0; i < tlen; ++i) { if (idx[target[i]] < 0) idx[target[i]] = alen++; t[i] = (uint8_t)idx[target[i]]; }
    if (qlen == 0 || tlen == 0) {                                         /* :165-183 */
        if (mode == LCD_EDLIB_MODE_NW) { res->edit_distance = imax(qlen, tlen); res->end_loc = tlen - 1; }
        else { res->edit_distance = qlen; res->end_loc = -1; }
        free(q); free(t);
        return 0;
    }
    const int NB = ceil_div(qlen, WS), W = NB * WS - qlen;
    Word *peq = build_peq(alen, q, qlen);
    IntVec pos = { NULL, 0, 0 };
    int k = WS, best = -1;
    do {                                                                  /* :199-217 */
        if (mode == LCD_EDLIB_MODE_NW) { AlignData *ad = NULL; nw(peq, W, NB, qlen, t, tlen, k, &best, 0, &ad, -1); }
        else semi_global(peq, W, NB, qlen, t, tlen, k, mode, &best, &pos);
        k *= 2;
    } while (best == -1);
    res->edit_distance = best;
    int end_loc = mode == LCD_EDLIB_MODE_NW ? tlen - 1 : pos.v[0], start_loc = 0;
    if (!want_path) start_loc = -1;                                       /* TASK_DISTANCE: no start locations (:227) */
    else if (mode == LCD_EDLIB_MODE_HW) {                                 /* :229-262 */
        if (end_loc == -1) start_loc = 0;
        else {
            uint8_t *rt = reverse_copy(t, tlen), *rq = reverse_copy(q, qlen);
            Word *rpeq = build_peq(alen, rq, qlen);
            IntVec p2 = { NULL, 0, 0 }; int b2;
            semi_global(rpeq, W, NB, qlen, rt + tlen - end_loc - 1, end_loc + 1, best, LCD_EDLIB_MODE_SHW, &b2, &p2);
            start_loc = end_loc - p2.v[p2.n - 1];
            free(p2.v); free(rpeq); free(rt); free(rq);
        }
    }
    res->start_loc = start_loc; res->end_loc = end_loc;
    int rc = 0;
    if (want_path) {                                                      /* :272-287 */
        const uint8_t *at = t + start_loc; const int atl = end_loc - start_loc + 1;
        uint8_t *rat = reverse_copy(at, atl), *rq = reverse_copy(q, qlen);
        uint8_t *buf = (uint8_t*)malloc((size_t)qlen + atl + 2);
        int n = 0;
        rc = obtain_alignment(q, rq, qlen, at, rat, atl, alen, best, buf, &n);
        if (!rc) { res->aln_len = n; if (aln) memcpy(aln, buf, n); } else res->status = 1;
        free(buf); free(rat); free(rq);
    }
    free(pos.v); free(peq); free(q); free(t);
    return rc;
}
