// sites_device.cuh -- device-side logic of the candidate-site list: collect_all_cand_var_sites (reference
// src/collect_var.c:1209-1254) = gather every collectible difference-list record of the kept reads (is_collectible_var_digar
// :1153-1160), sort by exact_comp_var_site (:1878-1898), de-duplicate with exact_comp_var_site_ins (:1901-1935, which also merges
// large insertions of similar length at one anchor).
//
// B200 design: the sort key starts with the anchor position, and a chunk's anchors fall in [reg_beg - 1, reg_end], so the sort is
// a counting sort on position bins (16 anchors per bin) followed by an exact per-bin pass:
//   count_read / scatter_read : one thread per read walks its records (already in HBM from K1) and counts / scatters the
//                               collectible ones into their bin's slice of the candidate array (atomics per bin);
//   group_bin                 : one thread per bin reduces its candidates to the distinct ones (most are the same variant seen by
//                               ~30 reads), insertion-sorts the handful that remain with the reference's comparator and applies the
//                               sequential large-insertion merge -- the order inside a bin does not depend on the scatter order
//                               because equal candidates are identical in every compared field;
//   emit_bin                  : one thread per bin writes its sites at the offset an exclusive scan of the per-bin counts gives.
// The sites refer to the alt bases of one of their records in K1's digar_alt array (nothing is copied), so K2 runs on them in place.
// The file compiles for the host as well (tests/emu).
#pragma once
#include <stdint.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace sites {

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
constexpr int BIN_SHIFT = 4;

struct __align__(16) Chunk {
    long long reg_beg, reg_end;       // collectible window (-1: open), chunk->reg_beg / reg_end
    long long lo;                     // anchor position of the chunk's first bin
    long long bin0, n_bins;           // the chunk's bins in the bin arrays
    long long alt_base;               // first digar_alt byte of the chunk (digar_alt_off is relative to it)
    int32_t min_sv_len, pad;
};

struct KernelArgs {
    const Chunk *chunks; long long n_reads_total, n_bins_total;
    const int32_t *read_chunk; const uint8_t *read_active, *read_dropped;
    const long long *digar_first; const int32_t *n_digar;
    const long long *digar_pos; const int8_t *digar_type; const int32_t *digar_len; const uint8_t *digar_low_qual; const long long *digar_alt_off; const uint8_t *digar_alt;
    const int32_t *bin_chunk_first;   // unused on the device (host bookkeeping)
    int32_t *bin_count; const long long *bin_first; int32_t *bin_cursor; long long *cand;
    int32_t *bin_keep; const long long *keep_first;
    long long *site_pos; int32_t *site_type, *site_ref_len, *site_alt_len; long long *site_src, *site_alt_off;
    int32_t *status;
};

__device__ __forceinline__ bool collectible(const KernelArgs &a, const Chunk &ch, long long d) {
    const int t = a.digar_type[d];
    if (t != CDIFF && t != CINS && t != CDEL) return false;
    if (a.digar_low_qual[d]) return false;
    const long long p = a.digar_pos[d];
    if (ch.reg_beg != -1 && p < ch.reg_beg) return false;
    if (ch.reg_end != -1 && p > ch.reg_end) return false;
    return true;
}
__device__ __forceinline__ long long anchor(const KernelArgs &a, long long d) { return a.digar_type[d] == CDIFF ? a.digar_pos[d] : a.digar_pos[d] - 1; }
__device__ __forceinline__ long long bin_of(const KernelArgs &a, const Chunk &ch, long long d) {
    long long b = (anchor(a, d) - ch.lo) >> BIN_SHIFT;
    if (b < 0) b = 0;
    if (b >= ch.n_bins) b = ch.n_bins - 1;      // (the plan sizes the bins to cover every read; clamping keeps a bad input in bounds)
    return ch.bin0 + b;
}

__device__ void count_read(const KernelArgs &a, long long g) {
    if (!a.read_active[g] || (a.read_dropped && a.read_dropped[g])) return;
    const Chunk ch = a.chunks[a.read_chunk[g]];
    for (long long d = a.digar_first[g], e = d + a.n_digar[g]; d < e; ++d)
        if (collectible(a, ch, d)) atomicAdd(a.bin_count + bin_of(a, ch, d), 1);
}

__device__ void scatter_read(const KernelArgs &a, long long g) {
    if (!a.read_active[g] || (a.read_dropped && a.read_dropped[g])) return;
    const Chunk ch = a.chunks[a.read_chunk[g]];
    for (long long d = a.digar_first[g], e = d + a.n_digar[g]; d < e; ++d)
        if (collectible(a, ch, d)) {
            const long long b = bin_of(a, ch, d);
            a.cand[a.bin_first[b] + atomicAdd(a.bin_cursor + b, 1)] = d;
        }
}

struct Key { long long anchor; int type, ref_len, alt_len; const uint8_t *alt; };
__device__ __forceinline__ Key key_of(const KernelArgs &a, long long alt_base, long long d) {
    Key k; const int t = a.digar_type[d], l = a.digar_len[d];
    k.anchor = t == CDIFF ? a.digar_pos[d] : a.digar_pos[d] - 1; k.type = t;
    k.ref_len = t == CINS ? 0 : (t == CDEL ? l : 1); k.alt_len = t == CDEL ? 0 : l;
    k.alt = a.digar_alt + alt_base + a.digar_alt_off[d];
    return k;
}
__device__ __forceinline__ int cmp_bytes(const uint8_t *x, const uint8_t *y, int n) {
    for (int i = 0; i < n; ++i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
// exact_comp_var_site, src/collect_var.c:1878-1898
__device__ __forceinline__ int cmp_exact(const Key &x, const Key &y) {
    if (x.anchor != y.anchor) return x.anchor < y.anchor ? -1 : 1;
    if (x.type != y.type) return x.type < y.type ? -1 : 1;
    if (x.ref_len != y.ref_len) return x.ref_len < y.ref_len ? -1 : 1;
    if (x.alt_len != y.alt_len) return x.alt_len < y.alt_len ? -1 : 1;
    if (x.type == CDIFF || x.type == CINS) return cmp_bytes(x.alt, y.alt, x.alt_len);
    return 0;
}
// exact_comp_var_site_ins, src/collect_var.c:1901-1935
__device__ __forceinline__ int cmp_ins(const Key &x, const Key &y, int min_sv_len) {
    if (x.anchor != y.anchor) return x.anchor < y.anchor ? -1 : 1;
    if (x.type != y.type) return x.type < y.type ? -1 : 1;
    if (x.ref_len != y.ref_len) return x.ref_len < y.ref_len ? -1 : 1;
    if (x.type == CDIFF || (x.type == CINS && x.alt_len < min_sv_len)) {
        if (x.alt_len != y.alt_len) return x.alt_len < y.alt_len ? -1 : 1;
        return cmp_bytes(x.alt, y.alt, x.alt_len);
    } else if (x.type == CINS) {
        const int mn = x.alt_len < y.alt_len ? x.alt_len : y.alt_len, mx = x.alt_len > y.alt_len ? x.alt_len : y.alt_len;
        if (mn >= mx * 0.8) return 0;
        return x.alt_len - y.alt_len;
    }
    return 0;
}

// one bin: candidates cand[f .. f+n) -> the bin's sites, in the reference's order, compacted to cand[f .. f+keep)
__device__ void group_bin(const KernelArgs &a, int chunk, long long bin) {
    const int n = a.bin_count[bin];
    if (n == 0) { a.bin_keep[bin] = 0; return; }
    const Chunk ch = a.chunks[chunk];
    long long *c = a.cand + a.bin_first[bin];
    int m = 0;
    for (int i = 0; i < n; ++i) {                    // distinct candidates, kept sorted (insertion at the first greater one)
        const long long d = c[i]; const Key k = key_of(a, ch.alt_base, d);
        int lo = 0, hi = m, eq = 0;
        while (lo < hi) {                            // binary search in the sorted distinct prefix
            const int mid = (lo + hi) >> 1;
            const int r = cmp_exact(key_of(a, ch.alt_base, c[mid]), k);
            if (r == 0) { eq = 1; if (d < c[mid]) c[mid] = d; break; }     // the lowest record index stands for the site: the result does not depend on the scatter order
            if (r < 0) lo = mid + 1; else hi = mid;
        }
        if (eq) continue;
        for (int j = m; j > lo; --j) c[j] = c[j - 1];   // m <= i: the slots up to i are free to reuse
        c[lo] = d; ++m;
    }
    int w = 1;                                       // the reference's sequential pass: compare with the last kept site
    for (int i = 1; i < m; ++i) {
        if (cmp_ins(key_of(a, ch.alt_base, c[w - 1]), key_of(a, ch.alt_base, c[i]), ch.min_sv_len) == 0) continue;
        c[w++] = c[i];
    }
    a.bin_keep[bin] = w;
}

__device__ void emit_bin(const KernelArgs &a, int chunk, long long bin) {
    const int w = a.bin_keep[bin];
    if (w == 0) return;
    const long long *c = a.cand + a.bin_first[bin]; const long long o = a.keep_first[bin];
    for (int k = 0; k < w; ++k) {
        const long long d = c[k]; const int t = a.digar_type[d], l = a.digar_len[d];
        a.site_pos[o + k] = a.digar_pos[d]; a.site_type[o + k] = t; a.site_ref_len[o + k] = t == CINS ? 0 : (t == CDEL ? l : 1);
        a.site_alt_len[o + k] = t == CDEL ? 0 : l; a.site_src[o + k] = d; a.site_alt_off[o + k] = a.digar_alt_off[d];
    }
}

} // namespace sites
} // namespace lcd
