// wfa_device.cuh -- device-side logic of K6 (see wfa_kernel.cu for the design notes).
// Pure group-parallel code: everything is written against Grp<G> (G = 32: a warp, G = 256: a CTA).
// tests/emu compiles this same header for the host with G = 1 (single-lane emulation) to check the
// logic against the oracle under valgrind/ASAN without a GPU; the product only instantiates 32/256.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace wfa {

#define WF_NULL (INT32_MIN / 2)
constexpr int RING = 32;            // >= max_score_scope (26 for longcallD's penalties)
constexpr int NCOMP = 5;            // M, I1, D1, I2, D2
enum { C_M = 0, C_I1 = 1, C_D1 = 2, C_I2 = 3, C_D2 = 4 };
constexpr int STATUS_ESCALATED = -100;   // internal: handed over to the CTA kernel
constexpr int WARP_GROUPS_PER_CTA = 8;
#ifndef LCD_WFA_CTA_THREADS
#define LCD_WFA_CTA_THREADS 256
#endif
constexpr int CTA_GROUP_THREADS = LCD_WFA_CTA_THREADS;
constexpr int WARP_SEQ_SMEM = 2560;   // bytes of staged pattern+text per warp group
constexpr int CTA_SEQ_SMEM = 96 * 1024;
constexpr uint32_t OVERFLOW_CHUNK_UNITS = (8u << 20) / 16;   // 8 MiB overflow chunks (16-byte units)

struct __align__(16) WfSet {     // descriptor of the wavefronts of one score (64 bytes)
    int32_t base;                // diagonal stored at element 0 of every component
    int32_t wpad;                // padded width of one component, in int32 words (multiple of 4)
    uint32_t data;               // slab address: 16-byte units from the pool base
    int32_t mask;                // bit c: component c allocated ("exists")
    int32_t lo[NCOMP];
    int32_t hi[NCOMP];
    int32_t pad[2];
};
static_assert(sizeof(WfSet) == 64, "WfSet must be 64 bytes");

struct __align__(16) Problem {   // device-side problem descriptor
    uint64_t pat, txt;           // byte offsets into the packed sequence buffer (16-byte aligned)
    uint64_t ops;                // byte offset into the ops buffer (capacity 2*(plen+tlen)+8)
    int32_t plen, tlen;
    int32_t s_cap;               // descriptor capacity needed (upper bound on the final score + 2)
    int32_t pad;
    lcd_wfa_params_t par;
    int32_t pad2;
};

struct __align__(16) DevResult {
    int32_t status, score, n_ops, end_v, end_h;
    int32_t ops_begin;           // first operation inside the problem's ops slot
    uint32_t cells_lo, cells_hi; // wavefront cells computed (work units)
};

struct KernelArgs {
    const Problem *problems;
    const int32_t *order;        // problem indices of this class, largest first
    int32_t n;
    uint32_t *queue;             // next position in order[]
    const uint8_t *seqs;
    char *ops;
    DevResult *results;
    int32_t *pool;               // base of the workspace pool (int32 words)
    WfSet *meta;                 // n_groups * meta_cap descriptors
    int32_t meta_cap;
    uint32_t arena_base;         // private arenas: 16-byte units from pool base
    uint32_t arena_units;        // per group
    // shared overflow pool: n_chunks chunks of OVERFLOW_CHUNK_UNITS, handed out through a bitmap
    // (bit set = in use) and returned when the problem that took them finishes
    uint32_t overflow_base, n_chunks;
    uint32_t *chunk_bitmap;
    // escalation: a warp group gives up a problem whose score reaches esc_score (quadratic work on 32
    // lanes) and queues it for the CTA kernel; n_dev (if set) overrides n with a device-side count
    int32_t esc_score;
    int32_t *esc_list; uint32_t *esc_count;
    const uint32_t *n_dev;
};

// ------------------------------------------------------------------------------------------
// group primitives: G == 32 (one warp per problem) or G == CTA_GROUP_THREADS (one CTA per problem)
template <int G> struct Grp {
    int lane;                 // thread index inside the group
    WfSet *ring;              // shared: RING descriptors
    int *red;                 // shared scratch for CTA-wide reductions [2][G/32][16]
    int phase;
    __device__ __forceinline__ void sync() const {
        if (G == 32) __syncwarp(); else __syncthreads();
    }
    // all-lanes max reduction of N values (min = negate outside)
    template <int N> __device__ __forceinline__ void reduce_max(int (&v)[N]) {
        // Explicit reconvergence first: the callers reach this point after loops whose trip count
        // differs per lane, and nvcc lowers __reduce_max_sync to the convergent CREDUX on sm_100a.
        // Measured on B200: without the WARPSYNC a lane still in its last iteration is left out.
        __syncwarp();
#ifdef LCD_REDUX_SHFL
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v[i] = max(v[i], __shfl_xor_sync(0xffffffffu, v[i], d));
        }
#else
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __reduce_max_sync(0xffffffffu, v[i]);
#endif
        if (G > 32) {
            const int w = lane >> 5, nw = G / 32;
            int *buf = red + phase * (nw * 16);
            if ((lane & 31) == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) buf[w * 16 + i] = v[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < N; ++i) {
                int m = buf[i];
                for (int ww = 1; ww < nw; ++ww) m = max(m, buf[ww * 16 + i]);
                v[i] = m;
            }
            phase ^= 1;
        }
    }
    __device__ __forceinline__ uint32_t bcast_u32(uint32_t x) {   // value of lane 0 to all
        if (G == 32) return __shfl_sync(0xffffffffu, x, 0);
        int v[1] = { lane == 0 ? (int)(x ^ 0x80000000u) : INT32_MIN };
        reduce_max<1>(v);
        return (uint32_t)v[0] ^ 0x80000000u;
    }
};

struct In { const int32_t *p; int lo, hi; };
__device__ __forceinline__ int ld(const In &in, int k) {
    return (k >= in.lo && k <= in.hi) ? in.p[k] : WF_NULL;
}
__device__ __forceinline__ const int32_t *comp_ptr(const int32_t *pool, const WfSet &w, int comp) {
    const int slot = __popc(w.mask & ((1 << comp) - 1));
    return pool + (size_t)w.data * 4 + (size_t)slot * w.wpad - w.base;
}
__device__ __forceinline__ In fetch_in(const WfSet *ring, const int32_t *pool, int s, int comp, bool enabled = true) {
    In r; r.p = nullptr; r.lo = 1; r.hi = -1;
    if (s < 0 || !enabled) return r;
    const WfSet &w = ring[s & (RING - 1)];
    if (!((w.mask >> comp) & 1)) return r;
    const int lo = w.lo[comp], hi = w.hi[comp];
    if (lo > hi) return r;
    r.p = comp_ptr(pool, w, comp); r.lo = lo; r.hi = hi;
    return r;
}

// unaligned 32-bit load from a 4-byte aligned byte array (shared or global)
__device__ __forceinline__ uint32_t ldu32(const uint8_t *base, int i) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (i >> 2);
    return __funnelshift_r(w[0], w[1], (i & 3) * 8);
}
// number of matching bases from pattern[v], text[h]; sentinels ('!' vs '?') stop the run at either end
__device__ __forceinline__ int match_run(const uint8_t *pat, const uint8_t *txt, int v, int h) {
    int n = 0;
    for (;;) {
        const uint32_t x = ldu32(pat, v + n) ^ ldu32(txt, h + n);
        if (x) return n + ((__ffs(x) - 1) >> 3);
        n += 4;
    }
}

template <int G> struct Aligner {
    static constexpr int BTL = G < 32 ? G : 32;   // lanes that run the backtrace
    Grp<G> g;
    const KernelArgs &a;
    int32_t *pool;
    WfSet *gmeta;
    const uint8_t *pat, *txt;
    int plen, tlen;
    lcd_wfa_params_t par;
    // arena cursor (uniform across the group)
    uint32_t cur, end;
    uint32_t chunk_head;        // lane 0: most recent overflow chunk of this problem
    bool oom;
    unsigned long long cells;

    __device__ Aligner(const KernelArgs &args) : a(args) {}

    __device__ uint32_t chunk_take() {            // one lane
        const uint32_t nwords = (a.n_chunks + 31) >> 5;
        const uint32_t start = (uint32_t)(cells * 2654435761ull >> 7) % max(nwords, 1u);
        for (uint32_t j = 0; j < nwords; ++j) {
            const uint32_t i = (start + j) % nwords;
            uint32_t w = a.chunk_bitmap[i];
            while (~w) {
                const int bit = __ffs(~w) - 1;
                const uint32_t id = i * 32 + bit;
                if (id >= a.n_chunks) break;
                const uint32_t old = atomicOr(a.chunk_bitmap + i, 1u << bit);
                if (!(old & (1u << bit))) return id;
                w = old | (1u << bit);
            }
        }
        return 0xffffffffu;
    }
    __device__ void chunks_release() {            // one lane, at the end of a problem
        uint32_t id = chunk_head;
        while (id != 0xffffffffu) {
            const uint32_t next = (uint32_t)pool[((size_t)a.overflow_base + (size_t)id * OVERFLOW_CHUNK_UNITS) * 4];
            atomicAnd(a.chunk_bitmap + (id >> 5), ~(1u << (id & 31)));
            id = next;
        }
        chunk_head = 0xffffffffu;
    }
    __device__ __forceinline__ uint32_t alloc_units(uint32_t units) {
        if (cur + units <= end) { const uint32_t r = cur; cur += units; return r; }
        // private arena exhausted: take a chunk of the shared overflow pool (its first 16 bytes link
        // to the chunk taken before it, so that the problem can give them all back)
        uint32_t got = 0xffffffffu;
        if (g.lane == 0 && units + 1 <= OVERFLOW_CHUNK_UNITS) {
            got = chunk_take();
            if (got != 0xffffffffu) {
                pool[((size_t)a.overflow_base + (size_t)got * OVERFLOW_CHUNK_UNITS) * 4] = (int32_t)chunk_head;
                chunk_head = got;
            }
        }
        got = g.bcast_u32(got);
        if (got == 0xffffffffu) { oom = true; return a.arena_base; }
        const uint32_t base = a.overflow_base + got * OVERFLOW_CHUNK_UNITS;
        cur = base + 1 + units; end = base + OVERFLOW_CHUNK_UNITS;
        return base + 1;
    }
    __device__ __forceinline__ void publish(int score, const WfSet &w) {   // lane 0 only
        g.ring[score & (RING - 1)] = w;
        gmeta[score] = w;
    }

    // ---- one score step: compute + trim + extend, fused (wavefront_compute_affine2p.c:334-369,
    //      wavefront_compute.c:40-86,407-494,579-613, wavefront_extend.c:90-128) ----
    // returns: bit0 = M exists, bit1 = end reached
    __device__ int step(int score, int &num_null_steps) {
        const int two = par.affine2p;
        const int s_x = score - par.mismatch;
        const int s_o1 = score - par.gap_open1 - par.gap_ext1, s_e1 = score - par.gap_ext1;
        const int s_o2 = score - par.gap_open2 - par.gap_ext2, s_e2 = score - par.gap_ext2;
        const In m_x = fetch_in(g.ring, pool, s_x, C_M), m_o1 = fetch_in(g.ring, pool, s_o1, C_M);
        const In i1_e = fetch_in(g.ring, pool, s_e1, C_I1), d1_e = fetch_in(g.ring, pool, s_e1, C_D1);
        const In m_o2 = fetch_in(g.ring, pool, s_o2, C_M, two);
        const In i2_e = fetch_in(g.ring, pool, s_e2, C_I2, two), d2_e = fetch_in(g.ring, pool, s_e2, C_D2, two);
        const bool n_x = !m_x.p, n_o1 = !m_o1.p, n_i1 = !i1_e.p, n_d1 = !d1_e.p;
        const bool n_o2 = !m_o2.p, n_i2 = !i2_e.p, n_d2 = !d2_e.p;
        WfSet out;
        out.pad[0] = out.pad[1] = 0;
        if (n_x && n_o1 && n_i1 && n_d1 && n_o2 && n_i2 && n_d2) {   // null step
            ++num_null_steps;
            out.base = 0; out.wpad = 0; out.data = 0; out.mask = 0;
#pragma unroll
            for (int c = 0; c < NCOMP; ++c) { out.lo[c] = 1; out.hi[c] = -1; }
            g.sync();                      // everybody has read the ring entries they need
            if (g.lane == 0) publish(score, out);
            g.sync();
            return 0;
        }
        num_null_steps = 0;
        int lo = m_x.lo, hi = m_x.hi;
        lo = min(lo, m_o1.lo - 1); hi = max(hi, m_o1.hi + 1);
        lo = min(lo, i1_e.lo + 1); hi = max(hi, i1_e.hi + 1);
        lo = min(lo, d1_e.lo - 1); hi = max(hi, d1_e.hi - 1);
        if (two) {
            lo = min(lo, m_o2.lo - 1); hi = max(hi, m_o2.hi + 1);
            lo = min(lo, i2_e.lo + 1); hi = max(hi, i2_e.hi + 1);
            lo = min(lo, d2_e.lo - 1); hi = max(hi, d2_e.hi - 1);
        }
        int mask = 1;
        if (!n_o1 || !n_i1) mask |= 1 << C_I1;
        if (!n_o1 || !n_d1) mask |= 1 << C_D1;
        if (two && (!n_o2 || !n_i2)) mask |= 1 << C_I2;
        if (two && (!n_o2 || !n_d2)) mask |= 1 << C_D2;
        const bool full2p = two && !(n_o2 && n_i2 && n_d2);
        const int width = hi - lo + 1;
        const int wpad = (width + 3) & ~3;
        const int ncomp = __popc(mask);
        out.base = lo; out.wpad = wpad; out.mask = mask;
        out.data = alloc_units((uint32_t)(ncomp * wpad) >> 2);
        cells += (unsigned long long)width;
        int32_t *slab = pool + (size_t)out.data * 4 - lo;
        int32_t *o_m = slab;
        int slot = 1;
        int32_t *o_i1 = (mask >> C_I1 & 1) ? slab + (slot++) * wpad : nullptr;
        int32_t *o_d1 = (mask >> C_D1 & 1) ? slab + (slot++) * wpad : nullptr;
        int32_t *o_i2 = (mask >> C_I2 & 1) ? slab + (slot++) * wpad : nullptr;
        int32_t *o_d2 = (mask >> C_D2 & 1) ? slab + (slot++) * wpad : nullptr;
        const uint32_t utlen = (uint32_t)tlen, uplen = (uint32_t)plen;
        const int alignment_k = tlen - plen;
        // reduction slots: [0..4] = -min valid k per component, [5..9] = max valid k, [10] = end flag
        int r[11];
        int vmin[NCOMP], vmax[NCOMP];   // first / last valid diagonal per component seen by this lane
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) { vmin[c] = hi + 1; vmax[c] = lo - 1; }
        r[10] = 0;
        if (!oom) {
            for (int k = lo + g.lane; k <= hi; k += G) {
                const int ins1 = max(ld(m_o1, k - 1), ld(i1_e, k - 1)) + 1;
                const int del1 = max(ld(m_o1, k + 1), ld(d1_e, k + 1));
                const int misms = ld(m_x, k) + 1;
                int mx = max(del1, max(misms, ins1));
                if (o_i1) { o_i1[k] = ins1; if ((uint32_t)ins1 <= utlen && (uint32_t)(ins1 - k) <= uplen) { vmin[C_I1] = min(vmin[C_I1], k); vmax[C_I1] = max(vmax[C_I1], k); } }
                if (o_d1) { o_d1[k] = del1; if ((uint32_t)del1 <= utlen && (uint32_t)(del1 - k) <= uplen) { vmin[C_D1] = min(vmin[C_D1], k); vmax[C_D1] = max(vmax[C_D1], k); } }
                if (full2p) {
                    const int ins2 = max(ld(m_o2, k - 1), ld(i2_e, k - 1)) + 1;
                    const int del2 = max(ld(m_o2, k + 1), ld(d2_e, k + 1));
                    if (o_i2) { o_i2[k] = ins2; if ((uint32_t)ins2 <= utlen && (uint32_t)(ins2 - k) <= uplen) { vmin[C_I2] = min(vmin[C_I2], k); vmax[C_I2] = max(vmax[C_I2], k); } }
                    if (o_d2) { o_d2[k] = del2; if ((uint32_t)del2 <= utlen && (uint32_t)(del2 - k) <= uplen) { vmin[C_D2] = min(vmin[C_D2], k); vmax[C_D2] = max(vmax[C_D2], k); } }
                    mx = max(mx, max(ins2, del2));
                }
                if ((uint32_t)mx > utlen || (uint32_t)(mx - k) > uplen) mx = WF_NULL;
                else {
                    vmin[C_M] = min(vmin[C_M], k); vmax[C_M] = max(vmax[C_M], k);
                    mx += match_run(pat, txt, mx - k, mx);          // extend (wavefront_extend_kernels.c:66-112)
                    if (k == alignment_k && mx >= tlen) r[10] = 1;  // wavefront_termination.c:46-57
                }
                o_m[k] = mx;
            }
        }
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) { r[c] = -vmin[c]; r[5 + c] = vmax[c]; }
        g.template reduce_max<11>(r);
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
            if ((mask >> c) & 1) {
                // no valid cell: the reference ends with hi = lo-1 and lo unchanged
                const int mn = -r[c], mxk = r[5 + c];
                if (mxk < lo) { out.lo[c] = lo; out.hi[c] = lo - 1; }
                else { out.lo[c] = mn; out.hi[c] = mxk; }
            } else { out.lo[c] = 1; out.hi[c] = -1; }
        }
        // the reduction above is a barrier for G==32 (shuffles) and for CTAs (its __syncthreads)
        if (g.lane == 0) publish(score, out);
        g.sync();
        return 1 | (r[10] << 1);
    }

    // ---- heuristics (wavefront_heuristic.c:509-570, wfadaptive :232-292, zdrop :297-331,400-452) ----
    struct Heur { int steps_wait, max_sw_score, max_wf_score, max_sw_score_k, max_sw_score_offset; };
    // returns true when z-drop fired; end position written to end_*
    __device__ bool heuristic_cutoff(int score, Heur &h, int &end_k, int &end_offset) {
        WfSet w = g.ring[score & (RING - 1)];
        if (!(w.mask & 1) || w.lo[C_M] > w.hi[C_M]) return false;
        --h.steps_wait;
        const int lo = w.lo[C_M], hi = w.hi[C_M];
        const int32_t *m = comp_ptr(pool, w, C_M);
        int new_lo = lo, new_hi = hi;
        if (par.heuristic == LCD_WFA_HEUR_ADAPTIVE) {
            if (h.steps_wait > 0) return false;
            if (hi - lo + 1 < par.min_wavefront_length) return false;
            // (accumulate the natural minimum and negate once after the loop: nvcc 12.9 was seen to
            //  drop the negation of a max(acc, -d) accumulator on the peeled second iteration)
            int dmin = max(plen, tlen);
            for (int k = lo + g.lane; k <= hi; k += G) {
                const int off = m[k];
                const int d = (off >= 0) ? max(plen - (off - k), tlen - off) : -WF_NULL;
                dmin = min(dmin, d);
#ifdef LCD_TRACE
                if (score == -38) printf("  lane %d k %d off %d d %d base %d wpad %d data %u\n", g.lane, k, off, d, w.base, w.wpad, w.data);
#endif
            }
            int mind[1] = { -dmin };
#ifdef LCD_TRACE
            const int mind_before = mind[0];
#endif
            g.template reduce_max<1>(mind);
#ifdef LCD_TRACE
            if (score >= 36 && score <= 40) printf("  lane %d s %d lo %d hi %d before %d after %d steps_wait %d\n", g.lane, score, lo, hi, mind_before, mind[0], h.steps_wait);
#endif
            const int min_distance = -mind[0], thr = par.max_distance_threshold;
            int first_ok = INT32_MAX, last_ok = INT32_MIN;
            for (int k = lo + g.lane; k <= hi; k += G) {
                const int off = m[k];
                const int d = (off >= 0) ? max(plen - (off - k), tlen - off) : -WF_NULL;
                if (d - min_distance <= thr) { first_ok = min(first_ok, k); last_ok = max(last_ok, k); }
            }
            int fl[2] = { first_ok == INT32_MAX ? INT32_MIN : -first_ok, last_ok };   // -first ok k, last ok k
            g.template reduce_max<2>(fl);
            const int alignment_k = tlen - plen;
            const int top_limit = min(alignment_k, hi);
            if (top_limit > lo) new_lo = (fl[0] == INT32_MIN) ? top_limit : min(-fl[0], top_limit);
            const int bottom_limit = max(alignment_k, new_lo);
            if (bottom_limit < hi) new_hi = (fl[1] == INT32_MIN) ? bottom_limit : max(fl[1], bottom_limit);
            h.steps_wait = par.steps_between_cutoffs;
#ifdef LCD_TRACE
            if (g.lane == 0) printf("heur s=%d lo=%d hi=%d mind=%d fl=%d,%d -> %d %d\n", score, lo, hi, min_distance, fl[0], fl[1], new_lo, new_hi);
#endif
        } else if (par.heuristic == LCD_WFA_HEUR_ZDROP) {
            if (h.steps_wait > 0) return false;
            // first k attaining the maximum SW score: key = (sw, -k)
            int kv[2] = { INT32_MIN, INT32_MIN };
            int best_k = 0;
            for (int k = lo + g.lane; k <= hi; k += G) {
                const int off = m[k];
                if (off < 0) continue;
                const int sw = (-(off - k + off) - score) / 2;
                if (sw > kv[0]) { kv[0] = sw; best_k = k; }
            }
            if (kv[0] != INT32_MIN) kv[1] = -best_k;
            int best[1] = { kv[0] };
            g.template reduce_max<1>(best);
            int bk[1] = { (kv[0] == best[0] && kv[0] != INT32_MIN) ? kv[1] : INT32_MIN };
            g.template reduce_max<1>(bk);
            int cmax = best[0], cmax_k = 0, cmax_off = 0;
            if (bk[0] != INT32_MIN) { cmax_k = -bk[0]; cmax_off = m[cmax_k]; }
            if (h.max_sw_score_k != INT32_MAX) {
                if (cmax > h.max_sw_score) {
                    h.max_sw_score = cmax; h.max_wf_score = score; h.max_sw_score_k = cmax_k; h.max_sw_score_offset = cmax_off;
                } else if (h.max_sw_score - cmax > par.zdrop) {
                    end_k = h.max_sw_score_k; end_offset = h.max_sw_score_offset;
                    return true;
                }
            } else {
                h.max_sw_score = cmax; h.max_wf_score = score; h.max_sw_score_k = cmax_k; h.max_sw_score_offset = cmax_off;
            }
            h.steps_wait = par.steps_between_cutoffs;
        }
        if (new_lo == lo && new_hi == hi) return false;
        w.lo[C_M] = new_lo; w.hi[C_M] = new_hi;
#pragma unroll
        for (int c = 1; c < NCOMP; ++c) {           // wf_heuristic_equate :161-172
            if (!((w.mask >> c) & 1)) continue;
            if (new_lo > w.lo[c]) w.lo[c] = new_lo;
            if (new_hi < w.hi[c]) w.hi[c] = new_hi;
        }
        g.sync();
        if (g.lane == 0) publish(score, w);
        g.sync();
        return false;
    }

    // ---- backtrace (wavefront_backtrace.c:65-222,320-539): warp 0, lanes in lock-step ----
    __device__ __forceinline__ long long bt_cand(int comp, int s, int k, int add, int type) const {
        if (s < 0) return WF_NULL;
        const WfSet *w = gmeta + s;
        const int4 hd = *reinterpret_cast<const int4 *>(w);     // base, wpad, data, mask
        if (!((hd.w >> comp) & 1)) return WF_NULL;
        if (k < w->lo[comp] || k > w->hi[comp]) return WF_NULL;
        const int slot = __popc(hd.w & ((1 << comp) - 1));
        const int off = pool[(size_t)(uint32_t)hd.z * 4 + (size_t)slot * hd.y + (k - hd.x)];
        return (((long long)(off + add)) * 16) | type;
    }
    struct Cig { char *ops; int cap, begin, end; };
    __device__ __forceinline__ void bt_push(Cig &c, char op, int lane) const {
        if (lane == 0 && c.begin >= 0) c.ops[c.begin] = op;
        c.begin--;
    }
    __device__ __forceinline__ void bt_fill(Cig &c, char op, int n, int lane) const {
        const int b = c.begin;
        for (int i = lane; i < n; i += BTL) if (b - i >= 0) c.ops[b - i] = op;
        c.begin -= n;
    }
    __device__ void backtrace(Cig &c, int score, int k, int offset, int lane) const {
        enum { I1O = 1, I1E, I2O, I2E, D1O, D1E, D2O, D2E, MM };
        int matrix = C_M;
        int h = offset, v = offset - k;
        if (v < plen) bt_fill(c, 'D', plen - v, lane);
        if (h < tlen) bt_fill(c, 'I', tlen - h, lane);
        while (v > 0 && h > 0 && score > 0) {
            const int mismatch = score - par.mismatch;
            const int go1 = score - par.gap_open1 - par.gap_ext1, ge1 = score - par.gap_ext1;
            const int go2 = score - par.gap_open2 - par.gap_ext2, ge2 = score - par.gap_ext2;
            long long best;
            if (matrix == C_M) {
                best = bt_cand(C_M, mismatch, k, 1, MM);
                best = max(best, max(bt_cand(C_M, go1, k - 1, 1, I1O), bt_cand(C_I1, ge1, k - 1, 1, I1E)));
                best = max(best, max(bt_cand(C_M, go1, k + 1, 0, D1O), bt_cand(C_D1, ge1, k + 1, 0, D1E)));
                if (par.affine2p) {
                    best = max(best, max(bt_cand(C_M, go2, k - 1, 1, I2O), bt_cand(C_I2, ge2, k - 1, 1, I2E)));
                    best = max(best, max(bt_cand(C_M, go2, k + 1, 0, D2O), bt_cand(C_D2, ge2, k + 1, 0, D2E)));
                }
            } else if (matrix == C_I1) best = max(bt_cand(C_M, go1, k - 1, 1, I1O), bt_cand(C_I1, ge1, k - 1, 1, I1E));
            else if (matrix == C_I2)   best = max(bt_cand(C_M, go2, k - 1, 1, I2O), bt_cand(C_I2, ge2, k - 1, 1, I2E));
            else if (matrix == C_D1)   best = max(bt_cand(C_M, go1, k + 1, 0, D1O), bt_cand(C_D1, ge1, k + 1, 0, D1E));
            else                       best = max(bt_cand(C_M, go2, k + 1, 0, D2O), bt_cand(C_D2, ge2, k + 1, 0, D2E));
            if (best < 0) break;
            if (matrix == C_M) {
                const int max_offset = (int)(best >> 4);
                bt_fill(c, 'M', offset - max_offset, lane);
                offset = max_offset;
                v = offset - k; h = offset;
                if (v <= 0 || h <= 0) break;
            }
            const int type = (int)(best & 0xF);
            switch (type) {
                case MM:  score = mismatch; matrix = C_M;  break;
                case I1O: score = go1;      matrix = C_M;  break;
                case I1E: score = ge1;      matrix = C_I1; break;
                case I2O: score = go2;      matrix = C_M;  break;
                case I2E: score = ge2;      matrix = C_I2; break;
                case D1O: score = go1;      matrix = C_M;  break;
                case D1E: score = ge1;      matrix = C_D1; break;
                case D2O: score = go2;      matrix = C_M;  break;
                default:  score = ge2;      matrix = C_D2; break;
            }
            if (type == MM) { bt_push(c, 'X', lane); --offset; }
            else if (type <= I2E) { bt_push(c, 'I', lane); --k; --offset; }
            else { bt_push(c, 'D', lane); ++k; }
            v = offset - k; h = offset;
        }
        if (matrix == C_M) {
            if (v > 0 && h > 0) { const int n = min(v, h); bt_fill(c, 'M', n, lane); v -= n; h -= n; }
            if (v > 0) bt_fill(c, 'D', v, lane);
            if (h > 0) bt_fill(c, 'I', h, lane);
        }
        ++c.begin;
    }

    // cigar_maxtrim_gap_affine / _affine2p (alignment/cigar.c:476-600); lane 0 of warp 0 only
    __device__ void maxtrim(Cig &c, int &score_out, int &end_v, int &end_h) const {
        const int b = c.begin, e = c.end;
        int max_score = 0, max_off = b, max_v = 0, max_h = 0, score = 0, ev = 0, eh = 0;
        if (!par.affine2p) {
            char last = 0;
            for (int i = b; i < e; ++i) {
                const char op = c.ops[i];
                if (op == 'M') { score += 1; ++ev; ++eh; }
                else if (op == 'X') { score -= par.mismatch; ++ev; ++eh; }
                else if (op == 'I') { score -= par.gap_ext1 + ((last == 'I') ? 0 : par.gap_open1); ++eh; }
                else if (op == 'D') { score -= par.gap_ext1 + ((last == 'D') ? 0 : par.gap_open1); ++ev; }
                last = op;
                if (max_score < score) { max_score = score; max_off = i; max_v = ev; max_h = eh; }
            }
        } else {
            if (b >= e) return;
            char last = 0; int op_len = 0;
            for (int i = b; i <= e; ++i) {
                const char op = (i < e) ? c.ops[i] : 0;
                if (op != last && last != 0) {
                    const int s1 = par.gap_open1 + par.gap_ext1 * op_len, s2 = par.gap_open2 + par.gap_ext2 * op_len;
                    if (last == 'M') { score += op_len; ev += op_len; eh += op_len; }
                    else if (last == 'X') { score -= par.mismatch * op_len; ev += op_len; eh += op_len; }
                    else if (last == 'D') { score -= min(s1, s2); ev += op_len; }
                    else { score -= min(s1, s2); eh += op_len; }
                    op_len = 0;
                    if (max_score < score) { max_score = score; max_off = i - 1; max_v = ev; max_h = eh; }
                }
                last = op; ++op_len;
            }
        }
        if (max_score == 0) { c.begin = c.end = 0; score_out = INT32_MIN; end_v = end_h = -1; }
        else { c.end = max_off + 1; score_out = max_score; end_v = max_v; end_h = max_h; }
    }

    // ---- whole alignment (wavefront_unialign.c:242-275 + terminate :146-236) ----
    __device__ void align(const Problem &pb, DevResult *res, uint8_t *seq_smem, int seq_smem_bytes,
                          uint32_t arena_lo, uint32_t arena_hi, int problem_index) {
        plen = pb.plen; tlen = pb.tlen; par = pb.par;
        cur = arena_lo; end = arena_hi; oom = false; cells = 0; chunk_head = 0xffffffffu;
        // stage sequences (+ sentinel padding) in shared memory when they fit
        const int pbytes = (plen + 12 + 15) & ~15, tbytes = (tlen + 12 + 15) & ~15;
        const uint8_t *gp = a.seqs + pb.pat, *gt = a.seqs + pb.txt;
        if (pbytes + tbytes <= seq_smem_bytes) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(gp), *st = reinterpret_cast<const uint4 *>(gt);
            uint4 *dp = reinterpret_cast<uint4 *>(seq_smem), *dt = reinterpret_cast<uint4 *>(seq_smem + pbytes);
            for (int i = g.lane; i < pbytes / 16; i += G) dp[i] = sp[i];
            for (int i = g.lane; i < tbytes / 16; i += G) dt[i] = st[i];
            pat = seq_smem; txt = seq_smem + pbytes;
        } else { pat = gp; txt = gt; }
        const int scope_indel = par.affine2p ? max(par.gap_open1 + par.gap_ext1, par.gap_open2 + par.gap_ext2)
                                             : par.gap_open1 + par.gap_ext1;
        const int max_score_scope = max(scope_indel, par.mismatch) + 1;
        Heur hs; hs.steps_wait = par.steps_between_cutoffs; hs.max_sw_score = 0; hs.max_wf_score = 0;
        hs.max_sw_score_offset = WF_NULL; hs.max_sw_score_k = INT32_MAX;
        int end_k = INT32_MAX, end_offset = WF_NULL;
        int status = LCD_WFA_STATUS_COMPLETED;
        bool unreachable = false;
        int score = 0, num_null_steps = 0;
        const int alignment_k = tlen - plen;
        g.sync();                              // staged sequences visible; previous problem's ring reads done
        if (max_score_scope > RING || pb.s_cap > a.meta_cap) status = LCD_WFA_STATUS_ERROR;
        else {
            // score 0: M_0[0] = 0, extended (wavefront_aligner.c:251-310)
            WfSet w0; w0.base = 0; w0.wpad = 4; w0.mask = 1; w0.pad[0] = w0.pad[1] = 0;
            w0.data = alloc_units(1);
#pragma unroll
            for (int c = 0; c < NCOMP; ++c) { w0.lo[c] = 1; w0.hi[c] = -1; }
            w0.lo[C_M] = w0.hi[C_M] = 0;
            const int off0 = match_run(pat, txt, 0, 0);
            cells += 1;
            if (g.lane == 0) { pool[(size_t)w0.data * 4] = off0; publish(0, w0); }
            g.sync();
            bool done = (alignment_k == 0 && off0 >= tlen);
            if (done) { end_k = alignment_k; end_offset = tlen; }
            else if (par.heuristic != LCD_WFA_HEUR_NONE && heuristic_cutoff(0, hs, end_k, end_offset)) { unreachable = true; done = true; }
            while (!done) {
                ++score;
                if (score >= pb.s_cap || oom) { status = oom ? LCD_WFA_STATUS_OOM : LCD_WFA_STATUS_ERROR; break; }
                if (a.esc_score > 0 && score >= a.esc_score) { status = STATUS_ESCALATED; break; }
                const int st = step(score, num_null_steps);
                if (!(st & 1)) {
                    if (num_null_steps > max_score_scope) { unreachable = true; break; }
                    continue;
                }
                if (st & 2) { end_k = alignment_k; end_offset = tlen; break; }
                if (par.heuristic != LCD_WFA_HEUR_NONE && heuristic_cutoff(score, hs, end_k, end_offset)) { unreachable = true; break; }
            }
            if (oom) status = LCD_WFA_STATUS_OOM;
        }
        // ---- terminate: backtrace + (partial) maxtrim; warp 0 of the group ----
        if (g.lane < BTL) {
            const int lane = g.lane;
            // capacity as in the reference (cigar_resize(2*(plen+tlen)), wavefront_aligner.c:403): a z-drop
            // backtrace starts from an inconsistent (score, k, offset) and can exceed plen+tlen operations
            Cig c; c.ops = a.ops + pb.ops; c.cap = max(2 * (plen + tlen) - 1, 0); c.end = c.cap; c.begin = c.cap - 1;
            int out_score = INT32_MIN, end_v = -1, end_h = -1;
            if (status >= 0) {
                if (end_offset != WF_NULL) backtrace(c, score, end_k, end_offset, lane);
                else { c.begin = c.end = 0; }
                __syncwarp();
                if (unreachable) {
                    // lane 0 scans what the whole warp wrote: make the stores visible first
                    __threadfence_block();
                    if (lane == 0) maxtrim(c, out_score, end_v, end_h);
                    c.begin = __shfl_sync(0xffffffffu, c.begin, 0); c.end = __shfl_sync(0xffffffffu, c.end, 0);
                    status = LCD_WFA_STATUS_PARTIAL;
                } else {
                    end_v = end_offset - end_k; end_h = end_offset; out_score = -score;
                }
            } else { c.begin = c.end = 0; }
            if (lane == 0) {
                DevResult r;
                r.status = status; r.score = out_score; r.n_ops = max(c.end - c.begin, 0);
                r.end_v = end_v; r.end_h = end_h; r.ops_begin = c.begin;
                r.cells_lo = (uint32_t)cells; r.cells_hi = (uint32_t)(cells >> 32);
                *res = r;
                chunks_release();
                if (status == STATUS_ESCALATED) a.esc_list[atomicAdd(a.esc_count, 1u)] = problem_index;
            }
        }
        g.sync();
    }
};


} // namespace wfa
} // namespace lcd
