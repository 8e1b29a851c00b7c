// poa_device.cuh -- device-side logic of K5/K5b: progressive partial-order alignment of the reads of
// one (noisy region, haplotype) into a consensus + row-column MSA, entirely on the GPU.
//
// What it replaces: the abPOA call sequence of abpoa_partial_aln_msa_cons (reference src/align.c:762-870)
// and abpoa_aln_msa_cons (:872-953) for full-cover reads with one consensus:
//   abpoa_align_sequence_to_subgraph  -> simd_abpoa_align_sequence_to_subgraph (abPOA/src/abpoa_align_simd.c:1250)
//   abpoa_add_subgraph_alignment      (abPOA/src/abpoa_graph.c:689) + abpoa_topological_sort (:322)
//   abpoa_output: abpoa_generate_rc_msa (abPOA/src/abpoa_output.c:149) + abpoa_most_frequent (:549)
//
// B200 design:
//   * one warp per POA problem (rows of the banded DP are 1-4 vectors wide, so a warp IS the natural
//     width); persistent warps pull problems, largest first, from a device queue.
//   * a warp lane is one int16 lane of the reference's AVX-512 vector: the band-snapping and the
//     2-then-1 lane F propagation of the SIMD code (which change results) are reproduced exactly, and
//     the F1/F2 prefix maxima are log-step warp-shuffle scans.
//   * only the band of each row is stored (5 int16 planes, 64-byte coalesced vector rows) in a per-warp
//     HBM arena that is recycled for every read, so the planes the next read touches are L2-resident.
//   * the graph never leaves the device: fusion of the alignment, the per-node exchange sort of the
//     edge lists, edge path scores, heaviest-path "remain" values and the final MSA / consensus all
//     run in the same kernel.  The topological order is kept as a linked list that is patched in O(1)
//     per new node instead of re-running the reference's BFS (any topological order yields the same DP).
//
// The file is written against a lane policy L so that tests/emu can run the identical logic on the
// host (32-element arrays instead of warp lanes) and diff it against the oracle without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/lcd_gpu.h"

#ifdef LCD_SIMT_EMU
static long simt_stat[8];       // test instrumentation: rows taken by chain segments / by the general path / ...
#endif
namespace lcd {
namespace poa {

constexpr int PN = 32;
constexpr int LOGN = 5;
constexpr int POA_SMEM_HALFS = 2 * 3 * 8 * 32;   // per-warp shared memory (int16): two row buffers x 3 planes x NVC vectors
constexpr int GARBAGE = 0x5555;        // what a read of a never-written DP cell returns (never decisive)

enum { ST_OK = 0, ST_INT32 = -1, ST_BAND = -2, ST_BACKTRACK = -3, ST_NOBASE = -4, ST_OOM = -5,
       ST_SUBGRAPH = -7, ST_SUB_UNSUPPORTED = -8 };      // bad anchors of a partially covering read / sub-graph alignment asked of a kernel variant without it

struct __align__(16) Problem {
    uint64_t seq_base;        // byte offset of the first read in the packed sequence buffer
    int32_t read_first;       // index of the first read in read_off[] / read_len[]
    int32_t n_reads;
    int32_t sum_len, max_len;
    int32_t cons_off;         // byte offset in the consensus buffer (capacity sum_len)
    int32_t node_cap;         // workspace budget of the first attempt (host estimate); a problem that outgrows
    int32_t edge_cap;         // it ends with ST_OOM and is re-run with worst-case budgets (KernelArgs.worst_case)
    int32_t min_w;            // max_n_cons = 2: MAX(2, ceil(n_reads * min_freq)) (abpoa_output.c:1141), computed on the host in double
    lcd_poa_params_t par;
};

struct __align__(16) DevResult {
    int32_t status;
    int32_t cons_len;
    int32_t msa_len;
    int32_t n_nodes;
    uint64_t msa_off;         // byte offset in the MSA output pool
    uint32_t cells_lo, cells_hi;
    int32_t n_cons, cons_len2; // max_n_cons = 2: number of read clusters (1 or 2) and the length of the second consensus (stored right after the first)
#ifdef LCD_POA_TIMING
    unsigned long long t_dp, t_bt, t_add, t_after, t_fin, t_seg, t_gen, n_seg, n_gen, t_pro;   // SM clock cycles per phase (debug builds); rows by chain segments / the general path
#endif
};

#if defined(LCD_POA_TIMING) && !defined(LCD_EMU)
#define LCD_T0() const long long t0_ = clock64()
#define LCD_T1(x) (x) += (unsigned long long)(clock64() - t0_)
#else
#define LCD_T0()
#define LCD_T1(x)
#endif

struct KernelArgs {
    const Problem *problems;
    const int32_t *order;
    int32_t n;
    uint32_t *queue;
    const uint8_t *seqs;
    const int64_t *read_off;      // relative to Problem.seq_base
    const int32_t *read_len;
    uint8_t *cons;
    uint8_t *msa; unsigned long long msa_cap; unsigned long long *msa_used;
    DevResult *results;
    int32_t *arena;               // per-group arenas
    uint64_t arena_words;         // int32 words per group
    int32_t worst_case;           // 1: ignore Problem.node_cap / edge_cap, size for the worst case
    // partially covering reads (nullptr: none): read r is aligned against the sub-graph between the nodes sub_beg[r] and sub_end[r] of the
    // first read (abpoa_subgraph_nodes; both 0: the whole graph), or left out (sub_beg[r] < 0).  Indexed like read_off / read_len.
    const int32_t *sub_beg, *sub_end;
    uint8_t *read_clu;            // max_n_cons = 2: cluster (0 / 1) of every read, indexed like read_off / read_len
};

// per-group workspace carved from the arena for one problem
struct WS {
    int N;                        // node capacity
    int *base, *in_off, *in_n, *in_cap, *out_off, *out_n, *out_cap, *n_read, *n_span;
    int *aln_off, *aln_n, *aln_cap, *next, *remain, *maxl, *maxr, *msa_rank;
    int *wsum, *order, *tmp, *s1, *s2, *s3, *s4, *s5, *s6, *s7;
    int *pos, *fp_id, *fp_ps;                        // inverse of order[]; first (heaviest) in-edge of a node: source and path score
    int4 *rinfo;                                     // per node: DP row descriptor of the current read (see Poa::pack)
    int4 *in_pool; int in_top, in_capacity;          // {from, w, ps, -}
    int *out_pool; int out_top, out_capacity, out_stride, rid_w;   // {to, w, rid[2*rid_w]}
    int *aln_pool; int aln_top, aln_capacity;
    int2 *cigar; int cigar_cap;                      // {op | len << 2, node_id}
    uint8_t *qs;                                     // the read shifted by one with sentinels: qs[j] = query[j-1] (strip rows)
    int4 *meta;                                      // per position of the topological order: {node, first in-edge source, base | min(n_in,255) << 8 | (path score & 0xff) << 16, remain}
    int16_t *qp; int qp_stride;                      // query profile of the current read: qp[b * qp_stride + j] = score of column j against node base b (b = 4: N)
    int16_t *dp; uint32_t dp_capacity;               // in cells
    int n_nodes;
    int oom;
};

// ---------------------------------------------------------------------------------------------
// lane policy for the GPU: one warp, lane l == int16 lane l of the reference's 512-bit vector
#ifndef LCD_EMU
struct WarpLanes {
    typedef int vec;                                  // this lane's cell (int16 value in an int)
    static constexpr int NT = 32, NW = 1;             // threads / warps cooperating on one problem
    static constexpr bool TWO_PHASE = false;
    static constexpr bool STRIP = true;               // rows are computed by Poa::strip_row (each lane owns C consecutive columns)
    __device__ static __forceinline__ int lane() { return threadIdx.x & 31; }
    __device__ static __forceinline__ int tid() { return threadIdx.x & 31; }
    __device__ static __forceinline__ int warp() { return 0; }
    __device__ static __forceinline__ void sync() { __syncwarp(); }
    __device__ static __forceinline__ vec load(const int16_t *p) { return p[lane()]; }
    __device__ static __forceinline__ vec load_m1(const int16_t *p, int first) {   // lane l <- p[l-1], lane 0 <- first
        return lane() == 0 ? first : (int)p[lane() - 1];
    }
    __device__ static __forceinline__ void store(int16_t *p, vec v) { p[lane()] = (int16_t)v; }
    __device__ static __forceinline__ vec set1(int x) { return x; }
    __device__ static __forceinline__ vec add(vec a, vec b) { return (int16_t)(a + b); }
    __device__ static __forceinline__ vec sub(vec a, vec b) { return (int16_t)(a - b); }
    __device__ static __forceinline__ vec vmax(vec a, vec b) { return a > b ? a : b; }
    __device__ static __forceinline__ vec shift_up(vec x, int n, int fill) {
        const int y = __shfl_up_sync(0xffffffffu, x, n);
        return lane() < n ? fill : y;
    }
    __device__ static __forceinline__ int lane_value(vec x, int l) { return __shfl_sync(0xffffffffu, x, l); }
    // keep lanes whose index is in [lo, hi], others <- fill
    __device__ static __forceinline__ vec keep(vec x, int lo, int hi, int fill) { return (lane() >= lo && lane() <= hi) ? x : fill; }
    // q score of this vector's columns (column of lane 0 = col0)
    template <class F> __device__ static __forceinline__ vec map_cols(int col0, F f) { return f(col0 + lane()); }
    // max over lanes lo..hi, and first / last lane attaining it (lanes outside ignored); false if empty
    __device__ static __forceinline__ bool row_max(vec x, int lo, int hi, int &m, int &first, int &last) {
        const bool in = lane() >= lo && lane() <= hi;
        if (lo > hi) return false;
        m = __reduce_max_sync(0xffffffffu, in ? x : INT32_MIN);
        const unsigned eq = __ballot_sync(0xffffffffu, in && x == m);
        first = __ffs(eq) - 1; last = 31 - __clz(eq);
        return true;
    }
};
// one CTA of W warps per problem (kilobase regions): same lane semantics per warp; the vectors of a DP row are
// dealt to the warps and the F carries are resolved through shared memory (Poa::align, two-phase rows)
template <int W> struct CtaLanes : WarpLanes {
    static constexpr int NT = 32 * W, NW = W;
    static constexpr bool TWO_PHASE = true;
    static constexpr bool STRIP = false;
    __device__ static __forceinline__ int tid() { return threadIdx.x; }
    __device__ static __forceinline__ int warp() { return threadIdx.x >> 5; }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
};
#endif


// ---------------------------------------------------------------------------------------------
// lane policy for one THREAD per problem (thousands of ~100-node problems): the 32 int16 lanes of a
// vector are packed two per 32-bit register and processed with the SIMD-in-word instructions
// (__vadd2 / __vsub2 / __vmaxs2 wrap / compare per halfword exactly like the epi16 instructions).
// Word i holds lane 2i (low half) and lane 2i+1 (high half).
struct ThreadLanes {
    struct vec { uint32_t w[16]; };
    static constexpr int NT = 1, NW = 1;
    static constexpr bool TWO_PHASE = false;
    static constexpr bool STRIP = false;
    __device__ static __forceinline__ int lane() { return 0; }
    __device__ static __forceinline__ int tid() { return 0; }
    __device__ static __forceinline__ int warp() { return 0; }
    __device__ static __forceinline__ void sync() {}
    __device__ static __forceinline__ vec load(const int16_t *p) {
        vec r; const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint4 t = q[i]; r.w[4 * i] = t.x; r.w[4 * i + 1] = t.y; r.w[4 * i + 2] = t.z; r.w[4 * i + 3] = t.w; }
        return r;
    }
    __device__ static __forceinline__ vec shift1(const vec &x, int first) {     // lane l <- x[l-1], lane 0 <- first
        vec r;
#pragma unroll
        for (int i = 15; i > 0; --i) r.w[i] = (x.w[i] << 16) | (x.w[i - 1] >> 16);
        r.w[0] = (x.w[0] << 16) | ((uint32_t)first & 0xffffu);
        return r;
    }
    __device__ static __forceinline__ vec load_m1(const int16_t *p, int first) { return shift1(load(p), first); }
    __device__ static __forceinline__ void store(int16_t *p, const vec &v) {
        uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i) { uint4 t; t.x = v.w[4 * i]; t.y = v.w[4 * i + 1]; t.z = v.w[4 * i + 2]; t.w = v.w[4 * i + 3]; q[i] = t; }
    }
    __device__ static __forceinline__ vec set1(int x) { vec r; const uint32_t v = ((uint32_t)x & 0xffffu) * 0x10001u;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = v; return r; }
    __device__ static __forceinline__ vec add(const vec &a, const vec &b) { vec r;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = __vadd2(a.w[i], b.w[i]); return r; }
    __device__ static __forceinline__ vec sub(const vec &a, const vec &b) { vec r;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = __vsub2(a.w[i], b.w[i]); return r; }
    __device__ static __forceinline__ vec vmax(const vec &a, const vec &b) { vec r;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = __vmaxs2(a.w[i], b.w[i]); return r; }
    __device__ static __forceinline__ vec shift_up(const vec &x, int n, int fill) {   // n in {1,2,4,8,16} (compile-time after unrolling)
        if (n == 1) return shift1(x, fill);
        vec r; const int k = n >> 1; const uint32_t f = ((uint32_t)fill & 0xffffu) * 0x10001u;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = i < k ? f : x.w[i - k < 0 ? 0 : i - k];
        return r;
    }
    __device__ static __forceinline__ int lane_value(const vec &x, int l) { return (int)(int16_t)(x.w[l >> 1] >> ((l & 1) * 16)); }
    __device__ static __forceinline__ vec keep(const vec &x, int lo, int hi, int fill) {
        vec r; const uint32_t f = (uint32_t)fill & 0xffffu;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t m = ((2 * i >= lo && 2 * i <= hi) ? 0xffffu : 0u) | ((2 * i + 1 >= lo && 2 * i + 1 <= hi) ? 0xffff0000u : 0u);
            r.w[i] = (x.w[i] & m) | ((f * 0x10001u) & ~m);
        }
        return r;
    }
    template <class F> __device__ static __forceinline__ vec map_cols(int col0, F f) {
        vec r;
#pragma unroll
        for (int i = 0; i < 16; ++i) r.w[i] = ((uint32_t)f(col0 + 2 * i) & 0xffffu) | ((uint32_t)f(col0 + 2 * i + 1) << 16);
        return r;
    }
    __device__ static __forceinline__ bool row_max(const vec &x, int lo, int hi, int &m, int &first, int &last) {
        if (lo > hi) return false;
        int mm = INT32_MIN;
#pragma unroll
        for (int l = 0; l < 32; ++l) { const int v = lane_value(x, l); if (l >= lo && l <= hi && v > mm) mm = v; }
        int fi = -1, la = -1;
#pragma unroll
        for (int l = 0; l < 32; ++l) { const int v = lane_value(x, l); if (l >= lo && l <= hi && v == mm) { if (fi < 0) fi = l; la = l; } }
        m = mm; first = fi; last = la;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
template <class L> struct Poa {
    typedef typename L::vec vec;
    WS w;
    lcd_poa_params_t par;
    int n_reads;
    int inf_min, oe1, oe2;
    unsigned long long cells;
    unsigned long long t_dp, t_bt, t_add, t_after, t_fin, t_seg, t_gen, n_seg, n_gen, t_pro;
    // optional per-group on-chip cache of the previous row's H/E1/E2 band (2 buffers x 3 planes x NVC vectors):
    // in a chain graph the only predecessor of a row is the row computed just before it, and reading it
    // back from HBM/L2 (~250 cycles) is the critical path of the whole DP.  nullptr = disabled.
    int16_t *row_cache = nullptr;
    static constexpr int NVC = 8;
    // the (sub-)graph the current read is aligned against: first / last node, rows of the order taken, and the in-edge lists as the alignment
    // sees them (a sub-graph alignment sees only predecessors inside the sub-graph: prepare_sub)
    int beg_id = 0, end_id = 1, n_rows = 0; bool sub = false;
    const int4 *ai_pool = nullptr; const int *ai_off = nullptr, *ai_n = nullptr;
    // Row descriptors of the order (w.meta, 16 B per row) reach the row loop through a two-half ring in shared memory that ONE lane keeps
    // filled with 1-D TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier per half): the next 32 rows' descriptors
    // land while the current 32 are computed, so the row loop never waits for a global load of its own.  nullptr: disabled (emulator, CTA kernel).
    int4 *ring = nullptr; unsigned long long *ring_bar = nullptr;
    int ring_block = -1, ring_issued = -1, ring_n = 0; unsigned ring_fill[2] = {0, 0}, ring_seen[2] = {0, 0};
    // scratch of a multi-warp group (shared memory): F carries of the row's vectors + row-max partials
    int *gs = nullptr;
    static constexpr int MAXV = 512, GS_INTS = 2 * MAXV + 4 + 3 * 32;

    // ---- workspace ---------------------------------------------------------------------------
    // columns of one base's row of the query profile: every vector a packed chain row can touch (dp_sn + 8) and 16-byte alignment
    __host__ __device__ static int qp_stride_of(int max_len) { return ((max_len + 32) / 32 + 9) * 32; }
    __device__ bool carve(int32_t *arena, uint64_t words, int N, int E, int max_len, int n_reads_) {
        uint64_t top = 0;
        w.N = N;
        int **arr[] = { &w.base, &w.in_off, &w.in_n, &w.in_cap, &w.out_off, &w.out_n, &w.out_cap, &w.n_read, &w.n_span,
                        &w.aln_off, &w.aln_n, &w.aln_cap, &w.next, &w.remain, &w.maxl, &w.maxr, &w.msa_rank,
                        &w.wsum, &w.order, &w.tmp, &w.s1, &w.s2, &w.s3, &w.s4, &w.s5, &w.s6, &w.s7,
                        &w.pos, &w.fp_id, &w.fp_ps };
        for (unsigned i = 0; i < sizeof(arr) / sizeof(arr[0]); ++i) { *arr[i] = arena + top; top += (uint64_t)((N + 3) & ~3); }
        w.rinfo = reinterpret_cast<int4 *>(arena + top); top += (uint64_t)N * 4;
        w.in_capacity = E; w.in_pool = reinterpret_cast<int4 *>(arena + top); top += (uint64_t)E * 4;
        w.rid_w = 1 + ((n_reads_ - 1) >> 6);
        w.out_stride = 2 + 2 * w.rid_w;
        w.out_capacity = E; w.out_pool = arena + top; top += ((uint64_t)E * w.out_stride + 3) & ~3ull;
        w.aln_capacity = E; w.aln_pool = arena + top; top += (uint64_t)((E + 3) & ~3);
        w.cigar_cap = max_len + N + 8; w.cigar = reinterpret_cast<int2 *>(arena + top); top += (uint64_t)w.cigar_cap * 2;
        w.qs = reinterpret_cast<uint8_t *>(arena + top); top += (uint64_t)(max_len + 192) / 4;
        top = (top + 3) & ~3ull;
        w.meta = reinterpret_cast<int4 *>(arena + top); top += (uint64_t)N * 4;
        w.qp_stride = qp_stride_of(max_len);
        w.qp = reinterpret_cast<int16_t *>(arena + top); top += (uint64_t)5 * w.qp_stride / 2;
        top = (top + 31) & ~31ull;
        if (top + 1024 > words) return false;
        w.dp = reinterpret_cast<int16_t *>(arena + top);
        const uint64_t cells_cap = (words - top) * 2;
        w.dp_capacity = cells_cap > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cells_cap;
        return true;
    }

    // ---- graph primitives (single lane) ---------------------------------------------------------
    __device__ int add_node(int b) {
        const int id = w.n_nodes;
        if (id >= w.N) { w.oom = 1; return w.N - 1; }
        w.base[id] = b; w.in_n[id] = w.in_cap[id] = w.out_n[id] = w.out_cap[id] = 0; w.in_off[id] = w.out_off[id] = 0;
        w.n_read[id] = w.n_span[id] = 0; w.aln_n[id] = w.aln_cap[id] = 0; w.aln_off[id] = 0; w.next[id] = -1;
        w.n_nodes = id + 1;
        return id;
    }
    __device__ int *out_entry(int node, int i) const { return w.out_pool + (size_t)(w.out_off[node] + i) * w.out_stride; }
    // abpoa_add_graph_edge, abpoa_graph.c:480-556 (weight 1, use_qv == 0)
    __device__ void add_edge(int from, int to, int check, int add_rid, int read_id) {
        int exist = 0, oi = -1;
        if (check) {
            int4 *ie = w.in_pool + w.in_off[to];
            for (int i = 0; i < w.in_n[to]; ++i) if (ie[i].x == from) { ie[i].y += 1; break; }
            for (int i = 0; i < w.out_n[from]; ++i) { int *e = out_entry(from, i); if (e[0] == to) { e[1] += 1; exist = 1; oi = i; break; } }
        }
        if (!exist) {
            if (w.in_n[to] == w.in_cap[to]) {
                const int nc = w.in_cap[to] ? w.in_cap[to] * 2 : 2;
                if (w.in_top + nc > w.in_capacity) { w.oom = 1; return; }
                int4 *src = w.in_pool + w.in_off[to], *dst = w.in_pool + w.in_top;
                for (int i = 0; i < w.in_n[to]; ++i) dst[i] = src[i];
                w.in_off[to] = w.in_top; w.in_top += nc; w.in_cap[to] = nc;
            }
            w.in_pool[w.in_off[to] + w.in_n[to]] = make_int4(from, 1, 0, 0);
            w.in_n[to]++;
            if (w.out_n[from] == w.out_cap[from]) {
                const int nc = w.out_cap[from] ? w.out_cap[from] * 2 : 2;
                if (w.out_top + nc > w.out_capacity) { w.oom = 1; return; }
                int *src = w.out_pool + (size_t)w.out_off[from] * w.out_stride, *dst = w.out_pool + (size_t)w.out_top * w.out_stride;
                const int nw = w.out_n[from] * w.out_stride;
                for (int i = 0; i < nw; ++i) dst[i] = src[i];
                w.out_off[from] = w.out_top; w.out_top += nc; w.out_cap[from] = nc;
            }
            oi = w.out_n[from];
            int *e = out_entry(from, oi);
            e[0] = to; e[1] = 1;
            for (int x = 0; x < 2 * w.rid_w; ++x) e[2 + x] = 0;
            w.out_n[from]++;
        }
        if (add_rid) out_entry(from, oi)[2 + (read_id >> 5)] |= 1 << (read_id & 31);
        w.n_read[from] += 1;
    }
    __device__ void add_aligned1(int node, int id) {
        if (w.aln_n[node] == w.aln_cap[node]) {
            const int nc = w.aln_cap[node] ? w.aln_cap[node] * 2 : 2;
            if (w.aln_top + nc > w.aln_capacity) { w.oom = 1; return; }
            for (int i = 0; i < w.aln_n[node]; ++i) w.aln_pool[w.aln_top + i] = w.aln_pool[w.aln_off[node] + i];
            w.aln_off[node] = w.aln_top; w.aln_top += nc; w.aln_cap[node] = nc;
        }
        w.aln_pool[w.aln_off[node] + w.aln_n[node]++] = id;
    }
    __device__ void add_aligned(int node, int aligned) {          // abpoa_add_graph_aligned_node :456-464
        for (int i = 0; i < w.aln_n[node]; ++i) {
            const int other = w.aln_pool[w.aln_off[node] + i];
            add_aligned1(other, aligned); add_aligned1(aligned, other);
        }
        add_aligned1(node, aligned); add_aligned1(aligned, node);
    }
    __device__ void list_insert_after(int after, int id) { w.next[id] = w.next[after]; w.next[after] = id; }
    // A node and the nodes aligned with it (one MSA column) stay contiguous in the list.  A new successor
    // of `id` goes after the whole column: fusion may later link ANY member of the column to it
    // (abpoa_get_aligned_id swaps the matched node for its aligned twin before abpoa_add_graph_edge).
    __device__ void list_insert_after_column(int id, int nw) {
        int pos = id;
        for (;;) {
            const int nx = w.next[pos];
            if (nx < 0) break;
            bool member = false;
            for (int i = 0; i < w.aln_n[id]; ++i) if (w.aln_pool[w.aln_off[id] + i] == nx) { member = true; break; }
            if (!member) break;
            pos = nx;
        }
        list_insert_after(pos, nw);
    }

    // abpoa_add_subgraph_alignment (abpoa_graph.c:689-774) for beg = SRC (0), end = SINK (1); single lane.
    // cigar[0..n_cig) is in backtrack (reverse) order.
    __device__ void add_alignment(const uint8_t *seq, int seq_l, int n_cig, int read_id) {
        const int inc = par.sub_aln ? 0 : 1;
        if (w.n_nodes == 2) {                         // abpoa_add_graph_sequence :573-593
            if (seq_l <= 0) return;
            int last = 0;
            for (int i = 0; i < seq_l; ++i) {
                const int cur = add_node(seq[i]);
                add_edge(last, cur, 0, 1, read_id);
                w.n_span[cur] = w.n_span[last];
                list_insert_after(last, cur);
                last = cur;
            }
            add_edge(last, 1, 0, 1, read_id);
            return;
        }
        if (n_cig == 0) return;
        int query_id = -1, last_new = 0, last_id = 0;
        for (int c = n_cig - 1; c >= 0; --c) {
            const int2 cg = w.cigar[c];
            const int op = cg.x & 3;
            if (op == 0) {
                const int node_id = cg.y;
                query_id++;
                const int add = (last_id != 0 || inc) ? 1 : 0;
                const int b = seq[query_id];
                if (w.base[node_id] != b) {
                    int aligned = -1;
                    for (int i = 0; i < w.aln_n[node_id]; ++i) { const int a = w.aln_pool[w.aln_off[node_id] + i]; if (w.base[a] == b) { aligned = a; break; } }
                    if (aligned != -1) {
                        add_edge(last_id, aligned, 1 - last_new, add, read_id);
                        if (!add) w.n_read[last_id]--;
                        last_id = aligned; last_new = 0;
                    } else {
                        const int nw = add_node(b);
                        add_edge(last_id, nw, 0, add, read_id);
                        w.n_span[nw] = w.n_span[last_id];
                        if (!add) w.n_read[last_id]--;
                        // keep the new node next to the column it is aligned with: a later read may enter
                        // it from ANY ancestor of node_id (abpoa_get_aligned_id), not only from last_id
                        list_insert_after(node_id, nw);
                        last_id = nw; last_new = 1;
                        add_aligned(node_id, nw);
                    }
                } else {
                    add_edge(last_id, node_id, 1 - last_new, add, read_id);
                    if (!add) w.n_read[last_id]--;
                    last_id = node_id; last_new = 0;
                }
            } else if (op == 1) {
                const int len = cg.x >> 2;
                query_id += len;
                for (int j = len - 1; j >= 0; --j) {
                    const int nw = add_node(seq[query_id - j]);
                    const int add = (last_id != 0 || inc) ? 1 : 0;
                    add_edge(last_id, nw, 0, add, read_id);
                    w.n_span[nw] = w.n_span[last_id];
                    if (!add) w.n_read[last_id]--;
                    list_insert_after_column(last_id, nw);
                    last_id = nw; last_new = 1;
                    if (w.oom) return;
                }
            }
            if (w.oom) return;
        }
        add_edge(last_id, 1, 1 - last_new, 1, read_id);
    }


#ifndef LCD_EMU
    // ---- warp policy: backtrack and graph fusion with all 32 lanes ------------------------------------------
    // The backtrack (simd_abpoa_cg_backtrack :309-458) yields, per query base, the graph node it is matched to
    // (path[q] = node id) or -1 (inserted base); deletions leave no trace, which is all abpoa_add_subgraph_alignment
    // needs.  Runs of plain diagonal matches through the heaviest in-edge -- almost every step for a read that
    // agrees with the graph -- are verified 32 at a time: lane l checks step l of the run along the topological
    // order (its node's first in-edge must come from the node before it in the order, and the MATCH test of the
    // reference must hold for that edge, which is then the first edge the reference would accept).  Any other step
    // is taken by the scalar rules, executed uniformly by the warp.
    __device__ int backtrack_warp(const uint8_t *query, int qlen, int *path) {
        const int lane = threadIdx.x & 31;
        const int e1 = par.gap_ext1, e2 = par.gap_ext2;
        for (int q = lane; q < qlen; q += 32) path[q] = -1;
        int best = inf_min, bi = 0, bj = 0;
        {
            const int4 *ie = ai_pool + ai_off[end_id];
            for (int k = 0; k < ai_n[end_id]; ++k) {
                const int r = ie[k].x;
                const Row rr = unpack(w.rinfo[r]);
                const int e = qlen > rr.end ? rr.end : qlen;
                const int s = cell(rr, 0, e);
                if (s > best) { best = s; bi = r; bj = e; }
            }
        }
        __syncwarp();
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int id = bi, j = bj, cur_op = ALL_OP, rc = 0;
        Row cur = unpack(w.rinfo[id]);
        while (id != beg_id && j > 0) {
            if (cur_op == ALL_OP) {
                const int pos = w.pos[id];
                int ok = 0, nl = -1, pl = -1;
                if (pos - lane >= 1 && j - lane >= 1) {
                    nl = w.order[pos - lane]; pl = w.order[pos - lane - 1];
                    if (w.fp_id[nl] == pl) {
                        const Row rn = lane == 0 ? cur : unpack(w.rinfo[nl]), rp = unpack(w.rinfo[pl]);
                        const int jl = j - lane;
                        if (jl - 1 >= rp.beg && jl - 1 <= rp.end) {
                            const int nb = w.base[nl], qb = query[jl - 1];
                            const int s = (nb > 3 || qb > 3) ? 0 : (nb == qb ? par.match : -par.mismatch);
                            ok = cell(rp, 0, jl - 1) + s + w.fp_ps[nl] == cell(rn, 0, jl);
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int run = __ffs(~m) - 1 < 0 ? 32 : __ffs(~m) - 1;
                if (run > 0) {
                    if (lane < run) path[j - lane - 1] = nl;
                    id = __shfl_sync(0xffffffffu, pl, run - 1);
                    j -= run;
                    cur = unpack(w.rinfo[id]);
                    continue;
                }
            }
            const int nb = w.base[id], qb = query[j - 1];
            const int s = (nb > 3 || qb > 3) ? 0 : (nb == qb ? par.match : -par.mismatch);
            const int4 *ie = ai_pool + ai_off[id];
            const int nin = ai_n[id];
            const int hj = cell(cur, 0, j);
            int hit = 0;
            for (int pass = 0; pass < 2 && !hit; ++pass) {
                if (pass == 1) {
                    if (cur_op & E_OP) {
                        const int e1j = cell(cur, 1, j), e2j = cell(cur, 2, j);
                        for (int k = 0; k < nin && !hit; ++k) {
                            const int4 e = ie[k];
                            const int p = e.x, ps = e.z;
                            const Row pr = unpack(w.rinfo[p]);
                            if (j < pr.beg || j > pr.end) continue;
                            const int phj = cell(pr, 0, j);
                            if (cur_op & E1_OP) {
                                const int pe = cell(pr, 1, j);
                                const int okk = (cur_op & M_OP) ? (hj == pe + ps) : (e1j == pe - e1 + ps);
                                if (okk) { cur_op = (phj - oe1 == pe) ? (M_OP | F_OP) : E1_OP; hit = 1; }
                            }
                            if (!hit && (cur_op & E2_OP)) {
                                const int pe = cell(pr, 2, j);
                                const int okk = (cur_op & M_OP) ? (hj == pe + ps) : (e2j == pe - e2 + ps);
                                if (okk) { cur_op = (phj - oe2 == pe) ? (M_OP | F_OP) : E2_OP; hit = 1; }
                            }
                            if (hit) { id = p; cur = pr; }                 // DEL: the node is skipped by the read
                        }
                    }
                    if (!hit && (cur_op & F_OP)) {
                        const int hj1 = cell(cur, 0, j - 1);
                        if (cur_op & F1_OP) {
                            const int f = cell(cur, 3, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe1 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(cur, 3, j - 1) - e1 == f) { cur_op = F1_OP; hit = 1; }
                            }
                        }
                        if (!hit && (cur_op & F2_OP)) {
                            const int f = cell(cur, 4, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe2 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(cur, 4, j - 1) - e2 == f) { cur_op = F2_OP; hit = 1; }
                            }
                        }
                        if (hit) --j;                                      // INS: path[j-1] stays -1
                    }
                    if (hit) break;
                }
                if (cur_op & M_OP) {
                    for (int k = 0; k < nin; ++k) {
                        const int4 e = ie[k];
                        const int p = e.x, ps = e.z;
                        const Row pr = unpack(w.rinfo[p]);
                        if (j - 1 < pr.beg || j - 1 > pr.end) continue;
                        if (cell(pr, 0, j - 1) + s + ps == hj) {
                            if (lane == 0) path[j - 1] = id;
                            id = p; cur = pr; --j; hit = 1; cur_op = ALL_OP;
                            break;
                        }
                    }
                }
            }
            if (!hit) { rc = ST_BACKTRACK; break; }
        }
        __syncwarp();
        return rc;
    }

    // abpoa_add_graph_edge (:480-556) for the parallel fusion: every node is the source of at most one and the
    // target of at most one edge of a read's path, so the 32 lanes update disjoint lists; only the bump cursors of
    // the two edge pools (kept in the arena: tmp[8], tmp[9]) are shared and advanced atomically.
    __device__ void add_edge_par(int from, int to, int check, int add_rid, int read_id) {
        int exist = 0, oi = -1;
        if (check) {
            int4 *ie = w.in_pool + w.in_off[to];
            for (int i = 0; i < w.in_n[to]; ++i) if (ie[i].x == from) { ie[i].y += 1; break; }
            for (int i = 0; i < w.out_n[from]; ++i) { int *e = out_entry(from, i); if (e[0] == to) { e[1] += 1; exist = 1; oi = i; break; } }
        }
        if (!exist) {
            if (w.in_n[to] == w.in_cap[to]) {
                const int nc = w.in_cap[to] ? w.in_cap[to] * 2 : 2;
                const int at = atomicAdd(&w.tmp[8], nc);
                if (at + nc > w.in_capacity) { w.tmp[1] = 1; return; }
                int4 *src = w.in_pool + w.in_off[to], *dst = w.in_pool + at;
                for (int i = 0; i < w.in_n[to]; ++i) dst[i] = src[i];
                w.in_off[to] = at; w.in_cap[to] = nc;
            }
            w.in_pool[w.in_off[to] + w.in_n[to]] = make_int4(from, 1, 0, 0);
            w.in_n[to]++;
            if (w.out_n[from] == w.out_cap[from]) {
                const int nc = w.out_cap[from] ? w.out_cap[from] * 2 : 2;
                const int at = atomicAdd(&w.tmp[9], nc);
                if (at + nc > w.out_capacity) { w.tmp[1] = 1; return; }
                int *src = w.out_pool + (size_t)w.out_off[from] * w.out_stride, *dst = w.out_pool + (size_t)at * w.out_stride;
                const int nw = w.out_n[from] * w.out_stride;
                for (int i = 0; i < nw; ++i) dst[i] = src[i];
                w.out_off[from] = at; w.out_cap[from] = nc;
            }
            oi = w.out_n[from];
            int *e = out_entry(from, oi);
            e[0] = to; e[1] = 1;
            for (int x = 0; x < 2 * w.rid_w; ++x) e[2 + x] = 0;
            w.out_n[from]++;
        }
        if (add_rid) { out_entry(from, oi)[2 + (read_id >> 5)] |= 1 << (read_id & 31); w.n_read[from] += 1; }
    }

    // abpoa_add_subgraph_alignment (:689-774) / abpoa_add_graph_sequence (:573-593) with the 32 lanes dealt over the
    // read's bases.  path[q] is the matched node of base q or -1 (inserted; every base of the first read).
    //   1. target node of every base: the matched node, its aligned twin carrying the read's base, or a NEW node;
    //      new ids are handed out in read order (ballot ranks), exactly the ids sequential creation would give.
    //   2. new nodes form chains (consecutive ids); the chain heads are spliced into the topological list and the
    //      aligned-node sets are updated by lane 0 (rare: read errors and the first read carrying an insertion).
    //   3. the edges (previous target -> target) are applied in parallel (add_edge_par).
    __device__ void add_alignment_warp(const uint8_t *seq, int seq_l, int *path, int read_id, bool first_read) {
        const int lane = threadIdx.x & 31;
        const int inc = par.sub_aln ? 0 : 1;
        const int n0 = w.n_nodes;
        if (seq_l <= 0) return;
        int *tgt = path + ((seq_l + 3) & ~3), *heads = w.s1;
        if (lane == 0) { w.tmp[8] = w.in_top; w.tmp[9] = w.out_top; w.tmp[1] = 0; }
        int n_new = 0, n_heads = 0, last_exist = beg_id, prev_t = beg_id, prev_node = beg_id;   // carries across 32-base chunks
        for (int base = 0; base < seq_l; base += 32) {
            const int q = base + lane;
            const bool in = q < seq_l;
            int node = -1, t = -1, b = 0;
            if (in) {
                b = seq[q];
                node = first_read ? -1 : path[q];
                if (node >= 0) {
                    if (w.base[node] == b) t = node;
                    else for (int i = 0; i < w.aln_n[node]; ++i) { const int a = w.aln_pool[w.aln_off[node] + i]; if (w.base[a] == b) { t = a; break; } }
                }
            }
            const bool isnew = in && t < 0;
            const unsigned mnew = __ballot_sync(0xffffffffu, isnew);
            if (isnew) t = n0 + n_new + __popc(mnew & ((1u << lane) - 1));
            // last existing node at or before this base (n_span of a new node is inherited from it)
            const unsigned mexist = __ballot_sync(0xffffffffu, in && !isnew) & ((2u << lane) - 1);
            const int src_lane = mexist ? 31 - __clz(mexist) : -1;
            const int le_t = __shfl_sync(0xffffffffu, t, src_lane < 0 ? 0 : src_lane);
            const int le = src_lane < 0 ? last_exist : le_t;
            // predecessor target in the path
            int pt = __shfl_up_sync(0xffffffffu, t, 1), pnode = __shfl_up_sync(0xffffffffu, node, 1);
            if (lane == 0) { pt = prev_t; pnode = prev_node; }
            // a new node continues the chain of the previous one only if that one is an INSERTED new node (no aligned
            // column to step over); anything else starts a chain that lane 0 splices in behind the right column
            const bool interior = isnew && node < 0 && q > 0 && pt >= n0 && pnode < 0;
            const bool head = isnew && !interior;
            const unsigned mhead = __ballot_sync(0xffffffffu, head);
            const bool fits = isnew && t < w.N;
            if (fits) {
                w.base[t] = b; w.in_n[t] = w.in_cap[t] = w.out_n[t] = w.out_cap[t] = 0; w.in_off[t] = w.out_off[t] = 0;
                w.n_read[t] = 0; w.aln_n[t] = w.aln_cap[t] = 0; w.aln_off[t] = 0;
                w.n_span[t] = w.n_span[le];
                w.next[t] = -1;
            }
            __syncwarp();
            if (fits) {
                if (head) heads[n_heads + __popc(mhead & ((1u << lane) - 1))] = q;
                else w.next[pt] = t;                                         // interior link of a chain
            }
            if (in) tgt[q] = t;
            n_new += __popc(mnew); n_heads += __popc(mhead);
            const unsigned mall = __ballot_sync(0xffffffffu, in && !isnew);
            if (mall) last_exist = __shfl_sync(0xffffffffu, t, 31 - __clz(mall));
            prev_t = __shfl_sync(0xffffffffu, t, 31); prev_node = __shfl_sync(0xffffffffu, node, 31);
        }
        if (n0 + n_new > w.N) { w.oom = 1; return; }
        w.n_nodes = n0 + n_new;
        __syncwarp();
        // 2. splice the chains into the list; aligned-node sets (lane 0; aln pool cursor stays in its registers)
        if (lane == 0) {
            for (int i = 0; i < n_heads; ++i) {
                const int q = heads[i], h = tgt[q];
                const int tail = (i + 1 < n_heads ? tgt[heads[i + 1]] : n0 + n_new) - 1;
                const int node = first_read ? -1 : path[q];
                int at;
                if (node >= 0) { at = node; add_aligned(node, h); }
                else {
                    const int last = q > 0 ? tgt[q - 1] : beg_id;
                    at = last;
                    for (;;) {                                               // end of the aligned column of `last`
                        const int nx = w.next[at];
                        if (nx < 0) break;
                        bool member = false;
                        for (int x = 0; x < w.aln_n[last]; ++x) if (w.aln_pool[w.aln_off[last] + x] == nx) { member = true; break; }
                        if (!member) break;
                        at = nx;
                    }
                }
                w.next[tail] = w.next[at]; w.next[at] = h;
            }
            if (w.oom) w.tmp[1] = 1;
        }
        __syncwarp();
        // 3. edges
        for (int q = lane; q <= seq_l; q += 32) {
            const int from = q == 0 ? beg_id : tgt[q - 1], to = q == seq_l ? end_id : tgt[q];
            const int check = (q > 0 && from >= n0) ? 0 : 1;
            const int add = (first_read || q == seq_l) ? 1 : ((from != beg_id || inc) ? 1 : 0);
            add_edge_par(from, to, check, add, read_id);
        }
        __syncwarp();
        w.in_top = w.tmp[8]; w.out_top = w.tmp[9]; w.oom = w.tmp[1];
        __syncwarp();
    }
#endif

    // after fusing a read (abpoa_topological_sort :322-357 + abpoa_update_node_n_span_reads :559-571):
    // per-node exchange sort of the edge lists, out-weight sums, edge path scores (abpoa_get_incre_path_score
    // :429-437), n_span; then the flattened topological order and the heaviest-path remain values
    // (abpoa_BFS_set_node_remain :268-309) by pointer jumping over the list / heaviest-successor links.
#ifndef LCD_EMU
    // abpoa_BFS_set_node_index (abpoa_graph.c:221-266): Kahn's order with a FIFO queue, a node entering only together with the nodes aligned
    // to it, over the out-edge lists as they stand BEFORE this round's edge sort.  Needed only by problems with partially covering reads: the
    // sub-graph of such a read and the nodes it spans are defined by ranges of this index.  The queue is the index -> node table (w.maxr),
    // w.maxl the node -> index table; both live until the next call.  The order is inherently sequential (on a chain the queue holds one node),
    // so one lane walks it -- but on records the whole warp prepared: {first out-edge target, out-degree, in-degree, aligned nodes} per node
    // in one 16-byte load (w.rinfo is free between two alignments).  A node with one out-edge into a node with one in-edge and no aligned
    // nodes -- a link of a chain, ~97 % of a region's graph -- then costs ONE dependent load (the successor's record) instead of five.
    __device__ void bfs_index() {
        const int n = w.n_nodes, lane = threadIdx.x & 31;
        int *deg = w.msa_rank, *q = w.maxr, *idx = w.maxl;
        int4 *rec = w.rinfo;
        for (int i = lane; i < n; i += 32) {
            const int on = w.out_n[i], in = w.in_n[i];
            rec[i] = make_int4(on > 0 ? out_entry(i, 0)[0] : -1, on, in, w.aln_n[i]);
            deg[i] = in; idx[i] = -1;
        }
        __syncwarp();
        if (lane == 0) {
            int qh = 0, qt = 0;
            q[qt++] = 0;
            int known = 0; int4 rc = rec[0];                   // the node at the queue's head and its record, when they are in registers already
            while (qh < qt) {
                const int cur = known >= 0 ? known : q[qh];
                if (known < 0) rc = rec[cur];
                known = -1;
                idx[cur] = qh; ++qh;
                if (cur == 1) break;
                if (rc.y == 1) {
                    const int out = rc.x; const int4 ro = rec[out];
                    if (ro.z == 1 && ro.w == 0) {                  // ready at once; nobody looks at its in-degree counter again
                        q[qt] = out;
                        if (qh == qt) { known = out; rc = ro; }
                        ++qt;
                        continue;
                    }
                }
                for (int i = 0; i < rc.y; ++i) {
                    const int out = i == 0 ? rc.x : out_entry(cur, i)[0];
                    if (--deg[out] == 0) {
                        bool ok = true;
                        for (int j = 0; j < w.aln_n[out]; ++j) if (deg[w.aln_pool[w.aln_off[out] + j]] != 0) { ok = false; break; }
                        if (!ok) continue;
                        q[qt++] = out;
                        for (int j = 0; j < w.aln_n[out]; ++j) q[qt++] = w.aln_pool[w.aln_off[out] + j];
                    }
                }
            }
        }
        __syncwarp();
    }

    // The sub-graph a partially covering read is aligned against (abpoa_subgraph_nodes, abpoa_graph.c:595-680, on the BFS index of the last
    // bfs_index()) and the alignment's view of it; the whole warp.  inc_beg / inc_end: the two anchor nodes (bases of the first read).
    //   * exc_beg / exc_end: the nodes at the outermost index reached by in-edges on the left / out-edges on the right of the anchors' range
    //     (the reference's closure loops, 32 indices of the range per step, min / max / any by warp reductions)
    //   * rows: the nodes of that index range that can be reached from exc_beg (index_map, abpoa_align_simd.c:1259-1269: a forward propagation
    //     in index order -- 32 indices per step, the dependencies inside a step resolved by OR-reductions of the lanes' target masks), kept in
    //     the order of the topological list (w.order / w.meta / w.pos are compacted in place; after_add rebuilds them after the read)
    //   * in-edges: only predecessors inside the sub-graph, compacted; the k-th one carries the path score of the node's k-th in-edge
    //     OVERALL, as the reference computes it (abpoa_get_incre_path_score is called with the index into the restricted list)
    // Results: beg_id, end_id, n_rows set (uniformly); returns the cells of the DP arena taken by the compacted lists (its top end), < 0: error.
    __device__ int prepare_sub(int inc_beg, int inc_end) {
        const int n = w.n_nodes, lane = threadIdx.x & 31;
        const int *idx = w.maxl, *at = w.maxr;
        if (inc_beg < 2 || inc_end < 2 || inc_beg >= n || inc_end >= n || idx[inc_beg] < 0 || idx[inc_end] < 0 || idx[inc_beg] > idx[inc_end]) return ST_SUBGRAPH;
        // every in-edge of the nodes at indices (up, down] comes from an index inside [min(up, b), max(down, e)]   (is_full_upstream_subgraph :595-606)
        auto full_up = [&](int up, int down, int b, int e) -> bool {
            const int mn = up < b ? up : b, mx = down > e ? down : e;
            bool bad = false;
            for (int i = up + 1 + lane; i <= down; i += 32) {
                const int id = at[i]; const int4 *ie = w.in_pool + w.in_off[id];
                for (int j = 0; j < w.in_n[id]; ++j) { const int x = idx[ie[j].x]; if (x < mn || x > mx) bad = true; }
            }
            return !__any_sync(0xffffffffu, bad);
        };
        int exc_b, exc_e;
        {
            int b = idx[inc_beg], e = idx[inc_end];
            for (;;) {                                                   // abpoa_upstream_index :608-628
                int mn = b;
                for (int i = b + lane; i <= e; i += 32) { const int id = at[i]; const int4 *ie = w.in_pool + w.in_off[id]; for (int j = 0; j < w.in_n[id]; ++j) { const int x = idx[ie[j].x]; if (x < mn) mn = x; } }
                mn = __reduce_min_sync(0xffffffffu, mn);
                if (full_up(mn, b, b, e)) { exc_b = mn; break; }
                e = b; b = mn;
            }
            b = idx[inc_beg]; e = idx[inc_end];
            for (;;) {                                                   // abpoa_downstream_index :642-662
                int mx = e;
                for (int i = b + lane; i <= e; i += 32) { const int id = at[i]; for (int j = 0; j < w.out_n[id]; ++j) { const int x = idx[out_entry(id, j)[0]]; if (x > mx) mx = x; } }
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (full_up(e, mx, b, e)) { exc_e = mx; break; }
                b = e; e = mx;
            }
        }
        if (exc_b < 0 || exc_e >= n) return ST_SUBGRAPH;
        int *in_sub = w.msa_rank;
        for (int i = lane; i < n; i += 32) in_sub[i] = 0;
        __syncwarp();
        if (lane == 0) { in_sub[at[exc_b]] = 1; in_sub[at[exc_e]] = 1; }
        __syncwarp();
        for (int c0 = exc_b; c0 < exc_e - 1; c0 += 32) {
            const int i = c0 + lane; const bool act = i < exc_e - 1;
            const int id = act ? at[i] : 0;
            bool mark = act && in_sub[id] != 0;
            unsigned tmask = 0;                                          // out-edge targets whose index lies in this step
            const int on = act ? w.out_n[id] : 0;
            for (int j = 0; j < on; ++j) { const int ti = idx[out_entry(id, j)[0]]; if (ti >= c0 && ti < c0 + 32) tmask |= 1u << (ti - c0); }
            for (;;) {
                const unsigned agg = __reduce_or_sync(0xffffffffu, mark ? tmask : 0u);
                const bool nm = mark || (act && ((agg >> lane) & 1u));
                const bool changed = nm != mark;
                mark = nm;
                if (!__any_sync(0xffffffffu, changed)) break;
            }
            if (mark) { in_sub[id] = 1; for (int j = 0; j < on; ++j) in_sub[out_entry(id, j)[0]] = 1; }
            __syncwarp();
        }
        // rows, restricted in-edge lists (from the top of the DP arena downwards), first in-edge / position tables
        int *so = w.s2, *sn = w.s3;
        int4 *pool = reinterpret_cast<int4 *>(w.dp + ((size_t)w.dp_capacity & ~(size_t)7)) - w.in_top - 8;
        int top = 0, k = 0;
        for (int o0 = 0; o0 < n; o0 += 32) {
            const int oi = o0 + lane;
            const int id = oi < n ? w.order[oi] : 0;
            const bool sel = oi < n && in_sub[id] && idx[id] >= exc_b && idx[id] <= exc_e;
            const int4 *ie = w.in_pool + w.in_off[id];
            int nf = 0;
            if (sel && idx[id] != exc_b)
                for (int j = 0; j < w.in_n[id]; ++j) { const int p = ie[j].x; if (in_sub[p] && idx[p] >= exc_b && idx[p] <= exc_e) ++nf; }
            const unsigned ms = __ballot_sync(0xffffffffu, sel);
            int pre = nf;                                                // exclusive prefix of nf over the lanes
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += y; }
            const int tot = __shfl_sync(0xffffffffu, pre, 31);
            pre -= nf;
            __syncwarp();                                                // every lane has read order[] of this step before it is overwritten
            if (sel) {
                const int my_k = k + __popc(ms & ((1u << lane) - 1)), my_top = top + pre;
                int f = 0;
                if (idx[id] != exc_b)
                    for (int j = 0; j < w.in_n[id]; ++j) { const int p = ie[j].x; if (in_sub[p] && idx[p] >= exc_b && idx[p] <= exc_e) { pool[my_top + f] = make_int4(p, ie[j].y, ie[f].z, 0); ++f; } }
                so[id] = my_top; sn[id] = nf;
                const int fp = nf ? pool[my_top].x : -1, ps0 = nf ? pool[my_top].z : 0;
                w.fp_id[id] = fp; w.fp_ps[id] = ps0; w.pos[id] = my_k; w.order[my_k] = id;
                w.meta[my_k] = make_int4(id, fp, w.base[id] | ((nf < 255 ? nf : 255) << 8) | ((ps0 & 0xff) << 16), w.remain[id]);
            }
            k += __popc(ms); top += tot;
            __syncwarp();
        }
        beg_id = at[exc_b]; end_id = at[exc_e]; n_rows = k;
        return (w.in_top + 8) * 8 + 8;
    }
#else
    __device__ void bfs_index() {}
    __device__ int prepare_sub(int, int) { return ST_SUB_UNSUPPORTED; }
#endif

    __device__ void after_add(int first_read, bool do_bfs = false, bool partial = false) {
        const int n = w.n_nodes, lane = L::tid();
        const int inc = par.sub_aln ? 0 : 1;
#ifndef LCD_EMU
        if (do_bfs) { if constexpr (L::STRIP) bfs_index(); L::sync(); }
#endif
        const int ib = partial ? w.maxl[beg_id] : 0, ie_ = partial ? w.maxl[end_id] : 0;
        int *jn_a = w.s1, *jn_b = w.s2, *dn_a = w.s3, *dn_b = w.s4;       // list links / distance to list end
        int *jh_a = w.s5, *jh_b = w.s6, *dh_a = w.remain, *dh_b = w.s7;   // heaviest-successor links / hops to SINK
        for (int i = lane; i < n; i += L::NT) {
            int4 *ie = w.in_pool + w.in_off[i];
            const int nin = w.in_n[i];
            for (int j = 0; j < nin - 1; ++j) for (int k = j + 1; k < nin; ++k) if (ie[j].y < ie[k].y) { const int4 t = ie[j]; ie[j] = ie[k]; ie[k] = t; }
            const int nout = w.out_n[i], S = w.out_stride;
            int ws = 0;
            for (int j = 0; j < nout - 1; ++j) for (int k = j + 1; k < nout; ++k) {
                int *a = out_entry(i, j), *b = out_entry(i, k);
                if (a[1] < b[1]) for (int x = 0; x < S; ++x) { const int t = a[x]; a[x] = b[x]; b[x] = t; }
            }
            for (int j = 0; j < nout; ++j) ws += out_entry(i, j)[1];
            w.wsum[i] = ws;
            // abpoa_update_node_n_span_reads :559-571: the nodes whose index lies between the alignment's first and last node
            if (partial ? ((w.maxl[i] > ib && w.maxl[i] < ie_) || (inc && (i == beg_id || i == end_id))) : (first_read || inc || i >= 2)) w.n_span[i] += 1;
            const int nx = w.next[i];
            jn_a[i] = nx < 0 ? i : nx; dn_a[i] = nx < 0 ? 0 : 1;
            // heaviest out edge = first entry after the descending exchange sort
            const int hv = (i == 1 || nout == 0) ? i : out_entry(i, 0)[0];
            jh_a[i] = hv; dh_a[i] = hv == i ? 0 : 1;
        }
        L::sync();
        for (int i = lane; i < n; i += L::NT) {
            int4 *ie = w.in_pool + w.in_off[i];
            for (int k = 0; k < w.in_n[i]; ++k) {
                const int node_w = w.wsum[ie[k].x], edge_w = ie[k].y;
                int ps = 0;
                if (node_w != 0 && edge_w != 0 && node_w != edge_w) { ps = (int)round(log((double)edge_w / (double)node_w)); if (ps < -20) ps = -20; }
                ie[k].z = ps;
            }
            w.fp_id[i] = w.in_n[i] > 0 ? ie[0].x : -1;
            w.fp_ps[i] = w.in_n[i] > 0 ? ie[0].z : 0;
        }
        if (L::NT == 1) {
            // one thread: walk the list once and follow the heaviest-successor links backwards
            int cnt = 0;
            for (int id = 0; id != -1; id = w.next[id]) { w.pos[id] = cnt; w.order[cnt++] = id; }
            w.remain[1] = -1;
            for (int i = cnt - 1; i >= 0; --i) { const int id = w.order[i]; if (id != 1) w.remain[id] = (jh_a[id] == id ? -1 : w.remain[jh_a[id]]) + 1; }
            return;
        }
        for (int span = 1; span < n; span <<= 1) {
            // 4 nodes per lane and step: the two dependent gathers of a node overlap with those of the other three
            for (int i0 = lane; i0 < n; i0 += 4 * L::NT) {
                int a[4], h[4], da[4], dh[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + u * L::NT; if (i < n) { a[u] = jn_a[i]; h[u] = jh_a[i]; da[u] = dn_a[i]; dh[u] = dh_a[i]; } }
                int a2[4], h2[4], da2[4], dh2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + u * L::NT; if (i < n) { da2[u] = dn_a[a[u]]; a2[u] = jn_a[a[u]]; dh2[u] = dh_a[h[u]]; h2[u] = jh_a[h[u]]; } }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + u * L::NT; if (i < n) { dn_b[i] = da[u] + da2[u]; jn_b[i] = a2[u]; dh_b[i] = dh[u] + dh2[u]; jh_b[i] = h2[u]; } }
            }
            L::sync();
            int *t;
            t = jn_a; jn_a = jn_b; jn_b = t;  t = dn_a; dn_a = dn_b; dn_b = t;
            t = jh_a; jh_a = jh_b; jh_b = t;  t = dh_a; dh_a = dh_b; dh_b = t;
        }
        for (int i = lane; i < n; i += L::NT) {
            const int at = n - 1 - dn_a[i], rem = dh_a[i] - 1;
            w.order[at] = i; w.pos[i] = at;
            w.remain[i] = rem;                     // remain[SINK] = -1, remain[x] = remain[heaviest successor] + 1
            // everything the DP needs to know about the row at this position of the order, in one 16-byte record
            const int nin = w.in_n[i];
            w.meta[at] = make_int4(i, w.fp_id[i], w.base[i] | ((nin < 255 ? nin : 255) << 8) | ((w.fp_ps[i] & 0xff) << 16), rem);
        }
        L::sync();
    }

    // ---- DP ---------------------------------------------------------------------------------
    // row descriptor: {offset of the row's planes in the DP arena (cells), dp_beg | dp_end << 16,
    //                  (left_max_i + 1) | (right_max_i + 1) << 16, -}
    struct Row { int off, beg, end, nv, left1, right1; };
    __device__ static __forceinline__ Row unpack(const int4 r) {
        Row x; x.off = r.x; x.beg = r.y & 0xffff; x.end = (r.y >> 16) & 0xffff; x.nv = (x.end >> 5) - (x.beg >> 5) + 2;
        x.left1 = r.z & 0xffff; x.right1 = (r.z >> 16) & 0xffff; return x;
    }
    __device__ static __forceinline__ int4 pack(int off, int beg, int end, int left, int right) {
        return make_int4(off, beg | (end << 16), (left + 1) | ((right + 1) << 16), 0);
    }
    // plane p (0 H, 1 E1, 2 E2, 3 F1, 4 F2) of a row, addressed by absolute column
    __device__ const int16_t *plane(const Row &r, int p) const { return w.dp + r.off + (size_t)p * r.nv * PN - (size_t)(r.beg >> 5) * PN; }
    __device__ int cell(const Row &r, int p, int col) const {       // bounds-checked scalar read (backtrack)
        const int lo = (r.beg >> 5) * PN, hi = lo + r.nv * PN;
        if (col < lo || col >= hi) return GARBAGE;
        return plane(r, p)[col];
    }
    // SIMD_SET_F, abpoa_align_simd.c:691-725
    __device__ __forceinline__ vec set_f(vec F, int set_num, int e) const {
        if (set_num == PN) {
#pragma unroll
            for (int s = 0; s < LOGN; ++s)
                F = L::vmax(F, L::shift_up(L::sub(F, L::set1((int16_t)(e << s))), 1 << s, inf_min));
            return F;
        }
        int cov = set_num;
#pragma unroll
        for (int s = 0; s < LOGN; ++s) {
            const int sh = 1 << s;
            if (s > 0) cov += sh;
            vec t = L::shift_up(L::sub(F, L::set1((int16_t)(e << s))), sh, inf_min);
            t = L::keep(t, 0, cov < PN - 1 ? cov : PN - 1, inf_min);
            F = L::vmax(F, t);
        }
        return F;
    }

    struct Pred { const int16_t *ph, *pe1, *pe2; int bsn, esn_m, esn_e, ps, from_mem; };
    __device__ __forceinline__ Pred make_pred(const int4 e, const Row &r, int beg_sn, int end_sn, int dp_sn, const int16_t *cache = nullptr) const {
        Pred p;
        if (cache) { p.ph = cache - (size_t)(r.beg >> 5) * PN; p.pe1 = p.ph + NVC * PN; p.pe2 = p.pe1 + NVC * PN; }
        else { p.ph = plane(r, 0); p.pe1 = plane(r, 1); p.pe2 = plane(r, 2); }
        p.ps = e.z;
        const int pre_beg_sn = r.beg >> 5, pre_end_sn = r.end >> 5;
        p.from_mem = pre_beg_sn < beg_sn;
        p.bsn = p.from_mem ? beg_sn : pre_beg_sn;
        int esn = (r.end + 1) >> 5; if (esn > end_sn) esn = end_sn; if (esn > dp_sn - 1) esn = dp_sn - 1;
        p.esn_m = esn;
        p.esn_e = pre_end_sn < end_sn ? pre_end_sn : end_sn;
        return p;
    }
    __device__ __forceinline__ void pred_accumulate(const Pred &p, int k, int sn, int col0, vec &h, vec &ve1, vec &ve2) const {
        if (sn >= p.bsn && sn <= p.esn_m) {
            const int first = (sn == p.bsn && !p.from_mem) ? inf_min : (int)p.ph[col0 - 1];
            const vec v = L::add(L::load_m1(p.ph + col0, first), L::set1(p.ps));
            h = k == 0 ? v : L::vmax(v, h);
        }
        if (sn >= p.bsn && sn <= p.esn_e) {
            const vec v1 = L::add(L::load(p.pe1 + col0), L::set1(p.ps)), v2 = L::add(L::load(p.pe2 + col0), L::set1(p.ps));
            ve1 = k == 0 ? v1 : L::vmax(v1, ve1);
            ve2 = k == 0 ? v2 : L::vmax(v2, ve2);
        }
    }

#ifndef LCD_EMU
    // ---- strip rows (warp policy) -------------------------------------------------------------------
    // The vectors sn in [beg_sn, sn_hi] of a row whose F scan is the plain prefix maximum (set_num == PN: every
    // vector up to the predecessors' last one) are computed in ONE pass: lane l owns the C consecutive columns
    // c0 + l*C .. c0 + l*C + C-1 (c0 = beg_sn * 32), walks them sequentially in registers, and the F carries of
    // the 32 strips are resolved by one max-plus warp scan per gap model instead of one 5-step scan per
    // 32-column vector.  Cell values are exactly those of the per-vector code below (same formulas per cell; the
    // F recurrence F[j] = max(Hm[j-1] - oe, F[j-1] - e) is associative, and no in-band value wraps in int16).
    // Vectors beyond the predecessors' last one keep the reference's 2-then-1 lane propagation: they are left to
    // the per-vector code, which continues from the carries (first1, first2) returned here.
    template <int C> __device__ static __forceinline__ void load_strip(const int16_t *p, int (&v)[C]) {
        if (C == 8) { const uint4 t = *reinterpret_cast<const uint4 *>(p);
            v[0] = (int16_t)(t.x & 0xffff); v[1] = (int)t.x >> 16; v[2 % C] = (int16_t)(t.y & 0xffff); v[3 % C] = (int)t.y >> 16;
            v[4 % C] = (int16_t)(t.z & 0xffff); v[5 % C] = (int)t.z >> 16; v[6 % C] = (int16_t)(t.w & 0xffff); v[7 % C] = (int)t.w >> 16; }
        else if (C == 4) { const uint2 t = *reinterpret_cast<const uint2 *>(p);
            v[0] = (int16_t)(t.x & 0xffff); v[1] = (int)t.x >> 16; v[2 % C] = (int16_t)(t.y & 0xffff); v[3 % C] = (int)t.y >> 16; }
        else { const uint32_t t = *reinterpret_cast<const uint32_t *>(p); v[0] = (int16_t)(t & 0xffff); v[1] = (int)t >> 16; }
    }
    template <int C> __device__ static __forceinline__ void store_strip(int16_t *p, const int (&v)[C]) {
        auto pk = [](int a, int b) -> uint32_t { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); };
        if (C == 8) *reinterpret_cast<uint4 *>(p) = make_uint4(pk(v[0], v[1]), pk(v[2 % C], v[3 % C]), pk(v[4 % C], v[5 % C]), pk(v[6 % C], v[7 % C]));
        else if (C == 4) *reinterpret_cast<uint2 *>(p) = make_uint2(pk(v[0], v[1]), pk(v[2 % C], v[3 % C]));
        else *reinterpret_cast<uint32_t *>(p) = pk(v[0], v[1]);
    }
    struct StripArgs {
        int beg, end, beg_sn, end_sn, sn_hi, nb;
        int16_t *H, *E1, *E2, *F1, *F2, *cwH;       // plane pointers addressed by absolute column; cwH: smem copy or nullptr
        bool banded;
    };
    // Everything of a strip row after the predecessor maxima: h = max_p(H_p[j-1] + ps), ve1/ve2 = max_p(E*_p[j] + ps).
    template <int C> __device__ __forceinline__ void strip_core(const StripArgs &a, int (&h)[C], int (&ve1)[C], int (&ve2)[C],
                                                                 int &first1, int &first2, int &mx, int &left, int &right) {
        const int lane = threadIdx.x & 31;
        const int c0 = a.beg_sn * PN, width = (a.sn_hi - a.beg_sn + 1) * PN;
        const int j0 = c0 + lane * C;
        const bool act = lane * C < width;
        const int e1 = par.gap_ext1, e2 = par.gap_ext2, o1 = par.gap_open1, o2 = par.gap_open2;
        // + query profile (qs[j] = query[j-1], sentinel 7 outside the read), band mask, Hm = max(M + q, E1, E2)
        const int s_eq = a.nb > 3 ? 0 : par.match, s_ne = a.nb > 3 ? 0 : -par.mismatch;
        const int lo = a.beg - j0, span = a.end - a.beg;                   // cell i is inside the band iff 0 <= i - lo <= span
        int hm[C];
        {
            uint32_t qw[2];
            if (C == 8) { const uint2 t = *reinterpret_cast<const uint2 *>(w.qs + j0); qw[0] = t.x; qw[1] = t.y; }
            else if (C == 4) { qw[0] = *reinterpret_cast<const uint32_t *>(w.qs + j0); qw[1] = 0; }
            else { qw[0] = *reinterpret_cast<const uint16_t *>(w.qs + j0); qw[1] = 0; }
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int qb = (int)((qw[i >> 2] >> (8 * (i & 3))) & 0xff);
                const int sc = qb > 3 ? 0 : (qb == a.nb ? s_eq : s_ne);
                const bool inb = (unsigned)(i - lo) <= (unsigned)span;
                const int x = inb ? h[i] + sc : inf_min;
                if (!inb) { ve1[i] = inf_min; ve2[i] = inf_min; }
                h[i] = x;
                int y = x > ve1[i] ? x : ve1[i]; y = y > ve2[i] ? y : ve2[i];
                hm[i] = y;
            }
        }
        const int row_first = __shfl_sync(0xffffffffu, h[0], 0);          // (M + q) of the row's first stored column
        // F: local pass without carry-in, then a max-plus scan of the strip aggregates
        const int lowf = inf_min - 20000;
        int hl = __shfl_up_sync(0xffffffffu, hm[C - 1], 1);
        if (lane == 0) hl = row_first;
        int f1[C], f2[C];
        {
            int g1 = lowf, g2 = lowf, prev = hl;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int a1 = prev - oe1, a2 = prev - oe2;
                g1 = g1 - e1 > a1 ? g1 - e1 : a1;
                g2 = g2 - e2 > a2 ? g2 - e2 : a2;
                f1[i] = g1; f2[i] = g2;
                prev = hm[i];
            }
            int s1 = g1, s2 = g2;                                            // F at the strip's last column without carry-in
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y1 = __shfl_up_sync(0xffffffffu, s1, d) - e1 * C * d, y2 = __shfl_up_sync(0xffffffffu, s2, d) - e2 * C * d;
                if (lane >= d) { s1 = s1 > y1 ? s1 : y1; s2 = s2 > y2 ? s2 : y2; }
            }
            int c1 = __shfl_up_sync(0xffffffffu, s1, 1), c2 = __shfl_up_sync(0xffffffffu, s2, 1);   // F at the column left of the strip
            if (lane == 0) { c1 = lowf; c2 = lowf; }
#pragma unroll
            for (int i = 0; i < C; ++i) {
                c1 -= e1; c2 -= e2;
                f1[i] = f1[i] > c1 ? f1[i] : c1;
                f2[i] = f2[i] > c2 ? f2[i] : c2;
            }
        }
        // carries for the vectors after sn_hi: lane 31 of max(Hm, F + o) in vector sn_hi
        if (a.sn_hi < a.end_sn) {
            const int last_lane = width / C - 1;
            const int x1 = hm[C - 1] > f1[C - 1] + o1 ? hm[C - 1] : f1[C - 1] + o1;
            const int x2 = hm[C - 1] > f2[C - 1] + o2 ? hm[C - 1] : f2[C - 1] + o2;
            first1 = __shfl_sync(0xffffffffu, x1, last_lane);
            first2 = __shfl_sync(0xffffffffu, x2, last_lane);
        }
        // H, stored E, row maximum.  Cells right of the band in the row's last vector are reset to INF_MIN.
        const int hi_m = ((j0 >> 5) == a.end_sn) ? a.end - j0 : C;          // cells i > hi_m are reset
        int lm = INT32_MIN, lf = 0, ll = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            int x = hm[i]; x = x > f1[i] ? x : f1[i]; x = x > f2[i] ? x : f2[i];
            if (i > hi_m) { x = inf_min; ve1[i] = inf_min; ve2[i] = inf_min; }
            h[i] = x;
            const int y1 = ve1[i] - e1, y2 = ve2[i] - e2;
            ve1[i] = y1 > x - oe1 ? y1 : x - oe1;
            ve2[i] = y2 > x - oe2 ? y2 : x - oe2;
            if ((unsigned)(i - lo) <= (unsigned)span) {
                if (x > lm) { lm = x; lf = i; ll = i; } else if (x == lm) ll = i;
            }
        }
        if (act) {
            store_strip<C>(a.H + j0, h); store_strip<C>(a.E1 + j0, ve1); store_strip<C>(a.E2 + j0, ve2);
            store_strip<C>(a.F1 + j0, f1); store_strip<C>(a.F2 + j0, f2);
            if (a.cwH) { store_strip<C>(a.cwH + j0, h); store_strip<C>(a.cwH + NVC * PN + j0, ve1); store_strip<C>(a.cwH + 2 * NVC * PN + j0, ve2); }
        }
        if (a.banded) {
            if (!act) lm = INT32_MIN;
            const int m = __reduce_max_sync(0xffffffffu, lm);
            if (m != INT32_MIN) {
                const int fi = __reduce_min_sync(0xffffffffu, lm == m ? j0 + lf : INT32_MAX);
                const int la = __reduce_max_sync(0xffffffffu, lm == m ? j0 + ll : -1);
                if (m > mx) { mx = m; left = fi; right = la; }
                else if (m == mx) right = la;
            }
        }
    }
    // general strip row: up to 4 predecessors anywhere in the arena
    template <int C> __device__ __forceinline__ void strip_row(const StripArgs &a, int nin, const Pred (&pd)[4],
                                                                 int &first1, int &first2, int &mx, int &left, int &right) {
        const int lane = threadIdx.x & 31;
        const int c0 = a.beg_sn * PN, width = (a.sn_hi - a.beg_sn + 1) * PN;
        const int j0 = c0 + lane * C, sn = j0 >> 5;
        const bool act = lane * C < width;
        const bool vfirst = (j0 & 31) == 0;
        int h[C], ve1[C], ve2[C];
#pragma unroll
        for (int i = 0; i < C; ++i) { h[i] = inf_min; ve1[i] = inf_min; ve2[i] = inf_min; }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < nin) {
            const Pred &p = pd[k];
            if (act && sn >= p.bsn && sn <= p.esn_m) {
                int v[C];
                load_strip<C>(p.ph + j0, v);
                const int lft = (vfirst && sn == p.bsn && !p.from_mem) ? inf_min : (int)p.ph[j0 - 1];
#pragma unroll
                for (int i = C - 1; i > 0; --i) v[i] = v[i - 1];
                v[0] = lft;
#pragma unroll
                for (int i = 0; i < C; ++i) { const int x = v[i] + p.ps; h[i] = k == 0 ? x : (x > h[i] ? x : h[i]); }
            }
            if (act && sn >= p.bsn && sn <= p.esn_e) {
                int v1[C], v2[C];
                load_strip<C>(p.pe1 + j0, v1); load_strip<C>(p.pe2 + j0, v2);
#pragma unroll
                for (int i = 0; i < C; ++i) {
                    const int x1 = v1[i] + p.ps, x2 = v2[i] + p.ps;
                    ve1[i] = k == 0 ? x1 : (x1 > ve1[i] ? x1 : ve1[i]);
                    ve2[i] = k == 0 ? x2 : (x2 > ve2[i] ? x2 : ve2[i]);
                }
            }
        }
        strip_core<C>(a, h, ve1, ve2, first1, first2, mx, left, right);
    }
    // chain row: the only predecessor is the row computed just before, whose H / E1 / E2 band sits in shared memory
    // (crd, first vector = pre_beg_sn) and covers every vector of this row
    template <int C> __device__ __forceinline__ void chain_row(const StripArgs &a, const int16_t *crd, int pre_beg_sn, int ps,
                                                                 int &first1, int &first2, int &mx, int &left, int &right) {
        const int lane = threadIdx.x & 31;
        const int j0 = a.beg_sn * PN + lane * C;
        const bool act = lane * C < (a.sn_hi - a.beg_sn + 1) * PN;
        const int16_t *ph = crd + (j0 - pre_beg_sn * PN);
        int h[C], ve1[C], ve2[C];
        if (act) { load_strip<C>(ph, h); load_strip<C>(ph + NVC * PN, ve1); load_strip<C>(ph + 2 * NVC * PN, ve2); }
        else {
#pragma unroll
            for (int i = 0; i < C; ++i) { h[i] = inf_min; ve1[i] = inf_min; ve2[i] = inf_min; }
        }
        int lft = __shfl_up_sync(0xffffffffu, h[C - 1], 1);
        if (lane == 0) lft = pre_beg_sn < a.beg_sn ? (int)ph[-1] : inf_min;
#pragma unroll
        for (int i = C - 1; i > 0; --i) h[i] = h[i - 1] + ps;
        h[0] = lft + ps;
#pragma unroll
        for (int i = 0; i < C; ++i) { ve1[i] += ps; ve2[i] += ps; }
        strip_core<C>(a, h, ve1, ve2, first1, first2, mx, left, right);
    }

    // ---- packed chain segments (warp policy) ---------------------------------------------------------
    // A run of chain rows -- one in-edge, from the row computed just before, band not reaching past the last vector of
    // that row (~95 % of all rows of a region's graph) -- is computed without touching memory on its critical path:
    //   * the previous row's H / E1 / E2 stay in REGISTERS, two int16 cells per 32-bit register, lane l owning the C
    //     consecutive columns c0 + l*C .. (c0 = 32 * the row's first vector); when the band's first vector moves right
    //     the registers are shifted across lanes;
    //   * cell arithmetic is the wrapping int16x2 SIMD-in-word set of sm_100a (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2,
    //     VIADDMNMX.S16x2 -- what _mm512_add_epi16 / _mm512_max_epi16 do per lane in the reference);
    //   * the two gap models' F values of one column share a register (F1 low half, F2 high half), so the strip-local
    //     recurrence and the max-plus warp scan of the strip aggregates cost one instruction stream for both;
    //   * the node's row of the query profile (qp) is one 2C-byte load; the row's meta data (node, in-edge, base, path
    //     score, remain) one 16-byte load fetched a row ahead.
    // Results (all five planes of every row, the row descriptor) go to the same HBM layout the general path and the
    // backtrack read.  Cell values are those of chain_row / strip_core (same formulas; see the notes there).
#if !defined(LCD_EMU) && !defined(LCD_SIMT_EMU)
    __device__ __forceinline__ void ring_issue(int k) {          // one lane: block k (rows 32k .. 32k+31) into half k & 1
        const int h = k & 1, rows = ring_n - 32 * k < 32 ? ring_n - 32 * k : 32;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(ring_bar + h), dst = (unsigned)__cvta_generic_to_shared(ring + 32 * h), bytes = 16u * (unsigned)rows;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(w.meta + 32 * k), "r"(bytes), "r"(bar) : "memory");
    }
    __device__ __forceinline__ void ring_wait(int h) {           // all lanes: the oldest outstanding fill of half h
        const unsigned bar = (unsigned)__cvta_generic_to_shared(ring_bar + h), parity = ring_seen[h] & 1u;
        unsigned ok = 0; int spins = 0;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (!ok && ++spins > (1 << 24)) __trap();            // a copy that never lands must not hang the GPU
        } while (!ok);
        ring_seen[h]++;
    }
    __device__ __forceinline__ void ring_start(int n) {           // at the start of an alignment: the order's descriptors were just (re)written
        if (!ring) return;
        ring_n = n; ring_block = -1;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");     // the warp's generic-proxy stores to w.meta before the async-proxy reads
            ring_issue(0); if (n > 32) ring_issue(1);
        }
        ring_fill[0]++; if (n > 32) ring_fill[1]++;
        ring_issued = n > 32 ? 1 : 0;
    }
    __device__ __forceinline__ int4 ring_get(int oi) {            // all lanes, rows are asked for in increasing order
        if (!ring) return w.meta[oi];
        const int k = oi >> 5;
        if (k != ring_block) {
            ring_wait(k & 1);
            ring_block = k;
            if (k + 1 > ring_issued && 32 * (k + 1) < ring_n) {   // the other half held block k - 1: consumed
                __syncwarp();
                if ((threadIdx.x & 31) == 0) ring_issue(k + 1);
                ring_fill[(k + 1) & 1]++; ring_issued = k + 1;
            }
        }
        return ring[32 * (k & 1) + (oi & 31)];
    }
    __device__ __forceinline__ void ring_drain() {                // fills issued but never read (the alignment ended early)
        if (!ring) return;
        for (int h = 0; h < 2; ++h) while (ring_seen[h] < ring_fill[h]) ring_wait(h);
    }
#else
    __device__ __forceinline__ void ring_start(int) {}
    __device__ __forceinline__ int4 ring_get(int oi) { return w.meta[oi]; }
    __device__ __forceinline__ void ring_drain() {}
#endif
    struct ChainState { int last_id; Row last_row; bool last_cached; int cache_buf; uint32_t dp_top; };
    __device__ static __forceinline__ uint32_t pk2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
    __device__ static __forceinline__ uint32_t dup2(int x) { return ((uint32_t)x & 0xffffu) * 0x10001u; }
    template <int C> __device__ static __forceinline__ void ldp(const int16_t *p, uint32_t (&v)[C / 2]) {
        if constexpr (C == 8) { const uint4 t = *reinterpret_cast<const uint4 *>(p); v[0] = t.x; v[1] = t.y; v[2 % (C / 2)] = t.z; v[3 % (C / 2)] = t.w; }
        else if constexpr (C == 4) { const uint2 t = *reinterpret_cast<const uint2 *>(p); v[0] = t.x; v[1 % (C / 2)] = t.y; }
        else v[0] = *reinterpret_cast<const uint32_t *>(p);
    }
    template <int C> __device__ static __forceinline__ void stp(int16_t *p, const uint32_t (&v)[C / 2]) {
        if constexpr (C == 8) *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2 % (C / 2)], v[3 % (C / 2)]);
        else if constexpr (C == 4) *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1 % (C / 2)]);
        else *reinterpret_cast<uint32_t *>(p) = v[0];
    }
    // band of a chain row below the row `pr` (GET_AD_DP_BEGIN / END, abpoa_align.h:34-35, abpoa_align_simd.c:946-960)
    __device__ __forceinline__ void chain_band(const Row &pr, int rem, int rem_end, int qlen, int n, int wband, bool banded,
                                               int &beg, int &end, int &beg_sn) const {
        beg = 0; end = qlen; beg_sn = 0;
        if (banded) {
            const int rr = qlen - (rem - rem_end - 1);
            const int maxl = pr.left1 < n ? pr.left1 : n, maxr = pr.right1 > 0 ? pr.right1 : 0;
            beg = (maxl < rr ? maxl : rr) - wband; if (beg < 0) beg = 0;
            end = (maxr > rr ? maxr : rr) + wband; if (end > qlen) end = qlen;
            beg_sn = beg >> 5;
            if (beg_sn < (pr.beg >> 5)) { beg = pr.beg; beg_sn = pr.beg >> 5; }
        }
    }
    // Computes rows oi, oi + 1, ... while they are chain rows that fit C columns per lane; returns the first position not taken.
    // first_sn: first vector of row oi's band (the register window starts there).
    template <int C> __device__ int chain_segment(int oi, const int first_sn, const int n, const int n_tot, const int qlen, const int dp_sn, const int wband, const bool banded,
                                                  const int rem_end, ChainState &st, int &err) {
        constexpr int P = C / 2;
        const int lane = threadIdx.x & 31;
        const uint32_t INF2 = dup2(inf_min), MIN2 = 0x80008000u, LOW2 = dup2(-32600);
        const int e1 = par.gap_ext1, e2 = par.gap_ext2;
        const uint32_t NE12 = pk2(-e1, -e2), NOE12 = pk2(-oe1, -oe2);
        const uint32_t NE1 = dup2(-e1), NE2 = dup2(-e2), NOE1 = dup2(-oe1), NOE2 = dup2(-oe2);
        const int rel = lane * C;                        // this lane's first column, relative to the row's first vector
        int c0sn = first_sn, pend_sn = st.last_row.end >> 5;
        // the previous row, in a window of C vectors starting at the first vector of row oi: from the on-chip cache a
        // general row left, else from its planes in the arena
        uint32_t hp[P], e1p[P], e2p[P], lh_next = INF2;
        {
            const int16_t *src; int stride;
            if (st.last_cached) { src = row_cache + st.cache_buf * 3 * NVC * PN; stride = NVC * PN; }
            else { src = w.dp + st.last_row.off; stride = st.last_row.nv * PN; }
            const int shift = (c0sn - (st.last_row.beg >> 5)) * PN;            // >= 0: a row's band never starts left of its predecessor's first vector
            src += shift;
            if (c0sn + (rel >> 5) <= pend_sn) { ldp<C>(src + rel, hp); ldp<C>(src + stride + rel, e1p); ldp<C>(src + 2 * stride + rel, e2p); }
            else {
#pragma unroll
                for (int k = 0; k < P; ++k) { hp[k] = INF2; e1p[k] = INF2; e2p[k] = INF2; }
            }
            if (shift > 0) lh_next = (uint32_t)(uint16_t)src[-1] << 16;
        }
        Row lr = st.last_row; int last_id = st.last_id;
        uint32_t dp_top = st.dp_top;
        int4 mt = ring_get(oi);
        int n_done = 0, narrow = 0;
        for (;;) {
            const int id = mt.x, nb = mt.z & 0xff, ps = (int)(int8_t)((mt.z >> 16) & 0xff);
            if (id == end_id || ((mt.z >> 8) & 0xff) != 1 || mt.y != last_id) break;
            int beg, end, beg_sn;
            chain_band(lr, mt.w, rem_end, qlen, n_tot, wband, banded, beg, end, beg_sn);
            const int end_sn = end >> 5, nvd = end_sn - beg_sn + 1, nv = nvd + 1;
            const int dl = (beg_sn - c0sn) * (32 / C);
            // a band reaching ONE vector past the predecessor's last one (every ~32nd row of a diagonal) is taken too: that
            // vector has the reference's truncated F propagation and is computed lane-per-column below
            const bool extra = end_sn == pend_sn + 1;
            if (end_sn > pend_sn + 1 || beg_sn > pend_sn || nvd > C || dl >= 32) break;
            if (C > 2) { narrow = 2 * nvd <= C ? narrow + 1 : 0; if (narrow >= 48) break; }       // a band that got narrower: re-enter with fewer columns per lane
            const uint32_t need = (uint32_t)nv * PN * 5;
            if (dp_top + need > w.dp_capacity) { err = ST_OOM; return oi; }
            const uint32_t off = dp_top; dp_top += need;
            cells += (unsigned long long)(end - beg + 1);
            if (oi + 1 < n) mt = ring_get(oi + 1);
            // the band's first vector moved right: shift the previous row across the lanes
            uint32_t lh = lh_next;                       // H of the previous row at column c0 - 1 (high half)
            lh_next = INF2;
            if (dl > 0) {
                lh = __shfl_sync(0xffffffffu, hp[P - 1], dl - 1);
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const uint32_t a = __shfl_down_sync(0xffffffffu, hp[k], dl), b = __shfl_down_sync(0xffffffffu, e1p[k], dl), c = __shfl_down_sync(0xffffffffu, e2p[k], dl);
                    const bool in = lane + dl < 32;
                    hp[k] = in ? a : INF2; e1p[k] = in ? b : INF2; e2p[k] = in ? c : INF2;
                }
                c0sn = beg_sn;
            }
            const int j0 = c0sn * PN + rel;
            // M candidates: H of the previous row one column to the left, + path score + query profile
            uint32_t q[P];
            ldp<C>(w.qp + (size_t)(nb > 3 ? 4 : nb) * w.qp_stride + j0, q);
            uint32_t up = __shfl_up_sync(0xffffffffu, hp[P - 1], 1);
            if (lane == 0) up = lh;
            uint32_t t[P], x1[P], x2[P];
#pragma unroll
            for (int k = 0; k < P; ++k) { t[k] = __vadd2(__byte_perm(k == 0 ? up : hp[k > 0 ? k - 1 : 0], hp[k], 0x5432), q[k]); x1[k] = e1p[k]; x2[k] = e2p[k]; }
            if (ps != 0) {
                const uint32_t ps2 = dup2(ps);
#pragma unroll
                for (int k = 0; k < P; ++k) { t[k] = __vadd2(t[k], ps2); x1[k] = __vadd2(x1[k], ps2); x2[k] = __vadd2(x2[k], ps2); }
            }
            // band mask, Hm = max(M + q, E1, E2)
            const int lo = beg - j0, span = end - beg;
            uint32_t m[P], hm[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const bool a = (unsigned)(2 * k - lo) <= (unsigned)span, b = (unsigned)(2 * k + 1 - lo) <= (unsigned)span;
                m[k] = (a ? 0xffffu : 0u) | (b ? 0xffff0000u : 0u);
                t[k] = (t[k] & m[k]) | (INF2 & ~m[k]); x1[k] = (x1[k] & m[k]) | (INF2 & ~m[k]); x2[k] = (x2[k] & m[k]) | (INF2 & ~m[k]);
                hm[k] = __vimax3_s16x2(t[k], x1[k], x2[k]);
            }
            // F: strip-local recurrence (F1 | F2 packed per column), max-plus scan of the strip aggregates, carry-in
            const uint32_t rf = __shfl_sync(0xffffffffu, t[0], 0);             // (M + q) of the row's first stored column
            uint32_t hl = __shfl_up_sync(0xffffffffu, hm[P - 1], 1);
            uint32_t prev2 = lane == 0 ? __byte_perm(rf, 0, 0x1010) : __byte_perm(hl, 0, 0x3232);
            uint32_t f[C], g = 0;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const uint32_t a2 = __vadd2(prev2, NOE12);
                g = i == 0 ? a2 : __viaddmax_s16x2(g, NE12, a2);
                f[i] = g;
                prev2 = __byte_perm(hm[i >> 1], 0, (i & 1) ? 0x3232 : 0x1010);
            }
            {
                uint32_t sg = g, dec = pk2(-e1 * C, -e2 * C);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __vadd2(__shfl_up_sync(0xffffffffu, sg, d), dec);
                    if (lane >= d) sg = __vmaxs2(sg, y);
                    dec = __vadd2(dec, dec);
                }
                uint32_t c = __shfl_up_sync(0xffffffffu, sg, 1);
                if (lane == 0) c = LOW2;
#pragma unroll
                for (int i = 0; i < C; ++i) { c = __vadd2(c, NE12); f[i] = __vmaxs2(f[i], c); }
            }
            // H, stored E, in-band row maximum
            uint32_t hn[P], f1[P], f2[P], hb[P], lm2 = MIN2;
            const bool in_strip = c0sn + (rel >> 5) <= pend_sn;          // false only for the lanes of an extra vector
            if (!in_strip) {
#pragma unroll
                for (int k = 0; k < P; ++k) m[k] = 0;
            }
#pragma unroll
            for (int k = 0; k < P; ++k) {
                f1[k] = __byte_perm(f[2 * k], f[2 * k + 1], 0x5410); f2[k] = __byte_perm(f[2 * k], f[2 * k + 1], 0x7632);
                uint32_t h = __vimax3_s16x2(hm[k], f1[k], f2[k]);
                h = (h & m[k]) | (INF2 & ~m[k]);
                hn[k] = h;
                x1[k] = __viaddmax_s16x2(x1[k], NE1, __vadd2(h, NOE1));
                x2[k] = __viaddmax_s16x2(x2[k], NE2, __vadd2(h, NOE2));
                hb[k] = (h & m[k]) | (MIN2 & ~m[k]);
                lm2 = __vmaxs2(lm2, hb[k]);
            }
            int left = -1, right = -1, mxrow = inf_min;
            if (banded) {
                const int lo16 = (int)(int16_t)(lm2 & 0xffffu), hi16 = (int)lm2 >> 16;
                const int mx = __reduce_max_sync(0xffffffffu, lo16 > hi16 ? lo16 : hi16);
                const uint32_t mx2 = dup2(mx);
                uint32_t bits = 0;
#pragma unroll
                for (int k = 0; k < P; ++k) { bool ph, pl; __vibmax_s16x2(hb[k], mx2, &ph, &pl); bits |= (pl ? 1u : 0u) << (2 * k) | (ph ? 1u : 0u) << (2 * k + 1); }
                const int fi = __reduce_min_sync(0xffffffffu, bits ? j0 + __ffs(bits) - 1 : INT32_MAX);
                const int la = __reduce_max_sync(0xffffffffu, bits ? j0 + 31 - __clz(bits) : -1);
                if (mx > inf_min) { mxrow = mx; left = fi; right = la; }
                else if (mx == inf_min) right = la;
            }
            // the five planes of the row (vectors beg_sn .. end_sn) and INF_MIN in H of the vector after the band
            {
                int16_t *H = w.dp + off;
                if (extra) {
                    // vector end_sn = pend_sn + 1, one column per lane (the per-vector code of the general path, set_num = 2):
                    // M only from the predecessor's last column (when that is the last lane of its vector), no E, F seeded by
                    // the strip pass' carries and propagated by the reference's truncated log-step scan
                    const int o1 = par.gap_open1, o2 = par.gap_open2;
                    const int last_lane = (pend_sn - c0sn + 1) * (32 / C) - 1;
                    const int f1l = (int)f1[P - 1] >> 16, f2l = (int)f2[P - 1] >> 16, hml = (int)hm[P - 1] >> 16;
                    const int first1 = __shfl_sync(0xffffffffu, hml > f1l + o1 ? hml : f1l + o1, last_lane);
                    const int first2 = __shfl_sync(0xffffffffu, hml > f2l + o2 ? hml : f2l + o2, last_lane);
                    const int hpl = __shfl_sync(0xffffffffu, (int)hp[P - 1] >> 16, last_lane);      // H of the previous row at column 32 * end_sn - 1
                    const int col0 = end_sn * PN, khi = end - col0;
                    int hv = inf_min;
                    if (((lr.end + 1) >> 5) >= end_sn) hv = (int16_t)((lane == 0 ? hpl : inf_min) + ps);
                    hv = (int16_t)(hv + w.qp[(size_t)(nb > 3 ? 4 : nb) * w.qp_stride + col0 + lane]);
                    if (lane > khi || hv < inf_min) hv = inf_min;                                   // band mask; max(h, E1, E2) with E = INF_MIN
                    int g1 = WarpLanes::sub(WarpLanes::shift_up(hv, 1, first1), oe1), g2 = WarpLanes::sub(WarpLanes::shift_up(hv, 1, first2), oe2);
                    g1 = set_f(g1, 2, e1); g2 = set_f(g2, 2, e2);
                    hv = hv > g1 ? hv : g1; hv = hv > g2 ? hv : g2;
                    if (lane > khi) hv = inf_min;
                    const int v1 = WarpLanes::vmax(WarpLanes::sub(inf_min, e1), WarpLanes::sub(hv, oe1)), v2 = WarpLanes::vmax(WarpLanes::sub(inf_min, e2), WarpLanes::sub(hv, oe2));
                    int16_t *X = H + (size_t)(nvd - 1) * PN + lane;
                    X[0] = (int16_t)hv; X[(size_t)nv * PN] = (int16_t)v1; X[(size_t)2 * nv * PN] = (int16_t)v2; X[(size_t)3 * nv * PN] = (int16_t)g1; X[(size_t)4 * nv * PN] = (int16_t)g2;
                    if (banded) {
                        int mv, fi, la;
                        if (WarpLanes::row_max(hv, 0, khi > PN - 1 ? PN - 1 : khi, mv, fi, la)) {
                            if (mv > mxrow) { left = col0 + fi; right = col0 + la; }
                            else if (mv == mxrow) right = col0 + la;
                        }
                    }
                    // into the register window: lanes L0 .. hold this vector, C columns each
                    const int L0 = (end_sn - c0sn) * (32 / C);
                    const bool mine = lane >= L0 && lane < L0 + 32 / C;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const int s0 = ((lane - L0) * C + 2 * k) & 31;
                        const uint32_t a = pk2(__shfl_sync(0xffffffffu, hv, s0), __shfl_sync(0xffffffffu, hv, s0 + 1));
                        const uint32_t b = pk2(__shfl_sync(0xffffffffu, v1, s0), __shfl_sync(0xffffffffu, v1, s0 + 1));
                        const uint32_t c = pk2(__shfl_sync(0xffffffffu, v2, s0), __shfl_sync(0xffffffffu, v2, s0 + 1));
                        if (mine) { hn[k] = a; x1[k] = b; x2[k] = c; }
                    }
                }
                if ((rel >> 5) < (extra ? nvd - 1 : nvd)) {
                    stp<C>(H + rel, hn); stp<C>(H + (size_t)nv * PN + rel, x1); stp<C>(H + (size_t)2 * nv * PN + rel, x2);
                    stp<C>(H + (size_t)3 * nv * PN + rel, f1); stp<C>(H + (size_t)4 * nv * PN + rel, f2);
                }
                if (end_sn + 1 <= dp_sn - 1) H[nvd * PN + lane] = (int16_t)inf_min;
                const int4 packed = pack((int)off, beg, end, left, right);
                if (lane == 0) w.rinfo[id] = packed;
                lr = unpack(packed);
            }
#pragma unroll
            for (int k = 0; k < P; ++k) { hp[k] = hn[k]; e1p[k] = x1[k]; e2p[k] = x2[k]; }
            last_id = id; pend_sn = end_sn; ++n_done;
#ifdef LCD_SIMT_EMU
            if (lane == 0) { simt_stat[0]++; simt_stat[2 + P]++; }
#endif
            if (++oi >= n) break;
        }
        if (n_done == 0) return oi;
        // leave the last row where a general row expects it
        st.last_id = last_id; st.last_row = lr; st.dp_top = dp_top;
        const int nvd = (lr.end >> 5) - (lr.beg >> 5) + 1;
        if (nvd + 1 <= NVC) {
            int16_t *dst = row_cache + st.cache_buf * 3 * NVC * PN;
            if ((rel >> 5) < nvd) { stp<C>(dst + rel, hp); stp<C>(dst + NVC * PN + rel, e1p); stp<C>(dst + 2 * NVC * PN + rel, e2p); }
            dst[nvd * PN + lane] = (int16_t)inf_min;
            st.last_cached = true;
        } else st.last_cached = false;
        __syncwarp();
        return oi;
    }
#endif

    // one sequence against the whole graph: simd_abpoa_cg_align_sequence_to_graph_core (:1200-1228)
    // returns number of cigar entries (reverse order) or <0
    __device__ int align(const uint8_t *query, int qlen) {
        const int n = n_rows, n_tot = w.n_nodes;       // rows of the (sub-)graph; node count of the whole graph (reset value of max_pos_left)
        const int dp_sn = (qlen + 1 + PN - 1) / PN;
        const int wband = par.wb < 0 ? qlen : par.wb + (int)(par.wf * qlen);
        const int o1 = par.gap_open1, e1 = par.gap_ext1, o2 = par.gap_open2, e2 = par.gap_ext2;
        const int match = par.match, mism = par.mismatch;
        const bool banded = par.wb >= 0;
#ifndef LCD_EMU
        if constexpr (L::STRIP) {       // the read shifted by one, padded with a never-matching sentinel (strip_core)
            const int qs_len = (dp_sn + 2) * PN + 16;
            for (int j = L::tid(); j < qs_len; j += L::NT) w.qs[j] = (j >= 1 && j <= qlen) ? query[j - 1] : 7;
            // query profile (simd_abpoa_cg_var :531-560): column 0 and the columns past the read score 0
            const int qst = w.qp_stride;
            for (int j = L::tid(); j < qst; j += L::NT) {
                const int qb = (j >= 1 && j <= qlen) ? query[j - 1] : 7;
#pragma unroll
                for (int b = 0; b < 4; ++b) w.qp[b * qst + j] = (int16_t)(qb > 3 ? 0 : (qb == b ? match : -mism));
                w.qp[4 * qst + j] = 0;
            }
        }
#endif
        uint32_t dp_top = 0;
        int last_id = beg_id, cache_buf = 0; bool last_cached = false; Row last_row;
        const int rem_end = banded ? w.remain[end_id] : 0;
        // ---- first row (SRC) :627-688 ; max_pos_left/right of SRC are 0, i.e. left_max_i = right_max_i = -1 + ...
        {
            int end0 = qlen;
            if (banded) {
                const int r = qlen - (w.remain[beg_id] - rem_end - 1);
                end0 = (0 > r ? 0 : r) + wband; if (end0 > qlen) end0 = qlen;
            }
            const int nv = (end0 >> 5) + 2;
            if ((uint64_t)nv * PN * 5 > w.dp_capacity) return ST_OOM;
            // successors of SRC start from max_pos_left = max_pos_right = 1: encode as left_max_i = right_max_i = 0
            if (L::tid() == 0) w.rinfo[beg_id] = pack(0, 0, end0, 0, 0);
            dp_top = (uint32_t)nv * PN * 5;
            int16_t *h = w.dp, *pe1 = h + (size_t)nv * PN, *pe2 = pe1 + (size_t)nv * PN, *pf1 = pe2 + (size_t)nv * PN, *pf2 = pf1 + (size_t)nv * PN;
            const int esn = ((end0 >> 5) + 1 < dp_sn - 1) ? (end0 >> 5) + 1 : dp_sn - 1;
            for (int sn = 0; sn < nv; ++sn) {
                const int col0 = sn * PN;
                const bool init = sn <= esn;
                vec vh = L::map_cols(col0, [&](int j) -> int {
                    if (j == 0) return 0;
                    if (j <= end0) { const int a = (int16_t)(-o1 - e1 * j), b = (int16_t)(-o2 - e2 * j); return a > b ? a : b; }
                    return init ? inf_min : GARBAGE; });
                vec ve1 = L::map_cols(col0, [&](int j) -> int { return j == 0 ? (int16_t)-oe1 : (init ? inf_min : GARBAGE); });
                vec ve2 = L::map_cols(col0, [&](int j) -> int { return j == 0 ? (int16_t)-oe2 : (init ? inf_min : GARBAGE); });
                vec vf1 = L::map_cols(col0, [&](int j) -> int { return j == 0 ? inf_min : (j <= end0 ? (int)(int16_t)(-o1 - e1 * j) : GARBAGE); });
                vec vf2 = L::map_cols(col0, [&](int j) -> int { return j == 0 ? inf_min : (j <= end0 ? (int)(int16_t)(-o2 - e2 * j) : GARBAGE); });
                L::store(h + col0, vh); L::store(pe1 + col0, ve1); L::store(pe2 + col0, ve2); L::store(pf1 + col0, vf1); L::store(pf2 + col0, vf2);
            }
            cells += (unsigned long long)(end0 + 1);
            last_row = unpack(pack(0, 0, end0, 0, 0));
            if (row_cache && nv <= NVC) {
                for (int sn = 0; sn < nv; ++sn) {
                    L::store(row_cache + sn * PN, L::load(h + sn * PN));
                    L::store(row_cache + (NVC + sn) * PN, L::load(pe1 + sn * PN));
                    L::store(row_cache + (2 * NVC + sn) * PN, L::load(pe2 + sn * PN));
                }
                last_cached = true;
            }
        }
        L::sync();
        // ---- rows in topological (list) order, SINK excluded
        int oi = 1;
        while (oi < n) {
#ifndef LCD_EMU
            if constexpr (L::STRIP) {
                // Runs of chain rows (one in-edge, from the row computed just before, band not reaching past the predecessor's
                // last vector) are computed by chain_segment: previous row in registers, int16x2 cells, no memory on the
                // critical path.  It returns the position of the first row it did not take.
                const int4 mt = ring_get(oi);
                if (mt.x != end_id && ((mt.z >> 8) & 0xff) == 1 && mt.y == last_id) {
                    int beg, end, beg_sn;
                    chain_band(last_row, mt.w, rem_end, qlen, n_tot, wband, banded, beg, end, beg_sn);
                    const int end_sn = end >> 5, nvd = end_sn - beg_sn + 1;
#ifdef LCD_SIMT_EMU
                    if (L::tid() == 0 && getenv("SIMT_WHY")) { if (!(end_sn <= (last_row.end >> 5) + 1)) fprintf(stderr, "why: end_sn %d pend_sn %d beg_sn %d pbeg_sn %d qlen %d nin-prev? left1 %d right1 %d prev beg %d end %d\n", end_sn, last_row.end >> 5, beg_sn, last_row.beg >> 5, qlen, last_row.left1, last_row.right1, last_row.beg, last_row.end); else if (!(beg_sn <= (last_row.end >> 5))) fprintf(stderr, "why: beg_sn\n"); else if (nvd > 8) fprintf(stderr, "why: nvd %d\n", nvd); }
#endif
                    if (end_sn <= (last_row.end >> 5) + 1 && beg_sn <= (last_row.end >> 5) && nvd <= 8) {
                        ChainState st; st.last_id = last_id; st.last_row = last_row; st.last_cached = last_cached; st.cache_buf = cache_buf; st.dp_top = dp_top;
                        int err = 0, noi;
                        LCD_T0();
                        if (nvd <= 2) noi = chain_segment<2>(oi, beg_sn, n, n_tot, qlen, dp_sn, wband, banded, rem_end, st, err);
                        else if (nvd <= 4) noi = chain_segment<4>(oi, beg_sn, n, n_tot, qlen, dp_sn, wband, banded, rem_end, st, err);
                        else noi = chain_segment<8>(oi, beg_sn, n, n_tot, qlen, dp_sn, wband, banded, rem_end, st, err);
                        LCD_T1(t_seg);
                        if (err) return err;
                        if (noi > oi) {
#ifdef LCD_POA_TIMING
                            n_seg += noi - oi;
#endif
                            last_id = st.last_id; last_row = st.last_row; last_cached = st.last_cached; cache_buf = st.cache_buf; dp_top = st.dp_top;
                            oi = noi;
                            continue;
                        }
                    }
                }
            }
#endif
            const int id = w.order[oi++];
            if (id == end_id) continue;
#ifdef LCD_POA_TIMING
            const long long tg0_ = clock64(); n_gen++;
#endif
#ifdef LCD_SIMT_EMU
            if (L::tid() == 0) { simt_stat[1]++; if (ai_n[id] == 1 && ai_pool[ai_off[id]].x == last_id) simt_stat[2]++; }
#endif
            const int nin = ai_n[id], nb = w.base[id], rem = banded ? w.remain[id] : 0;
            const int4 *ie = ai_pool + ai_off[id];
            const int4 ie0 = nin > 0 ? ie[0] : make_int4(0, 0, 0, 0);
            // predecessor descriptors: the first MAXP in registers, any further ones re-read per vector
            constexpr int MAXP = 4;
            int4 pe[MAXP]; Row pr[MAXP];
#pragma unroll
            for (int k = 0; k < MAXP; ++k) if (k < nin) pe[k] = k == 0 ? ie0 : ie[k];
#pragma unroll
            for (int k = 0; k < MAXP; ++k) if (k < nin) pr[k] = pe[k].x == last_id ? last_row : unpack(w.rinfo[pe[k].x]);
            int beg, end, beg_sn, end_sn, max_pre_end_sn;
            if (!banded) { beg = 0; end = qlen; beg_sn = 0; end_sn = end >> 5; max_pre_end_sn = end_sn; }
            else {
                // max_pos_left/right pulled from the predecessors' row maxima (simd_abpoa_ada_max_i :1121-1130
                // pushes left_max_i+1 / right_max_i+1 to every successor; reset values are node_n and 0)
                int maxl = n_tot, maxr = 0, min_pre_beg = INT32_MAX, min_pre_beg_sn = INT32_MAX; max_pre_end_sn = -1;
                for (int k = 0; k < nin; ++k) {
                    Row r;
                    if (k < MAXP) {
#pragma unroll
                        for (int q = 0; q < MAXP; ++q) if (q == k) r = pr[q];
                    } else r = unpack(w.rinfo[ie[k].x]);
                    if (r.left1 < maxl) maxl = r.left1;
                    if (r.right1 > maxr) maxr = r.right1;
                    if (min_pre_beg > r.beg) { min_pre_beg = r.beg; min_pre_beg_sn = r.beg >> 5; }
                    if (max_pre_end_sn < (r.end >> 5)) max_pre_end_sn = r.end >> 5;
                }
                const int rr = qlen - (rem - rem_end - 1);
                beg = (maxl < rr ? maxl : rr) - wband; if (beg < 0) beg = 0;
                end = (maxr > rr ? maxr : rr) + wband; if (end > qlen) end = qlen;
                beg_sn = beg >> 5;
                if (beg_sn < min_pre_beg_sn) { beg = min_pre_beg; beg_sn = min_pre_beg_sn; }
                end_sn = end >> 5;
            }
            const int nv = end_sn - beg_sn + 2;
            const uint32_t need = (uint32_t)nv * PN * 5;
            if (dp_top + need > w.dp_capacity) return ST_OOM;
            const uint32_t off = dp_top; dp_top += need;
            cells += (unsigned long long)(end - beg + 1);
            int16_t *H = w.dp + off - (size_t)beg_sn * PN, *E1 = H + (size_t)nv * PN, *E2 = E1 + (size_t)nv * PN;
            int16_t *F1 = E2 + (size_t)nv * PN, *F2 = F1 + (size_t)nv * PN;
            const int16_t *cache_rd = (row_cache && last_cached) ? row_cache + cache_buf * 3 * NVC * PN : nullptr;
            const bool cache_wr = row_cache && nv <= NVC;
            int16_t *cwH = cache_wr ? row_cache + (cache_buf ^ 1) * 3 * NVC * PN - (size_t)beg_sn * PN : nullptr;
            Pred pd[MAXP];
#pragma unroll
            for (int k = 0; k < MAXP; ++k) if (k < nin) pd[k] = make_pred(pe[k], pr[k], beg_sn, end_sn, dp_sn, pe[k].x == last_id ? cache_rd : nullptr);
            int first1 = 0, first2 = 0;
            int mx = inf_min, left = -1, right = -1;
            if (!L::TWO_PHASE) {
            int sn0 = beg_sn;
#ifndef LCD_EMU
            if constexpr (L::STRIP) {
                const int sn_hi = end_sn < max_pre_end_sn ? end_sn : max_pre_end_sn;
                const int nvs = sn_hi - beg_sn + 1;
                if (nin >= 1 && nin <= MAXP && nvs >= 1 && nvs <= 8) {
                    StripArgs sa;
                    sa.beg = beg; sa.end = end; sa.beg_sn = beg_sn; sa.end_sn = end_sn; sa.sn_hi = sn_hi; sa.nb = nb;
                    sa.H = H; sa.E1 = E1; sa.E2 = E2; sa.F1 = F1; sa.F2 = F2; sa.cwH = cwH; sa.banded = banded;
                    if (nin == 1 && cache_rd && pe[0].x == last_id && sn_hi == end_sn) {      // chain row, predecessor on chip
                        const int pbs = pr[0].beg >> 5;
                        if (nvs <= 2) chain_row<2>(sa, cache_rd, pbs, pe[0].z, first1, first2, mx, left, right);
                        else if (nvs <= 4) chain_row<4>(sa, cache_rd, pbs, pe[0].z, first1, first2, mx, left, right);
                        else chain_row<8>(sa, cache_rd, pbs, pe[0].z, first1, first2, mx, left, right);
                    }
                    else if (nvs <= 2) strip_row<2>(sa, nin, pd, first1, first2, mx, left, right);
                    else if (nvs <= 4) strip_row<4>(sa, nin, pd, first1, first2, mx, left, right);
                    else strip_row<8>(sa, nin, pd, first1, first2, mx, left, right);
                    sn0 = sn_hi + 1;
                }
            }
#endif
            for (int sn = sn0; sn <= end_sn; ++sn) {
                const int col0 = sn * PN;
                vec h = L::set1(inf_min), ve1 = L::set1(inf_min), ve2 = L::set1(inf_min);
#pragma unroll
                for (int k = 0; k < MAXP; ++k) if (k < nin) pred_accumulate(pd[k], k, sn, col0, h, ve1, ve2);
                for (int k = MAXP; k < nin; ++k) {
                    const int4 e = ie[k];
                    const Pred p = make_pred(e, unpack(w.rinfo[e.x]), beg_sn, end_sn, dp_sn);
                    pred_accumulate(p, k, sn, col0, h, ve1, ve2);
                }
                // + query profile, band mask
                const vec q = L::map_cols(col0, [&](int j) -> int {
                    if (j == 0 || j > qlen) return 0;
                    const int qb = query[j - 1];
                    return (nb > 3 || qb > 3) ? 0 : (nb == qb ? match : -mism); });
                h = L::add(h, q);
                const int klo = beg - col0, khi = end - col0;       // lanes inside the band
                if (klo > 0 || khi < PN - 1) { h = L::keep(h, klo, khi, inf_min); ve1 = L::keep(ve1, klo, khi, inf_min); ve2 = L::keep(ve2, klo, khi, inf_min); }
                if (sn == beg_sn) first1 = first2 = L::lane_value(h, 0);
                const int set_num = sn > max_pre_end_sn ? (sn == max_pre_end_sn + 1 ? 2 : 1) : PN;
                h = L::vmax(L::vmax(h, ve1), ve2);
                vec f1 = L::sub(L::shift_up(h, 1, first1), L::set1(oe1));
                vec f2 = L::sub(L::shift_up(h, 1, first2), L::set1(oe2));
                f1 = set_f(f1, set_num, e1);
                f2 = set_f(f2, set_num, e2);
                first1 = L::lane_value(L::vmax(h, L::add(f1, L::set1(o1))), PN - 1);
                first2 = L::lane_value(L::vmax(h, L::add(f2, L::set1(o2))), PN - 1);
                h = L::vmax(h, L::vmax(f1, f2));
                if (sn == end_sn && khi < PN - 1) { h = L::keep(h, -1, khi, inf_min); ve1 = L::keep(ve1, -1, khi, inf_min); ve2 = L::keep(ve2, -1, khi, inf_min); }
                ve1 = L::vmax(L::sub(ve1, L::set1(e1)), L::sub(h, L::set1(oe1)));
                ve2 = L::vmax(L::sub(ve2, L::set1(e2)), L::sub(h, L::set1(oe2)));
                L::store(H + col0, h); L::store(E1 + col0, ve1); L::store(E2 + col0, ve2); L::store(F1 + col0, f1); L::store(F2 + col0, f2);
                if (cache_wr) { L::store(cwH + col0, h); L::store(cwH + NVC * PN + col0, ve1); L::store(cwH + 2 * NVC * PN + col0, ve2); }
                if (banded) {                // simd_abpoa_max_in_row :1107-1119
                    int m, fi, la;
                    if (L::row_max(h, klo < 0 ? 0 : klo, khi > PN - 1 ? PN - 1 : khi, m, fi, la)) {
                        if (m > mx) { mx = m; left = col0 + fi; right = col0 + la; }
                        else if (m == mx) right = col0 + la;
                    }
                }
            }
            } else {
                // Two-phase row for a group of NW warps: the vectors of the row are dealt to the warps.  Phase A
                // computes everything that does not depend on the F carry of the vectors to the left: hm = max(M+q,
                // E1, E2) and G = SIMD_SET_F of the F start vector with a never-winning value in lane 0.  SIMD_SET_F
                // is max-plus affine, so the true F is max(G, carry - oe - e*lane), and the carry obeys
                // c' = max(max(hm[31], G[31] + o), c - 32 e)  -- a scalar recurrence resolved through shared memory.
                int *B1 = gs, *B2 = gs + MAXV, *misc = gs + 2 * MAXV;
                if (end_sn - beg_sn + 1 > MAXV) return ST_OOM;
                const int lowf = inf_min - 900;
                for (int sn = beg_sn + L::warp(); sn <= end_sn; sn += L::NW) {
                    const int col0 = sn * PN;
                    vec h = L::set1(inf_min), ve1 = L::set1(inf_min), ve2 = L::set1(inf_min);
#pragma unroll
                    for (int k = 0; k < MAXP; ++k) if (k < nin) pred_accumulate(pd[k], k, sn, col0, h, ve1, ve2);
                    for (int k = MAXP; k < nin; ++k) {
                        const int4 e = ie[k];
                        const Pred p = make_pred(e, unpack(w.rinfo[e.x]), beg_sn, end_sn, dp_sn);
                        pred_accumulate(p, k, sn, col0, h, ve1, ve2);
                    }
                    const vec q = L::map_cols(col0, [&](int j) -> int {
                        if (j == 0 || j > qlen) return 0;
                        const int qb = query[j - 1];
                        return (nb > 3 || qb > 3) ? 0 : (nb == qb ? match : -mism); });
                    h = L::add(h, q);
                    const int klo = beg - col0, khi = end - col0;
                    if (klo > 0 || khi < PN - 1) { h = L::keep(h, klo, khi, inf_min); ve1 = L::keep(ve1, klo, khi, inf_min); ve2 = L::keep(ve2, klo, khi, inf_min); }
                    if (sn == beg_sn) { const int c0 = L::lane_value(h, 0); if (L::lane() == 0) misc[0] = c0; }
                    const int set_num = sn > max_pre_end_sn ? (sn == max_pre_end_sn + 1 ? 2 : 1) : PN;
                    h = L::vmax(L::vmax(h, ve1), ve2);
                    vec g1 = L::keep(L::sub(L::shift_up(h, 1, 0), L::set1(oe1)), 1, PN - 1, lowf);
                    vec g2 = L::keep(L::sub(L::shift_up(h, 1, 0), L::set1(oe2)), 1, PN - 1, lowf);
                    g1 = set_f(g1, set_num, e1);
                    g2 = set_f(g2, set_num, e2);
                    const int b1 = L::lane_value(L::vmax(h, L::add(g1, L::set1(o1))), PN - 1);
                    const int b2 = L::lane_value(L::vmax(h, L::add(g2, L::set1(o2))), PN - 1);
                    if (L::lane() == 0) { B1[sn - beg_sn] = b1; B2[sn - beg_sn] = b2; }
                    L::store(H + col0, h); L::store(E1 + col0, ve1); L::store(E2 + col0, ve2); L::store(F1 + col0, g1); L::store(F2 + col0, g2);
                }
                L::sync();
                for (int sn = beg_sn + L::warp(); sn <= end_sn; sn += L::NW) {
                    const int col0 = sn * PN;
                    int c1 = misc[0], c2 = c1;
                    for (int u = 0; u < sn - beg_sn; ++u) {
                        const int x1 = B1[u], x2 = B2[u];
                        c1 = x1 > c1 - 32 * e1 ? x1 : c1 - 32 * e1;
                        c2 = x2 > c2 - 32 * e2 ? x2 : c2 - 32 * e2;
                    }
                    vec h = L::load(H + col0), ve1 = L::load(E1 + col0), ve2 = L::load(E2 + col0);
                    const int khi = end - col0, klo = beg - col0;
                    const vec f1 = L::vmax(L::load(F1 + col0), L::map_cols(0, [&](int l) -> int { return c1 - oe1 - e1 * l; }));
                    const vec f2 = L::vmax(L::load(F2 + col0), L::map_cols(0, [&](int l) -> int { return c2 - oe2 - e2 * l; }));
                    h = L::vmax(h, L::vmax(f1, f2));
                    if (sn == end_sn && khi < PN - 1) { h = L::keep(h, -1, khi, inf_min); ve1 = L::keep(ve1, -1, khi, inf_min); ve2 = L::keep(ve2, -1, khi, inf_min); }
                    ve1 = L::vmax(L::sub(ve1, L::set1(e1)), L::sub(h, L::set1(oe1)));
                    ve2 = L::vmax(L::sub(ve2, L::set1(e2)), L::sub(h, L::set1(oe2)));
                    L::store(H + col0, h); L::store(E1 + col0, ve1); L::store(E2 + col0, ve2); L::store(F1 + col0, f1); L::store(F2 + col0, f2);
                    if (banded) {
                        int m, fi, la;
                        if (L::row_max(h, klo < 0 ? 0 : klo, khi > PN - 1 ? PN - 1 : khi, m, fi, la)) {
                            if (m > mx) { mx = m; left = col0 + fi; right = col0 + la; }
                            else if (m == mx) right = col0 + la;
                        }
                    }
                }
                if (banded && L::NW > 1) {        // combine the per-warp row maxima: first / last column attaining the maximum
                    if (L::lane() == 0) { misc[4 + 3 * L::warp()] = mx; misc[5 + 3 * L::warp()] = left; misc[6 + 3 * L::warp()] = right; }
                    L::sync();
                    int M = inf_min;
                    for (int x = 0; x < L::NW; ++x) if (misc[4 + 3 * x] > M) M = misc[4 + 3 * x];
                    int lft = INT32_MAX, rgt = -1;
                    for (int x = 0; x < L::NW; ++x) if (misc[4 + 3 * x] == M) {
                        if (misc[5 + 3 * x] >= 0 && misc[5 + 3 * x] < lft) lft = misc[5 + 3 * x];
                        if (misc[6 + 3 * x] > rgt) rgt = misc[6 + 3 * x];
                    }
                    mx = M; left = lft == INT32_MAX ? -1 : lft; right = rgt;
                }
            }
            // the vector after the band: INF_MIN in H (read by successors' M), undefined elsewhere
            if (end_sn + 1 <= dp_sn - 1) {
                L::store(H + (end_sn + 1) * PN, L::set1(inf_min));
                if (cache_wr) L::store(cwH + (end_sn + 1) * PN, L::set1(inf_min));
            }
            if (L::tid() == 0) w.rinfo[id] = pack((int)off, beg, end, left, right);
            last_id = id; last_row = unpack(pack((int)off, beg, end, left, right));
            last_cached = cache_wr; if (cache_wr) cache_buf ^= 1;
            L::sync();
#ifdef LCD_POA_TIMING
            t_gen += (unsigned long long)(clock64() - tg0_);
#endif
        }
        // ---- best cell :1092-1105 and backtrack :309-458 (lane 0; result broadcast through memory)
#ifndef LCD_EMU
        if constexpr (L::STRIP) {
            LCD_T0(); const int rc = backtrack_warp(query, qlen, reinterpret_cast<int *>(w.cigar)); LCD_T1(t_bt);
            return rc;
        }
#endif
        { LCD_T0(); if (L::tid() == 0) backtrack(query, qlen);
        L::sync(); LCD_T1(t_bt); }
        return w.tmp[0];
    }

    __device__ void cig_push(int &n, int op, int len, int node) {
        if (n > 0 && op == 1 && (w.cigar[n - 1].x & 3) == 1) { w.cigar[n - 1].x += len << 2; return; }
        if (n >= w.cigar_cap) { w.oom = 1; return; }
        w.cigar[n++] = make_int2(op | (len << 2), node);
    }
    __device__ int backtrack(const uint8_t *query, int qlen) {
        const int e1 = par.gap_ext1, e2 = par.gap_ext2;
        int best = inf_min, bi = 0, bj = 0;
        {
            const int4 *ie = ai_pool + ai_off[end_id];
            for (int k = 0; k < ai_n[end_id]; ++k) {
                const int r = ie[k].x;
                const Row rr = unpack(w.rinfo[r]);
                const int e = qlen > rr.end ? rr.end : qlen;
                const int s = cell(rr, 0, e);
                if (s > best) { best = s; bi = r; bj = e; }
            }
        }
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int n = 0, id = bi, j = bj, cur_op = ALL_OP, rc = 0;
        if (bj < qlen) cig_push(n, 1, qlen - bj, -1);
        Row cur = unpack(w.rinfo[id]);
        while (id != beg_id && j > 0) {
            const int nb = w.base[id], qb = query[j - 1];
            const int s = (nb > 3 || qb > 3) ? 0 : (nb == qb ? par.match : -par.mismatch);
            const int4 *ie = ai_pool + ai_off[id];
            const int nin = ai_n[id];
            const int hj = cell(cur, 0, j);
            int hit = 0;
            for (int pass = 0; pass < 2 && !hit; ++pass) {
                if (pass == 1) {
                    if (cur_op & E_OP) {
                        const int e1j = cell(cur, 1, j), e2j = cell(cur, 2, j);
                        for (int k = 0; k < nin && !hit; ++k) {
                            const int4 e = ie[k];
                            const int p = e.x, ps = e.z;
                            const Row pr = unpack(w.rinfo[p]);
                            if (j < pr.beg || j > pr.end) continue;
                            const int phj = cell(pr, 0, j);
                            if (cur_op & E1_OP) {
                                const int pe = cell(pr, 1, j);
                                const int ok = (cur_op & M_OP) ? (hj == pe + ps) : (e1j == pe - e1 + ps);
                                if (ok) { cur_op = (phj - oe1 == pe) ? (M_OP | F_OP) : E1_OP; hit = 1; }
                            }
                            if (!hit && (cur_op & E2_OP)) {
                                const int pe = cell(pr, 2, j);
                                const int ok = (cur_op & M_OP) ? (hj == pe + ps) : (e2j == pe - e2 + ps);
                                if (ok) { cur_op = (phj - oe2 == pe) ? (M_OP | F_OP) : E2_OP; hit = 1; }
                            }
                            if (hit) { cig_push(n, 2, 1, id); id = p; cur = pr; }
                        }
                    }
                    if (!hit && (cur_op & F_OP)) {
                        const int hj1 = cell(cur, 0, j - 1);
                        if (cur_op & F1_OP) {
                            const int f = cell(cur, 3, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe1 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(cur, 3, j - 1) - e1 == f) { cur_op = F1_OP; hit = 1; }
                            }
                        }
                        if (!hit && (cur_op & F2_OP)) {
                            const int f = cell(cur, 4, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe2 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(cur, 4, j - 1) - e2 == f) { cur_op = F2_OP; hit = 1; }
                            }
                        }
                        if (hit) { cig_push(n, 1, 1, id); --j; }
                    }
                    if (hit) break;
                }
                if (cur_op & M_OP) {
                    for (int k = 0; k < nin; ++k) {
                        const int4 e = ie[k];
                        const int p = e.x, ps = e.z;
                        const Row pr = unpack(w.rinfo[p]);
                        if (j - 1 < pr.beg || j - 1 > pr.end) continue;
                        if (cell(pr, 0, j - 1) + s + ps == hj) {
                            cig_push(n, 0, 1, id);
                            id = p; cur = pr; --j; hit = 1; cur_op = ALL_OP;
                            break;
                        }
                    }
                }
            }
            if (!hit) { rc = ST_BACKTRACK; break; }
            if (w.oom) { rc = ST_OOM; break; }
        }
        if (rc == 0 && j > 0) cig_push(n, 1, j, -1);
        w.tmp[0] = rc ? rc : n;
        return w.tmp[0];
    }

    // ---- de-novo read clustering for max_n_cons = 2 (abpoa_multip_read_clu_kmedoids, abPOA/src/abpoa_output.c:676-1180) --------------
    // Works on the row-column MSA the warp / CTA has just written (a node's out-edge read sets restricted to a cluster, :345-352, are counts
    // over the MSA rows of the cluster's reads).  Candidate het columns (:676-720) are found by the lanes, one column each; a candidate is
    // kept as its column and a 6-nibble map base -> allele index in order of first appearance, so that a read's allele at a candidate is one
    // MSA byte and a shift; candidates with identical read partitions (allele_clu_exist :647-674) are found through a 64-bit fingerprint and
    // verified read by read.  Priorities (:772-791: a stable sort) are ranks counted by the lanes, the distance matrix (:802-862) is dealt
    // pair by pair; the k-medoids rounds (:865-1121; two clusters: two dozen reads) run on one lane.  Scratch: the DP arena, free by now.
    __device__ static __forceinline__ int cl_alle(const uint8_t *msa, int ml, int col, unsigned map, int r) {
        return (int)((map >> (4 * msa[(size_t)r * ml + col])) & 15u) - 1;
    }
    // returns the number of clusters (1 or 2) or ST_OOM; with two clusters the consensus rows n_seq, n_seq + 1 of the MSA, the two consensus
    // sequences (back to back in cons) and read_clu are (re)written
    __device__ int cluster(uint8_t *msa, int n_seq, int ml, int min_w, uint8_t *cons, uint8_t *read_clu, int *cons_len0, int *cons_len1) {
        const int lane = L::tid();
        int *S = reinterpret_cast<int *>(w.dp);
        const uint64_t need = 10ull * (uint64_t)ml + (uint64_t)n_seq * n_seq + 2ull * n_seq + 64;
        if (need * 2 > w.dp_capacity) return ST_OOM;
        int *c_map = S, *c_info = S + ml, *c_hlo = S + 2 * (size_t)ml, *c_hhi = S + 3 * (size_t)ml, *list = S + 4 * (size_t)ml, *rep = S + 5 * (size_t)ml;
        int *h_k = S + 6 * (size_t)ml, *h_cnt = S + 7 * (size_t)ml, *h_vt = S + 8 * (size_t)ml, *prio = S + 9 * (size_t)ml;
        int *dis = S + 10 * (size_t)ml, *clu = dis + (size_t)n_seq * n_seq;
        const int min_het = min_w / 2 > 2 ? min_w / 2 : 2, min_hom = n_seq - min_het;
        for (int r = lane; r < n_seq; r += L::NT) read_clu[r] = 0;
        for (int i = lane; i < ml; i += L::NT) {
            int depth[6] = {0, 0, 0, 0, 0, 0}, first[6] = {0, 0, 0, 0, 0, 0};
            for (int r = 0; r < n_seq; ++r) { const int b = msa[(size_t)r * ml + i]; if (++depth[b] == 1) first[b] = r; }
            int al[6], nu = 0, vt = 0, tot = 0;
            for (int j = 0; j < 6; ++j) if (depth[j] >= min_het && depth[j] <= min_hom) { al[nu++] = j; tot += depth[j]; if (j == 5) vt = 1; }
            if (nu < 2) { c_info[i] = 0; continue; }
            for (int j = 0; j < nu - 1; ++j) for (int k = j + 1; k < nu; ++k) if (first[al[j]] > first[al[k]]) { const int t = al[j]; al[j] = al[k]; al[k] = t; }
            unsigned map = 0;
            for (int j = 0; j < nu; ++j) map |= (unsigned)(j + 1) << (4 * al[j]);
            unsigned long long h = 1469598103934665603ull;
            for (int r = 0; r < n_seq; ++r) h = (h ^ ((map >> (4 * msa[(size_t)r * ml + i])) & 15u)) * 1099511628211ull;
            c_map[i] = (int)map; c_info[i] = 1 | (vt << 1) | (nu << 4) | (tot << 8); c_hlo[i] = (int)(unsigned)h; c_hhi[i] = (int)(unsigned)(h >> 32);
        }
        L::sync();
        if (lane == 0) { int nc = 0; for (int i = 0; i < ml; ++i) if (c_info[i]) list[nc++] = i; w.tmp[0] = nc; }
        L::sync();
        const int nc = w.tmp[0];
        if (nc == 0) return 1;
        for (int k = lane; k < nc; k += L::NT) {
            const int ck = list[k]; int rp = k;
            for (int j = 0; j < k; ++j) {
                const int cj = list[j];
                if (c_hlo[cj] != c_hlo[ck] || c_hhi[cj] != c_hhi[ck] || ((c_info[cj] ^ c_info[ck]) & 0xf0)) continue;
                bool eq = true;
                for (int r = 0; r < n_seq && eq; ++r) if (cl_alle(msa, ml, cj, (unsigned)c_map[cj], r) != cl_alle(msa, ml, ck, (unsigned)c_map[ck], r)) eq = false;
                if (eq) { rp = j; break; }
            }
            rep[k] = rp;
        }
        L::sync();
        if (lane == 0) {             // the merged candidates in column order: count, X-if-any-X (:722-727)
            int nh = 0;
            for (int k = 0; k < nc; ++k) {
                const int ck = list[k], vt = (c_info[ck] >> 1) & 1, r0 = rep[k];
                if (r0 == k) { h_k[nh] = ck; h_cnt[nh] = 1; h_vt[nh] = vt; rep[k] = nh++; }
                else { const int e = rep[r0]; h_cnt[e]++; if (!vt) h_vt[e] = 0; }
            }
            w.tmp[0] = nh;
        }
        L::sync();
        const int nh = w.tmp[0];
        for (int e = lane; e < nh; e += L::NT) {
            const int ce = h_cnt[e], de = c_info[h_k[e]] >> 8, ve = h_vt[e]; int rank = 0;
            for (int f = 0; f < nh; ++f) {
                if (f == e) continue;
                const int cf = h_cnt[f], df = c_info[h_k[f]] >> 8, vf = h_vt[f];
                if (cf > ce || (cf == ce && (df > de || (df == de && (vf < ve || (vf == ve && f < e)))))) ++rank;
            }
            prio[rank] = e;
        }
        for (int p = lane; p < n_seq * n_seq; p += L::NT) {
            const int i = p / n_seq, j = p - i * n_seq;
            if (i == j) dis[p] = 0;
            if (i >= j) continue;
            int d = 0;
            for (int e = 0; e < nh; ++e) {
                const int col = h_k[e]; const unsigned map = (unsigned)c_map[col];
                const int a = cl_alle(msa, ml, col, map, i), b = cl_alle(msa, ml, col, map, j);
                if (a >= 0 && b >= 0 && a != b) d += (h_vt[e] == 0 ? 2 : 1) * h_cnt[e];
            }
            dis[i * n_seq + j] = d; dis[j * n_seq + i] = d;
        }
        L::sync();
        if (lane == 0) {
            int n_clu = 1, med0 = -1, med1 = -1, have = 0;
            for (int t = 0; t < nh && !have; ++t) {             // abpoa_collect_2medoids on the candidates in priority order (:865-890,996-1003)
                const int e = prio[t], col = h_k[e], nu = (c_info[col] >> 4) & 15; const unsigned map = (unsigned)c_map[col];
                int max_dis = 0;
                for (int a = 0; a < nu - 1; ++a) for (int b = a + 1; b < nu; ++b)
                    for (int r1 = 0; r1 < n_seq; ++r1) {
                        if (cl_alle(msa, ml, col, map, r1) != a) continue;
                        for (int r2 = 0; r2 < n_seq; ++r2) {
                            if (cl_alle(msa, ml, col, map, r2) != b) continue;
                            if (dis[r1 * n_seq + r2] > max_dis) { max_dis = dis[r1 * n_seq + r2]; med0 = r1; med1 = r2; }
                        }
                    }
                if (max_dis > 0) have = 1;
            }
            int ncs0 = 0, ncs1 = 0;
            if (have) {
                for (int iter = 0; iter < 10; ++iter) {          // abpoa_update_kmedoids (:1030-1093)
                    ncs0 = ncs1 = 0;
                    for (int i = 0; i < n_seq; ++i) {
                        const int d0 = dis[i * n_seq + med0], d1 = dis[i * n_seq + med1];
                        const int c = d1 < d0 ? 1 : d1 == d0 ? (ncs0 < ncs1 ? 0 : 1) : 0;
                        if (c) clu[n_seq + ncs1++] = i; else clu[ncs0++] = i;
                    }
                    int nm[2] = {-1, -1};
                    for (int c = 0; c < 2; ++c) {
                        const int *cr = clu + c * n_seq, cn = c ? ncs1 : ncs0; int best = INT32_MAX;
                        for (int j = 0; j < cn; ++j) {
                            int sum = 0;
                            for (int k = 0; k < cn; ++k) if (k != j) sum += dis[cr[j] * n_seq + cr[k]];
                            if (sum < best) { best = sum; nm[c] = cr[j]; }
                        }
                    }
                    if (nm[0] > nm[1]) { const int t = nm[0]; nm[0] = nm[1]; nm[1] = t; }
                    int changed = 0;
                    if (nm[0] != -1 && nm[1] != -1) changed = (nm[0] != med0 || nm[1] != med1) ? 1 : 0;
                    med0 = nm[0]; med1 = nm[1];
                    if (!changed) break;
                }
                if (ncs0 >= min_w && ncs1 >= min_w) { n_clu = 2; for (int j = 0; j < ncs1; ++j) read_clu[clu[n_seq + j]] = 1; }
            }
            w.tmp[0] = n_clu; w.tmp[1] = ncs0; w.tmp[2] = ncs1;
        }
        L::sync();
        const int n_clu = w.tmp[0];
        if (n_clu != 2) return 1;
        const int csz[2] = { w.tmp[1], w.tmp[2] };
        for (int i = lane; i < ml; i += L::NT) {                  // abpoa_set_major_voting_cons per cluster, sub_aln = 0 (:393-424)
            int cnt[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            for (int r = 0; r < n_seq; ++r) { const int b = msa[(size_t)r * ml + i]; if (b < 4) cnt[read_clu[r]][b]++; }
            for (int c = 0; c < 2; ++c) {
                int max_c = 0, total = 0, max_base = 5;
                for (int j = 0; j < 4; ++j) { if (cnt[c][j] > max_c) { max_c = cnt[c][j]; max_base = j; } total += cnt[c][j]; }
                msa[(size_t)(n_seq + c) * ml + i] = (uint8_t)(max_c >= csz[c] - total ? max_base : 5);
            }
        }
        L::sync();
        if (lane == 0) {
            int at = 0;
            for (int c = 0; c < 2; ++c) {
                const uint8_t *row = msa + (size_t)(n_seq + c) * ml; const int a0 = at;
                for (int i = 0; i < ml; ++i) if (row[i] != 5) cons[at++] = row[i];
                w.tmp[1 + c] = at - a0;
            }
        }
        L::sync();
        *cons_len0 = w.tmp[1]; *cons_len1 = w.tmp[2];
        return 2;
    }

    // ---- output: MSA ranks (abpoa_DFS_set_msa_rank :359-410), most-frequent consensus, RC-MSA ----------
    __device__ int finish(int n_seq, uint8_t *cons, int *cons_len_out, const KernelArgs &a, unsigned long long *msa_off_out, int *msa_len_out,
                          int min_w, uint8_t *read_clu, int *n_cons_out, int *cons_len2_out) {
        const int n = w.n_nodes, lane = L::tid();
        int *deg = w.tmp, *st = w.order;          // order[] is free now; DFS stack needs <= n entries... see below
        if (lane == 0) {
            for (int i = 0; i < n; ++i) deg[i] = w.in_n[i];
            // the stack can hold every node once
            int sp = 0, rank = 0;
            st[sp++] = 0; w.msa_rank[0] = -1;
            while (sp > 0) {
                const int cur = st[--sp];
                if (w.msa_rank[cur] < 0) {
                    w.msa_rank[cur] = rank;
                    for (int i = 0; i < w.aln_n[cur]; ++i) w.msa_rank[w.aln_pool[w.aln_off[cur] + i]] = rank;
                    rank++;
                }
                if (cur == 1) break;
                for (int i = 0; i < w.out_n[cur]; ++i) {
                    const int out = out_entry(cur, i)[0];
                    if (--deg[out] == 0) {
                        int ok = 1;
                        for (int j = 0; j < w.aln_n[out]; ++j) if (deg[w.aln_pool[w.aln_off[out] + j]] != 0) { ok = 0; break; }
                        if (!ok) continue;
                        st[sp++] = out; w.msa_rank[out] = -1;
                        for (int j = 0; j < w.aln_n[out]; ++j) { const int al = w.aln_pool[w.aln_off[out] + j]; st[sp++] = al; w.msa_rank[al] = -1; }
                    }
                }
            }
        }
        L::sync();
        const int ml = w.msa_rank[1] - 1;
        // column of every node (max rank over its aligned group) -> remain[] reused; per-column counts in the DP arena
        int *col = w.remain;
        int *cnt = reinterpret_cast<int *>(w.dp), *nid = cnt + (size_t)ml * 5;
        if ((uint64_t)ml * 10 * 2 + 64 > w.dp_capacity) return ST_OOM;
        for (int i = lane; i < ml * 5; i += L::NT) { cnt[i] = 0; nid[i] = 0; }
        for (int i = lane; i < n; i += L::NT) {
            int r = w.msa_rank[i];
            for (int j = 0; j < w.aln_n[i]; ++j) { const int rr = w.msa_rank[w.aln_pool[w.aln_off[i] + j]]; if (rr > r) r = rr; }
            col[i] = r - 1;
        }
        L::sync();
        for (int i = 2 + lane; i < n; i += L::NT) { cnt[col[i] * 5 + w.base[i]] = w.n_read[i]; nid[col[i] * 5 + w.base[i]] = i; }
        L::sync();
        // voting (abpoa_set_major_voting_cons :393-424); ordered compaction by lane 0 over per-column flags
        int *emit = w.maxl;        // per column: consensus node id or -1   (ml <= N)
        int bad = 0;
        for (int i = lane; i < ml; i += L::NT) {
            int max_c = 0, total = 0, max_base = 5;
            for (int j = 0; j < 4; ++j) { const int c = cnt[i * 5 + j]; if (c > max_c) { max_c = c; max_base = j; } total += c; }
            if (max_base == 5) { if (par.sub_aln) bad = 1; emit[i] = -1; continue; }      // sub_aln: the reference would read out of bounds; else 0 >= n_seq fails: no base
            const int gap_c = (par.sub_aln ? w.n_span[nid[i * 5 + max_base]] : n_seq) - total;
            emit[i] = max_c >= gap_c ? nid[i * 5 + max_base] : -1;
        }
        L::sync();
        int *cons_ids = w.maxr;
        if (lane == 0) {
            int cl = 0;
            for (int i = 0; i < ml; ++i) if (emit[i] >= 0) { cons_ids[cl] = emit[i]; cons[cl] = (uint8_t)w.base[emit[i]]; cl++; }
            w.tmp[0] = cl;
            // MSA output slot
            const unsigned long long bytes = ((unsigned long long)(n_seq + (par.max_n_cons == 2 ? 2 : 1)) * ml + 15) & ~15ull;
            const unsigned long long off = atomicAdd(a.msa_used, bytes);
            w.tmp[1] = (off + bytes <= a.msa_cap) ? 1 : 0;
            w.tmp[2] = (int)(off & 0xffffffffu); w.tmp[3] = (int)(off >> 32);
        }
        L::sync();
        const int cl = w.tmp[0];
        *cons_len_out = cl; *msa_len_out = ml;
        const unsigned long long off = ((unsigned long long)(unsigned)w.tmp[3] << 32) | (unsigned)w.tmp[2];
        *msa_off_out = off;
        if (bad) return ST_NOBASE;
        if (!w.tmp[1]) return ST_OOM;
        uint8_t *msa = a.msa + off;
        const long long tot = (long long)(n_seq + 1) * ml;
        for (long long i = lane; i < tot; i += L::NT) msa[i] = 5;
        L::sync();
        for (int i = 2 + lane; i < n; i += L::NT) {
            const int c = col[i], b = w.base[i];
            for (int e = 0; e < w.out_n[i]; ++e) {
                const int *ent = out_entry(i, e);
                for (int x = 0; x < 2 * w.rid_w; ++x) {
                    unsigned bits = (unsigned)ent[2 + x];
                    while (bits) { const int r = __ffs(bits) - 1; bits &= bits - 1; msa[(size_t)(x * 32 + r) * ml + c] = (uint8_t)b; }
                }
            }
        }
        for (int i = lane; i < cl; i += L::NT) msa[(size_t)n_seq * ml + col[cons_ids[i]]] = cons[i];
        L::sync();
        *n_cons_out = 1; *cons_len2_out = 0;
        if (par.max_n_cons == 2) {
            const int nc = cluster(msa, n_seq, ml, min_w, cons, read_clu, cons_len_out, cons_len2_out);
            if (nc < 0) return nc;
            *n_cons_out = nc;
        }
        return ST_OK;
    }

    // ---- whole problem ---------------------------------------------------------------------------
    __device__ void run(const KernelArgs &a, const Problem &pb, DevResult *res, int32_t *arena) {
        par = pb.par; n_reads = pb.n_reads; cells = 0; t_dp = t_bt = t_add = t_after = t_fin = t_seg = t_gen = n_seg = n_gen = t_pro = 0;
        oe1 = par.gap_open1 + par.gap_ext1; oe2 = par.gap_open2 + par.gap_ext2;
        int status = ST_OK, cons_len = 0, msa_len = 0, n_cons = 0, cons_len2 = 0; unsigned long long msa_off = 0;
        {
            const int m1 = INT16_MIN + par.mismatch, m2 = INT16_MIN + oe1, m3 = INT16_MIN + oe2;
            int m = m1 > m2 ? m1 : m2; if (m3 > m) m = m3;
            inf_min = (int16_t)(m + 512 * (par.gap_ext1 > par.gap_ext2 ? par.gap_ext1 : par.gap_ext2));
        }
        L::sync();
        const int ncap = a.worst_case ? pb.sum_len + 34 : pb.node_cap, ecap = a.worst_case ? 3 * (pb.sum_len + pb.n_reads) + 64 : pb.edge_cap;
        if (!carve(arena, a.arena_words, ncap, ecap, pb.max_len, pb.n_reads) || (par.max_n_cons != 1 && !(par.max_n_cons == 2 && !par.sub_aln && a.read_clu))) status = ST_OOM;
        else {
            w.in_top = w.out_top = w.aln_top = 0; w.n_nodes = 0; w.oom = 0;
            if (L::tid() == 0) { add_node(0); add_node(0); w.next[0] = 1; w.next[1] = -1; }
            else w.n_nodes = 2;
            L::sync();
            const uint8_t *seqs = a.seqs + pb.seq_base;
            const int32_t *sbp = a.sub_beg ? a.sub_beg + pb.read_first : nullptr, *sep = a.sub_end ? a.sub_end + pb.read_first : nullptr;
            bool any_sub = false;                  // a problem with partially covering reads keeps the reference's BFS node index up to date
            if (sbp) for (int r = 1; r < pb.n_reads; ++r) if (sbp[r] > 0) any_sub = true;
            for (int r = 0; r < pb.n_reads && status == ST_OK; ++r) {
                if (r > 0 && sbp && sbp[r] < 0) continue;        // left out of the graph (src/align.c:800)
                const uint8_t *q = seqs + a.read_off[pb.read_first + r];
                const int ql = a.read_len[pb.read_first + r];
                int n_cig = 0;
                const int first_read = (w.n_nodes == 2);
                beg_id = 0; end_id = 1; n_rows = w.n_nodes; sub = false;
                ai_pool = w.in_pool; ai_off = w.in_off; ai_n = w.in_n;
                const uint32_t dp_cap0 = w.dp_capacity;
                if (!first_read && r > 0 && sbp && sbp[r] > 0) {
#ifdef LCD_EMU
                    status = ST_SUB_UNSUPPORTED; break;      // (the sub-graph alignment lives in the warp policy)
#else
                    if (!L::STRIP) { status = ST_SUB_UNSUPPORTED; break; }
#endif
                    int taken;
                    { LCD_T0(); taken = prepare_sub(sbp[r], sep[r]); L::sync(); LCD_T1(t_pro); }
                    if (taken < 0) { status = taken; break; }
                    sub = true;
                    if ((uint32_t)taken + 1024 > w.dp_capacity) { status = ST_OOM; break; }
                    ai_pool = reinterpret_cast<const int4 *>(w.dp + ((size_t)w.dp_capacity & ~(size_t)7)) - w.in_top - 8; ai_off = w.s2; ai_n = w.s3;
                    w.dp_capacity -= (uint32_t)taken;
                }
                if (!first_read) {
                    const int gn = sub ? w.maxl[end_id] - w.maxl[beg_id] + 1 : w.n_nodes, len = ql > gn ? ql : gn;
                    const int ms = (ql * par.match > len * par.gap_ext1 + par.gap_open1) ? ql * par.match : len * par.gap_ext1 + par.gap_open1;
                    if (!(ms <= INT16_MAX - par.mismatch - oe1 - oe2)) { status = ST_INT32; break; }
#ifndef LCD_EMU
                    { LCD_T0(); ring_start(n_rows); n_cig = align(q, ql); ring_drain(); LCD_T1(t_dp); }
#else
                    n_cig = align(q, ql);
#endif
                    w.dp_capacity = dp_cap0;
                    if (n_cig < 0) { status = n_cig; break; }
                }
                bool fused = false;
#ifndef LCD_EMU
                if constexpr (L::STRIP) {
                    LCD_T0(); add_alignment_warp(q, ql, reinterpret_cast<int *>(w.cigar), r, first_read != 0); LCD_T1(t_add);
                    fused = true;
                }
#endif
                if (!fused) {
                { LCD_T0();
                if (L::tid() == 0) {
                    add_alignment(q, ql, n_cig, r);
                    w.tmp[0] = w.n_nodes; w.tmp[1] = w.oom;
                }
                L::sync(); LCD_T1(t_add); }
                w.n_nodes = w.tmp[0]; w.oom = w.tmp[1];
                }
                // pool cursors are only used by lane 0; n_nodes and oom are shared through tmp[]
                L::sync();
                if (w.oom) { status = ST_OOM; break; }
                if (w.n_nodes > 2) {
                    // the BFS index is needed when this read went against a sub-graph (the nodes it spans) or the next read that joins the graph will
                    bool bfs = sub;
                    if (any_sub && !bfs) for (int r2 = r + 1; r2 < pb.n_reads; ++r2) if (sbp[r2] >= 0) { bfs = sbp[r2] > 0; break; }
                    LCD_T0(); after_add(first_read, bfs, sub); LCD_T1(t_after);
                }
            }
            if (status == ST_OK && w.n_nodes > 2) {
                LCD_T0();
                status = finish(pb.n_reads, a.cons + pb.cons_off, &cons_len, a, &msa_off, &msa_len, pb.min_w, a.read_clu ? a.read_clu + pb.read_first : nullptr, &n_cons, &cons_len2);
                LCD_T1(t_fin);
            }
        }
        if (L::tid() == 0) {
            DevResult r;
            r.status = status; r.cons_len = cons_len; r.msa_len = msa_len; r.n_nodes = w.n_nodes; r.msa_off = msa_off;
            r.cells_lo = (uint32_t)cells; r.cells_hi = (uint32_t)(cells >> 32);
            r.n_cons = n_cons; r.cons_len2 = cons_len2;
#ifdef LCD_POA_TIMING
            r.t_dp = t_dp - t_bt; r.t_bt = t_bt; r.t_add = t_add; r.t_after = t_after; r.t_fin = t_fin; r.t_seg = t_seg; r.t_gen = t_gen; r.n_seg = n_seg; r.n_gen = n_gen; r.t_pro = t_pro;
#endif
            *res = r;
        }
        L::sync();
    }
};

} // namespace poa
} // namespace lcd
