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

namespace lcd {
namespace poa {

constexpr int PN = 32;
constexpr int LOGN = 5;
constexpr int GARBAGE = 0x5555;        // what a read of a never-written DP cell returns (never decisive)

enum { ST_OK = 0, ST_INT32 = -1, ST_BAND = -2, ST_BACKTRACK = -3, ST_NOBASE = -4, ST_OOM = -5 };

struct __align__(16) Problem {
    uint64_t seq_base;        // byte offset of the first read in the packed sequence buffer
    int32_t read_first;       // index of the first read in read_off[] / read_len[]
    int32_t n_reads;
    int32_t sum_len, max_len;
    int32_t cons_off;         // byte offset in the consensus buffer (capacity sum_len)
    int32_t pad;
    lcd_poa_params_t par;
};

struct __align__(16) DevResult {
    int32_t status;
    int32_t cons_len;
    int32_t msa_len;
    int32_t n_nodes;
    uint64_t msa_off;         // byte offset in the MSA output pool
    uint32_t cells_lo, cells_hi;
};

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
};

// per-group workspace carved from the arena for one problem
struct WS {
    int N;                        // node capacity
    int *base, *in_off, *in_n, *in_cap, *out_off, *out_n, *out_cap, *n_read, *n_span;
    int *aln_off, *aln_n, *aln_cap, *next, *remain, *maxl, *maxr, *msa_rank;
    int *row_off, *dp_beg, *dp_end, *wsum, *order, *tmp;
    int4 *in_pool; int in_top, in_capacity;          // {from, w, ps, -}
    int *out_pool; int out_top, out_capacity, out_stride, rid_w;   // {to, w, rid[2*rid_w]}
    int *aln_pool; int aln_top, aln_capacity;
    int2 *cigar; int cigar_cap;                      // {op | len << 2, node_id}
    int16_t *dp; uint32_t dp_top, dp_capacity;       // in cells
    int n_nodes;
    int oom;
};

// ---------------------------------------------------------------------------------------------
// lane policy for the GPU: one warp, lane l == int16 lane l of the reference's 512-bit vector
#ifndef LCD_EMU
struct WarpLanes {
    typedef int vec;                                  // this lane's cell (int16 value in an int)
    static constexpr int STRIDE = 32;
    __device__ static __forceinline__ int lane() { return threadIdx.x & 31; }
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
#endif

// ---------------------------------------------------------------------------------------------
template <class L> struct Poa {
    typedef typename L::vec vec;
    WS w;
    lcd_poa_params_t par;
    int n_reads;
    int inf_min, oe1, oe2;
    unsigned long long cells;

    // ---- workspace ---------------------------------------------------------------------------
    __device__ bool carve(int32_t *arena, uint64_t words, int sum_len, int max_len, int n_reads_) {
        uint64_t top = 0;
        const int N = sum_len + 2 + 32;
        w.N = N;
        int **arr[] = { &w.base, &w.in_off, &w.in_n, &w.in_cap, &w.out_off, &w.out_n, &w.out_cap, &w.n_read, &w.n_span,
                        &w.aln_off, &w.aln_n, &w.aln_cap, &w.next, &w.remain, &w.maxl, &w.maxr, &w.msa_rank,
                        &w.row_off, &w.dp_beg, &w.dp_end, &w.wsum, &w.order, &w.tmp };
        for (unsigned i = 0; i < sizeof(arr) / sizeof(arr[0]); ++i) { *arr[i] = arena + top; top += (uint64_t)((N + 3) & ~3); }
        const int E = 4 * (sum_len + n_reads_) + 64;
        w.in_capacity = E; w.in_pool = reinterpret_cast<int4 *>(arena + top); top += (uint64_t)E * 4;
        w.rid_w = 1 + ((n_reads_ - 1) >> 6);
        w.out_stride = 2 + 2 * w.rid_w;
        w.out_capacity = E; w.out_pool = arena + top; top += ((uint64_t)E * w.out_stride + 3) & ~3ull;
        w.aln_capacity = E; w.aln_pool = arena + top; top += (uint64_t)E;
        w.cigar_cap = max_len + N + 8; w.cigar = reinterpret_cast<int2 *>(arena + top); top += (uint64_t)w.cigar_cap * 2;
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
                const int nc = w.in_cap[to] ? w.in_cap[to] * 2 : 4;
                if (w.in_top + nc > w.in_capacity) { w.oom = 1; return; }
                int4 *src = w.in_pool + w.in_off[to], *dst = w.in_pool + w.in_top;
                for (int i = 0; i < w.in_n[to]; ++i) dst[i] = src[i];
                w.in_off[to] = w.in_top; w.in_top += nc; w.in_cap[to] = nc;
            }
            w.in_pool[w.in_off[to] + w.in_n[to]] = make_int4(from, 1, 0, 0);
            w.in_n[to]++;
            if (w.out_n[from] == w.out_cap[from]) {
                const int nc = w.out_cap[from] ? w.out_cap[from] * 2 : 4;
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
            const int nc = w.aln_cap[node] ? w.aln_cap[node] * 2 : 4;
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

    // after fusing a read (abpoa_topological_sort :322-357 + abpoa_update_node_n_span_reads :559-571):
    // per-node exchange sort of the edge lists, out-weight sums, edge path scores (abpoa_get_incre_path_score
    // :429-437), n_span, band seeds, flattened topological order and heaviest-path remain values.
    __device__ void after_add(int first_read) {
        const int n = w.n_nodes, lane = L::lane();
        const int inc = par.sub_aln ? 0 : 1;
        for (int i = lane; i < n; i += L::STRIDE) {
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
            w.maxr[i] = 0; w.maxl[i] = n;
            if (first_read || inc || i >= 2) w.n_span[i] += 1;
        }
        L::sync();
        for (int i = lane; i < n; i += L::STRIDE) {
            int4 *ie = w.in_pool + w.in_off[i];
            for (int k = 0; k < w.in_n[i]; ++k) {
                const int node_w = w.wsum[ie[k].x], edge_w = ie[k].y;
                int ps = 0;
                if (node_w != 0 && edge_w != 0) { ps = (int)round(log((double)edge_w / (double)node_w)); if (ps < -20) ps = -20; }
                ie[k].z = ps;
            }
        }
        if (lane == 0) {
            int cnt = 0;
            for (int id = 0; id != -1; id = w.next[id]) w.order[cnt++] = id;
            if (par.wb >= 0) {                      // abpoa_BFS_set_node_remain :268-309
                w.remain[1] = -1;
                for (int i = cnt - 1; i >= 0; --i) {
                    const int id = w.order[i];
                    if (id == 1) continue;
                    // heaviest out edge = first entry after the descending exchange sort
                    w.remain[id] = (w.out_n[id] > 0 ? w.remain[out_entry(id, 0)[0]] : -1) + 1;
                }
            }
        }
        L::sync();
    }

    // ---- DP ---------------------------------------------------------------------------------
    __device__ int nvec(int id) const { return (w.dp_end[id] >> 5) - (w.dp_beg[id] >> 5) + 2; }
    // plane p (0 H, 1 E1, 2 E2, 3 F1, 4 F2) of row `id`, addressed by absolute column
    __device__ const int16_t *plane(int id, int p) const { return w.dp + w.row_off[id] + (size_t)p * nvec(id) * PN - (size_t)(w.dp_beg[id] >> 5) * PN; }
    __device__ int cell(int id, int p, int col) const {       // bounds-checked scalar read (backtrack)
        const int lo = (w.dp_beg[id] >> 5) * PN, hi = lo + nvec(id) * PN;
        if (col < lo || col >= hi) return GARBAGE;
        return plane(id, p)[col];
    }
    __device__ bool alloc_row(int id) {
        const uint32_t need = (uint32_t)nvec(id) * PN * 5;
        if (w.dp_top + need > w.dp_capacity) { w.oom = 1; return false; }
        w.row_off[id] = (int)w.dp_top; w.dp_top += need;
        return true;
    }
    // SIMD_SET_F, abpoa_align_simd.c:691-725
    __device__ vec set_f(vec F, int set_num, int e) const {
        int cov = set_num;
#pragma unroll
        for (int s = 0; s < LOGN; ++s) {
            const int sh = 1 << s;
            if (set_num != PN && s > 0) cov += sh;
            vec t = L::shift_up(L::sub(F, L::set1((int16_t)(e << s))), sh, inf_min);
            if (set_num != PN) t = L::keep(t, 0, cov < PN - 1 ? cov : PN - 1, inf_min);
            F = L::vmax(F, t);
        }
        return F;
    }

    // one sequence against the whole graph: simd_abpoa_cg_align_sequence_to_graph_core (:1200-1228)
    // returns number of cigar entries (reverse order) or <0
    __device__ int align(const uint8_t *query, int qlen) {
        const int n = w.n_nodes;
        const int dp_sn = (qlen + 1 + PN - 1) / PN;
        const int wband = par.wb < 0 ? qlen : par.wb + (int)(par.wf * qlen);
        const int o1 = par.gap_open1, e1 = par.gap_ext1, o2 = par.gap_open2, e2 = par.gap_ext2;
        const int match = par.match, mism = par.mismatch;
        w.dp_top = 0;
        const int rem_end = par.wb >= 0 ? w.remain[1] : 0;
        // ---- first row (SRC) :627-688
        {
            if (par.wb >= 0) {
                L::sync();
                if (L::lane() == 0) {
                    w.maxl[0] = w.maxr[0] = 0;
                    for (int i = 0; i < w.out_n[0]; ++i) { const int o = out_entry(0, i)[0]; w.maxl[o] = w.maxr[o] = 1; }
                }
                L::sync();
                const int r = qlen - (w.remain[0] - rem_end - 1);
                int e = (w.maxr[0] > r ? w.maxr[0] : r) + wband; if (e > qlen) e = qlen;
                w.dp_beg[0] = 0; w.dp_end[0] = e;
            } else { w.dp_beg[0] = 0; w.dp_end[0] = qlen; }
            L::sync();
            if (!alloc_row(0)) return ST_OOM;
            const int end0 = w.dp_end[0], nv = nvec(0);
            int16_t *h = const_cast<int16_t *>(plane(0, 0)), *pe1 = const_cast<int16_t *>(plane(0, 1)), *pe2 = const_cast<int16_t *>(plane(0, 2));
            int16_t *pf1 = const_cast<int16_t *>(plane(0, 3)), *pf2 = const_cast<int16_t *>(plane(0, 4));
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
        }
        L::sync();
        // ---- rows in topological (list) order, SINK excluded
        for (int oi = 1; oi < n; ++oi) {
            const int id = w.order[oi];
            if (id == 1) continue;
            const int nin = w.in_n[id];
            const int4 *ie = w.in_pool + w.in_off[id];
            int beg, end, beg_sn, end_sn, min_pre_beg_sn, max_pre_end_sn;
            if (par.wb < 0) { beg = 0; end = qlen; beg_sn = 0; end_sn = end >> 5; min_pre_beg_sn = 0; max_pre_end_sn = end_sn; }
            else {
                const int r = qlen - (w.remain[id] - rem_end - 1);
                beg = (w.maxl[id] < r ? w.maxl[id] : r) - wband; if (beg < 0) beg = 0;
                end = (w.maxr[id] > r ? w.maxr[id] : r) + wband; if (end > qlen) end = qlen;
                beg_sn = beg >> 5;
                int min_pre_beg = INT32_MAX; min_pre_beg_sn = INT32_MAX; max_pre_end_sn = -1;
                for (int k = 0; k < nin; ++k) {
                    const int p = ie[k].x;
                    if (min_pre_beg > w.dp_beg[p]) { min_pre_beg = w.dp_beg[p]; min_pre_beg_sn = w.dp_beg[p] >> 5; }
                    if (max_pre_end_sn < (w.dp_end[p] >> 5)) max_pre_end_sn = w.dp_end[p] >> 5;
                }
                if (beg_sn < min_pre_beg_sn) { beg = min_pre_beg; beg_sn = min_pre_beg_sn; }
                end_sn = end >> 5;
            }
            L::sync();                       // everyone has read the predecessors' band before this row's is published
            w.dp_beg[id] = beg; w.dp_end[id] = end;
            L::sync();
            if (!alloc_row(id)) return ST_OOM;
            if (beg_sn < min_pre_beg_sn) return ST_BAND;
            cells += (unsigned long long)(end - beg + 1);
            int16_t *H = const_cast<int16_t *>(plane(id, 0)), *E1 = const_cast<int16_t *>(plane(id, 1)), *E2 = const_cast<int16_t *>(plane(id, 2));
            int16_t *F1 = const_cast<int16_t *>(plane(id, 3)), *F2 = const_cast<int16_t *>(plane(id, 4));
            const int nb = w.base[id];
            int first1 = 0, first2 = 0;
            int mx = inf_min, left = -1, right = -1;
            for (int sn = beg_sn; sn <= end_sn; ++sn) {
                const int col0 = sn * PN;
                vec h = L::set1(inf_min), ve1 = L::set1(inf_min), ve2 = L::set1(inf_min);
                for (int k = 0; k < nin; ++k) {
                    const int p = ie[k].x, ps = ie[k].z;
                    const int pre_end = w.dp_end[p], pre_beg_sn = w.dp_beg[p] >> 5, pre_end_sn = pre_end >> 5;
                    int bsn, esn;
                    const bool from_mem = pre_beg_sn < beg_sn;
                    bsn = from_mem ? beg_sn : pre_beg_sn;
                    esn = (pre_end + 1) >> 5; if (esn > end_sn) esn = end_sn; if (esn > dp_sn - 1) esn = dp_sn - 1;
                    if (sn >= bsn && sn <= esn) {
                        const int16_t *ph = plane(p, 0);
                        const int first = (sn == bsn && !from_mem) ? inf_min : (int)ph[col0 - 1];
                        const vec v = L::add(L::load_m1(ph + col0, first), L::set1(ps));
                        h = k == 0 ? v : L::vmax(v, h);
                    }
                    esn = pre_end_sn < end_sn ? pre_end_sn : end_sn;
                    if (sn >= bsn && sn <= esn) {
                        const vec v1 = L::add(L::load(plane(p, 1) + col0), L::set1(ps)), v2 = L::add(L::load(plane(p, 2) + col0), L::set1(ps));
                        ve1 = k == 0 ? v1 : L::vmax(v1, ve1);
                        ve2 = k == 0 ? v2 : L::vmax(v2, ve2);
                    }
                }
                // + query profile, band mask
                const vec q = L::map_cols(col0, [&](int j) -> int {
                    if (j == 0 || j > qlen) return 0;
                    const int qb = query[j - 1];
                    return (nb > 3 || qb > 3) ? 0 : (nb == qb ? match : -mism); });
                h = L::add(h, q);
                const int klo = beg - col0, khi = end - col0;       // lanes inside the band
                h = L::keep(h, klo, khi, inf_min); ve1 = L::keep(ve1, klo, khi, inf_min); ve2 = L::keep(ve2, klo, khi, inf_min);
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
                if (sn == end_sn) { h = L::keep(h, -1, khi, inf_min); ve1 = L::keep(ve1, -1, khi, inf_min); ve2 = L::keep(ve2, -1, khi, inf_min); }
                ve1 = L::vmax(L::sub(ve1, L::set1(e1)), L::sub(h, L::set1(oe1)));
                ve2 = L::vmax(L::sub(ve2, L::set1(e2)), L::sub(h, L::set1(oe2)));
                L::store(H + col0, h); L::store(E1 + col0, ve1); L::store(E2 + col0, ve2); L::store(F1 + col0, f1); L::store(F2 + col0, f2);
                if (par.wb >= 0) {           // simd_abpoa_max_in_row :1107-1119
                    int m, fi, la;
                    if (L::row_max(h, klo < 0 ? 0 : klo, khi > PN - 1 ? PN - 1 : khi, m, fi, la)) {
                        if (m > mx) { mx = m; left = col0 + fi; right = col0 + la; }
                        else if (m == mx) right = col0 + la;
                    }
                }
            }
            // the vector after the band: INF_MIN in H (read by successors' M), undefined elsewhere
            if (end_sn + 1 <= dp_sn - 1) L::store(H + (end_sn + 1) * PN, L::set1(inf_min));
            if (par.wb >= 0) {               // simd_abpoa_ada_max_i :1121-1130
                L::sync();
                if (L::lane() == 0) {
                    for (int i = 0; i < w.out_n[id]; ++i) {
                        const int o = out_entry(id, i)[0];
                        if (right + 1 > w.maxr[o]) w.maxr[o] = right + 1;
                        if (left + 1 < w.maxl[o]) w.maxl[o] = left + 1;
                    }
                }
            }
            L::sync();
        }
        // ---- best cell :1092-1105 and backtrack :309-458 (lane 0; result broadcast through memory)
        int n_cig = 0;
        if (L::lane() == 0) n_cig = backtrack(query, qlen);
        L::sync();
        n_cig = w.tmp[0];
        return n_cig;
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
            const int4 *ie = w.in_pool + w.in_off[1];
            for (int k = 0; k < w.in_n[1]; ++k) {
                const int r = ie[k].x;
                const int e = qlen > w.dp_end[r] ? w.dp_end[r] : qlen;
                const int s = cell(r, 0, e);
                if (s > best) { best = s; bi = r; bj = e; }
            }
        }
        enum { M_OP = 1, E1_OP = 2, E2_OP = 4, E_OP = 6, F1_OP = 8, F2_OP = 16, F_OP = 24, ALL_OP = 31 };
        int n = 0, id = bi, j = bj, cur_op = ALL_OP, rc = 0;
        if (bj < qlen) cig_push(n, 1, qlen - bj, -1);
        while (id != 0 && j > 0) {
            const int nb = w.base[id], qb = query[j - 1];
            const int s = (nb > 3 || qb > 3) ? 0 : (nb == qb ? par.match : -par.mismatch);
            const int4 *ie = w.in_pool + w.in_off[id];
            const int nin = w.in_n[id];
            const int hj = cell(id, 0, j);
            int hit = 0;
            for (int pass = 0; pass < 2 && !hit; ++pass) {
                if (pass == 1) {
                    if (cur_op & E_OP) {
                        const int e1j = cell(id, 1, j), e2j = cell(id, 2, j);
                        for (int k = 0; k < nin && !hit; ++k) {
                            const int p = ie[k].x, ps = ie[k].z;
                            if (j < w.dp_beg[p] || j > w.dp_end[p]) continue;
                            const int phj = cell(p, 0, j);
                            if (cur_op & E1_OP) {
                                const int pe = cell(p, 1, j);
                                const int ok = (cur_op & M_OP) ? (hj == pe + ps) : (e1j == pe - e1 + ps);
                                if (ok) { cur_op = (phj - oe1 == pe) ? (M_OP | F_OP) : E1_OP; hit = 1; }
                            }
                            if (!hit && (cur_op & E2_OP)) {
                                const int pe = cell(p, 2, j);
                                const int ok = (cur_op & M_OP) ? (hj == pe + ps) : (e2j == pe - e2 + ps);
                                if (ok) { cur_op = (phj - oe2 == pe) ? (M_OP | F_OP) : E2_OP; hit = 1; }
                            }
                            if (hit) { cig_push(n, 2, 1, id); id = p; }
                        }
                    }
                    if (!hit && (cur_op & F_OP)) {
                        const int hj1 = cell(id, 0, j - 1);
                        if (cur_op & F1_OP) {
                            const int f = cell(id, 3, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe1 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(id, 3, j - 1) - e1 == f) { cur_op = F1_OP; hit = 1; }
                            }
                        }
                        if (!hit && (cur_op & F2_OP)) {
                            const int f = cell(id, 4, j);
                            if (!(cur_op & M_OP) || hj == f) {
                                if (hj1 - oe2 == f) { cur_op = M_OP | E_OP; hit = 1; }
                                else if (cell(id, 4, j - 1) - e2 == f) { cur_op = F2_OP; hit = 1; }
                            }
                        }
                        if (hit) { cig_push(n, 1, 1, id); --j; }
                    }
                    if (hit) break;
                }
                if (cur_op & M_OP) {
                    for (int k = 0; k < nin; ++k) {
                        const int p = ie[k].x, ps = ie[k].z;
                        if (j - 1 < w.dp_beg[p] || j - 1 > w.dp_end[p]) continue;
                        if (cell(p, 0, j - 1) + s + ps == hj) {
                            cig_push(n, 0, 1, id);
                            id = p; --j; hit = 1; cur_op = ALL_OP;
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

    // ---- output: MSA ranks (abpoa_DFS_set_msa_rank :359-410), most-frequent consensus, RC-MSA ----------
    __device__ int finish(int n_seq, uint8_t *cons, int *cons_len_out, const KernelArgs &a, unsigned long long *msa_off_out, int *msa_len_out) {
        const int n = w.n_nodes, lane = L::lane();
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
        for (int i = lane; i < ml * 5; i += L::STRIDE) { cnt[i] = 0; nid[i] = 0; }
        for (int i = lane; i < n; i += L::STRIDE) {
            int r = w.msa_rank[i];
            for (int j = 0; j < w.aln_n[i]; ++j) { const int rr = w.msa_rank[w.aln_pool[w.aln_off[i] + j]]; if (rr > r) r = rr; }
            col[i] = r - 1;
        }
        L::sync();
        for (int i = 2 + lane; i < n; i += L::STRIDE) { cnt[col[i] * 5 + w.base[i]] = w.n_read[i]; nid[col[i] * 5 + w.base[i]] = i; }
        L::sync();
        // voting (abpoa_set_major_voting_cons :393-424); ordered compaction by lane 0 over per-column flags
        int *emit = w.maxl;        // per column: consensus node id or -1   (ml <= N)
        int bad = 0;
        for (int i = lane; i < ml; i += L::STRIDE) {
            int max_c = 0, total = 0, max_base = 5;
            for (int j = 0; j < 4; ++j) { const int c = cnt[i * 5 + j]; if (c > max_c) { max_c = c; max_base = j; } total += c; }
            if (max_base == 5) { bad = 1; emit[i] = -1; continue; }
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
            const unsigned long long bytes = ((unsigned long long)(n_seq + 1) * ml + 15) & ~15ull;
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
        for (long long i = lane; i < tot; i += L::STRIDE) msa[i] = 5;
        L::sync();
        for (int i = 2 + lane; i < n; i += L::STRIDE) {
            const int c = col[i], b = w.base[i];
            for (int e = 0; e < w.out_n[i]; ++e) {
                const int *ent = out_entry(i, e);
                for (int x = 0; x < 2 * w.rid_w; ++x) {
                    unsigned bits = (unsigned)ent[2 + x];
                    while (bits) { const int r = __ffs(bits) - 1; bits &= bits - 1; msa[(size_t)(x * 32 + r) * ml + c] = (uint8_t)b; }
                }
            }
        }
        for (int i = lane; i < cl; i += L::STRIDE) msa[(size_t)n_seq * ml + col[cons_ids[i]]] = cons[i];
        L::sync();
        return ST_OK;
    }

    // ---- whole problem ---------------------------------------------------------------------------
    __device__ void run(const KernelArgs &a, const Problem &pb, DevResult *res, int32_t *arena) {
        par = pb.par; n_reads = pb.n_reads; cells = 0;
        oe1 = par.gap_open1 + par.gap_ext1; oe2 = par.gap_open2 + par.gap_ext2;
        int status = ST_OK, cons_len = 0, msa_len = 0; unsigned long long msa_off = 0;
        {
            const int m1 = INT16_MIN + par.mismatch, m2 = INT16_MIN + oe1, m3 = INT16_MIN + oe2;
            int m = m1 > m2 ? m1 : m2; if (m3 > m) m = m3;
            inf_min = (int16_t)(m + 512 * (par.gap_ext1 > par.gap_ext2 ? par.gap_ext1 : par.gap_ext2));
        }
        L::sync();
        if (!carve(arena, a.arena_words, pb.sum_len, pb.max_len, pb.n_reads) || par.max_n_cons != 1) status = ST_OOM;
        else {
            w.in_top = w.out_top = w.aln_top = 0; w.n_nodes = 0; w.oom = 0; w.dp_top = 0;
            if (L::lane() == 0) { add_node(0); add_node(0); w.next[0] = 1; w.next[1] = -1; }
            else w.n_nodes = 2;
            L::sync();
            const uint8_t *seqs = a.seqs + pb.seq_base;
            for (int r = 0; r < pb.n_reads && status == ST_OK; ++r) {
                const uint8_t *q = seqs + a.read_off[pb.read_first + r];
                const int ql = a.read_len[pb.read_first + r];
                int n_cig = 0;
                const int first_read = (w.n_nodes == 2);
                if (!first_read) {
                    const int gn = w.n_nodes, len = ql > gn ? ql : gn;
                    const int ms = (ql * par.match > len * par.gap_ext1 + par.gap_open1) ? ql * par.match : len * par.gap_ext1 + par.gap_open1;
                    if (!(ms <= INT16_MAX - par.mismatch - oe1 - oe2)) { status = ST_INT32; break; }
                    n_cig = align(q, ql);
                    if (n_cig < 0) { status = n_cig; break; }
                }
                if (L::lane() == 0) {
                    add_alignment(q, ql, n_cig, r);
                    w.tmp[0] = w.n_nodes; w.tmp[1] = w.oom;
                }
                L::sync();
                w.n_nodes = w.tmp[0]; w.oom = w.tmp[1];
                // pool cursors are only used by lane 0; n_nodes and oom are shared through tmp[]
                L::sync();
                if (w.oom) { status = ST_OOM; break; }
                if (w.n_nodes > 2) after_add(first_read);
            }
            if (status == ST_OK && w.n_nodes > 2)
                status = finish(pb.n_reads, a.cons + pb.cons_off, &cons_len, a, &msa_off, &msa_len);
        }
        if (L::lane() == 0) {
            DevResult r;
            r.status = status; r.cons_len = cons_len; r.msa_len = msa_len; r.n_nodes = w.n_nodes; r.msa_off = msa_off;
            r.cells_lo = (uint32_t)cells; r.cells_hi = (uint32_t)(cells >> 32);
            *res = r;
        }
        L::sync();
    }
};

} // namespace poa
} // namespace lcd
