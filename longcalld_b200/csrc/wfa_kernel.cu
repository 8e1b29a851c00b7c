// wfa_kernel.cu -- K6: gap-affine / gap-affine-2p end-to-end wavefront alignment on sm_100a.
//
// What it replaces: wavefront_aligner_new + wavefront_align + wf_aligner->cigar as driven by
// wfa_end2end_aln (reference src/align.c:374-460) and is_diff_between_ref_hap_aln
// (src/assign_hap.c:972), i.e. WFA2-lib's wavefront_unialign loop
// (WFA2-lib/wavefront/wavefront_unialign.c:242-275) in memory_high mode with heuristic
// none / wf-adaptive / z-drop and the full backtrace (wavefront_backtrace.c:320-539).
//
// B200 design (not a translation of the CPU code):
//   * one *group* of threads per alignment problem; two group shapes run concurrently on two
//     streams: a warp (32 lanes, __syncwarp only, several problems per CTA) for the thousands of
//     small problems and a whole CTA (256 threads) for the few kilobase-scale ones.  Grids are
//     persistent (a multiple of the SM count) and pull problems, sorted by size, from a device queue.
//   * compute, trim and match-extension of score s are ONE fused pass over the diagonals: each lane
//     owns diagonals lo+lane, lo+lane+G, ... so every wavefront load/store is a coalesced int32 row;
//     the five output components of a score live side by side in one 16-byte aligned slab.
//   * the descriptors of the last 32 scores (the recurrence looks back at most o2+e2 = 25) sit in
//     a shared-memory ring; sequences are staged in shared memory when they fit; match extension
//     compares 4 bases per step with funnel-shifted 32-bit words against sentinel padding.
//   * all wavefronts are retained (the reference's backtrace needs them) in a per-group HBM arena
//     that is recycled from problem to problem, so the working set of the many small problems stays
//     in the 126 MB L2; very large problems spill to a shared overflow pool (atomic bump).
//   * lo/hi trimming, termination and the heuristics are warp-shuffle (redux.sync) reductions.
//   * backtrace is run by warp 0 of the group with all lanes in lock-step (uniform loads are
//     broadcasts) so that runs of matches are written 32 bytes at a time.
#include "lcd_common.cuh"
#include "wfa_device.cuh"
#include <algorithm>
#include <numeric>

namespace lcd {
namespace wfa {

// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(G == 32 ? 32 * WARP_GROUPS_PER_CTA : G)
wfa_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int GROUPS = (G == 32) ? WARP_GROUPS_PER_CTA : 1;
    constexpr int SEQ_BYTES = (G == 32) ? WARP_SEQ_SMEM : CTA_SEQ_SMEM;
    WfSet *rings = reinterpret_cast<WfSet *>(smem);
    int *red = reinterpret_cast<int *>(smem + sizeof(WfSet) * RING * GROUPS);
    uint8_t *seqbuf = smem + sizeof(WfSet) * RING * GROUPS + ((G == 32) ? 16 : 2 * (G / 32) * 16 * 4);
    __shared__ uint32_t next_item[GROUPS];

    const int gi = (G == 32) ? (threadIdx.x >> 5) : 0;
    const int group_id = blockIdx.x * GROUPS + gi;
    Aligner<G> al(a);
    al.g.lane = (G == 32) ? (threadIdx.x & 31) : threadIdx.x;
    al.g.ring = rings + gi * RING;
    al.g.red = red;
    al.g.phase = 0;
    al.pool = a.pool;
    al.gmeta = a.meta + (size_t)group_id * a.meta_cap;
    const uint32_t arena_lo = a.arena_base + (uint32_t)group_id * a.arena_units;
    const uint32_t arena_hi = arena_lo + a.arena_units;
    const uint32_t n_items = a.n_dev ? *a.n_dev : (uint32_t)a.n;
    for (;;) {
        if (al.g.lane == 0) next_item[gi] = atomicAdd(a.queue, 1u);
        al.g.sync();
        const uint32_t item = next_item[gi];
        al.g.sync();
        if (item >= n_items) break;
        const int pi = a.order[item];
        al.align(a.problems[pi], a.results + pi, seqbuf + gi * SEQ_BYTES, SEQ_BYTES, arena_lo, arena_hi, pi);
    }
}

static size_t smem_bytes(int G) {
    const int groups = (G == 32) ? WARP_GROUPS_PER_CTA : 1;
    const int seq = (G == 32) ? WARP_SEQ_SMEM : CTA_SEQ_SMEM;
    return sizeof(WfSet) * RING * groups + ((G == 32) ? 16 : 2 * (G / 32) * 16 * 4) + (size_t)seq * groups;
}

// ------------------------------------------------------------------------------------------
// host side: plan = packed sequences + descriptors resident in HBM
static inline int gap_cost(const lcd_wfa_params_t &p, long len) {
    if (len <= 0) return 0;
    long c1 = p.gap_open1 + (long)p.gap_ext1 * len;
    if (p.affine2p) { long c2 = p.gap_open2 + (long)p.gap_ext2 * len; if (c2 < c1) c1 = c2; }
    return (int)std::min<long>(c1, INT32_MAX / 4);
}
// Upper bound on the optimal score (+ slack): either "delete everything, insert everything" or
// "gap to the target diagonal, then mismatches"; only the second survives band heuristics, which
// never prune the target diagonal (wavefront_heuristic.c:232-255).
static int score_cap(const lcd_wfa_params_t &p, int plen, int tlen) {
    const long a = (long)gap_cost(p, plen) + gap_cost(p, tlen);
    const long b = (long)gap_cost(p, std::abs(tlen - plen)) + (long)p.mismatch * std::min(plen, tlen);
    long cap = (p.heuristic == LCD_WFA_HEUR_NONE) ? std::min(a, b) : b;
    cap += 64;
    return (int)std::min<long>(cap, INT32_MAX / 2);
}

constexpr int ESC_SCORE = 160;      // warp groups hand problems over to the CTA kernel at this score

struct WfaPlan : Plan {
    int pool_window() const override { return 1; }
    DevBuf<uint8_t> d_seqs;
    DevBuf<Problem> d_problems;
    DevBuf<int32_t> d_order_small, d_order_large;
    DevBuf<char> d_ops;
    DevBuf<DevResult> d_results;
    DevBuf<uint32_t> d_queue;     // [0] small, [1] large, [2] escalated, [3] number of escalated problems
    DevBuf<int32_t> d_esc_list;
    std::vector<Problem> problems;
    std::vector<int64_t> ops_dev_off;
    int n_small = 0, n_large = 0;
    int cap_small = 0, cap_large = 0;
    size_t ops_bytes = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<DevResult> h_results;
    std::vector<char> h_ops;

    ~WfaPlan() override {
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }

    int build(int n_, const uint8_t *seqs, size_t seqs_len, const int64_t *pat_off, const int32_t *plen,
              const int64_t *txt_off, const int32_t *tlen, const lcd_wfa_params_t *params) {
        n = n_;
        Context &c = ctx();
        problems.resize(n);
        ops_dev_off.resize(n);
        size_t seq_bytes = 0, ops_total = 0;
        for (int i = 0; i < n; ++i) {
            if (plen[i] < 0 || tlen[i] < 0 || pat_off[i] < 0 || txt_off[i] < 0 ||
                (size_t)pat_off[i] + plen[i] > seqs_len || (size_t)txt_off[i] + tlen[i] > seqs_len) {
                set_error("lcd_wfa: problem %d has an invalid sequence range", i); return -1;
            }
            Problem &p = problems[i];
            memset(&p, 0, sizeof(p));
            p.plen = plen[i]; p.tlen = tlen[i]; p.par = params[i];
            p.pat = seq_bytes; seq_bytes += ((size_t)plen[i] + 12 + 15) & ~(size_t)15;
            p.txt = seq_bytes; seq_bytes += ((size_t)tlen[i] + 12 + 15) & ~(size_t)15;
            p.ops = ops_total; ops_dev_off[i] = (int64_t)ops_total;
            ops_total += (2 * ((size_t)plen[i] + tlen[i]) + 8 + 15) & ~(size_t)15;
            p.s_cap = score_cap(params[i], plen[i], tlen[i]);
        }
        ops_bytes = ops_total;
        // pack sequences with sentinel padding ('!' after the pattern, '?' after the text:
        // wavefront_sequences.c:37-39)
        std::vector<uint8_t> packed(seq_bytes + 16);
        for (int i = 0; i < n; ++i) {
            const Problem &p = problems[i];
            uint8_t *dp = packed.data() + p.pat, *dt = packed.data() + p.txt;
            memcpy(dp, seqs + pat_off[i], p.plen);
            memset(dp + p.plen, '!', (((size_t)p.plen + 12 + 15) & ~(size_t)15) - p.plen);
            memcpy(dt, seqs + txt_off[i], p.tlen);
            memset(dt + p.tlen, '?', (((size_t)p.tlen + 12 + 15) & ~(size_t)15) - p.tlen);
        }
        // classes + processing order (largest first)
        std::vector<int32_t> small, large;
        for (int i = 0; i < n; ++i) {
            const Problem &p = problems[i];
            // a length difference alone already costs gap_cost(|d|): such problems never finish below the
            // escalation score, so they go to the CTA kernel directly
            const bool is_small = ((p.plen + 27) & ~15) + ((p.tlen + 27) & ~15) <= WARP_SEQ_SMEM && p.s_cap <= 4096 &&
                                  gap_cost(params[i], std::abs(p.tlen - p.plen)) < ESC_SCORE;
            (is_small ? small : large).push_back(i);
        }
        auto by_size = [&](int32_t x, int32_t y) {
            const long sx = (long)problems[x].plen + problems[x].tlen, sy = (long)problems[y].plen + problems[y].tlen;
            return sx != sy ? sx > sy : x < y;
        };
        std::sort(small.begin(), small.end(), by_size);
        std::sort(large.begin(), large.end(), by_size);
        n_small = (int)small.size(); n_large = (int)large.size();
        for (int32_t i : small) cap_small = std::max(cap_small, problems[i].s_cap);
        for (int32_t i : large) cap_large = std::max(cap_large, problems[i].s_cap);
        cudaStream_t s = cur_stream();
        if (d_seqs.upload(packed.data(), packed.size(), s)) return -1;
        if (d_problems.upload(problems.data(), problems.size(), s)) return -1;
        if (d_order_small.upload(small.data(), small.size(), s)) return -1;
        if (d_order_large.upload(large.data(), large.size(), s)) return -1;
        if (d_ops.alloc(ops_bytes + 16)) return -1;
        if (d_results.alloc(n)) return -1;
        if (d_queue.alloc(4)) return -1;
        if (d_esc_list.alloc(std::max(n_small, 1))) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));      // host staging vectors go out of scope
        LCD_CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        LCD_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        LCD_CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        // pool layout (16-byte units): [meta small | meta large | arenas small | arenas large | overflow]
        const int occ_small = 4;                                   // CTAs per SM of the warp kernel
        int grid_small = n_small ? std::min((n_small + WARP_GROUPS_PER_CTA - 1) / WARP_GROUPS_PER_CTA, c.dp_sms() * occ_small) : 0;
        // CTA groups serve the pre-classified large problems and, afterwards, whatever the warp kernel escalates
        int grid_large = (n_large || n_small) ? c.dp_sms() * (CTA_GROUP_THREADS > 256 ? 1 : 2) : 0;
        const int cap_large = std::max(this->cap_large, n_small ? cap_small : 0);
        const size_t groups_small = (size_t)grid_small * WARP_GROUPS_PER_CTA, groups_large = grid_large;
        const size_t pool_units = win->words / 4;
        const size_t meta_small_units = groups_small * (size_t)cap_small * (sizeof(WfSet) / 16);
        const size_t meta_large_units = groups_large * (size_t)cap_large * (sizeof(WfSet) / 16);
        if (meta_small_units + meta_large_units > pool_units / 2) {
            set_error("lcd_wfa: workspace pool too small for score descriptors (%zu MiB needed)",
                      (meta_small_units + meta_large_units) * 16 >> 20);
            return -1;
        }
        size_t rest = pool_units - meta_small_units - meta_large_units;
        // private arenas: up to 1 MiB per warp group, up to 64 MiB per CTA group, at most half of the rest
        size_t arena_small = groups_small ? std::min<size_t>((1u << 20) / 16, rest / 4 / groups_small) : 0;
        size_t arena_large = groups_large ? std::min<size_t>((64u << 20) / 16, rest / 4 / groups_large) : 0;
        const size_t arenas = arena_small * groups_small + arena_large * groups_large;
        if (pool_units > 0xffffffffull) { set_error("lcd_wfa: pool larger than 64 GiB is not addressable"); return -1; }
        const size_t n_chunks = std::min<size_t>((rest - arenas) / wfa::OVERFLOW_CHUNK_UNITS, (size_t)Context::BITMAP_WORDS * 32);
        KernelArgs ka;
        ka.problems = d_problems.p; ka.seqs = d_seqs.p; ka.ops = d_ops.p; ka.results = d_results.p;
        ka.pool = c.pool + win->off; ka.chunk_bitmap = win->bitmap;
        ka.overflow_base = (uint32_t)(meta_small_units + meta_large_units + arenas);
        ka.n_chunks = (uint32_t)n_chunks;
        ka.esc_score = 0; ka.esc_list = d_esc_list.p; ka.esc_count = d_queue.p + 3; ka.n_dev = nullptr;
        LCD_CUDA_OK(cudaMemsetAsync(d_queue.p, 0, 4 * sizeof(uint32_t), s));
        LCD_CUDA_OK(cudaMemsetAsync(win->bitmap, 0, sizeof(uint32_t) * Context::BITMAP_WORDS, s));
        static bool attr_set = false;
        if (!attr_set) {
            LCD_CUDA_OK(cudaFuncSetAttribute(wfa_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(32)));
            LCD_CUDA_OK(cudaFuncSetAttribute(wfa_kernel<CTA_GROUP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(CTA_GROUP_THREADS)));
            attr_set = true;
        }
        KernelArgs kl = ka;
        kl.meta = reinterpret_cast<WfSet *>(c.pool + win->off + meta_small_units * 4); kl.meta_cap = cap_large;
        kl.arena_base = (uint32_t)(meta_small_units + meta_large_units + arena_small * groups_small);
        kl.arena_units = (uint32_t)arena_large;
        if (n_large) {             // large problems on the side stream, concurrently with the small ones
            LCD_CUDA_OK(cudaEventRecord(ev_fork, s));
            LCD_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0));
            kl.order = d_order_large.p; kl.n = n_large; kl.queue = d_queue.p + 1;
            wfa_kernel<CTA_GROUP_THREADS><<<std::min(n_large, grid_large), CTA_GROUP_THREADS, smem_bytes(CTA_GROUP_THREADS), side>>>(kl);
            LCD_CUDA_OK(cudaGetLastError());
            LCD_CUDA_OK(cudaEventRecord(ev_join, side));
            c.launches++;
        }
        if (grid_small) {
            KernelArgs ks = ka;
            ks.order = d_order_small.p; ks.n = n_small; ks.queue = d_queue.p;
            ks.meta = reinterpret_cast<WfSet *>(c.pool + win->off); ks.meta_cap = cap_small;
            ks.arena_base = (uint32_t)(meta_small_units + meta_large_units);
            ks.arena_units = (uint32_t)arena_small;
            ks.esc_score = ESC_SCORE;
            wfa_kernel<32><<<grid_small, 32 * WARP_GROUPS_PER_CTA, smem_bytes(32), s>>>(ks);
            LCD_CUDA_OK(cudaGetLastError());
            c.launches++;
        }
        if (n_large) LCD_CUDA_OK(cudaStreamWaitEvent(s, ev_join, 0));
        if (grid_small) {          // the escalated problems, one CTA each (count read on the device)
            kl.order = d_esc_list.p; kl.n = 0; kl.n_dev = d_queue.p + 3; kl.queue = d_queue.p + 2;
            wfa_kernel<CTA_GROUP_THREADS><<<grid_large, CTA_GROUP_THREADS, smem_bytes(CTA_GROUP_THREADS), s>>>(kl);
            LCD_CUDA_OK(cudaGetLastError());
            c.launches++;
        }
        return 0;
    }

    int download(cudaStream_t s) {
        h_results.resize(n);
        h_ops.resize(ops_bytes + 16);
        if (n == 0) return 0;
        LCD_DRAIN(s);
        LCD_CUDA_OK(cudaMemcpyAsync(h_results.data(), d_results.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_ops.data(), d_ops.p, ops_bytes, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int work_units(cudaStream_t s, uint64_t *units) override {
        if (download(s)) return -1;
        uint64_t t = 0;
        for (int i = 0; i < n; ++i) t += ((uint64_t)h_results[i].cells_hi << 32) | h_results[i].cells_lo;
        *units = t;
        return 0;
    }

    int fetch(cudaStream_t s, char *ops, const int64_t *ops_off, lcd_wfa_result_t *results) {
        if (download(s)) return -1;
        int bad = 0;
        for (int i = 0; i < n; ++i) {
            const DevResult &r = h_results[i];
            results[i].status = r.status; results[i].score = r.score; results[i].n_ops = r.n_ops;
            results[i].end_v = r.end_v; results[i].end_h = r.end_h;
            if (r.status < 0) { ++bad; continue; }
            if (ops && ops_off) {
                memcpy(ops + ops_off[i], h_ops.data() + ops_dev_off[i] + r.ops_begin, r.n_ops);
                ops[ops_off[i] + r.n_ops] = '\0';
            }
        }
        if (bad) { set_error("lcd_wfa: %d of %d alignments failed on the device (workspace pool exhausted or score cap exceeded)", bad, n); return -2; }
        return 0;
    }
};

} // namespace wfa
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_wfa_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                const int64_t *pat_off, const int32_t *plen,
                                const int64_t *txt_off, const int32_t *tlen,
                                const lcd_wfa_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (n < 0 || (n > 0 && (!seqs || !pat_off || !plen || !txt_off || !tlen || !params))) {
        set_error("lcd_wfa_plan_create: invalid arguments"); return nullptr;
    }
    wfa::WfaPlan *p = new wfa::WfaPlan();
    if (p->build(n, seqs, seqs_len, pat_off, plen, txt_off, tlen, params)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_wfa_plan_fetch(lcd_plan_t *plan, void *stream, char *ops, const int64_t *ops_off,
                       lcd_wfa_result_t *results) {
    wfa::WfaPlan *p = dynamic_cast<wfa::WfaPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !results) { set_error("lcd_wfa_plan_fetch: not a WFA plan / null results"); return -1; }
    return p->fetch(pick_stream(stream), ops, ops_off, results);
}

int lcd_wfa_batch(int n, const uint8_t *seqs, size_t seqs_len,
                  const int64_t *pat_off, const int32_t *plen,
                  const int64_t *txt_off, const int32_t *tlen,
                  const lcd_wfa_params_t *params,
                  char *ops, const int64_t *ops_off, lcd_wfa_result_t *results) {
    lcd_plan_t *plan = lcd_wfa_plan_create(n, seqs, seqs_len, pat_off, plen, txt_off, tlen, params);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_wfa_plan_fetch(plan, nullptr, ops, ops_off, results);
    lcd_plan_destroy(plan);
    return rc;
}

}
