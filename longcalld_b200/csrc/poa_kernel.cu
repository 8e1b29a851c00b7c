// poa_kernel.cu -- K5/K5b launcher and host plan: batched partial-order alignment (consensus + MSA).
// Device logic and design notes: poa_device.cuh.
#include "lcd_common.cuh"
#include "poa_device.cuh"
#include <algorithm>
#include <stdlib.h>

namespace lcd {
namespace poa {

constexpr int WARPS_PER_CTA = 4;
constexpr int POA_CARVEOUT_PCT = 50;      // % of the SM's L1 / shared memory kept as shared memory where the persistent grid sits: with the driver's choice (just what the grid needs) a kernel with its own shared memory (K1's histogram: 8 KB) cannot join a half-free SM and waits for the grid to retire (measured: 196 ms instead of 23 ms)

#ifndef POA_MIN_CTAS
#define POA_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(32 * WARPS_PER_CTA, POA_MIN_CTAS)
poa_kernel(const KernelArgs a) {
    const int gi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group_id = blockIdx.x * WARPS_PER_CTA + gi;
    int32_t *arena = a.arena + (size_t)group_id * a.arena_words;
    __shared__ __align__(16) int16_t row_cache[WARPS_PER_CTA][2 * 3 * Poa<WarpLanes>::NVC * PN];
    __shared__ __align__(128) int4 meta_ring[WARPS_PER_CTA][64];             // two halves of 32 row descriptors per warp, filled by TMA bulk copies
    __shared__ __align__(8) unsigned long long meta_bar[WARPS_PER_CTA][2];
    Poa<WarpLanes> poa;
    poa.row_cache = row_cache[gi];
#ifndef POA_NO_TMA_RING
    if (lane == 0) {
        for (int h = 0; h < 2; ++h) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&meta_bar[gi][h])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    poa.ring = meta_ring[gi]; poa.ring_bar = meta_bar[gi];
#endif
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.queue, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (uint32_t)a.n) break;
        const int pi = a.order[item];
        poa.run(a, a.problems[pi], a.results + pi, arena);
    }
}

// one CTA of CTA_WARPS warps per problem (kilobase regions): the vectors of every DP row are dealt to the warps
constexpr int CTA_WARPS = 4;
__global__ void __launch_bounds__(32 * CTA_WARPS)
poa_cta_kernel(const KernelArgs a) {
    typedef Poa<CtaLanes<CTA_WARPS>> P;
    __shared__ int gs[P::GS_INTS];
    __shared__ uint32_t next_item;
    int32_t *arena = a.arena + (size_t)blockIdx.x * a.arena_words;
    P poa;
    poa.gs = gs;
    for (;;) {
        if (threadIdx.x == 0) next_item = atomicAdd(a.queue, 1u);
        __syncthreads();
        const uint32_t item = next_item;
        __syncthreads();
        if (item >= (uint32_t)a.n) break;
        const int pi = a.order[item];
        poa.run(a, a.problems[pi], a.results + pi, arena);
    }
}

// D2H of the consensus sequences: the device layout reserves the worst case (sum of the read lengths) per problem;
// the bytes actually produced are packed back to back first, so the copy moves ~7 % of the buffer.
__global__ void __launch_bounds__(256)
poa_gather_cons_kernel(const uint8_t *cons, const Problem *problems, const DevResult *results, const unsigned long long *dst_off,
                       uint8_t *packed, int n) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = wid; i < n; i += nw) {
        if (results[i].status != ST_OK) continue;
        const uint8_t *src = cons + problems[i].cons_off;
        uint8_t *dst = packed + dst_off[i];
        for (int k = lane; k < results[i].cons_len + results[i].cons_len2; k += 32) dst[k] = src[k];
    }
}

// arena words a problem needs (mirrors Poa::carve) for given node / edge / DP cell budgets
static uint64_t arena_need_words(uint64_t N, uint64_t E, int max_len, int n_reads, uint64_t dp_cells) {
    uint64_t top = 30 * ((N + 3) & ~3ull) + N * 4;
    const uint64_t stride = 2 + 2 * (1 + ((n_reads - 1) >> 6));
    top += E * 4 + ((E * stride + 3) & ~3ull) + ((E + 3) & ~3ull) + 2 * ((uint64_t)max_len + N + 8) + ((uint64_t)max_len + 192) / 4;
    top = (top + 3) & ~3ull;
    top += N * 4 + (uint64_t)5 * Poa<WarpLanes>::qp_stride_of(max_len) / 2;          // row meta + query profile
    top = (top + 31) & ~31ull;
    return top + 1024 + (dp_cells + 1) / 2;
}

struct PoaPlan : Plan {
    DevBuf<uint8_t> d_seqs, d_cons, d_msa, d_read_clu;
    bool any_ncons = false;                // a problem asks for two consensus sequences (de-novo clustering)
    int n_total_reads_ = 0;
    DevBuf<Problem> d_problems;
    DevBuf<int64_t> d_read_off;
    DevBuf<int32_t> d_read_len, d_order, d_sub_beg, d_sub_end;
    std::vector<uint8_t> has_sub;          // per problem: a read is aligned against a sub-graph (partially covering reads)
    DevBuf<DevResult> d_results;
    DevBuf<uint32_t> d_queue;
    DevBuf<unsigned long long> d_msa_used, d_pack_off;
    DevBuf<uint8_t> d_pack;
    std::vector<unsigned long long> pack_off;
    std::vector<Problem> problems;
    std::vector<int64_t> cons_dev_off;
    std::vector<uint64_t> need_small;      // arena words with the estimated DP budget
    std::vector<uint64_t> need_full;       // arena words with the worst-case (full matrix) DP budget
    std::vector<int32_t> order_all;
    size_t cons_bytes = 0, msa_pool_bytes = 0;
    std::vector<DevResult> h_results;
    std::vector<uint8_t> h_cons, h_msa;
    int n_rescued = 0;
    std::vector<int32_t> cls[3];
    bool pending = false;                  // run() has launched; the statuses (and a rescue launch, if any) are still to be looked at
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    ~PoaPlan() override {
        for (int k = 0; k < 2; ++k) { if (side[k]) cudaStreamDestroy(side[k]); if (ev_join[k]) cudaEventDestroy(ev_join[k]); }
        if (ev_fork) cudaEventDestroy(ev_fork);
    }

    int build(int n_, const uint8_t *seqs, size_t seqs_len, const int32_t *first_read, const int32_t *n_reads,
              const int64_t *read_off, const int32_t *read_len, int n_total_reads, const lcd_poa_params_t *params,
              const int32_t *sub_beg = nullptr, const int32_t *sub_end = nullptr, const double *min_freq = nullptr) {
        n = n_; n_total_reads_ = n_total_reads;
        has_sub.assign(n, 0);
        Context &c = ctx();
        problems.resize(n); cons_dev_off.resize(n); need_small.resize(n); need_full.resize(n);
        std::vector<double> work(n);
        for (int r = 0; r < n_total_reads; ++r)
            if (read_len[r] < 0 || read_off[r] < 0 || (size_t)read_off[r] + read_len[r] > seqs_len) { set_error("lcd_poa: read %d has an invalid range", r); return -1; }
        size_t cons_total = 0; double msa_est = 0;
        for (int i = 0; i < n; ++i) {
            if (n_reads[i] < 1 || first_read[i] < 0 || first_read[i] + n_reads[i] > n_total_reads) { set_error("lcd_poa: problem %d has an invalid read range", i); return -1; }
            if (params[i].max_n_cons != 1 && !(params[i].max_n_cons == 2 && min_freq && !params[i].sub_aln)) {
                set_error("lcd_poa: problem %d asks for max_n_cons=%d; two consensus sequences come from lcd_poa_ncons_* (sub_aln = 0), more are not implemented", i, params[i].max_n_cons); return -1;
            }
            Problem &p = problems[i];
            memset(&p, 0, sizeof(p));
            if (params[i].max_n_cons == 2) {           // abpoa_output.c:1141, in double as the reference computes it
                any_ncons = true;
                const int cw = (int)ceil((double)n_reads[i] * min_freq[i]);
                p.min_w = cw > 2 ? cw : 2;
            }
            p.seq_base = 0; p.read_first = first_read[i]; p.n_reads = n_reads[i]; p.par = params[i];
            int mn = INT32_MAX;
            for (int r = 0; r < n_reads[i]; ++r) {
                const int l = read_len[first_read[i] + r]; p.sum_len += l; p.max_len = std::max(p.max_len, l);
                const bool part = sub_beg && r > 0 && sub_beg[first_read[i] + r] != 0;
                if (!part) mn = std::min(mn, l);                    // (a partially covering read is shorter by design: it does not widen the band estimate)
                if (sub_beg && r > 0 && sub_beg[first_read[i] + r] > 0) {
                    has_sub[i] = 1;
                    if (!sub_end || sub_end[first_read[i] + r] < sub_beg[first_read[i] + r] || !params[i].sub_aln || params[i].wb < 0) { set_error("lcd_poa: problem %d read %d has invalid sub-graph anchors (they need sub_aln = 1, a band, and beg <= end)", i, r); return -1; }
                }
            }
            if (mn == INT32_MAX) mn = p.max_len;
            p.cons_off = (int32_t)cons_total; cons_dev_off[i] = (int64_t)cons_total;
            cons_total += ((size_t)p.sum_len + 15) & ~(size_t)15;
            if (cons_total > 0x7fffffffull) { set_error("lcd_poa: batch too large (consensus buffer > 2 GiB); split it"); return -1; }
            // first-attempt budgets: nodes ~ longest read + branches, 3 edge slots per node, DP rows ~ nodes,
            // vectors per row ~ band / 32; the worst case (every base a new node, full matrix) is the rescue budget
            p.node_cap = (int32_t)std::min<int64_t>((int64_t)p.sum_len + 34, 2ll * p.max_len + 64 + 2ll * p.n_reads);
            p.edge_cap = 3 * p.node_cap;
            const double rows = 1.15 * p.max_len + 24;
            const int dp_sn = (p.max_len + 32) / 32;
            const int wband = params[i].wb < 0 ? p.max_len : params[i].wb + (int)(params[i].wf * p.max_len);
            double nv = params[i].wb < 0 ? dp_sn + 1 : std::min<double>(dp_sn + 1, (2.0 * wband + 2.0 * (p.max_len - mn) + 64) / 32 + 3);
            need_small[i] = arena_need_words(p.node_cap, p.edge_cap, p.max_len, p.n_reads, (uint64_t)(rows * nv * 160));
            need_full[i] = arena_need_words((uint64_t)p.sum_len + 34, 3ull * (p.sum_len + p.n_reads) + 64, p.max_len, p.n_reads,
                                            (uint64_t)((double)(p.sum_len + 34) * (dp_sn + 1) * 160));
            work[i] = (double)p.n_reads * rows * nv;
            msa_est += (double)(p.n_reads + params[i].max_n_cons) * (1.5 * p.max_len + 64);
        }
        cons_bytes = cons_total;
        msa_pool_bytes = ((size_t)(msa_est * 1.5) + 4096 + 15) & ~(size_t)15;
        order_all.resize(n);
        for (int i = 0; i < n; ++i) order_all[i] = i;
        std::sort(order_all.begin(), order_all.end(), [&](int32_t x, int32_t y) { return work[x] != work[y] ? work[x] > work[y] : x < y; });
        cudaStream_t s = cur_stream();
        if (d_seqs.upload(seqs, std::max<size_t>(seqs_len, 1), s)) return -1;
        if (d_problems.upload(problems.data(), n, s)) return -1;
        if (d_read_off.upload(read_off, n_total_reads, s)) return -1;
        if (d_read_len.upload(read_len, n_total_reads, s)) return -1;
        if (sub_beg && (d_sub_beg.upload(sub_beg, n_total_reads, s) || d_sub_end.upload(sub_end, n_total_reads, s))) return -1;
        if (d_order.alloc(std::max(n, 1))) return -1;
        if (d_cons.alloc(cons_bytes + 16)) return -1;
        if (d_msa.alloc(msa_pool_bytes)) return -1;
        if (d_results.alloc(std::max(n, 1))) return -1;
        if (d_queue.alloc(4)) return -1;
        for (int k = 0; k < 2; ++k) {
            LCD_CUDA_OK(cudaStreamCreateWithFlags(&side[k], cudaStreamNonBlocking));
            LCD_CUDA_OK(cudaEventCreateWithFlags(&ev_join[k], cudaEventDisableTiming));
        }
        LCD_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        if (d_msa_used.alloc(1)) return -1;
        if (any_ncons) { if (d_read_clu.alloc(std::max(n_total_reads, 1))) return -1; LCD_CUDA_OK(cudaMemsetAsync(d_read_clu.p, 0, std::max(n_total_reads, 1), s)); }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    // one launch over `idx` with per-group arenas of `words`
    // kind: 1 = poa_kernel (one problem per warp), 2 = poa_cta_kernel (one per CTA: unbanded problems, rescue launches).
    // The launch uses pool words [pool_lo, pool_hi) for its arenas and reports how many it took in *used.
    int launch(cudaStream_t s, const std::vector<int32_t> &idx, int32_t *d_idx, uint32_t *d_q, uint64_t words, int max_groups,
               int kind, bool worst_case, uint64_t pool_lo, uint64_t pool_hi, uint64_t *used) {
        Context &c = ctx();
        if (used) *used = 0;
        if (idx.empty()) return 0;
        const int per_cta = kind == 1 ? WARPS_PER_CTA : 1;
        const uint64_t avail = pool_hi > pool_lo ? pool_hi - pool_lo : 0;
        const uint64_t fit = avail / words;
        if (fit == 0) { set_error("lcd_poa: a problem needs %zu MiB of workspace but the pool has %zu MiB", (size_t)(words * 4 >> 20), (size_t)(win->words * 4 >> 20)); return -1; }
        int groups = (int)std::min<uint64_t>(std::min<uint64_t>(fit, (uint64_t)max_groups), idx.size());
        int grid = (groups + per_cta - 1) / per_cta;
        if ((uint64_t)grid * per_cta > fit) grid = (int)(fit / per_cta);
        if (grid == 0) { grid = 1; }
        if ((uint64_t)grid * per_cta * words > avail) {
            set_error("lcd_poa: workspace pool too small for one CTA of %d arenas of %zu MiB", per_cta, (size_t)(words * 4 >> 20)); return -1;
        }
        if (used) *used = (uint64_t)grid * per_cta * words;
        LCD_CUDA_OK(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        LCD_CUDA_OK(cudaMemsetAsync(d_q, 0, sizeof(uint32_t), s));
        KernelArgs ka;
        ka.problems = d_problems.p; ka.order = d_idx; ka.n = (int)idx.size(); ka.queue = d_q;
        ka.seqs = d_seqs.p; ka.read_off = d_read_off.p; ka.read_len = d_read_len.p;
        ka.cons = d_cons.p; ka.msa = d_msa.p; ka.msa_cap = msa_pool_bytes; ka.msa_used = d_msa_used.p;
        ka.sub_beg = d_sub_beg.p; ka.sub_end = d_sub_end.p; ka.read_clu = any_ncons ? d_read_clu.p : nullptr;
        ka.results = d_results.p; ka.arena = c.pool + win->off + pool_lo; ka.arena_words = words; ka.worst_case = worst_case ? 1 : 0;
        static int carve_set = -2;          // shared-memory carve-out of the SMs the persistent grid sits on (see lcd_gpu_reserve_sms)
        if (carve_set == -2) {
            const char *cv = getenv("LCD_POA_CARVEOUT");
            carve_set = cv ? atoi(cv) : POA_CARVEOUT_PCT;
            if (carve_set >= 0) LCD_CUDA_OK(cudaFuncSetAttribute(poa_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve_set));
        }
        if (kind == 1) poa_kernel<<<grid, 32 * WARPS_PER_CTA, 0, s>>>(ka);
        else poa_cta_kernel<<<grid, 32 * CTA_WARPS, 0, s>>>(ka);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }

    // run() only enqueues the launch and returns: the host is free to start other engines' plans while the persistent grid works.
    // Problems that outgrew their first-attempt workspace are found and re-run by finish() (lcd_plan_sync / fetch / work_units).
    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        if (pending && finish_locked(s)) return -1;
        LCD_CUDA_OK(cudaMemsetAsync(d_msa_used.p, 0, sizeof(unsigned long long), s));
        // class W (1): one problem per warp -- every banded problem; class C (2): one per CTA -- rows wider than the warp kernel's on-chip row
        // cache (unbanded POA of reads over 224 bp); rescue: problems that outgrew their first-attempt budget, re-run with the worst-case
        // budget.  The two classes run concurrently (side streams) in disjoint parts of the plan's pool window.
        for (int k = 0; k < 3; ++k) cls[k].clear();         // (member: the launch's index upload reads them after run() has returned)
        uint64_t cw[3] = {0, 0, 0};
        for (int32_t i : order_all) {
            // kilobase problems run on the warp kernel as well: its strip rows, 32-wide backtrack and parallel fusion beat the CTA kernel's
            // two-phase rows (measured: profiles/README.md, r1 v5)
            const int k = (problems[i].par.wb >= 0 || problems[i].max_len <= 224 || has_sub[i]) ? 1 : 2;
            cls[k].push_back(i); cw[k] = std::max(cw[k], need_small[i]);
        }
        for (int k = 1; k < 3; ++k) cw[k] = (cw[k] + 63) & ~63ull;
        const int max_groups[3] = { 0, c.dp_sms() * WARPS_PER_CTA * POA_MIN_CTAS, c.sm_count * 4 };
        // pool split: CTAs and warps share the window in proportion to demand
        uint64_t want[3] = {0, 0, 0};
        const int per_cta[3] = { 1, WARPS_PER_CTA, 1 };
        for (int k = 1; k < 3; ++k) {
            const uint64_t g = std::min<uint64_t>(cls[k].size(), (uint64_t)max_groups[k]);
            want[k] = (g + per_cta[k] - 1) / per_cta[k] * per_cta[k] * cw[k];       // launches round up to whole CTAs
        }
        uint64_t lo[4]; lo[0] = 0; lo[1] = 0;
        const uint64_t rest_pool = win->words;
        const uint64_t w12 = want[1] + want[2];
        lo[2] = lo[1] + (w12 <= rest_pool ? want[1] : (uint64_t)((double)rest_pool * ((double)want[1] / (double)w12)));
        lo[3] = win->words;
        for (int k = 1; k < 4; ++k) lo[k] &= ~63ull;
        LCD_CUDA_OK(cudaEventRecord(ev_fork, s));
        for (int k = 2; k >= 1; --k) {            // big problems first
            if (cls[k].empty()) continue;
            cudaStream_t ss = side[k - 1];
            LCD_CUDA_OK(cudaStreamWaitEvent(ss, ev_fork, 0));
            const size_t off = k == 1 ? cls[0].size() : cls[0].size() + cls[1].size();
            if (launch(ss, cls[k], d_order.p + off, d_queue.p + k, cw[k], max_groups[k], k, false, lo[k], lo[k + 1], nullptr)) return -1;
            LCD_CUDA_OK(cudaEventRecord(ev_join[k - 1], ss));
        }
        for (int k = 1; k <= 2; ++k) if (!cls[k].empty()) LCD_CUDA_OK(cudaStreamWaitEvent(s, ev_join[k - 1], 0));
        pending = true;
        return 0;
    }

    int finish(cudaStream_t s) override {
        if (!pending) return 0;
        Context &c = ctx();
        std::lock_guard<std::mutex> lk(win->mu);     // a rescue launch carves from the plan's pool window again
        return finish_locked(s);
    }

    // statuses back: anything that ran out of workspace is re-run with the full-matrix budget
    int finish_locked(cudaStream_t s) {
        Context &c = ctx();
        if (!pending) return 0;
        pending = false;
        h_results.resize(n);
        LCD_DRAIN(s);
        LCD_CUDA_OK(cudaMemcpyAsync(h_results.data(), d_results.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        std::vector<int32_t> rescue; uint64_t rescue_words = 0;
        for (int32_t i : order_all) if (h_results[i].status == ST_OOM) { rescue.push_back(i); rescue_words = std::max(rescue_words, need_full[i]); }
        n_rescued = (int)rescue.size();
        if (!rescue.empty()) {
            rescue_words = std::min<uint64_t>((rescue_words + 63) & ~63ull, (win->words / WARPS_PER_CTA) & ~63ull);
            LCD_CUDA_OK(cudaStreamWaitEvent(s, win->done, 0));
            bool sub_rescue = false;           // sub-graph alignment lives in the warp kernel only
            for (int32_t i : rescue) if (has_sub[i]) sub_rescue = true;
            if (launch(s, rescue, d_order.p, d_queue.p, rescue_words, sub_rescue ? c.dp_sms() * WARPS_PER_CTA * POA_MIN_CTAS : c.sm_count * 4, sub_rescue ? 1 : 2, true, 0, win->words, nullptr)) return -1;
            LCD_CUDA_OK(cudaEventRecord(win->done, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
        }
        return 0;
    }

    // results always; consensus bytes (packed on the device first) and the MSA pool only when asked for
    int download(cudaStream_t s, bool want_cons, bool want_msa) {
        Context &c = ctx();
        h_results.resize(n);
        if (n == 0) return 0;
        if (finish(s)) return -1;
        LCD_DRAIN(s);
        unsigned long long used = 0;
        LCD_CUDA_OK(cudaMemcpyAsync(h_results.data(), d_results.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(&used, d_msa_used.p, sizeof(used), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        if (want_cons) {
            pack_off.resize(n + 1);
            unsigned long long tot = 0;
            for (int i = 0; i < n; ++i) { pack_off[i] = tot; if (h_results[i].status == ST_OK) tot += (unsigned long long)h_results[i].cons_len + (unsigned long long)h_results[i].cons_len2; }
            pack_off[n] = tot;
            h_cons.resize(tot + 16);
            if (tot) {
                if (d_pack.n < tot + 16 && d_pack.alloc(tot + 16 + tot / 8)) return -1;
                if (d_pack_off.n < (size_t)n + 1 && d_pack_off.alloc(n + 1)) return -1;
                LCD_CUDA_OK(cudaStreamSynchronize(cur_stream()));                 // allocation ordered before use on s
                LCD_CUDA_OK(cudaMemcpyAsync(d_pack_off.p, pack_off.data(), sizeof(unsigned long long) * (n + 1), cudaMemcpyHostToDevice, s));
                poa_gather_cons_kernel<<<c.sm_count * 4, 256, 0, s>>>(d_cons.p, d_problems.p, d_results.p, d_pack_off.p, d_pack.p, n);
                LCD_CUDA_OK(cudaGetLastError());
                c.launches++;
                LCD_CUDA_OK(cudaMemcpyAsync(h_cons.data(), d_pack.p, tot, cudaMemcpyDeviceToHost, s));
            }
        }
        if (want_msa) {
            used = std::min<unsigned long long>(used, msa_pool_bytes);
            h_msa.resize(used + 16);
            if (used) LCD_CUDA_OK(cudaMemcpyAsync(h_msa.data(), d_msa.p, used, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

#ifdef LCD_POA_TIMING
    void timing_report() {      // debug builds: phase cycles of the 12 slowest problems and the batch totals
        std::vector<int> idx(n); for (int i = 0; i < n; ++i) idx[i] = i;
        auto tot = [&](int i) { const DevResult &r = h_results[i]; return r.t_dp + r.t_bt + r.t_add + r.t_after + r.t_fin + r.t_pro; };
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { return tot(x) > tot(y); });
        unsigned long long T[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i) { const DevResult &r = h_results[i]; T[0] += r.t_dp; T[1] += r.t_bt; T[2] += r.t_add; T[3] += r.t_after; T[4] += r.t_fin; T[5] += r.t_seg; T[6] += r.t_gen; T[7] += r.n_seg; T[8] += r.n_gen; T[9] += r.t_pro; }
        fprintf(stderr, "[poa timing] batch Mcycles: dp %.1f (segments %.1f for %.2f Mrows, general %.1f for %.2f Mrows) bt %.1f add %.1f after %.1f (incl. BFS index) fin %.1f sub-graph set-up %.1f\n", T[0] / 1e6, T[5] / 1e6, T[7] / 1e6, T[6] / 1e6, T[8] / 1e6, T[1] / 1e6, T[2] / 1e6, T[3] / 1e6, T[4] / 1e6, T[9] / 1e6);
        for (int k = 0; k < std::min(n, 12); ++k) { const int i = idx[k]; const DevResult &r = h_results[i];
            fprintf(stderr, "[poa timing] #%d reads %d max_len %d nodes %d cells %u : dp %.1f (seg %.1f / %llu rows, gen %.1f / %llu rows) bt %.1f add %.1f after %.1f fin %.1f sub %.1f Mcycles\n", i, problems[i].n_reads,
                    problems[i].max_len, r.n_nodes, r.cells_lo, r.t_dp / 1e6, r.t_seg / 1e6, r.n_seg, r.t_gen / 1e6, r.n_gen, r.t_bt / 1e6, r.t_add / 1e6, r.t_after / 1e6, r.t_fin / 1e6, r.t_pro / 1e6); }
    }
#endif
    int work_units(cudaStream_t s, uint64_t *units) override {
        if (download(s, false, false)) return -1;
#ifdef LCD_POA_TIMING
        timing_report();
#endif
        uint64_t t = 0;
        for (int i = 0; i < n; ++i) t += ((uint64_t)h_results[i].cells_hi << 32) | h_results[i].cells_lo;
        *units = t;
        return 0;
    }

    int fetch(cudaStream_t s, uint8_t *cons, const int64_t *cons_off, uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
              lcd_poa_result_t *results) {
        if (download(s, cons && cons_off, msa && msa_off && msa_cap)) return -1;
#ifdef LCD_POA_TIMING
        if (getenv("LCD_POA_TIMING_PRINT")) timing_report();
#endif
        int bad = 0, first_bad = 0;
        for (int i = 0; i < n; ++i) {
            const DevResult &r = h_results[i];
            results[i].status = r.status; results[i].cons_len = r.cons_len; results[i].msa_len = r.msa_len; results[i].n_nodes = r.n_nodes;
            if (r.status != ST_OK) { if (!bad) first_bad = r.status; ++bad; continue; }
            if (cons && cons_off) memcpy(cons + cons_off[i], h_cons.data() + pack_off[i], (size_t)r.cons_len + r.cons_len2);
            if (msa && msa_off && msa_cap) {
                const int64_t bytes = (int64_t)(problems[i].n_reads + (r.n_cons == 2 ? 2 : 1)) * r.msa_len;
                if (bytes > msa_cap[i]) { results[i].status = LCD_POA_MSA_CAP; if (!bad) first_bad = LCD_POA_MSA_CAP; ++bad; continue; }
                memcpy(msa + msa_off[i], h_msa.data() + r.msa_off, bytes);
            }
        }
        if (bad) { set_error("lcd_poa: %d of %d problems failed on the device (first status %d; see LCD_POA_* in lcd_gpu.h)", bad, n, first_bad); return -2; }
        return 0;
    }

    // after fetch(): the clusters of the problems that asked for two consensus sequences
    int fetch_clusters(cudaStream_t s, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster) {
        if ((int)h_results.size() != n) { set_error("lcd_poa_plan_fetch_clusters: call lcd_poa_plan_fetch first"); return -1; }
        for (int i = 0; i < n; ++i) { if (n_cons) n_cons[i] = h_results[i].status == ST_OK ? h_results[i].n_cons : 0; if (cons_len2) cons_len2[i] = h_results[i].status == ST_OK ? h_results[i].cons_len2 : 0; }
        if (read_cluster && n_total_reads_ > 0) {
            if (any_ncons) { LCD_DRAIN(s); LCD_CUDA_OK(cudaMemcpyAsync(read_cluster, d_read_clu.p, n_total_reads_, cudaMemcpyDeviceToHost, s)); LCD_CUDA_OK(cudaStreamSynchronize(s)); }
            else memset(read_cluster, 0, n_total_reads_);
        }
        return 0;
    }
};

} // namespace poa
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_poa_sub_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                    const int32_t *first_read, const int32_t *n_reads,
                                    const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                    const int32_t *sub_beg, const int32_t *sub_end,
                                    const lcd_poa_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (n < 0 || (n > 0 && (!seqs || !first_read || !n_reads || !read_off || !read_len || !params)) || ((sub_beg == nullptr) != (sub_end == nullptr))) {
        set_error("lcd_poa_plan_create: invalid arguments"); return nullptr;
    }
    poa::PoaPlan *p = new poa::PoaPlan();
    if (p->build(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, params, sub_beg, sub_end)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}
lcd_plan_t *lcd_poa_ncons_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                      const int32_t *first_read, const int32_t *n_reads,
                                      const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                      const lcd_poa_params_t *params, const double *min_freq) {
    if (ensure_ready()) return nullptr;
    if (n < 0 || (n > 0 && (!seqs || !first_read || !n_reads || !read_off || !read_len || !params || !min_freq))) { set_error("lcd_poa_ncons_plan_create: invalid arguments"); return nullptr; }
    poa::PoaPlan *p = new poa::PoaPlan();
    if (p->build(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, params, nullptr, nullptr, min_freq)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}
int lcd_poa_plan_fetch_clusters(lcd_plan_t *plan, void *stream, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster) {
    poa::PoaPlan *p = dynamic_cast<poa::PoaPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p) { set_error("lcd_poa_plan_fetch_clusters: not a POA plan"); return -1; }
    return p->fetch_clusters(pick_stream(stream), n_cons, cons_len2, read_cluster);
}
int lcd_poa_ncons_batch(int n, const uint8_t *seqs, size_t seqs_len,
                        const int32_t *first_read, const int32_t *n_reads,
                        const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                        const lcd_poa_params_t *params, const double *min_freq,
                        uint8_t *cons, const int64_t *cons_off,
                        uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                        lcd_poa_result_t *results, int32_t *n_cons, int32_t *cons_len2, uint8_t *read_cluster) {
    lcd_plan_t *plan = lcd_poa_ncons_plan_create(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, params, min_freq);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_poa_plan_fetch(plan, nullptr, cons, cons_off, msa, msa_off, msa_cap, results);
    if (!rc || rc == -2) { const int rc2 = lcd_poa_plan_fetch_clusters(plan, nullptr, n_cons, cons_len2, read_cluster); if (!rc) rc = rc2; }
    lcd_plan_destroy(plan);
    return rc;
}

lcd_plan_t *lcd_poa_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                const int32_t *first_read, const int32_t *n_reads,
                                const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                                const lcd_poa_params_t *params) {
    return lcd_poa_sub_plan_create(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, nullptr, nullptr, params);
}

int lcd_poa_plan_fetch(lcd_plan_t *plan, void *stream, uint8_t *cons, const int64_t *cons_off,
                       uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap, lcd_poa_result_t *results) {
    poa::PoaPlan *p = dynamic_cast<poa::PoaPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !results) { set_error("lcd_poa_plan_fetch: not a POA plan / null results"); return -1; }
    return p->fetch(pick_stream(stream), cons, cons_off, msa, msa_off, msa_cap, results);
}

int lcd_poa_sub_batch(int n, const uint8_t *seqs, size_t seqs_len,
                      const int32_t *first_read, const int32_t *n_reads,
                      const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                      const int32_t *sub_beg, const int32_t *sub_end,
                      const lcd_poa_params_t *params,
                      uint8_t *cons, const int64_t *cons_off,
                      uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                      lcd_poa_result_t *results) {
    lcd_plan_t *plan = lcd_poa_sub_plan_create(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, sub_beg, sub_end, params);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_poa_plan_fetch(plan, nullptr, cons, cons_off, msa, msa_off, msa_cap, results);
    lcd_plan_destroy(plan);
    return rc;
}

int lcd_poa_batch(int n, const uint8_t *seqs, size_t seqs_len,
                  const int32_t *first_read, const int32_t *n_reads,
                  const int64_t *read_off, const int32_t *read_len, int n_total_reads,
                  const lcd_poa_params_t *params,
                  uint8_t *cons, const int64_t *cons_off,
                  uint8_t *msa, const int64_t *msa_off, const int64_t *msa_cap,
                  lcd_poa_result_t *results) {
    lcd_plan_t *plan = lcd_poa_plan_create(n, seqs, seqs_len, first_read, n_reads, read_off, read_len, n_total_reads, params);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_poa_plan_fetch(plan, nullptr, cons, cons_off, msa, msa_off, msa_cap, results);
    lcd_plan_destroy(plan);
    return rc;
}

}
