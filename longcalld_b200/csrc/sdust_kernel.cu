// sdust_kernel.cu -- K0 launcher and host plan: the low-complexity intervals (symmetric DUST) of the reference windows of many chunks: six launches over all of them.
// Device logic and design notes: sdust_device.cuh.
#include "lcd_common.cuh"
#include "sdust_device.cuh"
#include <algorithm>

namespace lcd {
namespace sdust {

constexpr int THREADS = 1024;       // a chunk's segment search: one CTA
constexpr int RTHREADS = 128;       // the replays: one thread per segment, the segments of all chunks side by side
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// step 0 (prevvalid) and step 2 (the segments) of a chunk on one CTA; step 1 in between is a thread per run of DECIDE_RUN positions over all chunks
__global__ void __launch_bounds__(THREADS)
sdust_segments_kernel(const Chunk *chunks, int n, int step) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const Chunk c = chunks[i];
        if (step == 0) scan_valid(c, (int)threadIdx.x, THREADS, CtaSync()); else collect_segments(c, (int)threadIdx.x, THREADS, CtaSync());
        __syncthreads();
    }
}
__global__ void __launch_bounds__(RTHREADS)
sdust_decide_kernel(const Chunk *chunks) {
    __shared__ Hot hot[RTHREADS];
    const Chunk c = chunks[blockIdx.y];
    const int i0 = (int)(blockIdx.x * RTHREADS + threadIdx.x) * DECIDE_RUN;
    if (i0 < c.n) decide_run(c, hot[threadIdx.x], i0, i0 + DECIDE_RUN < c.n ? i0 + DECIDE_RUN : c.n);
}

// the chunk's segments, taken from the chunk's counter by the threads of blockIdx.y's CTAs
__global__ void __launch_bounds__(RTHREADS)
sdust_replay_kernel(const Chunk *chunks) {
    __shared__ Hot hot[RTHREADS];
    const Chunk c = chunks[blockIdx.y];
    if (*c.status != ST_OK) return;
    const int ns = c.ctr[0];
    replay_segments(c, hot[threadIdx.x], [&]() { const int s = atomicAdd(&c.ctr[1], 1); return s < ns ? s : -1; });
}
__global__ void __launch_bounds__(RTHREADS)
sdust_pack_kernel(const Chunk *chunks) {
    const Chunk c = chunks[blockIdx.y];
    if (*c.status != ST_OK) return;
    const int s = (int)(blockIdx.x * RTHREADS + threadIdx.x);
    if (s < c.ctr[0]) pack_segment(c, s);
}

// finish_counts() by one warp per chunk: the segments' counts scanned 32 at a time
__global__ void __launch_bounds__(128)
sdust_offsets_kernel(const Chunk *chunks, int n) {
    const int i = (int)(blockIdx.x * 4 + threadIdx.x / 32), lane = (int)(threadIdx.x & 31);
    if (i >= n) return;
    const Chunk c = chunks[i];
    if (*c.status != ST_OK) return;
    const int ns = c.ctr[0];
    long long tot = 0; int bad = 0;
    for (int b0 = 0; b0 < ns; b0 += 32) {
        const int s = b0 + lane, v = s < ns ? c.seg_cnt[s] : 0;
        const unsigned neg = __ballot_sync(0xffffffffu, v < 0);
        if (neg) { bad = __shfl_sync(0xffffffffu, v, __ffs(neg) - 1); break; }      // (the first failed segment's status, as the one-thread version reports it)
        int x = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (s < ns) c.seg_off[s] = (int)(tot + x - v);
        tot += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) { if (bad) *c.status = bad; else if (tot > c.cap) *c.status = ST_CAP; *c.n_out = tot; }
}

struct SdustPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks; std::vector<size_t> hdr_off, beg_off, end_off; std::vector<long long> caps;
    DevBuf<uint8_t> d_seq, d_work; DevBuf<Chunk> d_chunks;
    long long tot_bases = 0; int max_seg_cap = 0, max_len = 0;

    int build(int n_, const lcd_sdust_input_t *in) {
        n = n_;
        if (n == 0) return 0;
        if (n > 65535) { set_error("lcd_sdust: %d chunks in one batch (at most 65535)", n); return -1; }
        chunks.resize(n); hdr_off.resize(n); beg_off.resize(n); end_off.resize(n); caps.resize(n);
        auto take = [](size_t &top, size_t bytes) { const size_t at = top; top += (bytes + 15) & ~(size_t)15; return at; };
        size_t seq_bytes = 0, work = 0;
        std::vector<size_t> seq_off(n);
        std::vector<std::vector<size_t>> wk(n);
        for (int i = 0; i < n; ++i) {
            const lcd_sdust_input_t &x = in[i];
            if (x.l_seq < 0 || (x.l_seq > 0 && !x.seq)) { set_error("lcd_sdust: chunk %d has no sequence", i); return -1; }
            if (x.W < WLEN + 1 || x.W > MAX_W || x.T < 1) { set_error("lcd_sdust: chunk %d asks for T = %d, W = %d; windows of %d .. %d bases are supported (the reference runs T = 5, W = 20)", i, x.T, x.W, WLEN + 1, MAX_W); return -1; }
            seq_off[i] = take(seq_bytes, (size_t)x.l_seq + 1);
            hdr_off[i] = take(work, 16);
        }
        for (int i = 0; i < n; ++i) {
            const size_t L = (size_t)in[i].l_seq, cap = L / 4 + 16, seg_cap = L / (size_t)(in[i].W + 20) + THREADS + 16;
            caps[i] = (long long)cap;
            beg_off[i] = take(work, cap * 8); end_off[i] = take(work, cap * 8);
            const size_t stage_cap = L / 4 + seg_cap * (size_t)(in[i].W / 4 + 3) + 16;
            for (size_t b : { (L + 1) * 4, L + 1, seg_cap * 4, seg_cap * 4, seg_cap * 4, (size_t)16, stage_cap * 8, stage_cap * 8 }) wk[i].push_back(take(work, b));
            tot_bases += (long long)L;
        }
        std::vector<uint8_t> h(seq_bytes + 16, 0);
        for (int i = 0; i < n; ++i) if (in[i].l_seq) memcpy(h.data() + seq_off[i], in[i].seq, (size_t)in[i].l_seq);
        cudaStream_t s = cur_stream();
        if (d_seq.upload(h.data(), h.size(), s) || d_work.alloc(work + 16)) return -1;
        for (int i = 0; i < n; ++i) {
            Chunk &c = chunks[i]; memset(&c, 0, sizeof(c));
            uint8_t *w = d_work.p;
            c.seq = (const char *)(d_seq.p + seq_off[i]); c.n = in[i].l_seq; c.T = in[i].T; c.W = in[i].W; c.base = in[i].base;
            c.out_beg = (long long *)(w + beg_off[i]); c.out_end = (long long *)(w + end_off[i]); c.cap = caps[i];
            c.n_out = (long long *)(w + hdr_off[i]); c.status = (int *)(w + hdr_off[i] + 8);
            c.prevvalid = (int *)(w + wk[i][0]); c.trig = w + wk[i][1]; c.seg_start = (int *)(w + wk[i][2]); c.seg_cnt = (int *)(w + wk[i][3]); c.seg_off = (int *)(w + wk[i][4]);
            c.seg_cap = (int)((size_t)in[i].l_seq / (size_t)(in[i].W + 20) + THREADS + 16); c.ctr = (int *)(w + wk[i][5]);
            max_seg_cap = std::max(max_seg_cap, c.seg_cap); max_len = std::max(max_len, c.n);
            c.stage_beg = (long long *)(w + wk[i][6]); c.stage_end = (long long *)(w + wk[i][7]); c.stage_cap = (long long)((size_t)in[i].l_seq / 4 + (size_t)c.seg_cap * (size_t)(in[i].W / 4 + 3) + 16);
        }
        if (d_chunks.upload(chunks.data(), n, s)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));      // host staging vector goes out of scope
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        const dim3 rgrid((unsigned)((max_seg_cap + RTHREADS - 1) / RTHREADS), (unsigned)n);
        const dim3 dgrid((unsigned)std::max(1, (max_len + RTHREADS * DECIDE_RUN - 1) / (RTHREADS * DECIDE_RUN)), (unsigned)n);
        sdust_segments_kernel<<<std::min(n, c.sm_count * 2), THREADS, 0, s>>>(d_chunks.p, n, 0);
        sdust_decide_kernel<<<dgrid, RTHREADS, 0, s>>>(d_chunks.p);
        sdust_segments_kernel<<<std::min(n, c.sm_count * 2), THREADS, 0, s>>>(d_chunks.p, n, 1);
        static const int spt = getenv("LCD_SDUST_SPT") ? std::max(1, atoi(getenv("LCD_SDUST_SPT"))) : 4;      // segment slots per replay thread
        const dim3 pgrid((unsigned)((max_seg_cap / spt + RTHREADS - 1) / RTHREADS), (unsigned)n);
        sdust_replay_kernel<<<pgrid, RTHREADS, 0, s>>>(d_chunks.p);
        sdust_offsets_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_chunks.p, n);
        sdust_pack_kernel<<<rgrid, RTHREADS, 0, s>>>(d_chunks.p);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches += 6;
        return 0;
    }
    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_bases; return 0; }      // reference bases scanned

    int fetch(cudaStream_t s, lcd_sdust_output_t *out) {
        if (n == 0) return 0;
        LCD_DRAIN(s);
        std::vector<uint8_t> hdr((size_t)16 * n);
        LCD_CUDA_OK(cudaMemcpyAsync(hdr.data(), d_work.p + hdr_off[0], hdr.size(), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) {
            long long k; int st; memcpy(&k, hdr.data() + 16 * (size_t)i, 8); memcpy(&st, hdr.data() + 16 * (size_t)i + 8, 4);
            if (st != ST_OK) { set_error("lcd_sdust: chunk %d failed on the device (status %d: %s)", i, st, st == ST_PLIST ? "more perfect intervals alive than the per-thread list holds" : "interval / segment capacity"); return -2; }
            out[i].n = k;
            if (k > out[i].cap) { set_error("lcd_sdust: chunk %d has %lld intervals, the caller's arrays hold %lld", i, k, (long long)out[i].cap); return -3; }
            if (k) {
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].beg, d_work.p + beg_off[i], 8 * (size_t)k, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].end, d_work.p + end_off[i], 8 * (size_t)k, cudaMemcpyDeviceToHost, s));
            }
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }
};

} // namespace sdust

int sdust_plan_view(Plan *plan, SdustView *v) {
    sdust::SdustPlan *p = dynamic_cast<sdust::SdustPlan *>(plan);
    if (!p) { set_error("not an sdust (K0) plan"); return -1; }
    v->n_chunks = p->n; v->beg.resize(p->n); v->end.resize(p->n); v->n_out.resize(p->n); v->status.resize(p->n); v->cap = p->caps;
    for (int i = 0; i < p->n; ++i) { v->beg[i] = p->chunks[i].out_beg; v->end[i] = p->chunks[i].out_end; v->n_out[i] = p->chunks[i].n_out; v->status[i] = p->chunks[i].status; }
    return 0;
}
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_sdust_plan_create(int n_chunks, const lcd_sdust_input_t *in) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && !in)) { set_error("lcd_sdust_plan_create: invalid arguments"); return nullptr; }
    sdust::SdustPlan *p = new sdust::SdustPlan();
    if (p->build(n_chunks, in)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}
int lcd_sdust_plan_fetch(lcd_plan_t *plan, void *stream, lcd_sdust_output_t *out) {
    sdust::SdustPlan *p = dynamic_cast<sdust::SdustPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_sdust_plan_fetch: not an sdust plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}
int lcd_sdust_batch(int n_chunks, const lcd_sdust_input_t *in, lcd_sdust_output_t *out) {
    lcd_plan_t *plan = lcd_sdust_plan_create(n_chunks, in);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_sdust_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
