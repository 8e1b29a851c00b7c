// noisyreg_kernel.cu -- K2c launcher and host plan: the noisy-region set of many chunks and the candidate sites that stay clean-region
// candidates, one CTA per chunk.  Device logic and design notes: noisyreg_device.cuh.
#include "lcd_common.cuh"
#include "noisyreg_device.cuh"
#include <algorithm>

namespace lcd {
namespace noisyreg {

constexpr int THREADS = 256;
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

__global__ void __launch_bounds__(THREADS)
noisyreg_kernel(const Chunk *chunks, int n) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) { run_chunk(chunks[i], (int)threadIdx.x, THREADS, CtaSync()); __syncthreads(); }
}

struct NoisyRegPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks;
    std::vector<long long> out_cate_off, out_keep_off, out_reg_off, out_nreg_off, out_status_off;     // byte offsets in the work blob
    std::vector<int> n_sites, reg_cap;
    DevBuf<uint8_t> d_in, d_work; DevBuf<Chunk> d_chunks;
    std::vector<uint8_t> h_in;
    long long tot_sites = 0;

    int build(int n_, const lcd_noisyreg_input_t *in) {
        n = n_;
        if (n == 0) return 0;
        chunks.resize(n); n_sites.resize(n); reg_cap.resize(n);
        out_cate_off.resize(n); out_keep_off.resize(n); out_reg_off.resize(n); out_nreg_off.resize(n); out_status_off.resize(n);
        // layout pass: input blob (one upload) and work blob (outputs + scratch)
        size_t in_bytes = 0, work_bytes = 0;
        auto take = [](size_t &top, size_t bytes) { const size_t at = top; top += (bytes + 15) & ~(size_t)15; return at; };
        struct Seg { size_t off; const void *src; size_t bytes; };
        std::vector<Seg> segs;
        std::vector<std::vector<size_t>> in_off(n), wk_off(n);
        for (int i = 0; i < n; ++i) {
            const lcd_noisyreg_input_t &x = in[i];
            if (x.n_sites < 0 || x.n_reads < 0 || x.n_cnreg < 0 || x.n_low < 0) { set_error("lcd_noisyreg: chunk %d has negative sizes", i); return -1; }
            if (x.is_ont < 0 || x.noisy_reg_flank_len < 0) { set_error("lcd_noisyreg: chunk %d has invalid options", i); return -1; }
            for (int s = 1; s < x.n_sites; ++s) {          // collect_all_cand_var_sites' order: by anchor (exact_comp_var_site, src/collect_var.c:1878)
                const long long a0 = x.site_pos[s - 1] - (x.site_type[s - 1] == 8 ? 0 : 1), a1 = x.site_pos[s] - (x.site_type[s] == 8 ? 0 : 1);
                if (a1 < a0) { set_error("lcd_noisyreg: chunk %d: candidate sites must ascend by anchor position (site %d)", i, s); return -1; }
            }
            for (long long k = 0; k < x.n_cnreg; ++k) if (x.cnreg_label[k] < 0) { set_error("lcd_noisyreg: chunk %d: noisy interval %lld has a negative label", i, k); return -1; }
            for (long long k = 1; k < x.n_low; ++k)
                if (x.low_beg[k] < x.low_beg[k - 1]) { set_error("lcd_noisyreg: chunk %d: low-complexity intervals must ascend by start (interval %lld)", i, k); return -1; }
            long long nd = 0, nn = 0;
            for (int r = 0; r < x.n_reads; ++r) {
                if (x.n_digar[r] < 0 || x.n_nreg[r] < 0 || (x.n_digar[r] > 0 && x.digar_first[r] < 0) || (x.n_nreg[r] > 0 && x.nreg_first[r] < 0)) { set_error("lcd_noisyreg: chunk %d read %d has invalid record ranges", i, r); return -1; }
                if (x.n_digar[r] > 0) nd = std::max<long long>(nd, x.digar_first[r] + x.n_digar[r]);
                if (x.n_nreg[r] > 0) nn = std::max<long long>(nn, x.nreg_first[r] + x.n_nreg[r]);
            }
            const size_t ns = (size_t)x.n_sites, nr = (size_t)x.n_reads, nc = (size_t)x.n_cnreg, nl = (size_t)x.n_low;
            const Seg s_[] = {
                {0, x.site_pos, ns * 8}, {0, x.site_type, ns * 4}, {0, x.site_ref_len, ns * 4}, {0, x.var_cate, ns * 4},
                {0, x.cnreg_beg, nc * 8}, {0, x.cnreg_end, nc * 8}, {0, x.cnreg_label, nc * 4}, {0, x.low_beg, nl * 8}, {0, x.low_end, nl * 8},
                {0, x.is_skipped, nr}, {0, x.read_beg, nr * 8}, {0, x.read_end, nr * 8}, {0, x.digar_first, nr * 8}, {0, x.n_digar, nr * 4},
                {0, x.digar_pos, (size_t)nd * 8}, {0, x.digar_type, (size_t)nd}, {0, x.digar_len, (size_t)nd * 4},
                {0, x.nreg_first, nr * 8}, {0, x.n_nreg, nr * 4}, {0, x.nreg_beg, (size_t)nn * 8}, {0, x.nreg_end, (size_t)nn * 8} };
            for (const Seg &sg : s_) {
                if (sg.bytes && !sg.src) { set_error("lcd_noisyreg: chunk %d has a null input array", i); return -1; }
                const size_t at = take(in_bytes, sg.bytes); in_off[i].push_back(at); segs.push_back({at, sg.src, sg.bytes});
            }
            const size_t cap = nc + ns + 8;
            n_sites[i] = x.n_sites; reg_cap[i] = (int)cap; tot_sites += x.n_sites;
            // work blob: var_cate, keep, out_beg, out_end, out_label, n_regs, status | A (3 x cap), B (3 x cap), low_pmax, vp_pmax, tot, noi, ctr
            const size_t wsz[] = { ns * 4, ns, cap * 8, cap * 8, cap * 4, 8, 4, cap * 4, cap * 4, cap * 4, cap * 4, cap * 4, cap * 4, nl * 4, ns * 4, cap * 4, cap * 4, 16 };
            for (size_t b : wsz) wk_off[i].push_back(take(work_bytes, b));
            out_cate_off[i] = (long long)wk_off[i][0]; out_keep_off[i] = (long long)wk_off[i][1]; out_reg_off[i] = (long long)wk_off[i][2];
            out_nreg_off[i] = (long long)wk_off[i][5]; out_status_off[i] = (long long)wk_off[i][6];
        }
        h_in.assign(in_bytes + 16, 0);
        for (const Seg &sg : segs) if (sg.bytes) memcpy(h_in.data() + sg.off, sg.src, sg.bytes);
        cudaStream_t s = cur_stream();
        if (d_in.upload(h_in.data(), in_bytes + 16, s) || d_work.alloc(work_bytes + 16)) return -1;
        for (int i = 0; i < n; ++i) {
            const lcd_noisyreg_input_t &x = in[i];
            Chunk &c = chunks[i]; memset(&c, 0, sizeof(c));
            c.reg_beg = x.reg_beg; c.reg_end = x.reg_end; c.min_af = x.min_af; c.min_alt_dp = x.min_alt_dp; c.flank = x.noisy_reg_flank_len; c.is_ont = x.is_ont ? 1 : 0;
            c.n_sites = x.n_sites; c.n_reads = x.n_reads; c.n_cnreg = (int)x.n_cnreg; c.n_low = (int)x.n_low; c.cap = reg_cap[i];
            const uint8_t *b = d_in.p; const std::vector<size_t> &o = in_off[i];
            c.site_pos = (const long long *)(b + o[0]); c.site_type = (const int *)(b + o[1]); c.site_ref_len = (const int *)(b + o[2]); c.var_cate_in = (const int *)(b + o[3]);
            c.cn_beg = (const long long *)(b + o[4]); c.cn_end = (const long long *)(b + o[5]); c.cn_label = (const int *)(b + o[6]);
            c.low_beg = (const long long *)(b + o[7]); c.low_end = (const long long *)(b + o[8]);
            c.is_skipped = b + o[9]; c.read_beg = (const long long *)(b + o[10]); c.read_end = (const long long *)(b + o[11]); c.digar_first = (const long long *)(b + o[12]);
            c.n_digar = (const int *)(b + o[13]); c.digar_pos = (const long long *)(b + o[14]); c.digar_type = (const signed char *)(b + o[15]); c.digar_len = (const int *)(b + o[16]);
            c.nreg_first = (const long long *)(b + o[17]); c.n_nreg = (const int *)(b + o[18]); c.nreg_beg = (const long long *)(b + o[19]); c.nreg_end = (const long long *)(b + o[20]);
            uint8_t *w = d_work.p; const std::vector<size_t> &k = wk_off[i];
            c.var_cate = (int *)(w + k[0]); c.keep = w + k[1]; c.out_beg = (long long *)(w + k[2]); c.out_end = (long long *)(w + k[3]); c.out_label = (int *)(w + k[4]);
            c.reg_cap = reg_cap[i]; c.n_regs = (long long *)(w + k[5]); c.status = (int *)(w + k[6]);
            c.A.st = (int *)(w + k[7]); c.A.en = (int *)(w + k[8]); c.A.label = (int *)(w + k[9]); c.B.st = (int *)(w + k[10]); c.B.en = (int *)(w + k[11]); c.B.label = (int *)(w + k[12]);
            c.low_pmax = (int *)(w + k[13]); c.vp_pmax = (int *)(w + k[14]); c.tot = (int *)(w + k[15]); c.noi = (int *)(w + k[16]); c.ctr = (int *)(w + k[17]);
        }
        if (d_chunks.upload(chunks.data(), n, s)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        noisyreg_kernel<<<std::min(n, c.sm_count * 4), THREADS, 0, s>>>(d_chunks.p, n);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }
    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_sites; return 0; }   // candidate sites examined

    int fetch(cudaStream_t s, lcd_noisyreg_output_t *out) {
        if (n == 0) return 0;
        LCD_DRAIN(s);
        std::vector<long long> nreg(n); std::vector<int> st(n);
        for (int i = 0; i < n; ++i) {
            LCD_CUDA_OK(cudaMemcpyAsync(&nreg[i], d_work.p + out_nreg_off[i], 8, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(&st[i], d_work.p + out_status_off[i], 4, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) {
            if (st[i] != ST_OK) { set_error("lcd_noisyreg: chunk %d failed on the device (status %d: interval list capacity)", i, st[i]); return -2; }
            out[i].n_regs = nreg[i];
            if (nreg[i] > out[i].reg_cap) { set_error("lcd_noisyreg: chunk %d has %lld noisy regions, the caller's arrays hold %lld", i, nreg[i], (long long)out[i].reg_cap); return -3; }
            const Chunk &c = chunks[i];
            if (n_sites[i]) {
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].var_cate, c.var_cate, sizeof(int32_t) * n_sites[i], cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].keep, c.keep, (size_t)n_sites[i], cudaMemcpyDeviceToHost, s));
            }
            if (nreg[i]) {
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].reg_beg, c.out_beg, 8 * (size_t)nreg[i], cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].reg_end, c.out_end, 8 * (size_t)nreg[i], cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].reg_label, c.out_label, 4 * (size_t)nreg[i], cudaMemcpyDeviceToHost, s));
            }
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }
};

} // namespace noisyreg
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_noisyreg_plan_create(int n_chunks, const lcd_noisyreg_input_t *in) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && !in)) { set_error("lcd_noisyreg_plan_create: invalid arguments"); return nullptr; }
    noisyreg::NoisyRegPlan *p = new noisyreg::NoisyRegPlan();
    if (p->build(n_chunks, in)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}
int lcd_noisyreg_plan_fetch(lcd_plan_t *plan, void *stream, lcd_noisyreg_output_t *out) {
    noisyreg::NoisyRegPlan *p = dynamic_cast<noisyreg::NoisyRegPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_noisyreg_plan_fetch: not a noisy-region plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}
int lcd_noisyreg_batch(int n_chunks, const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) {
    lcd_plan_t *plan = lcd_noisyreg_plan_create(n_chunks, in);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_noisyreg_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
