// noisyreg_kernel.cu -- K2c launcher and host plan: the noisy-region set of many chunks and the candidate sites that stay clean-region
// candidates, one CTA per chunk.  Device logic and design notes: noisyreg_device.cuh.
#include "lcd_common.cuh"
#include "noisyreg_device.cuh"
#include <algorithm>

namespace lcd {
namespace noisyreg {

constexpr int THREADS = 128;       // a chunk's critical path is its one-thread phases: small CTAs, so that more chunks are resident in the CTA slots the DP grids leave free
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

__global__ void __launch_bounds__(THREADS, 4)
noisyreg_kernel(const Chunk *chunks, int n) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) { run_chunk(chunks[i], (int)threadIdx.x, THREADS, CtaSync()); __syncthreads(); }
}

struct NoisyRegPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks;
    // work blob: [per-chunk headers: n_regs (8 B), status (4 B)] [var_cate | keep of every chunk] [region lists] [scratch]: a fetch is one copy of the
    // headers, one of the categories + keep flags of all chunks, and one per chunk of its n_regs x 20 bytes
    std::vector<long long> out_cate_off, out_keep_off, out_reg_off;     // byte offsets in the work blob
    size_t ck_beg = 0, ck_end = 0;                                      // the categories + keep flags of all chunks
    std::vector<uint8_t> h_ck;
    std::vector<int> n_sites, reg_cap;
    DevBuf<uint8_t> d_in, d_work; DevBuf<Chunk> d_chunks;
    std::vector<uint8_t> h_in;
    long long tot_sites = 0;

    struct WorkOff { size_t hdr, cate, keep, regs, a[3], b[3], low_pmax, vp_pmax, tot, noi, ctr; };
    std::vector<WorkOff> wk;
    // offsets of every chunk's outputs and scratch in the work blob; returns its size
    size_t layout_work(const std::vector<size_t> &ns, const std::vector<size_t> &cap, const std::vector<size_t> &nl) {
        auto take = [](size_t &top, size_t bytes) { const size_t at = top; top += (bytes + 15) & ~(size_t)15; return at; };
        size_t top = 0;
        wk.resize(n);
        for (int i = 0; i < n; ++i) wk[i].hdr = take(top, 16);
        ck_beg = top;
        for (int i = 0; i < n; ++i) { wk[i].cate = take(top, ns[i] * 4); wk[i].keep = take(top, ns[i]); }
        ck_end = top;
        for (int i = 0; i < n; ++i) wk[i].regs = take(top, cap[i] * 20);
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) wk[i].a[k] = take(top, cap[i] * 4);
            for (int k = 0; k < 3; ++k) wk[i].b[k] = take(top, cap[i] * 4);
            wk[i].low_pmax = take(top, nl[i] * 4); wk[i].vp_pmax = take(top, ns[i] * 4); wk[i].tot = take(top, cap[i] * 4); wk[i].noi = take(top, cap[i] * 4); wk[i].ctr = take(top, 16);
            out_cate_off[i] = (long long)wk[i].cate; out_keep_off[i] = (long long)wk[i].keep; out_reg_off[i] = (long long)wk[i].regs;
        }
        return top;
    }
    void wire_work(Chunk &c, int i) {
        uint8_t *w = d_work.p; const WorkOff &k = wk[i];
        c.var_cate = (int *)(w + k.cate); c.keep = w + k.keep; c.out_regs = (long long *)(w + k.regs);
        c.reg_cap = reg_cap[i]; c.n_regs = (long long *)(w + k.hdr); c.status = (int *)(w + k.hdr + 8);
        c.A.st = (int *)(w + k.a[0]); c.A.en = (int *)(w + k.a[1]); c.A.label = (int *)(w + k.a[2]); c.B.st = (int *)(w + k.b[0]); c.B.en = (int *)(w + k.b[1]); c.B.label = (int *)(w + k.b[2]);
        c.low_pmax = (int *)(w + k.low_pmax); c.vp_pmax = (int *)(w + k.vp_pmax); c.tot = (int *)(w + k.tot); c.noi = (int *)(w + k.noi); c.ctr = (int *)(w + k.ctr);
    }

    int build(int n_, const lcd_noisyreg_input_t *in) {
        n = n_;
        if (n == 0) return 0;
        chunks.resize(n); n_sites.resize(n); reg_cap.resize(n);
        out_cate_off.resize(n); out_keep_off.resize(n); out_reg_off.resize(n);
        // layout pass: input blob (one upload) and work blob (outputs + scratch)
        size_t in_bytes = 0, work_bytes = 0;
        auto take = [](size_t &top, size_t bytes) { const size_t at = top; top += (bytes + 15) & ~(size_t)15; return at; };
        struct Seg { size_t off; const void *src; size_t bytes; };
        std::vector<Seg> segs;
        std::vector<std::vector<size_t>> in_off(n); std::vector<size_t> v_ns, v_cap, v_nl;
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
            v_ns.push_back(ns); v_cap.push_back(cap); v_nl.push_back(nl);
        }
        work_bytes = layout_work(v_ns, v_cap, v_nl);
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
            wire_work(c, i);
        }
        if (d_chunks.upload(chunks.data(), n, s)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    // K2c on what K1 and K2b left in HBM: the reads' spans, records and noisy intervals (the chunk's own list is gathered from them on the device),
    // the sites and their categories; only the options and the low-complexity intervals are uploaded.
    // sdust != nullptr: the low-complexity intervals are read where a K0 plan leaves them (their number too: nothing of them passes through the host)
    int build_on(Plan *digar, Plan *classify, int n_, const lcd_noisyreg_params_t *par, Plan *sdust = nullptr) {
        n = n_;
        DigarView dv; ClassifyView cv; SdustView sv;
        if (sdust) { if (sdust_plan_view(sdust, &sv)) return -1; if (sv.n_chunks != n_) { set_error("lcd_noisyreg: %d parameter sets for an sdust plan of %d chunks", n_, sv.n_chunks); return -1; } }
        dv.want_host_spans = false;
        if (digar_plan_view(digar, cur_stream(), &dv) || classify_plan_view(classify, &cv)) return -1;
        if (dv.n_chunks != n || cv.n_chunks != n) { set_error("lcd_noisyreg: %d parameter sets for a digar plan of %d and a classify plan of %d chunks", n, dv.n_chunks, cv.n_chunks); return -1; }
        if (n == 0) return 0;
        chunks.resize(n); n_sites.resize(n); reg_cap.resize(n);
        out_cate_off.resize(n); out_keep_off.resize(n); out_reg_off.resize(n);
        size_t in_bytes = 0, work_bytes = 0;
        auto take = [](size_t &top, size_t bytes) { const size_t at = top; top += (bytes + 15) & ~(size_t)15; return at; };
        std::vector<std::vector<size_t>> in_off(n); std::vector<size_t> v_ns, v_cap, v_nl;
        for (int i = 0; i < n; ++i) {
            const lcd_noisyreg_params_t &x = par[i];
            if (x.n_low < 0 || (x.n_low > 0 && (!x.low_beg || !x.low_end)) || x.noisy_reg_flank_len < 0) { set_error("lcd_noisyreg: chunk %d has invalid options", i); return -1; }
            for (long long k = 1; k < x.n_low; ++k) if (x.low_beg[k] < x.low_beg[k - 1]) { set_error("lcd_noisyreg: chunk %d: low-complexity intervals must ascend by start (interval %lld)", i, k); return -1; }
            if (sdust && x.n_low) { set_error("lcd_noisyreg: chunk %d passes low-complexity intervals of its own to a plan chained to an sdust plan", i); return -1; }
            const size_t nl = sdust ? (size_t)sv.cap[i] : (size_t)x.n_low, ns = (size_t)(cv.site_off[i + 1] - cv.site_off[i]);      // (chained: scratch for as many as K0 can write)
            in_off[i].push_back(take(in_bytes, (size_t)x.n_low * 8)); in_off[i].push_back(take(in_bytes, (size_t)x.n_low * 8));
            const size_t cap = (size_t)dv.nreg_total[i] + ns + 8;
            n_sites[i] = (int)ns; reg_cap[i] = (int)cap; tot_sites += (long long)ns;
            v_ns.push_back(ns); v_cap.push_back(cap); v_nl.push_back(nl);
        }
        work_bytes = layout_work(v_ns, v_cap, v_nl);
        h_in.assign(in_bytes + 16, 0);
        for (int i = 0; i < n; ++i) if (par[i].n_low) { memcpy(h_in.data() + in_off[i][0], par[i].low_beg, (size_t)par[i].n_low * 8); memcpy(h_in.data() + in_off[i][1], par[i].low_end, (size_t)par[i].n_low * 8); }
        cudaStream_t s = cur_stream();
        if (d_in.upload(h_in.data(), in_bytes + 16, s) || d_work.alloc(work_bytes + 16)) return -1;
        for (int i = 0; i < n; ++i) {
            const lcd_noisyreg_params_t &x = par[i];
            Chunk &c = chunks[i]; memset(&c, 0, sizeof(c));
            const long long r0 = dv.read_off[i], s0 = cv.site_off[i];
            c.reg_beg = dv.reg_beg[i]; c.reg_end = dv.reg_end[i]; c.min_af = x.min_af; c.min_alt_dp = x.min_alt_dp; c.flank = x.noisy_reg_flank_len; c.is_ont = x.is_ont ? 1 : 0;
            c.n_sites = n_sites[i]; c.n_reads = (int)(dv.read_off[i + 1] - r0); c.n_cnreg = 0; c.n_low = (int)x.n_low; c.cap = reg_cap[i]; c.cn_from_reads = 1;
            c.site_pos = cv.spos + s0; c.site_type = cv.stype + s0; c.site_ref_len = cv.sref + s0; c.var_cate_in = cv.cate + s0;
            c.low_beg = (const long long *)(d_in.p + in_off[i][0]); c.low_end = (const long long *)(d_in.p + in_off[i][1]);
            if (sdust) { c.low_beg = sv.beg[i]; c.low_end = sv.end[i]; c.n_low_dev = sv.n_out[i]; c.low_status = sv.status[i]; }
            c.is_skipped = dv.dropped + r0; c.active = dv.active + r0; c.read_beg = dv.beg + r0; c.read_end = dv.end + r0; c.digar_first = dv.dfirst + r0; c.n_digar = dv.ndig + r0;
            c.digar_pos = dv.dpos; c.digar_type = (const signed char *)dv.dtype; c.digar_len = dv.dlen;
            c.nreg_first = dv.nfirst + r0; c.n_nreg = dv.nnreg + r0; c.nreg_beg = dv.nbeg; c.nreg_end = dv.nend; c.nreg_label = dv.nlabel;
            wire_work(c, i);
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
        std::vector<uint8_t> hdr((size_t)16 * n);
        h_ck.resize(ck_end - ck_beg + 16);
        LCD_CUDA_OK(cudaMemcpyAsync(hdr.data(), d_work.p + wk[0].hdr, hdr.size(), cudaMemcpyDeviceToHost, s));
        if (ck_end > ck_beg) LCD_CUDA_OK(cudaMemcpyAsync(h_ck.data(), d_work.p + ck_beg, ck_end - ck_beg, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        std::vector<std::vector<uint8_t>> regs(n);
        for (int i = 0; i < n; ++i) {
            long long nreg; int st; memcpy(&nreg, hdr.data() + 16 * (size_t)i, 8); memcpy(&st, hdr.data() + 16 * (size_t)i + 8, 4);
            if (st != ST_OK) { set_error("lcd_noisyreg: chunk %d failed on the device (status %d: %s)", i, st, st == ST_LOW ? "the sdust plan it reads its low-complexity intervals from failed for this chunk" : "interval list capacity"); return -2; }
            out[i].n_regs = nreg;
            if (nreg > out[i].reg_cap) { set_error("lcd_noisyreg: chunk %d has %lld noisy regions, the caller's arrays hold %lld", i, nreg, (long long)out[i].reg_cap); return -3; }
            if (n_sites[i]) {
                memcpy(out[i].var_cate, h_ck.data() + (out_cate_off[i] - (long long)ck_beg), sizeof(int32_t) * (size_t)n_sites[i]);
                memcpy(out[i].keep, h_ck.data() + (out_keep_off[i] - (long long)ck_beg), (size_t)n_sites[i]);
            }
            if (nreg) { regs[i].resize((size_t)nreg * 20); LCD_CUDA_OK(cudaMemcpyAsync(regs[i].data(), d_work.p + out_reg_off[i], (size_t)nreg * 20, cudaMemcpyDeviceToHost, s)); }
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) {
            const size_t k = (size_t)out[i].n_regs;
            if (!k) continue;
            memcpy(out[i].reg_beg, regs[i].data(), 8 * k); memcpy(out[i].reg_end, regs[i].data() + 8 * k, 8 * k); memcpy(out[i].reg_label, regs[i].data() + 16 * k, 4 * k);
        }
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
lcd_plan_t *lcd_noisyreg_plan_create_on_classify(lcd_plan_t *digar_plan, lcd_plan_t *classify_plan, int n_chunks, const lcd_noisyreg_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || !classify_plan || n_chunks < 0 || (n_chunks > 0 && !params)) { set_error("lcd_noisyreg_plan_create_on_classify: invalid arguments"); return nullptr; }
    noisyreg::NoisyRegPlan *p = new noisyreg::NoisyRegPlan();
    if (p->build_on(reinterpret_cast<Plan *>(digar_plan), reinterpret_cast<Plan *>(classify_plan), n_chunks, params)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}
lcd_plan_t *lcd_noisyreg_plan_create_on_sdust(lcd_plan_t *digar_plan, lcd_plan_t *classify_plan, lcd_plan_t *sdust_plan, int n_chunks, const lcd_noisyreg_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || !classify_plan || !sdust_plan || n_chunks < 0 || (n_chunks > 0 && !params)) { set_error("lcd_noisyreg_plan_create_on_sdust: invalid arguments"); return nullptr; }
    noisyreg::NoisyRegPlan *p = new noisyreg::NoisyRegPlan();
    if (p->build_on(reinterpret_cast<Plan *>(digar_plan), reinterpret_cast<Plan *>(classify_plan), n_chunks, params, reinterpret_cast<Plan *>(sdust_plan))) { delete p; return nullptr; }
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
