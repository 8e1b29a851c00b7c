// classify_kernel.cu -- K2b launcher and host plan: the category of every candidate site of many chunks, one thread per site.
// Device logic and design notes: classify_device.cuh.
#include "lcd_common.cuh"
#include "classify_device.cuh"
#include <algorithm>
#include <math.h>

namespace lcd {
namespace classify {

constexpr int THREADS = 128;

__global__ void __launch_bounds__(THREADS)
classify_kernel(const KernelArgs a) {
    for (long long s = (long long)blockIdx.x * THREADS + threadIdx.x; s < a.n_sites_total; s += (long long)gridDim.x * THREADS)
        classify_site(a, s);
}

struct ClassifyPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks; std::vector<long long> site_off;
    long long tot_sites = 0;
    // the site arrays and counters: this plan's own buffers, or a pileup plan's (K2 -> K2b in place)
    const long long *p_spos = nullptr, *p_saoff = nullptr; const int32_t *p_stype = nullptr, *p_sref = nullptr, *p_salt = nullptr, *p_counts = nullptr; const uint8_t *p_site_alt = nullptr;
    DevBuf<Chunk> d_chunks; DevBuf<int32_t> d_site_chunk, d_stype, d_sref, d_salt, d_counts, d_cate; DevBuf<long long> d_spos, d_saoff; DevBuf<uint8_t> d_site_alt; DevBuf<char> d_ref; DevBuf<int32_t> d_status; DevBuf<double> d_lgamma;

    int build(int n_, const lcd_classify_input_t *in) {
        n = n_;
        if (n == 0) return 0;
        chunks.resize(n); site_off.assign(n + 1, 0);
        long long tot_ref = 0, tot_alt = 0;
        std::vector<long long> alt_n(n, 0), ref_n(n, 0);
        std::vector<int32_t> site_chunk;
        for (int i = 0; i < n; ++i) {
            const lcd_classify_input_t &x = in[i];
            if (x.n_sites < 0 || x.ref_end < x.ref_beg || (x.n_sites > 0 && !x.ref_seq)) { set_error("lcd_classify: chunk %d has invalid sizes", i); return -1; }
            Chunk &k = chunks[i]; memset(&k, 0, sizeof(k));
            k.min_dp = x.min_dp; k.min_alt_dp = x.min_alt_dp; k.max_xgaps = x.max_xgaps; k.is_ont = x.is_ont ? 1 : 0; k.min_af = x.min_af; k.max_af = x.max_af;
            k.ref_beg = x.ref_beg; k.ref_end = x.ref_end; k.ref_off = tot_ref; k.alt_base = tot_alt;
            site_off[i] = tot_sites;
            for (int s = 0; s < x.n_sites; ++s) {
                const int t = x.site_type[s];
                if (t != CDIFF && t != CINS && t != CDEL) { set_error("lcd_classify: chunk %d site %d has type %d", i, s, t); return -1; }
                if (t == CDIFF || t == CINS) alt_n[i] = std::max<long long>(alt_n[i], x.site_alt_off[s] + x.site_alt_len[s]);
                const int len = t == CINS ? x.site_alt_len[s] : x.site_ref_len[s];
                if (t != CDIFF && len <= x.max_xgaps && (x.site_pos[s] - REF_MARGIN < x.ref_beg || x.site_pos[s] + len + REF_MARGIN > x.ref_end)) {
                    set_error("lcd_classify: chunk %d site %d (pos %lld) lies within %d bases of the reference window's ends [%lld, %lld]", i, s, (long long)x.site_pos[s],
                              REF_MARGIN, (long long)x.ref_beg, (long long)x.ref_end);
                    return -1;
                }
                site_chunk.push_back(i);
            }
            ref_n[i] = x.ref_end - x.ref_beg + 1;
            tot_sites += x.n_sites; tot_ref += (ref_n[i] + 15) & ~15ll; tot_alt += alt_n[i];
        }
        site_off[n] = tot_sites;
        site_chunk.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_site_chunk.upload(site_chunk.data(), site_chunk.size(), s)) return -1;
        // the site lists, counters and reference windows go straight from the caller's arrays into their slice of the device arrays
        if (d_spos.alloc(tot_sites + 1) || d_stype.alloc(tot_sites + 1) || d_sref.alloc(tot_sites + 1) || d_salt.alloc(tot_sites + 1) || d_saoff.alloc(tot_sites + 1) ||
            d_counts.alloc(8 * (size_t)tot_sites + 8) || d_cate.alloc(tot_sites + 1) || d_site_alt.alloc(tot_alt + 1) || d_ref.alloc(tot_ref + 16)) return -1;
        for (int i = 0; i < n; ++i) {
            const lcd_classify_input_t &x = in[i]; const long long o = site_off[i]; const size_t ns = (size_t)x.n_sites;
            if (ns) {
                LCD_CUDA_OK(cudaMemcpyAsync(d_spos.p + o, x.site_pos, sizeof(long long) * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_stype.p + o, x.site_type, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_sref.p + o, x.site_ref_len, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_salt.p + o, x.site_alt_len, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_saoff.p + o, x.site_alt_off, sizeof(long long) * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_counts.p + 8 * o, x.site_counts, sizeof(int32_t) * 8 * ns, cudaMemcpyHostToDevice, s));
                LCD_CUDA_OK(cudaMemcpyAsync(d_ref.p + chunks[i].ref_off, x.ref_seq, (size_t)ref_n[i], cudaMemcpyHostToDevice, s));
            }
            if (alt_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_site_alt.p + chunks[i].alt_base, x.site_alt, (size_t)alt_n[i], cudaMemcpyHostToDevice, s));
        }
        p_spos = d_spos.p; p_saoff = d_saoff.p; p_stype = d_stype.p; p_sref = d_sref.p; p_salt = d_salt.p; p_counts = d_counts.p; p_site_alt = d_site_alt.p;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    // K2b on the sites and counters a pileup plan holds in HBM: only the thresholds and the reference windows are uploaded.  The reference
    // reads the window unchecked; here the sites' distance from the window's ends is not known to the host, so the window must cover the
    // chunk's reads with the reference's own flank (checked on the device side by nothing: documented precondition of the call).
    int build_on_pileup(Plan *pileup, int n_, const lcd_classify_params_t *par) {
        n = n_;
        PileupView v;
        if (pileup_plan_view(pileup, &v)) return -1;
        if (v.n_chunks != n) { set_error("lcd_classify: %d parameter sets for a pileup plan of %d chunks", n, v.n_chunks); return -1; }
        if (n == 0) return 0;
        chunks.resize(n); site_off = v.site_off; tot_sites = v.site_off[n];
        long long tot_ref = 0;
        std::vector<int32_t> site_chunk; std::vector<long long> ref_n(n, 0);
        for (int i = 0; i < n; ++i) {
            const lcd_classify_params_t &x = par[i];
            if (x.ref_end < x.ref_beg || !x.ref_seq) { set_error("lcd_classify: chunk %d has no reference window", i); return -1; }
            Chunk &k = chunks[i]; memset(&k, 0, sizeof(k));
            k.min_dp = x.min_dp; k.min_alt_dp = x.min_alt_dp; k.max_xgaps = x.max_xgaps; k.is_ont = x.is_ont ? 1 : 0; k.min_af = x.min_af; k.max_af = x.max_af;
            k.ref_beg = x.ref_beg; k.ref_end = x.ref_end; k.ref_off = tot_ref; k.alt_base = v.salt_base[i];
            ref_n[i] = x.ref_end - x.ref_beg + 1; tot_ref += (ref_n[i] + 15) & ~15ll;
            for (long long s = site_off[i]; s < site_off[i + 1]; ++s) site_chunk.push_back(i);
        }
        site_chunk.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_site_chunk.upload(site_chunk.data(), site_chunk.size(), s) || d_cate.alloc(tot_sites + 1) || d_ref.alloc(tot_ref + 16)) return -1;
        for (int i = 0; i < n; ++i) LCD_CUDA_OK(cudaMemcpyAsync(d_ref.p + chunks[i].ref_off, par[i].ref_seq, (size_t)ref_n[i], cudaMemcpyHostToDevice, s));
        p_spos = v.spos; p_saoff = v.saoff; p_stype = v.stype; p_sref = v.sref; p_salt = v.salt; p_counts = v.counts; p_site_alt = v.site_alt;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    // opt->lgamma_cache (initialize_lgamma_cache, src/math_utils.c:6-11): the host's libm, as the reference fills it
    int upload_lgamma(cudaStream_t s) {
        if (d_lgamma.p) return 0;
        std::vector<double> h(LGAMMA_MAX_I + 1);
        for (int i = 0; i <= LGAMMA_MAX_I; ++i) h[i] = lgamma((double)i);
        if (d_lgamma.upload(h.data(), h.size(), s)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(d_lgamma.st));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0 || tot_sites == 0) return 0;
        if (upload_lgamma(s)) return -1;
        KernelArgs a; memset(&a, 0, sizeof(a));
        a.chunks = d_chunks.p; a.n_sites_total = tot_sites; a.site_chunk = d_site_chunk.p; a.site_pos = p_spos; a.site_type = p_stype; a.site_ref_len = p_sref;
        a.site_alt_len = p_salt; a.site_alt_off = p_saoff; a.site_alt = p_site_alt; a.site_counts = p_counts; a.ref = d_ref.p; a.var_cate = d_cate.p; a.lgamma_cache = d_lgamma.p;
        if (!d_status.p && d_status.alloc(1)) return -1;
        LCD_CUDA_OK(cudaMemsetAsync(d_status.p, 0, sizeof(int32_t), s));
        a.status = d_status.p;
        const int grid = (int)std::min<long long>((tot_sites + THREADS - 1) / THREADS, (long long)c.sm_count * 16);
        classify_kernel<<<grid, THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }

    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_sites; return 0; }   // sites classified

    int fetch(cudaStream_t s, lcd_classify_output_t *out) {
        if (n == 0) return 0;
        LCD_DRAIN(s);
        if (d_status.p) {
            int32_t st = 0;
            LCD_CUDA_OK(cudaMemcpyAsync(&st, d_status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
            if (st) { set_error("lcd_classify: a small indel lies within %d bases of its chunk's reference window ends", REF_MARGIN); return -3; }
        }
        for (int i = 0; i < n; ++i) {
            const long long ns = site_off[i + 1] - site_off[i];
            if (ns) LCD_CUDA_OK(cudaMemcpyAsync(out[i].var_cate, d_cate.p + site_off[i], sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }
};

} // namespace classify

int classify_plan_view(Plan *plan, ClassifyView *v) {
    classify::ClassifyPlan *p = dynamic_cast<classify::ClassifyPlan *>(plan);
    if (!p) { set_error("not a classify (K2b) plan"); return -1; }
    v->n_chunks = p->n; v->site_off = p->site_off;
    if ((int)v->site_off.size() != p->n + 1) v->site_off.assign(p->n + 1, 0);
    v->spos = p->p_spos; v->stype = p->p_stype; v->sref = p->p_sref; v->cate = p->d_cate.p;
    return 0;
}
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_classify_plan_create(int n_chunks, const lcd_classify_input_t *in) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && !in)) { set_error("lcd_classify_plan_create: invalid arguments"); return nullptr; }
    classify::ClassifyPlan *p = new classify::ClassifyPlan();
    if (p->build(n_chunks, in)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

lcd_plan_t *lcd_classify_plan_create_on_pileup(lcd_plan_t *pileup_plan, int n_chunks, const lcd_classify_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (!pileup_plan || n_chunks < 0 || (n_chunks > 0 && !params)) { set_error("lcd_classify_plan_create_on_pileup: invalid arguments"); return nullptr; }
    classify::ClassifyPlan *p = new classify::ClassifyPlan();
    if (p->build_on_pileup(reinterpret_cast<Plan *>(pileup_plan), n_chunks, params)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_classify_plan_fetch(lcd_plan_t *plan, void *stream, lcd_classify_output_t *out) {
    classify::ClassifyPlan *p = dynamic_cast<classify::ClassifyPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_classify_plan_fetch: not a classify plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}

int lcd_classify_batch(int n_chunks, const lcd_classify_input_t *in, lcd_classify_output_t *out) {
    lcd_plan_t *plan = lcd_classify_plan_create(n_chunks, in);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_classify_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
