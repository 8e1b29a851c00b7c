// sites_kernel.cu -- K1b launchers and host plan: the candidate-site list of many chunks as a counting sort on position bins
// followed by an exact pass per bin.  Device logic and design notes: sites_device.cuh; the bin offsets come from scan.cuh.
#include "lcd_common.cuh"
#include "sites_device.cuh"
#include "scan.cuh"
#include <algorithm>

namespace lcd {
namespace sites {

constexpr int THREADS = 128;

__global__ void __launch_bounds__(THREADS)
sites_count_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        count_read(a, g);
}

__global__ void __launch_bounds__(THREADS)
sites_scatter_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        scatter_read(a, g);
}

// grid.y = chunk, grid.x strides over the chunk's bins
__global__ void __launch_bounds__(THREADS)
sites_group_kernel(const KernelArgs a) {
    const int chunk = blockIdx.y; const long long b0 = a.chunks[chunk].bin0, nb = a.chunks[chunk].n_bins;
    for (long long b = (long long)blockIdx.x * THREADS + threadIdx.x; b < nb; b += (long long)gridDim.x * THREADS)
        group_bin(a, chunk, b0 + b);
}

__global__ void __launch_bounds__(THREADS)
sites_emit_kernel(const KernelArgs a, long long *chunk_site_off, int n_chunks) {
    const int chunk = blockIdx.y; const long long b0 = a.chunks[chunk].bin0, nb = a.chunks[chunk].n_bins;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        chunk_site_off[chunk] = a.keep_first[b0];
        if (chunk == n_chunks - 1) chunk_site_off[n_chunks] = a.keep_first[a.n_bins_total];
    }
    for (long long b = (long long)blockIdx.x * THREADS + threadIdx.x; b < nb; b += (long long)gridDim.x * THREADS)
        emit_bin(a, chunk, b0 + b);
}

template <typename T, typename U> static void append(std::vector<T> &dst, const U *src, size_t n, long long add = 0) {
    const size_t o = dst.size(); dst.resize(o + n);
    for (size_t i = 0; i < n; ++i) dst[o + i] = (T)(src[i] + (U)add);
}

struct SitesPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks; std::vector<long long> ev0;      // ev0[i]: first record of chunk i in the concatenated record arrays
    long long tot_reads = 0, tot_bins = 0, tot_events = 0, max_bins = 0;
    long long tot_cand = -1, tot_sites = -1;
    Plan *digar_plan = nullptr;
    // own copies of the read / record arrays (plans created on host lists)
    DevBuf<int32_t> d_read_chunk, d_ndig, d_dlen; DevBuf<uint8_t> d_active, d_dlow, d_dalt; DevBuf<int8_t> d_dtype; DevBuf<long long> d_dfirst, d_dpos, d_daoff;
    DevBuf<Chunk> d_chunks; DevBuf<int32_t> d_bin_count, d_bin_cursor, d_bin_keep, d_stype, d_sref, d_salt; DevBuf<long long> d_bin_first, d_keep_first, d_cand, d_spos, d_ssrc, d_saoff, d_tmp, d_chunk_site_off;
    KernelArgs base;
    std::vector<long long> h_site_off; bool have_index = false;

    // the chunk's bins: anchors of collectible records lie in [first read start - 1, last read end], cut to [reg_beg - 1, reg_end]
    void set_bins(Chunk &k, bool any, long long lo, long long hi, const lcd_sites_params_t &p) {
        if (!any) { lo = 0; hi = 0; }
        if (p.reg_beg != -1 && p.reg_beg - 1 > lo) lo = p.reg_beg - 1;
        if (p.reg_end != -1 && p.reg_end < hi) hi = p.reg_end;
        k.reg_beg = p.reg_beg; k.reg_end = p.reg_end; k.lo = lo; k.bin0 = tot_bins; k.n_bins = hi >= lo ? ((hi - lo) >> BIN_SHIFT) + 1 : 1;
        k.min_sv_len = p.min_sv_len; k.pad = 0;
        tot_bins += k.n_bins; max_bins = std::max<long long>(max_bins, k.n_bins);
    }

    int alloc_bins(cudaStream_t s) {
        if (d_chunks.upload(chunks.data(), n, s)) return -1;
        if (d_bin_count.alloc(tot_bins + 1) || d_bin_cursor.alloc(tot_bins + 1) || d_bin_keep.alloc(tot_bins + 1) || d_bin_first.alloc(tot_bins + 2) ||
            d_keep_first.alloc(tot_bins + 2) || d_chunk_site_off.alloc(n + 1)) return -1;
        return 0;
    }

    int build_on_digar(Plan *digar, int n_, const lcd_sites_params_t *par) {
        n = n_; digar_plan = digar;
        DigarView v;
        if (digar_plan_view(digar, cur_stream(), &v)) return -1;
        if (v.n_chunks != n) { set_error("lcd_sites: %d parameter sets for a digar plan of %d chunks", n, v.n_chunks); return -1; }
        if (n == 0) return 0;
        tot_reads = v.n_reads_total; tot_events = v.tot_events;
        chunks.resize(n); ev0.assign(n, 0);
        std::vector<int32_t> read_chunk; std::vector<long long> first(n);
        for (int i = 0; i < n; ++i) {
            long long lo = 0, hi = 0; bool any = false;
            for (long long g = v.read_off[i]; g < v.read_off[i + 1]; ++g) {
                read_chunk.push_back(i);
                if (!v.h_active[g]) continue;
                if (!any || v.h_beg[g] - 1 < lo) lo = v.h_beg[g] - 1;
                if (!any || v.h_end[g] + 1 > hi) hi = v.h_end[g] + 1;
                any = true;
            }
            set_bins(chunks[i], any, lo, hi, par[i]);
            chunks[i].alt_base = v.alt_base[i];
        }
        read_chunk.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s) || alloc_bins(s)) return -1;
        // first record of every chunk: the digar plan's scan output at the chunk's first read
        if (tot_reads) for (int i = 0; i < n; ++i) LCD_CUDA_OK(cudaMemcpyAsync(&ev0[i], v.dfirst + v.read_off[i], sizeof(long long), cudaMemcpyDeviceToHost, s));
        memset(&base, 0, sizeof(base));
        base.read_chunk = d_read_chunk.p; base.read_active = v.active; base.read_dropped = v.dropped; base.digar_first = v.dfirst; base.n_digar = v.ndig;
        base.digar_pos = v.dpos; base.digar_type = v.dtype; base.digar_len = v.dlen; base.digar_low_qual = v.dlow; base.digar_alt_off = v.daoff; base.digar_alt = v.dalt;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int build(int n_, const lcd_pileup_input_t *in, const lcd_sites_params_t *par) {
        n = n_;
        if (n == 0) return 0;
        std::vector<int32_t> read_chunk, ndig, dlen; std::vector<uint8_t> active, dlow, dalt; std::vector<int8_t> dtype; std::vector<long long> dfirst, dpos, daoff;
        chunks.resize(n); ev0.assign(n, 0);
        for (int i = 0; i < n; ++i) {
            const lcd_pileup_input_t &x = in[i];
            if (x.n_reads < 0) { set_error("lcd_sites: chunk %d has a negative read count", i); return -1; }
            long long n_ev = 0, n_alt = 0;
            for (int r = 0; r < x.n_reads; ++r) {
                if (x.n_digar[r] < 0 || x.digar_first[r] < 0) { set_error("lcd_sites: chunk %d read %d has an invalid record range", i, r); return -1; }
                n_ev = std::max<long long>(n_ev, x.digar_first[r] + x.n_digar[r]);
            }
            for (long long d = 0; d < n_ev; ++d) { const int t = x.digar_type[d]; if (t == CDIFF || t == CINS) n_alt = std::max<long long>(n_alt, x.digar_alt_off[d] + x.digar_len[d]); }
            std::vector<uint8_t> listed(x.n_reads, 0);
            for (int r = 0; r < x.n_reads; ++r) { const int id = x.ordered_read_ids[r]; if (id >= 0 && id < x.n_reads) listed[id] = 1; }
            long long lo = 0, hi = 0; bool any = false;
            for (int r = 0; r < x.n_reads; ++r) {
                const bool act = listed[r] && !x.is_skipped[r];
                read_chunk.push_back(i); active.push_back(act);
                if (!act) continue;
                if (!any || x.read_beg[r] - 1 < lo) lo = x.read_beg[r] - 1;
                if (!any || x.read_end[r] + 1 > hi) hi = x.read_end[r] + 1;
                any = true;
            }
            set_bins(chunks[i], any, lo, hi, par[i]);
            ev0[i] = (long long)dpos.size(); chunks[i].alt_base = (long long)dalt.size();
            append(dfirst, x.digar_first, x.n_reads, ev0[i]); append(ndig, x.n_digar, x.n_reads);
            append(dpos, x.digar_pos, (size_t)n_ev); append(dtype, x.digar_type, (size_t)n_ev); append(dlen, x.digar_len, (size_t)n_ev);
            append(dlow, x.digar_low_qual, (size_t)n_ev); append(daoff, x.digar_alt_off, (size_t)n_ev); append(dalt, x.digar_alt, (size_t)n_alt);
            tot_reads += x.n_reads; tot_events += n_ev;
        }
        auto pad = [](auto &v) { v.push_back(0); };
        pad(read_chunk); pad(active); pad(dfirst); pad(ndig); pad(dpos); pad(dtype); pad(dlen); pad(dlow); pad(daoff); pad(dalt);
        cudaStream_t s = cur_stream();
        if (d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s) || d_active.upload(active.data(), active.size(), s) || d_dfirst.upload(dfirst.data(), dfirst.size(), s) ||
            d_ndig.upload(ndig.data(), ndig.size(), s) || d_dpos.upload(dpos.data(), dpos.size(), s) || d_dtype.upload(dtype.data(), dtype.size(), s) ||
            d_dlen.upload(dlen.data(), dlen.size(), s) || d_dlow.upload(dlow.data(), dlow.size(), s) || d_daoff.upload(daoff.data(), daoff.size(), s) ||
            d_dalt.upload(dalt.data(), dalt.size(), s) || alloc_bins(s)) return -1;
        memset(&base, 0, sizeof(base));
        base.read_chunk = d_read_chunk.p; base.read_active = d_active.p; base.digar_first = d_dfirst.p; base.n_digar = d_ndig.p;
        base.digar_pos = d_dpos.p; base.digar_type = d_dtype.p; base.digar_len = d_dlen.p; base.digar_low_qual = d_dlow.p; base.digar_alt_off = d_daoff.p; base.digar_alt = d_dalt.p;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    void args(KernelArgs &a) {
        a = base;
        a.chunks = d_chunks.p; a.n_reads_total = tot_reads; a.n_bins_total = tot_bins;
        a.bin_count = d_bin_count.p; a.bin_first = d_bin_first.p; a.bin_cursor = d_bin_cursor.p; a.cand = d_cand.p; a.bin_keep = d_bin_keep.p; a.keep_first = d_keep_first.p;
        a.site_pos = d_spos.p; a.site_type = d_stype.p; a.site_ref_len = d_sref.p; a.site_alt_len = d_salt.p; a.site_src = d_ssrc.p; a.site_alt_off = d_saoff.p;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        have_index = false;
        if (n == 0) return 0;
        LCD_CUDA_OK(cudaMemsetAsync(d_bin_count.p, 0, sizeof(int32_t) * (tot_bins + 1), s));
        LCD_CUDA_OK(cudaMemsetAsync(d_bin_cursor.p, 0, sizeof(int32_t) * (tot_bins + 1), s));
        KernelArgs a; args(a);
        const int grid = (int)std::max<long long>(1, std::min<long long>((tot_reads + THREADS - 1) / THREADS, (long long)c.sm_count * 16));
        sites_count_kernel<<<grid, THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        if (scan::exclusive_scan(d_bin_count.p, tot_bins, d_bin_first.p, d_tmp, s)) return -1;
        if (tot_cand < 0) {         // first run: the candidate array is sized from the scan total (the sizes do not change between runs)
            LCD_DRAIN(s);
            LCD_CUDA_OK(cudaMemcpyAsync(&tot_cand, d_bin_first.p + tot_bins, sizeof(long long), cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
            if (d_cand.alloc(tot_cand + 1)) return -1;
            args(a);
        }
        sites_scatter_kernel<<<grid, THREADS, 0, s>>>(a);
        const dim3 bgrid((unsigned)std::max<long long>(1, std::min<long long>((max_bins + THREADS - 1) / THREADS, 1024)), (unsigned)n);
        sites_group_kernel<<<bgrid, THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches += 2;
        if (scan::exclusive_scan(d_bin_keep.p, tot_bins, d_keep_first.p, d_tmp, s)) return -1;
        if (tot_sites < 0) {
            LCD_DRAIN(s);
            LCD_CUDA_OK(cudaMemcpyAsync(&tot_sites, d_keep_first.p + tot_bins, sizeof(long long), cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
            if (d_spos.alloc(tot_sites + 1) || d_stype.alloc(tot_sites + 1) || d_sref.alloc(tot_sites + 1) || d_salt.alloc(tot_sites + 1) || d_ssrc.alloc(tot_sites + 1) || d_saoff.alloc(tot_sites + 1)) return -1;
            args(a);
        }
        sites_emit_kernel<<<bgrid, THREADS, 0, s>>>(a, d_chunk_site_off.p, n);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }

    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_events; return 0; }   // difference-list records filtered

    int index(cudaStream_t s) {
        if (have_index) return 0;
        h_site_off.assign(n + 1, 0);
        if (n) {
            if (tot_sites < 0) { set_error("lcd_sites: the plan has not been run"); return -1; }
            LCD_DRAIN(s);
            LCD_CUDA_OK(cudaMemcpyAsync(h_site_off.data(), d_chunk_site_off.p, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
        }
        have_index = true;
        return 0;
    }

    int fetch(cudaStream_t s, lcd_sites_output_t *out) {
        if (n == 0) return 0;
        if (index(s)) return -1;
        LCD_DRAIN(s);
        for (int i = 0; i < n; ++i) {
            const long long o = h_site_off[i], ns = h_site_off[i + 1] - o;
            out[i].n_sites = ns;
            if (ns > out[i].cap) { cudaStreamSynchronize(s); set_error("lcd_sites: chunk %d has %lld sites, the caller provided room for %lld (lcd_sites_plan_sizes)", i, ns, (long long)out[i].cap); return -3; }
            if (!ns) continue;
            LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_pos, d_spos.p + o, sizeof(long long) * ns, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_type, d_stype.p + o, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_ref_len, d_sref.p + o, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_alt_len, d_salt.p + o, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_src, d_ssrc.p + o, sizeof(long long) * ns, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i)         // record indices relative to the chunk's own lists
            for (long long k = 0; k < out[i].n_sites; ++k) out[i].site_src[k] -= ev0[i];
        return 0;
    }
};

} // namespace sites

int sites_plan_view(Plan *plan, Plan *digar, cudaStream_t s, SitesView *v) {
    sites::SitesPlan *p = dynamic_cast<sites::SitesPlan *>(plan);
    if (!p) { set_error("not a sites plan"); return -1; }
    if (p->digar_plan != digar) { set_error("the sites plan was not created on this digar plan"); return -1; }
    if (p->index(s)) return -1;
    v->n_chunks = p->n; v->site_off = p->h_site_off; v->min_sv_len.assign(p->n, 0);
    for (int i = 0; i < p->n; ++i) v->min_sv_len[i] = p->chunks[i].min_sv_len;
    v->spos = p->d_spos.p; v->saoff = p->d_saoff.p; v->stype = p->d_stype.p; v->sref = p->d_sref.p; v->salt = p->d_salt.p;
    return 0;
}
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_sites_plan_create(int n_chunks, const lcd_pileup_input_t *in, const lcd_sites_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && (!in || !params))) { set_error("lcd_sites_plan_create: invalid arguments"); return nullptr; }
    sites::SitesPlan *p = new sites::SitesPlan();
    if (p->build(n_chunks, in, params)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

lcd_plan_t *lcd_sites_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_sites_params_t *params) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || n_chunks < 0 || (n_chunks > 0 && !params)) { set_error("lcd_sites_plan_create_on_digar: invalid arguments"); return nullptr; }
    sites::SitesPlan *p = new sites::SitesPlan();
    if (p->build_on_digar(reinterpret_cast<Plan *>(digar_plan), n_chunks, params)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_sites_plan_sizes(lcd_plan_t *plan, void *stream, int chunk, int64_t *n_sites) {
    sites::SitesPlan *p = dynamic_cast<sites::SitesPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !n_sites || chunk < 0 || chunk >= p->n) { set_error("lcd_sites_plan_sizes: not a sites plan / chunk out of range / null output"); return -1; }
    if (p->index(pick_stream(stream))) return -1;
    *n_sites = p->h_site_off[chunk + 1] - p->h_site_off[chunk];
    return 0;
}

int lcd_sites_plan_fetch(lcd_plan_t *plan, void *stream, lcd_sites_output_t *out) {
    sites::SitesPlan *p = dynamic_cast<sites::SitesPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_sites_plan_fetch: not a sites plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}

int lcd_sites_batch(int n_chunks, const lcd_pileup_input_t *in, const lcd_sites_params_t *params, lcd_sites_output_t *out) {
    lcd_plan_t *plan = lcd_sites_plan_create(n_chunks, in, params);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_sites_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
