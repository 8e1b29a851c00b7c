// pileup_kernel.cu -- K2 launcher and host plan: batched per-site coverage pass, one thread per read.
// Device logic and design notes: pileup_device.cuh.
#include "lcd_common.cuh"
#include "pileup_device.cuh"
#include <algorithm>

namespace lcd {
namespace pileup {

constexpr int THREADS = 128;

__global__ void __launch_bounds__(THREADS)
pileup_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        process_read(a, g);
}

__global__ void __launch_bounds__(THREADS)
profile_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        profile_read(a, g);
}

template <typename T, typename U> static void append(std::vector<T> &dst, const U *src, size_t n, long long add = 0) {
    const size_t o = dst.size(); dst.resize(o + n);
    for (size_t i = 0; i < n; ++i) dst[o + i] = (T)(src[i] + (U)add);
}

struct PileupPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks; std::vector<long long> site_off;
    long long tot_reads = 0, tot_sites = 0, tot_events = 0;
    DevBuf<Chunk> d_chunks; DevBuf<int32_t> d_read_chunk, d_ndig, d_dlen, d_dqi, d_stype, d_sref, d_salt, d_counts;
    DevBuf<uint8_t> d_active, d_rev, d_qual, d_dlow, d_dalt, d_site_alt; DevBuf<int8_t> d_dtype;
    DevBuf<long long> d_beg, d_end, d_dfirst, d_qoff, d_dpos, d_daoff, d_spos, d_saoff;
    std::vector<int32_t> h_counts;
    // read x variant profile (K3)
    bool profile = false;
    std::vector<long long> read_off, row_off, chunk_row0; std::vector<int32_t> row_cap;
    DevBuf<int32_t> d_cate, d_nnreg, d_row_cap, d_pstart, d_pend, d_altqi, d_status; DevBuf<long long> d_nfirst, d_nbeg, d_nend, d_row_off, d_aoff;
    DevBuf<int8_t> d_alleles;
    long long tot_rows = 0;
    KernelArgs base;                 // device pointers of the read / event arrays: this plan's own buffers, or a digar plan's (K1 -> K2 / K3 in place)
    std::vector<long long> h_beg, h_end;
    // the site arrays: this plan's own buffers, or a sites plan's (K1b -> K2 in place)
    const long long *p_spos = nullptr, *p_saoff = nullptr; const int32_t *p_stype = nullptr, *p_sref = nullptr, *p_salt = nullptr; const uint8_t *p_site_alt = nullptr;
    void own_sites() { p_spos = d_spos.p; p_saoff = d_saoff.p; p_stype = d_stype.p; p_sref = d_sref.p; p_salt = d_salt.p; p_site_alt = d_site_alt.p; }

    void own_pointers() {
        memset(&base, 0, sizeof(base));
        base.read_chunk = d_read_chunk.p; base.read_active = d_active.p; base.read_beg = d_beg.p; base.read_end = d_end.p; base.read_is_rev = d_rev.p;
        base.digar_first = d_dfirst.p; base.n_digar = d_ndig.p; base.qual_off = d_qoff.p; base.qual = d_qual.p; base.digar_pos = d_dpos.p; base.digar_type = d_dtype.p;
        base.digar_len = d_dlen.p; base.digar_qi = d_dqi.p; base.digar_low_qual = d_dlow.p; base.digar_alt_off = d_daoff.p; base.digar_alt = d_dalt.p;
        base.nreg_first = d_nfirst.p; base.n_nreg = d_nnreg.p; base.nreg_beg = d_nbeg.p; base.nreg_end = d_nend.p;
    }

    // K2 / K3 on the difference lists a digar plan left in HBM: only the site lists (and categories) are uploaded
    int build_on_digar(Plan *digar, int n_, const lcd_site_list_t *sl, bool want_profile) {
        n = n_; profile = want_profile;
        Context &c = ctx();
        DigarView v;
        if (digar_plan_view(digar, cur_stream(), &v)) return -1;
        if (v.n_chunks != n) { set_error("lcd_pileup: %d site lists for a digar plan of %d chunks", n, v.n_chunks); return -1; }
        if (n == 0) return 0;
        std::vector<int32_t> read_chunk; std::vector<long long> salt_n(n, 0);
        chunks.resize(n); site_off.resize(n); read_off = v.read_off;
        tot_reads = v.n_reads_total; tot_events = v.tot_events;
        long long tot_salt = 0;
        for (int i = 0; i < n; ++i) {
            const lcd_site_list_t &x = sl[i];
            if (x.n_sites < 0 || (want_profile && !x.var_cate)) { set_error("lcd_pileup: chunk %d has an invalid site list", i); return -1; }
            Chunk &k = chunks[i];
            k.n_sites = x.n_sites; k.min_bq = v.min_bq[i]; k.min_sv_len = x.min_sv_len; k.pad = 0; k.site_off = tot_sites; k.alt_base = v.alt_base[i];
            k.salt_base = tot_salt; k.pad2 = 0; site_off[i] = tot_sites;
            long long n_salt = 0;
            for (int s = 0; s < x.n_sites; ++s) if (x.site_type[s] == CDIFF || x.site_type[s] == CINS) n_salt = std::max<long long>(n_salt, x.site_alt_off[s] + x.site_alt_len[s]);
            salt_n[i] = n_salt; tot_salt += n_salt;
            if (want_profile)
                for (int s = 0; s < x.n_sites; ++s) if (x.var_cate[s] == CAND_SOMATIC_VAR) { set_error("lcd_profile: chunk %d holds candidate somatic variants (-s); only the germline path is implemented on the GPU", i); return -1; }
            for (long long g = read_off[i]; g < read_off[i + 1]; ++g) read_chunk.push_back(i);
            tot_sites += x.n_sites;
        }
        if (want_profile) {
            row_off.assign(tot_reads + 1, 0); row_cap.assign(tot_reads + 1, 0); chunk_row0.assign(n + 1, 0);
            for (int i = 0; i < n; ++i) {
                chunk_row0[i] = tot_rows;
                const long long *sp = (const long long *)sl[i].site_pos; const int32_t *st = sl[i].site_type; const long long ns = sl[i].n_sites;
                for (long long g = read_off[i]; g < read_off[i + 1]; ++g) {
                    row_off[g] = tot_rows;
                    if (!v.h_active[g]) continue;
                    const long long v0 = first_site(sp, st, 0, ns, v.h_beg[g]);
                    const long long v1 = row_end_site(sp, st, v0, ns, v.h_end[g]);
                    row_cap[g] = (int32_t)(v1 - v0); tot_rows += v1 - v0;
                }
            }
            chunk_row0[n] = tot_rows;
        }
        read_chunk.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s)) return -1;
        // the site lists go straight from the caller's arrays into their slice of the device arrays (offsets stay chunk-relative)
        if (d_spos.alloc(tot_sites + 1) || d_stype.alloc(tot_sites + 1) || d_sref.alloc(tot_sites + 1) || d_salt.alloc(tot_sites + 1) || d_saoff.alloc(tot_sites + 1) ||
            d_site_alt.alloc(tot_salt + 1) || (want_profile && d_cate.alloc(tot_sites + 1))) return -1;
        for (int i = 0; i < n; ++i) {
            const lcd_site_list_t &x = sl[i]; const long long o = site_off[i]; const size_t ns = (size_t)x.n_sites;
            if (!ns) continue;
            LCD_CUDA_OK(cudaMemcpyAsync(d_spos.p + o, x.site_pos, sizeof(long long) * ns, cudaMemcpyHostToDevice, s));
            LCD_CUDA_OK(cudaMemcpyAsync(d_stype.p + o, x.site_type, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
            LCD_CUDA_OK(cudaMemcpyAsync(d_sref.p + o, x.site_ref_len, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
            LCD_CUDA_OK(cudaMemcpyAsync(d_salt.p + o, x.site_alt_len, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
            LCD_CUDA_OK(cudaMemcpyAsync(d_saoff.p + o, x.site_alt_off, sizeof(long long) * ns, cudaMemcpyHostToDevice, s));
            if (salt_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_site_alt.p + chunks[i].salt_base, x.site_alt, (size_t)salt_n[i], cudaMemcpyHostToDevice, s));
            if (want_profile) LCD_CUDA_OK(cudaMemcpyAsync(d_cate.p + o, x.var_cate, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, s));
        }
        if (d_counts.alloc(8 * (size_t)tot_sites + 8)) return -1;
        if (want_profile) {
            if (d_row_off.upload(row_off.data(), row_off.size(), s) || d_row_cap.upload(row_cap.data(), row_cap.size(), s)) return -1;
            if (d_pstart.alloc(tot_reads + 1) || d_pend.alloc(tot_reads + 1) || d_aoff.alloc(tot_reads + 1) || d_alleles.alloc(tot_rows + 16) ||
                d_altqi.alloc(tot_rows + 16) || d_status.alloc(1)) return -1;
        }
        memset(&base, 0, sizeof(base));
        base.read_chunk = d_read_chunk.p; base.read_active = v.active; base.read_dropped = v.dropped; base.read_beg = v.beg; base.read_end = v.end; base.read_is_rev = v.rev;
        base.digar_first = v.dfirst; base.n_digar = v.ndig; base.qual_off = v.qoff; base.qual = v.qual; base.digar_pos = v.dpos; base.digar_type = v.dtype;
        base.digar_len = v.dlen; base.digar_qi = v.dqi; base.digar_low_qual = v.dlow; base.digar_alt_off = v.daoff; base.digar_alt = v.dalt;
        base.nreg_first = v.nfirst; base.n_nreg = v.nnreg; base.nreg_beg = v.nbeg; base.nreg_end = v.nend;
        own_sites();
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    // K2 on the site lists a sites plan left in HBM (created on the same digar plan): nothing is uploaded but the chunk table
    int build_on_sites(Plan *digar, Plan *sites) {
        DigarView v; SitesView sv;
        if (digar_plan_view(digar, cur_stream(), &v) || sites_plan_view(sites, digar, cur_stream(), &sv)) return -1;
        n = v.n_chunks; profile = false;
        if (sv.n_chunks != n) { set_error("lcd_pileup: a sites plan of %d chunks for a digar plan of %d chunks", sv.n_chunks, n); return -1; }
        if (n == 0) return 0;
        std::vector<int32_t> read_chunk;
        chunks.resize(n); site_off.resize(n); read_off = v.read_off;
        tot_reads = v.n_reads_total; tot_events = v.tot_events; tot_sites = sv.site_off[n];
        for (int i = 0; i < n; ++i) {
            Chunk &k = chunks[i];
            k.n_sites = (int32_t)(sv.site_off[i + 1] - sv.site_off[i]); k.min_bq = v.min_bq[i]; k.min_sv_len = sv.min_sv_len[i]; k.pad = 0; k.site_off = sv.site_off[i];
            k.alt_base = v.alt_base[i]; k.salt_base = v.alt_base[i]; k.pad2 = 0; site_off[i] = sv.site_off[i];      // a site's alt bases are those of one of its records
            for (long long g = read_off[i]; g < read_off[i + 1]; ++g) read_chunk.push_back(i);
        }
        read_chunk.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s) || d_counts.alloc(8 * (size_t)tot_sites + 8)) return -1;
        memset(&base, 0, sizeof(base));
        base.read_chunk = d_read_chunk.p; base.read_active = v.active; base.read_dropped = v.dropped; base.read_beg = v.beg; base.read_end = v.end; base.read_is_rev = v.rev;
        base.digar_first = v.dfirst; base.n_digar = v.ndig; base.qual_off = v.qoff; base.qual = v.qual; base.digar_pos = v.dpos; base.digar_type = v.dtype;
        base.digar_len = v.dlen; base.digar_qi = v.dqi; base.digar_low_qual = v.dlow; base.digar_alt_off = v.daoff; base.digar_alt = v.dalt;
        p_spos = sv.spos; p_saoff = sv.saoff; p_stype = sv.stype; p_sref = sv.sref; p_salt = sv.salt; p_site_alt = v.dalt;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int build(int n_, const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex = nullptr) {
        n = n_; profile = ex != nullptr;
        Context &c = ctx();
        if (n == 0) return 0;
        std::vector<int32_t> read_chunk, ndig, dlen, dqi, stype, sref, salt;
        std::vector<uint8_t> active, rev, qual, dlow, dalt, site_alt; std::vector<int8_t> dtype;
        std::vector<long long> beg, end, dfirst, qoff, dpos, daoff, spos, saoff, nfirst, nbeg, nend;
        std::vector<int32_t> cate, nnreg;
        chunks.resize(n); site_off.resize(n); read_off.resize(n + 1);
        for (int i = 0; i < n; ++i) {
            const lcd_pileup_input_t &x = in[i];
            if (x.n_reads < 0 || x.n_sites < 0) { set_error("lcd_pileup: chunk %d has negative sizes", i); return -1; }
            Chunk &k = chunks[i]; read_off[i] = tot_reads;
            k.n_sites = x.n_sites; k.min_bq = x.min_bq; k.min_sv_len = x.min_sv_len; k.pad = 0; k.site_off = tot_sites; k.alt_base = 0; k.salt_base = 0; k.pad2 = 0; site_off[i] = tot_sites;
            long long n_ev = 0, n_q = 0, n_alt = 0, n_salt = 0;
            for (int r = 0; r < x.n_reads; ++r) {
                if (x.n_digar[r] < 0 || x.digar_first[r] < 0) { set_error("lcd_pileup: chunk %d read %d has an invalid event range", i, r); return -1; }
                n_ev = std::max<long long>(n_ev, x.digar_first[r] + x.n_digar[r]);
            }
            for (long long d = 0; d < n_ev; ++d) {
                const int t = x.digar_type[d];
                if (t == CDIFF || t == CINS) n_alt = std::max<long long>(n_alt, x.digar_alt_off[d] + x.digar_len[d]);
            }
            for (int r = 0; r < x.n_reads; ++r) {                   // qualities: up to the last read base an event of the read points at
                long long hi = 0;
                for (long long d = x.digar_first[r]; d < x.digar_first[r] + x.n_digar[r]; ++d) {
                    const int t = x.digar_type[d];
                    const long long last = t == CDEL ? x.digar_qi[d] : (long long)x.digar_qi[d] + x.digar_len[d] - 1;
                    if (t == CDIFF || t == CINS || t == CDEL) hi = std::max(hi, last + 1);
                }
                n_q = std::max(n_q, x.qual_off[r] + hi);
            }
            for (int s = 0; s < x.n_sites; ++s) if (x.site_type[s] == CDIFF || x.site_type[s] == CINS) n_salt = std::max<long long>(n_salt, x.site_alt_off[s] + x.site_alt_len[s]);
            std::vector<uint8_t> listed(x.n_reads, 0);
            for (int r = 0; r < x.n_reads; ++r) { const int id = x.ordered_read_ids[r]; if (id >= 0 && id < x.n_reads) listed[id] = 1; }
            const long long ev0 = (long long)dpos.size(), q0 = (long long)qual.size(), a0 = (long long)dalt.size(), sa0 = (long long)site_alt.size();
            for (int r = 0; r < x.n_reads; ++r) { read_chunk.push_back(i); active.push_back(listed[r] && !x.is_skipped[r]); }
            append(beg, x.read_beg, x.n_reads); append(end, x.read_end, x.n_reads); append(rev, x.read_is_rev, x.n_reads);
            append(dfirst, x.digar_first, x.n_reads, ev0); append(ndig, x.n_digar, x.n_reads); append(qoff, x.qual_off, x.n_reads, q0);
            append(qual, x.qual, (size_t)n_q);
            append(dpos, x.digar_pos, (size_t)n_ev); append(dtype, x.digar_type, (size_t)n_ev); append(dlen, x.digar_len, (size_t)n_ev);
            append(dqi, x.digar_qi, (size_t)n_ev); append(dlow, x.digar_low_qual, (size_t)n_ev); append(daoff, x.digar_alt_off, (size_t)n_ev, a0);
            append(dalt, x.digar_alt, (size_t)n_alt);
            append(spos, x.site_pos, x.n_sites); append(stype, x.site_type, x.n_sites); append(sref, x.site_ref_len, x.n_sites);
            append(salt, x.site_alt_len, x.n_sites); append(saoff, x.site_alt_off, x.n_sites, sa0); append(site_alt, x.site_alt, (size_t)n_salt);
            if (profile) {
                long long n_iv = 0;
                for (int r = 0; r < x.n_reads; ++r) n_iv = std::max<long long>(n_iv, ex[i].nreg_first[r] + ex[i].n_nreg[r]);
                for (int v = 0; v < x.n_sites; ++v) if (ex[i].var_cate[v] == CAND_SOMATIC_VAR) { set_error("lcd_profile: chunk %d holds candidate somatic variants (-s); only the germline path is implemented on the GPU", i); return -1; }
                const long long iv0 = (long long)nbeg.size();
                append(cate, ex[i].var_cate, x.n_sites); append(nfirst, ex[i].nreg_first, x.n_reads, iv0); append(nnreg, ex[i].n_nreg, x.n_reads);
                append(nbeg, ex[i].nreg_beg, (size_t)n_iv); append(nend, ex[i].nreg_end, (size_t)n_iv);
            }
            tot_reads += x.n_reads; tot_sites += x.n_sites; tot_events += n_ev;
        }
        read_off[n] = tot_reads;
        if (profile) {       // profile rows: per read the candidate sites its merge-join can visit
            row_off.assign(tot_reads + 1, 0); row_cap.assign(tot_reads + 1, 0); chunk_row0.assign(n + 1, 0);
            for (int i = 0; i < n; ++i) {
                chunk_row0[i] = tot_rows;
                const long long s0 = site_off[i], s1 = s0 + chunks[i].n_sites;
                for (long long g = read_off[i]; g < read_off[i + 1]; ++g) {
                    const long long v0 = first_site(spos.data(), stype.data(), s0, s1, beg[g]);
                    const long long v1 = row_end_site(spos.data(), stype.data(), v0, s1, end[g]);
                    row_off[g] = tot_rows; row_cap[g] = (int32_t)(v1 - v0); tot_rows += v1 - v0;
                }
            }
            chunk_row0[n] = tot_rows;
        }
        auto pad = [](auto &v) { v.push_back(0); };
        pad(read_chunk); pad(active); pad(beg); pad(end); pad(rev); pad(dfirst); pad(ndig); pad(qoff); pad(qual); pad(dpos); pad(dtype); pad(dlen); pad(dqi);
        pad(dlow); pad(daoff); pad(dalt); pad(spos); pad(stype); pad(sref); pad(salt); pad(saoff); pad(site_alt);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s) || d_active.upload(active.data(), active.size(), s) ||
            d_beg.upload(beg.data(), beg.size(), s) || d_end.upload(end.data(), end.size(), s) || d_rev.upload(rev.data(), rev.size(), s) ||
            d_dfirst.upload(dfirst.data(), dfirst.size(), s) || d_ndig.upload(ndig.data(), ndig.size(), s) || d_qoff.upload(qoff.data(), qoff.size(), s) ||
            d_qual.upload(qual.data(), qual.size(), s) || d_dpos.upload(dpos.data(), dpos.size(), s) || d_dtype.upload(dtype.data(), dtype.size(), s) ||
            d_dlen.upload(dlen.data(), dlen.size(), s) || d_dqi.upload(dqi.data(), dqi.size(), s) || d_dlow.upload(dlow.data(), dlow.size(), s) ||
            d_daoff.upload(daoff.data(), daoff.size(), s) || d_dalt.upload(dalt.data(), dalt.size(), s) || d_spos.upload(spos.data(), spos.size(), s) ||
            d_stype.upload(stype.data(), stype.size(), s) || d_sref.upload(sref.data(), sref.size(), s) || d_salt.upload(salt.data(), salt.size(), s) ||
            d_saoff.upload(saoff.data(), saoff.size(), s) || d_site_alt.upload(site_alt.data(), site_alt.size(), s)) return -1;
        if (d_counts.alloc(8 * (size_t)tot_sites + 8)) return -1;
        if (profile) {
            pad(cate); pad(nfirst); pad(nnreg); pad(nbeg); pad(nend);
            if (d_cate.upload(cate.data(), cate.size(), s) || d_nfirst.upload(nfirst.data(), nfirst.size(), s) || d_nnreg.upload(nnreg.data(), nnreg.size(), s) ||
                d_nbeg.upload(nbeg.data(), nbeg.size(), s) || d_nend.upload(nend.data(), nend.size(), s) ||
                d_row_off.upload(row_off.data(), row_off.size(), s) || d_row_cap.upload(row_cap.data(), row_cap.size(), s)) return -1;
            if (d_pstart.alloc(tot_reads + 1) || d_pend.alloc(tot_reads + 1) || d_aoff.alloc(tot_reads + 1) || d_alleles.alloc(tot_rows + 16) ||
                d_altqi.alloc(tot_rows + 16) || d_status.alloc(1)) return -1;
        }
        own_pointers(); own_sites();
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0 || tot_reads == 0) return 0;
        LCD_CUDA_OK(cudaMemsetAsync(d_counts.p, 0, sizeof(int32_t) * 8 * (size_t)tot_sites, s));
        KernelArgs a = base;
        a.chunks = d_chunks.p; a.n_reads_total = tot_reads;
        a.site_pos = p_spos; a.site_type = p_stype; a.site_ref_len = p_sref; a.site_alt_len = p_salt; a.site_alt_off = p_saoff; a.site_alt = p_site_alt;
        a.site_counts = d_counts.p;
        const int grid = (int)std::min<long long>((tot_reads + THREADS - 1) / THREADS, (long long)c.sm_count * 16);
        if (profile) {
            a.var_cate = d_cate.p;
            a.row_off = d_row_off.p; a.row_cap = d_row_cap.p; a.prof_start = d_pstart.p; a.prof_end = d_pend.p; a.allele_off = d_aoff.p;
            a.alleles = d_alleles.p; a.alt_qi = d_altqi.p; a.status = d_status.p;
            LCD_CUDA_OK(cudaMemsetAsync(d_status.p, 0, sizeof(int32_t), s));
            profile_kernel<<<grid, THREADS, 0, s>>>(a);
            LCD_CUDA_OK(cudaGetLastError());
            c.launches++;
            return 0;
        }
        pileup_kernel<<<grid, THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }

    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_events; return 0; }   // difference-list entries walked

    int fetch(cudaStream_t s, lcd_pileup_output_t *out) {
        if (n == 0) return 0;
        LCD_DRAIN(s);
        for (int i = 0; i < n; ++i)       // straight into the caller's arrays
            if (chunks[i].n_sites) LCD_CUDA_OK(cudaMemcpyAsync(out[i].site_counts, d_counts.p + 8 * site_off[i], sizeof(int32_t) * 8 * (size_t)chunks[i].n_sites, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    long long capacity(int i) const { return chunk_row0[i + 1] - chunk_row0[i]; }

    int fetch_profile(cudaStream_t s, lcd_profile_output_t *out) {
        if (n == 0) return 0;
        LCD_DRAIN(s);
        std::vector<int32_t> ps(tot_reads + 1), pe(tot_reads + 1), qi(tot_rows + 16); std::vector<long long> ao(tot_reads + 1); std::vector<int8_t> al(tot_rows + 16);
        int32_t status = 0;
        LCD_CUDA_OK(cudaMemcpyAsync(ps.data(), d_pstart.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(pe.data(), d_pend.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(ao.data(), d_aoff.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
        if (tot_rows) {
            LCD_CUDA_OK(cudaMemcpyAsync(al.data(), d_alleles.p, (size_t)tot_rows, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(qi.data(), d_altqi.p, sizeof(int32_t) * (size_t)tot_rows, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaMemcpyAsync(&status, d_status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        if (status) { set_error("lcd_profile: a profile row overflowed its capacity on the device (status %d)", status); return -2; }
        for (int i = 0; i < n; ++i) {
            const long long r0 = read_off[i], nr = read_off[i + 1] - r0, row0 = chunk_row0[i], rows = capacity(i);
            if (rows > out[i].alleles_cap) { set_error("lcd_profile: chunk %d needs %lld profile entries, the caller provided %lld (lcd_profile_capacity)", i, rows, (long long)out[i].alleles_cap); return -3; }
            for (long long r = 0; r < nr; ++r) { out[i].prof_start[r] = ps[r0 + r]; out[i].prof_end[r] = pe[r0 + r]; out[i].allele_off[r] = ao[r0 + r] - row0; }
            if (rows) { memcpy(out[i].alleles, al.data() + row0, (size_t)rows); memcpy(out[i].alt_qi, qi.data() + row0, sizeof(int32_t) * (size_t)rows); }
            out[i].n_alleles = rows;
        }
        return 0;
    }
};

} // namespace pileup

int pileup_plan_view(Plan *plan, PileupView *v) {
    pileup::PileupPlan *p = dynamic_cast<pileup::PileupPlan *>(plan);
    if (!p || p->profile) { set_error("not a pileup (K2) plan"); return -1; }
    v->n_chunks = p->n; v->site_off.assign(p->n + 1, 0); v->salt_base.assign(p->n, 0);
    for (int i = 0; i < p->n; ++i) { v->site_off[i] = p->site_off[i]; v->salt_base[i] = p->chunks[i].salt_base; }
    v->site_off[p->n] = p->tot_sites;
    v->spos = p->p_spos; v->saoff = p->p_saoff; v->stype = p->p_stype; v->sref = p->p_sref; v->salt = p->p_salt; v->site_alt = p->p_site_alt; v->counts = p->d_counts.p;
    return 0;
}
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_pileup_plan_create(int n_chunks, const lcd_pileup_input_t *in) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && !in)) { set_error("lcd_pileup_plan_create: invalid arguments"); return nullptr; }
    pileup::PileupPlan *p = new pileup::PileupPlan();
    if (p->build(n_chunks, in)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_pileup_plan_fetch(lcd_plan_t *plan, void *stream, lcd_pileup_output_t *out) {
    pileup::PileupPlan *p = dynamic_cast<pileup::PileupPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_pileup_plan_fetch: not a pileup plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}

lcd_plan_t *lcd_profile_plan_create(int n_chunks, const lcd_pileup_input_t *in, const lcd_profile_extra_t *extra) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && (!in || !extra))) { set_error("lcd_profile_plan_create: invalid arguments"); return nullptr; }
    pileup::PileupPlan *p = new pileup::PileupPlan();
    if (p->build(n_chunks, in, extra)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_profile_plan_fetch(lcd_plan_t *plan, void *stream, lcd_profile_output_t *out) {
    pileup::PileupPlan *p = dynamic_cast<pileup::PileupPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !p->profile || !out) { set_error("lcd_profile_plan_fetch: not a profile plan / null outputs"); return -1; }
    return p->fetch_profile(pick_stream(stream), out);
}

lcd_plan_t *lcd_pileup_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_site_list_t *sites) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || n_chunks < 0 || (n_chunks > 0 && !sites)) { set_error("lcd_pileup_plan_create_on_digar: invalid arguments"); return nullptr; }
    pileup::PileupPlan *p = new pileup::PileupPlan();
    if (p->build_on_digar(reinterpret_cast<Plan *>(digar_plan), n_chunks, sites, false)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

lcd_plan_t *lcd_pileup_plan_create_on_sites(lcd_plan_t *digar_plan, lcd_plan_t *sites_plan) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || !sites_plan) { set_error("lcd_pileup_plan_create_on_sites: invalid arguments"); return nullptr; }
    pileup::PileupPlan *p = new pileup::PileupPlan();
    if (p->build_on_sites(reinterpret_cast<Plan *>(digar_plan), reinterpret_cast<Plan *>(sites_plan))) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

lcd_plan_t *lcd_profile_plan_create_on_digar(lcd_plan_t *digar_plan, int n_chunks, const lcd_site_list_t *sites) {
    if (ensure_ready()) return nullptr;
    if (!digar_plan || n_chunks < 0 || (n_chunks > 0 && !sites)) { set_error("lcd_profile_plan_create_on_digar: invalid arguments"); return nullptr; }
    pileup::PileupPlan *p = new pileup::PileupPlan();
    if (p->build_on_digar(reinterpret_cast<Plan *>(digar_plan), n_chunks, sites, true)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int64_t lcd_profile_plan_capacity(lcd_plan_t *plan, int chunk) {
    pileup::PileupPlan *p = dynamic_cast<pileup::PileupPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !p->profile || chunk < 0 || chunk >= p->n) { set_error("lcd_profile_plan_capacity: not a profile plan / chunk out of range"); return -1; }
    return p->capacity(chunk);
}

int lcd_profile_batch(int n_chunks, const lcd_pileup_input_t *in, const lcd_profile_extra_t *extra, lcd_profile_output_t *out) {
    lcd_plan_t *plan = lcd_profile_plan_create(n_chunks, in, extra);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_profile_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

int64_t lcd_profile_capacity(const lcd_pileup_input_t *in) {
    if (!in) return -1;
    int64_t tot = 0;
    for (int r = 0; r < in->n_reads; ++r) {
        const long long v0 = pileup::first_site((const long long *)in->site_pos, in->site_type, 0, in->n_sites, in->read_beg[r]);
        tot += pileup::row_end_site((const long long *)in->site_pos, in->site_type, v0, in->n_sites, in->read_end[r]) - v0;
    }
    return tot;
}

int lcd_pileup_batch(int n_chunks, const lcd_pileup_input_t *in, lcd_pileup_output_t *out) {
    lcd_plan_t *plan = lcd_pileup_plan_create(n_chunks, in);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_pileup_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
