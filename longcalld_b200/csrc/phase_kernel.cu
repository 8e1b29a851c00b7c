// phase_kernel.cu -- K4 launcher and host plan: batched read -> haplotype assignment / phasing, one CTA per chunk.
// Device logic and design notes: phase_device.cuh.
#include "lcd_common.cuh"
#include "phase_device.cuh"
#include <algorithm>

namespace lcd {
namespace phase {

constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS)
phase_kernel(const KernelArgs a) {
    __shared__ int sh[8];
    for (int ci = blockIdx.x; ci < a.n_chunks; ci += gridDim.x) {
        Phaser p;
        p.run(a, a.chunks[ci], sh);
        __syncthreads();
    }
}

// ---- host glue: the order in which cgranges hands the reads of a chunk back (reference src/cgranges.c:13-64 cr_index ->
// radix_sort_cr_intv, an in-place MSD radix sort by interval start that is not stable, then an in-order walk :449-490).
// The seed pass visits reads in exactly this order, so the marshalling code reproduces the sort.
struct Intv { uint64_t key; int32_t label; };
static void insertion_sort(Intv *beg, Intv *end) {
    for (Intv *i = beg + 1; i < end; ++i)
        if (i->key < (i - 1)->key) {
            Intv tmp = *i, *j;
            for (j = i; j > beg && tmp.key < (j - 1)->key; --j) *j = *(j - 1);
            *j = tmp;
        }
}
static void flag_sort(Intv *beg, Intv *end, int shift) {
    Intv *head[256], *tail[256];
    size_t cnt[256] = {0};
    for (Intv *i = beg; i != end; ++i) cnt[(i->key >> shift) & 255]++;
    Intv *p = beg;
    for (int k = 0; k < 256; ++k) { head[k] = p; p += cnt[k]; tail[k] = p; }
    for (int k = 0; k < 256;) {
        if (head[k] == tail[k]) { ++k; continue; }
        int l = (int)((head[k]->key >> shift) & 255);
        if (l == k) { ++head[k]; continue; }
        Intv carry = *head[k];
        do { std::swap(carry, *head[l]); ++head[l]; l = (int)((carry.key >> shift) & 255); } while (l != k);
        *head[k]++ = carry;
    }
    if (shift == 0) return;
    const int next = shift > 8 ? shift - 8 : 0;
    Intv *b0 = beg;
    for (int k = 0; k < 256; ++k) {
        Intv *e0 = tail[k];
        if (e0 - b0 > 64) flag_sort(b0, e0, next); else if (e0 - b0 > 1) insertion_sort(b0, e0);
        b0 = e0;
    }
}

struct PhasePlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks;
    int64_t tot_reads = 0, tot_vars = 0, tot_alleles = 0;
    DevBuf<Chunk> d_chunks;
    DevBuf<int32_t> d_ordered, d_pstart, d_pend, d_cr_order, d_cr_pmax, d_haps, d_agree, d_conflict;
    DevBuf<uint8_t> d_skipped; DevBuf<int64_t> d_aoff; DevBuf<int8_t> d_alleles; DevBuf<long long> d_psets, d_pos, d_var_ps;
    DevBuf<int32_t> d_cate, d_type, d_hp, d_nuniq, d_covs, d_tcov, d_cons, d_prof, d_valid, d_flags, d_nag, d_ncf, d_snap;
    std::vector<int32_t> h_haps, h_agree, h_conflict, h_cons, h_prof; std::vector<long long> h_psets, h_var_ps;

    int build(int n_, const lcd_phase_input_t *in, const lcd_phase_output_t *out) {
        n = n_;
        Context &c = ctx();
        if (n == 0) return 0;
        chunks.resize(n);
        for (int i = 0; i < n; ++i) {
            if (in[i].n_reads < 0 || in[i].n_vars < 0) { set_error("lcd_phase: chunk %d has negative sizes", i); return -1; }
            Chunk &k = chunks[i];
            memset(&k, 0, sizeof(k));
            k.n_reads = in[i].n_reads; k.n_vars = in[i].n_vars; k.target = in[i].target_var_cate; k.is_ont = in[i].is_ont;
            k.read_off = tot_reads; k.var_off = tot_vars;
            tot_reads += k.n_reads; tot_vars += k.n_vars;
        }
        std::vector<int32_t> ordered(tot_reads + 1), pstart(tot_reads + 1), pend(tot_reads + 1), cr_order(tot_reads + 1), cr_pmax(tot_reads + 1);
        std::vector<uint8_t> skipped(tot_reads + 1); std::vector<int64_t> aoff(tot_reads + 1);
        std::vector<int8_t> alleles;
        std::vector<int32_t> cate(tot_vars + 1), type(tot_vars + 1), hp(tot_vars + 1), nuniq(tot_vars + 1), covs(4 * tot_vars + 4), tcov(tot_vars + 1);
        std::vector<long long> pos(tot_vars + 1);
        h_haps.resize(tot_reads + 1); h_agree.resize(tot_reads + 1); h_conflict.resize(tot_reads + 1); h_psets.resize(tot_reads + 1);
        h_cons.resize(3 * tot_vars + 3); h_prof.resize(12 * tot_vars + 12); h_var_ps.resize(tot_vars + 1);
        std::vector<Intv> iv;
        for (int i = 0; i < n; ++i) {
            Chunk &k = chunks[i]; const lcd_phase_input_t &x = in[i];
            const int64_t ro = k.read_off, vo = k.var_off;
            iv.clear();
            for (int r = 0; r < k.n_reads; ++r) {
                const int rid = x.ordered_read_ids[r];
                if (rid < 0 || rid >= k.n_reads) { set_error("lcd_phase: chunk %d: ordered_read_ids[%d] out of range", i, r); return -1; }
                ordered[ro + r] = rid; skipped[ro + r] = x.is_skipped[r]; pstart[ro + r] = x.prof_start[r]; pend[ro + r] = x.prof_end[r];
                const int len = x.prof_end[r] - x.prof_start[r] + 1;
                if (len > 0 && (x.prof_start[r] < 0 || x.prof_end[r] >= k.n_vars)) { set_error("lcd_phase: chunk %d: read %d spans variants outside the chunk", i, r); return -1; }
                aoff[ro + r] = (int64_t)alleles.size();
                if (len > 0) alleles.insert(alleles.end(), x.alleles + x.allele_off[r], x.alleles + x.allele_off[r] + len);
                h_haps[ro + r] = out[i].haps[r]; h_psets[ro + r] = out[i].phase_sets[r];
                h_agree[ro + r] = out[i].n_clean_agree_snps[r]; h_conflict[ro + r] = out[i].n_clean_conflict_snps[r];
            }
            // read_var_cr as collect_read_var_profile fills it (reference src/collect_var.c:1407-1431), then cr_index
            for (int r = 0; r < k.n_reads; ++r) {
                const int rid = x.ordered_read_ids[r];
                if (x.is_skipped[rid] || x.prof_start[rid] < 0 || x.prof_end[rid] < 0) continue;
                iv.push_back(Intv{(uint64_t)(uint32_t)x.prof_start[rid], rid});
            }
            if (iv.size() <= 64) insertion_sort(iv.data(), iv.data() + iv.size()); else flag_sort(iv.data(), iv.data() + iv.size(), 56);
            k.n_cr = (int32_t)iv.size();
            int32_t run_max = INT32_MIN;
            for (size_t t = 0; t < iv.size(); ++t) {
                cr_order[ro + t] = iv[t].label;
                run_max = std::max(run_max, x.prof_end[iv[t].label]);
                cr_pmax[ro + t] = run_max;
            }
            for (int v = 0; v < k.n_vars; ++v) {
                cate[vo + v] = x.var_cate[v]; type[vo + v] = x.var_type[v]; hp[vo + v] = x.is_hp_indel[v]; nuniq[vo + v] = std::min(4, x.n_uniq_alles[v]);
                for (int a = 0; a < 4; ++a) covs[4 * (vo + v) + a] = x.alle_covs[4 * v + a];
                tcov[vo + v] = x.total_cov[v]; pos[vo + v] = x.pos[v];
                for (int h = 0; h < 3; ++h) h_cons[3 * (vo + v) + h] = out[i].hap_to_cons_alle[3 * v + h];
                for (int q = 0; q < 12; ++q) h_prof[12 * (vo + v) + q] = out[i].hap_to_alle_profile[12 * v + q];
                h_var_ps[vo + v] = out[i].var_phase_set[v];
            }
        }
        tot_alleles = (int64_t)alleles.size();
        alleles.push_back(0);
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), std::max(n, 1), s)) return -1;
        if (d_ordered.upload(ordered.data(), ordered.size(), s) || d_skipped.upload(skipped.data(), skipped.size(), s) ||
            d_pstart.upload(pstart.data(), pstart.size(), s) || d_pend.upload(pend.data(), pend.size(), s) ||
            d_aoff.upload(aoff.data(), aoff.size(), s) || d_alleles.upload(alleles.data(), alleles.size(), s) ||
            d_cr_order.upload(cr_order.data(), cr_order.size(), s) || d_cr_pmax.upload(cr_pmax.data(), cr_pmax.size(), s) ||
            d_cate.upload(cate.data(), cate.size(), s) || d_type.upload(type.data(), type.size(), s) || d_hp.upload(hp.data(), hp.size(), s) ||
            d_nuniq.upload(nuniq.data(), nuniq.size(), s) || d_covs.upload(covs.data(), covs.size(), s) || d_tcov.upload(tcov.data(), tcov.size(), s) ||
            d_pos.upload(pos.data(), pos.size(), s)) return -1;
        // in/out state: what the call does not touch keeps the caller's values
        if (d_haps.upload(h_haps.data(), h_haps.size(), s) || d_psets.upload(h_psets.data(), h_psets.size(), s) ||
            d_agree.upload(h_agree.data(), h_agree.size(), s) || d_conflict.upload(h_conflict.data(), h_conflict.size(), s) ||
            d_cons.upload(h_cons.data(), h_cons.size(), s) || d_prof.upload(h_prof.data(), h_prof.size(), s) || d_var_ps.upload(h_var_ps.data(), h_var_ps.size(), s)) return -1;
        if (d_valid.alloc(tot_vars + 1) || d_flags.alloc(tot_vars + 1) || d_nag.alloc(tot_vars + 1) || d_ncf.alloc(tot_vars + 1) || d_snap.alloc(2 * tot_vars + 2)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        KernelArgs a;
        a.chunks = d_chunks.p; a.n_chunks = n;
        a.ordered_ids = d_ordered.p; a.is_skipped = d_skipped.p; a.pstart = d_pstart.p; a.pend = d_pend.p; a.allele_off = d_aoff.p; a.alleles = d_alleles.p;
        a.cr_order = d_cr_order.p; a.cr_pmax_end = d_cr_pmax.p;
        a.haps = d_haps.p; a.phase_sets = d_psets.p; a.agree = d_agree.p; a.conflict = d_conflict.p;
        a.cate = d_cate.p; a.type = d_type.p; a.hp = d_hp.p; a.nuniq = d_nuniq.p; a.alle_covs = d_covs.p; a.total_cov = d_tcov.p; a.pos = d_pos.p;
        a.cons = d_cons.p; a.prof = d_prof.p; a.var_ps = d_var_ps.p;
        a.valid = d_valid.p; a.flags = d_flags.p; a.n_agree = d_nag.p; a.n_conf = d_ncf.p; a.snap = d_snap.p;
        phase_kernel<<<std::min(n, c.sm_count * 4), THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        return 0;
    }

    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_alleles; return 0; }   // (read, variant) pairs

    int fetch(cudaStream_t s, lcd_phase_output_t *out) {
        LCD_DRAIN(s);
        if (n == 0) return 0;
        LCD_CUDA_OK(cudaMemcpyAsync(h_haps.data(), d_haps.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_psets.data(), d_psets.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_agree.data(), d_agree.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_conflict.data(), d_conflict.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_cons.data(), d_cons.p, sizeof(int32_t) * 3 * tot_vars, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_prof.data(), d_prof.p, sizeof(int32_t) * 12 * tot_vars, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(h_var_ps.data(), d_var_ps.p, sizeof(long long) * tot_vars, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) {
            const Chunk &k = chunks[i];
            for (int r = 0; r < k.n_reads; ++r) {
                out[i].haps[r] = h_haps[k.read_off + r]; out[i].phase_sets[r] = h_psets[k.read_off + r];
                out[i].n_clean_agree_snps[r] = h_agree[k.read_off + r]; out[i].n_clean_conflict_snps[r] = h_conflict[k.read_off + r];
            }
            for (int v = 0; v < k.n_vars; ++v) {
                for (int h = 0; h < 3; ++h) out[i].hap_to_cons_alle[3 * v + h] = h_cons[3 * (k.var_off + v) + h];
                for (int q = 0; q < 12; ++q) out[i].hap_to_alle_profile[12 * v + q] = h_prof[12 * (k.var_off + v) + q];
                out[i].var_phase_set[v] = h_var_ps[k.var_off + v];
            }
        }
        return 0;
    }
};

} // namespace phase
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_phase_plan_create(int n_chunks, const lcd_phase_input_t *in, const lcd_phase_output_t *state) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && (!in || !state))) { set_error("lcd_phase_plan_create: invalid arguments"); return nullptr; }
    phase::PhasePlan *p = new phase::PhasePlan();
    if (p->build(n_chunks, in, state)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_phase_plan_fetch(lcd_plan_t *plan, void *stream, lcd_phase_output_t *out) {
    phase::PhasePlan *p = dynamic_cast<phase::PhasePlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_phase_plan_fetch: not a phasing plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}

int lcd_phase_batch(int n_chunks, const lcd_phase_input_t *in, lcd_phase_output_t *out) {
    lcd_plan_t *plan = lcd_phase_plan_create(n_chunks, in, out);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_phase_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
