// edlib_kernel.cu -- K7 launcher and host plan: batched edlib NW / HW alignment with path.
// Device logic and design notes: edlib_device.cuh.
#include "lcd_common.cuh"
#include "edlib_device.cuh"
#include <algorithm>

namespace lcd {
namespace edlib {

constexpr int THREADS_PER_CTA = 32;     // one problem per thread; small CTAs spread a batch over all SMs

__global__ void __launch_bounds__(THREADS_PER_CTA)
edlib_kernel(const KernelArgs a) {
    Aligner al;
    for (;;) {
        const uint32_t item = atomicAdd(a.queue, 1u);
        if (item >= (uint32_t)a.n) break;
        const int pi = a.order[item];
        const Problem pb = a.problems[pi];
        al.align(pb, a.seqs, a.aln + pb.aln_off, a.pool + pb.ws_off, a.results + pi);
    }
}

struct EdlibPlan : Plan {
    int pool_window() const override { return 1; }
    DevBuf<uint8_t> d_seqs, d_aln;
    DevBuf<Problem> d_problems;
    DevBuf<int32_t> d_order;
    DevBuf<DevResult> d_results;
    DevBuf<uint32_t> d_queue;
    std::vector<Problem> problems;
    std::vector<int32_t> order;
    std::vector<std::pair<int, int>> launches;     // [begin, end) ranges of `order` whose workspaces fit the pool together
    std::vector<int64_t> aln_dev_off;
    size_t aln_bytes = 0;
    std::vector<DevResult> h_results;
    std::vector<uint8_t> h_aln;

    int build(int n_, const uint8_t *seqs, size_t seqs_len, const int64_t *q_off, const int32_t *qlen,
              const int64_t *t_off, const int32_t *tlen, const int32_t *mode, const int32_t *want_path) {
        n = n_;
        Context &c = ctx();
        problems.resize(n); aln_dev_off.resize(n); order.resize(n);
        size_t aln_total = 0;
        for (int i = 0; i < n; ++i) {
            if (qlen[i] < 0 || tlen[i] < 0 || q_off[i] < 0 || t_off[i] < 0 ||
                (size_t)q_off[i] + qlen[i] > seqs_len || (size_t)t_off[i] + tlen[i] > seqs_len) {
                set_error("lcd_edlib: problem %d has an invalid sequence range", i); return -1;
            }
            if (mode[i] != LCD_EDLIB_MODE_NW && mode[i] != LCD_EDLIB_MODE_HW && mode[i] != LCD_EDLIB_MODE_SHW) {
                set_error("lcd_edlib: problem %d has an unknown mode %d", i, mode[i]); return -1;
            }
            Problem &p = problems[i];
            memset(&p, 0, sizeof(p));
            p.q_off = (uint64_t)q_off[i]; p.t_off = (uint64_t)t_off[i]; p.qlen = qlen[i]; p.tlen = tlen[i];
            p.mode = mode[i]; p.want_path = want_path[i] ? 1 : 0;
            p.aln_off = aln_total; aln_dev_off[i] = (int64_t)aln_total;
            if (p.want_path) aln_total += ((size_t)qlen[i] + tlen[i] + 2 + 15) & ~(size_t)15;
            p.ws_words = workspace_words(p.qlen, p.tlen, p.want_path);
            order[i] = i;
        }
        aln_bytes = aln_total;
        // processing order: similar sizes side by side (the 32 threads of a warp run in lock-step), largest first
        std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
            const uint64_t wx = (uint64_t)(problems[x].qlen / 64 + 1) * (problems[x].tlen + 1), wy = (uint64_t)(problems[y].qlen / 64 + 1) * (problems[y].tlen + 1);
            return wx != wy ? wx > wy : x < y; });
        // workspace slices; a batch that does not fit the pool at once is split into several launches
        const uint64_t pool_w = c.class_words(1) / 2;           // pool is counted in int32 words
        uint64_t top = 0; int begin = 0;
        for (int k = 0; k < n; ++k) {
            Problem &p = problems[order[k]];
            if (p.ws_words > pool_w) { set_error("lcd_edlib: a problem needs %zu MiB of workspace but the pool has %zu MiB", (size_t)(p.ws_words * 8 >> 20), (size_t)(pool_w * 8 >> 20)); return -1; }
            if (top + p.ws_words > pool_w) { launches.push_back({begin, k}); begin = k; top = 0; }
            p.ws_off = top; top += p.ws_words;
        }
        if (n > begin) launches.push_back({begin, n});
        cudaStream_t s = cur_stream();
        if (d_seqs.upload(seqs, std::max<size_t>(seqs_len, 1), s)) return -1;
        if (d_problems.upload(problems.data(), n, s)) return -1;
        if (d_order.upload(order.data(), n, s)) return -1;
        if (d_aln.alloc(aln_bytes + 16)) return -1;
        if (d_results.alloc(std::max(n, 1))) return -1;
        if (d_queue.alloc(std::max<size_t>(launches.size(), 1))) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        if (n == 0) return 0;
        LCD_CUDA_OK(cudaMemsetAsync(d_queue.p, 0, sizeof(uint32_t) * launches.size(), s));
        for (size_t l = 0; l < launches.size(); ++l) {
            const int cnt = launches[l].second - launches[l].first;
            KernelArgs ka;
            ka.problems = d_problems.p; ka.order = d_order.p + launches[l].first; ka.n = cnt; ka.queue = d_queue.p + l;
            ka.seqs = d_seqs.p; ka.aln = d_aln.p; ka.results = d_results.p; ka.pool = reinterpret_cast<Word *>(c.pool + win->off);
            const int grid = std::min((cnt + THREADS_PER_CTA - 1) / THREADS_PER_CTA, c.sm_count * 16);
            edlib_kernel<<<grid, THREADS_PER_CTA, 0, s>>>(ka);
            LCD_CUDA_OK(cudaGetLastError());
            c.launches++;
        }
        return 0;
    }

    int download(cudaStream_t s, bool want_aln) {
        LCD_DRAIN(s);
        h_results.resize(n);
        if (n == 0) return 0;
        LCD_CUDA_OK(cudaMemcpyAsync(h_results.data(), d_results.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, s));
        if (want_aln) { h_aln.resize(aln_bytes + 16); if (aln_bytes) LCD_CUDA_OK(cudaMemcpyAsync(h_aln.data(), d_aln.p, aln_bytes, cudaMemcpyDeviceToHost, s)); }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }

    int work_units(cudaStream_t s, uint64_t *units) override {
        if (download(s, false)) return -1;
        uint64_t t = 0;
        for (int i = 0; i < n; ++i) t += ((uint64_t)h_results[i].units_hi << 32) | h_results[i].units_lo;
        *units = t;
        return 0;
    }

    int fetch(cudaStream_t s, uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results) {
        if (download(s, aln != nullptr)) return -1;
        int bad = 0, first_bad = 0;
        for (int i = 0; i < n; ++i) {
            const DevResult &r = h_results[i];
            results[i].status = r.status; results[i].edit_distance = r.edit_distance; results[i].start_loc = r.start_loc;
            results[i].end_loc = r.end_loc; results[i].aln_len = r.aln_len;
            if (r.status != ST_OK) { if (!bad) first_bad = r.status; ++bad; continue; }
            if (aln && aln_off && problems[i].want_path) memcpy(aln + aln_off[i], h_aln.data() + aln_dev_off[i], r.aln_len);
        }
        if (bad) { set_error("lcd_edlib: %d of %d alignments failed on the device (first status %d; see LCD_EDLIB_STATUS_* in lcd_gpu.h)", bad, n, first_bad); return -2; }
        return 0;
    }
};

} // namespace edlib
} // namespace lcd

using namespace lcd;

extern "C" {

lcd_plan_t *lcd_edlib_plan_create(int n, const uint8_t *seqs, size_t seqs_len,
                                  const int64_t *query_off, const int32_t *qlen,
                                  const int64_t *target_off, const int32_t *tlen,
                                  const int32_t *mode, const int32_t *want_path) {
    if (ensure_ready()) return nullptr;
    if (n < 0 || (n > 0 && (!seqs || !query_off || !qlen || !target_off || !tlen || !mode || !want_path))) {
        set_error("lcd_edlib_plan_create: invalid arguments"); return nullptr;
    }
    edlib::EdlibPlan *p = new edlib::EdlibPlan();
    if (p->build(n, seqs, seqs_len, query_off, qlen, target_off, tlen, mode, want_path)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_edlib_plan_fetch(lcd_plan_t *plan, void *stream, uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results) {
    edlib::EdlibPlan *p = dynamic_cast<edlib::EdlibPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !results) { set_error("lcd_edlib_plan_fetch: not an edlib plan / null results"); return -1; }
    return p->fetch(pick_stream(stream), aln, aln_off, results);
}

int lcd_edlib_batch(int n, const uint8_t *seqs, size_t seqs_len,
                    const int64_t *query_off, const int32_t *qlen,
                    const int64_t *target_off, const int32_t *tlen,
                    const int32_t *mode, const int32_t *want_path,
                    uint8_t *aln, const int64_t *aln_off, lcd_edlib_result_t *results) {
    lcd_plan_t *plan = lcd_edlib_plan_create(n, seqs, seqs_len, query_off, qlen, target_off, tlen, mode, want_path);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_edlib_plan_fetch(plan, nullptr, aln, aln_off, results);
    lcd_plan_destroy(plan);
    return rc;
}

}
