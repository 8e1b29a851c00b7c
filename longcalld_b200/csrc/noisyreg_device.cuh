// noisyreg_device.cuh -- K2c: the chunk's noisy-region set and the candidate sites that stay clean-region candidates
// (second half).  One CTA per region chunk.
//
// What it replaces: pre_process_noisy_regs (src/collect_var.c:557-643) and classify_cand_vars after its first loop (:925-1033, out_somatic = 0)
// with cr_extend_noisy_regs_with_low_comp / low_comp_cr_start_end (:466-479,538-553), cr_add_var_cr (:754-778), var_noisy_reads_ratio and its
// caches (:657-751), post_process_noisy_regs + collect_noisy_reg_start_end (:482-535,646-655) and the interval algebra of src/cgranges.c
// (cr_index order, cr_overlap, cr_cluster0 / cr_merge / cr_merge2 :225-335, cr_is_contained :512-530).
//
// Layout: interval lists are three int32 arrays (st, en, label) in the chunk's scratch; candidate sites, the reads' spans, difference
// records and noisy intervals are read where K1 / K1b / K2b left them.  The phases run in the reference's order, each dealt over the
// threads of the CTA where its items are independent:
//   * sorting a list (cr_index: by st << 32 | en) is a rank sort (a few hundred to a few thousand intervals);
//   * the low-complexity extension, the reads' votes per region, the per-site rules and the final containment test are one item per thread,
//     with binary searches into the sorted lists (the merged region set is disjoint; the low-complexity intervals carry a running maximum
//     of their ends so that "every interval overlapping [a, b)" is a walk back from a lower bound);
//   * var_noisy_reads_ratio is asked only for sites that overlap another site: the thread checks the ~1 000 read spans and binary-searches
//     every read's position-sorted records for an X / I / D record on the query (the reference's merged error
//     intervals are unions of overlapping records: a query meets the union iff it meets a record);
//   * cr_cluster0 is order-dependent (an interval absorbs every later one that starts within min(label, label') of its growing end) and runs on
//     one thread -- with the exact early exit that once an interval starts past end + label no later one can be absorbed until the next
//     absorption (labels are windows >= 0, the list is sorted by start), so a pass is linear;
//   * collect_noisy_reg_start_end's two-pointer sweep over (regions, sites) runs on one thread, the flank extension per region.
#pragma once
#include <stdint.h>

namespace lcd {
namespace noisyreg {

enum { CINS = 1, CDEL = 2, CDIFF = 8 };
enum { NON_VAR = 0x800, LOW_COV_VAR = 0x001, STRAND_BIAS_VAR = 0x002, LOW_AF_VAR = 0x400, REP_HET_VAR = 0x010 };
constexpr int NOT_CAND = NON_VAR | LOW_COV_VAR | STRAND_BIAS_VAR;
enum { ST_OK = 0, ST_REG_CAP = -5, ST_LOW = -7 };      // ST_LOW: the sdust plan the chunk reads its low-complexity intervals from failed for it

struct Ivs { int *st, *en, *label; };

struct Chunk {
    long long reg_beg, reg_end;
    double min_af;
    int min_alt_dp, flank, is_ont, n_sites, n_reads, n_cnreg, n_low, cap;          // cap: capacity of every interval list of the chunk
    const long long *site_pos; const int *site_type, *site_ref_len, *var_cate_in;
    const long long *cn_beg, *cn_end; const int *cn_label;
    const long long *low_beg, *low_end;                                             // ascending starts
    const long long *n_low_dev; const int *low_status;                              // chained to K0 (sdust_device.cuh): the number of intervals and K0's status, read on the device
    const unsigned char *is_skipped, *active;                                       // a read counts when !is_skipped[r] (K1's skip) and, where given, active[r] (not skipped by the loader)
    const long long *read_beg, *read_end, *digar_first; const int *n_digar;
    const long long *digar_pos; const signed char *digar_type; const int *digar_len;
    const long long *nreg_first; const int *n_nreg; const long long *nreg_beg, *nreg_end;
    const int *nreg_label; int cn_from_reads;                                      // 1: chunk_noisy_regs is gathered from the reads' intervals that touch the region (src/bam_utils.c:819-832), as K1 left them in HBM
    // outputs
    int *var_cate; unsigned char *keep; long long *out_regs; long long reg_cap; long long *n_regs; int *status;     // out_regs: n_regs starts, n_regs ends (int64), n_regs labels (int32), back to back
    // scratch
    Ivs A, B; int *low_pmax, *vp_pmax, *tot, *noi, *ctr;                             // ctr[0]: list length, ctr[1]: appended intervals
};

__device__ __forceinline__ bool skipped(const Chunk &c, int r) { return c.is_skipped[r] || (c.active && !c.active[r]); }
__device__ __forceinline__ unsigned long long key_of(int st, int en) { return ((unsigned long long)(long long)st << 32) | (unsigned long long)(long long)en; }

// cr_index: B <- A sorted by key (ties: input order)
__device__ __forceinline__ void rank_sort(const Ivs &A, const Ivs &B, int n, int tid, int nt) {
    for (int i = tid; i < n; i += nt) {
        const unsigned long long k = key_of(A.st[i], A.en[i]); int r = 0;
        for (int j = 0; j < n; ++j) { const unsigned long long kj = key_of(A.st[j], A.en[j]); r += (kj < k || (kj == k && j < i)) ? 1 : 0; }
        B.st[r] = A.st[i]; B.en[r] = A.en[i]; B.label[r] = A.label[i];
    }
}

// cr_merge (:290-301): cr_cluster0 passes until the list stops shrinking; one thread; src sorted; the result ends in `a` (lists swap per pass).
// Returns the length.  `m` is a scratch flag array (>= n ints).
__device__ __forceinline__ int merge_list(Ivs &a, Ivs &b, int n, int fixed_win, int *m) {
    for (;;) {
        for (int j = 0; j < n; ++j) m[j] = 0;
        int o = 0;
        for (int j = 0; j < n; ++j) {
            if (m[j]) continue;
            unsigned long long ms = (unsigned long long)(long long)a.st[j], me = (unsigned long long)(long long)a.en[j]; int ml = a.label[j];
            for (int k = j + 1; k < n; ++k) {
                const unsigned long long ns = (unsigned long long)(long long)a.st[k];
                if (ns > me + (unsigned long long)(long long)(fixed_win < 0 ? ml : fixed_win)) break;          // exact: min(ml, nl) <= ml, starts ascend
                if (m[k]) continue;
                const unsigned long long ne = (unsigned long long)(long long)a.en[k]; const int nl = a.label[k];
                const int win = fixed_win < 0 ? (ml < nl ? ml : nl) : fixed_win;
                if (me + (unsigned long long)(long long)win >= ns) { ml = ml > nl ? ml : nl; ms = ms < ns ? ms : ns; me = me > ne ? me : ne; m[k] = 1; }
            }
            b.st[o] = (int)ms; b.en[o] = (int)me; b.label[o] = ml; ++o;
        }
        // cr_index of the pass' output: already in key order (starts ascend strictly: an interval that starts at or before a kept one's end
        // was absorbed by it)
        Ivs t = a; a = b; b = t;
        if (o == n) return n;
        n = o;
    }
}

// last index in a sorted list whose start is < x, plus one (the intervals [0, ub) start before x)
__device__ __forceinline__ int starts_below(const int *st, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (st[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
// does the sorted, disjoint region list R overlap [qs, qe)?
__device__ __forceinline__ bool regs_overlap(const Ivs &R, int n, int qs, int qe) {
    const int ub = starts_below(R.st, n, qe);            // candidates start before qe; disjoint + sorted: the last of them has the largest end
    return ub > 0 && qs < R.en[ub - 1];
}

// var_noisy_reads_ratio (:718-751)
__device__ __forceinline__ float noisy_reads_ratio(const Chunk &c, long long var_start, long long var_end) {
    const int qs = (int)(var_start - 1), qe = (int)var_end;
    int total = 0, noisy = 0;
    for (int r = 0; r < c.n_reads; ++r) {
        if (skipped(c, r) || c.n_digar[r] <= 0 || c.read_beg[r] > c.read_end[r]) continue;
        const int rb = (int)(c.read_beg[r] - 1), re = (int)c.read_end[r];
        if (rb < qe && qs < re) ++total;
        // (a read is noisy here whether or not its span counts: an insertion after its last base lies one past read_end; K1's records lie in
        // [read_beg, read_end + 1], so a read whose widened span misses the query has no record on it)
        if (!(rb - 1 < qe && qs < re + 2)) continue;
        // the read's records are in reference order: the last one starting before qe, then back while they still reach past qs
        const long long f = c.digar_first[r]; int lo = 0, hi = c.n_digar[r];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(c.digar_pos[f + mid] - 1) < qe) lo = mid + 1; else hi = mid; }
        bool hit = false;
        for (int j = lo - 1; j >= 0 && !hit; --j) {
            const int t = c.digar_type[f + j]; const int cs = (int)(c.digar_pos[f + j] - 1);
            int ce = (int)c.digar_pos[f + j];
            if (t != CINS) ce += c.digar_len[f + j] - 1;          // (=, X, D, N, clips with a reference span; an insertion sits on one base)
            if (t == CDIFF || t == CINS || t == CDEL) { if (cs < qe && qs < ce) hit = true; }
            if (ce <= qs && t != CINS && c.digar_len[f + j] > 0) break;      // a reference-consuming record that ends before the query: nothing further back reaches it
        }
        noisy += hit ? 1 : 0;
    }
    if (total == 0) return 0.0f;
    return (float)((float)noisy / (total + 0.0));
}

// cr_add_var_cr (:754-778): the site's span, widened by the low-complexity intervals it touches; appended to list A at ctr[1]
__device__ __forceinline__ void add_var_cr(const Chunk &c, int i, bool check_ratio, int base_n) {
    long long vs = c.site_pos[i], ve = c.site_type[i] == CINS ? c.site_pos[i] : c.site_pos[i] + c.site_ref_len[i] - 1;
    const int qs = (int)(vs - 1), qe = (int)ve;
    // low[j].beg < qe: a prefix of the sorted list; walk back while the running maximum of the ends still passes qs
    int lo = 0, hi = c.n_low;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)c.low_beg[mid] < qe) lo = mid + 1; else hi = mid; }
    for (int j = lo - 1; j >= 0 && c.low_pmax[j] > qs; --j)
        if (qs < (int)c.low_end[j]) { const int s = (int)c.low_beg[j] + 1, e = (int)c.low_end[j]; if (s < vs) vs = s; if (e > ve) ve = e; }
    if (!check_ratio || noisy_reads_ratio(c, vs, ve) >= c.min_af) {
        const int at = base_n + atomicAdd(&c.ctr[1], 1);
        if (at < c.cap) { c.A.st[at] = (int)(vs - 1); c.A.en[at] = (int)ve; c.A.label[at] = 1; }
    }
}

// The whole chunk.  tid / nt: this thread and the number of threads working on the chunk; SYNC: barrier between phases.
template <class SyncF> __device__ void run_chunk(Chunk c, int tid, int nt, SyncF SYNC) {
    if (c.n_low_dev) {
        if (*c.low_status != 0) { if (tid == 0) { *c.status = ST_LOW; *c.n_regs = 0; } return; }
        c.n_low = (int)*c.n_low_dev;
    }
    Ivs A = c.A, B = c.B;
    const int n = c.n_sites;
    // running maximum of the low-complexity ends (for the walk-back overlap queries)
    if (tid == 0) { int m = INT32_MIN; for (int j = 0; j < c.n_low; ++j) { const int e = (int)c.low_end[j]; if (e > m) m = e; c.low_pmax[j] = m; } c.ctr[0] = 0; c.ctr[1] = 0; c.ctr[2] = 0; *c.status = ST_OK; }
    int nR = c.n_cnreg;
    if (c.cn_from_reads) {
        SYNC();
        for (int r = tid; r < c.n_reads; r += nt) {
            if (skipped(c, r)) continue;
            for (int x = 0; x < c.n_nreg[r]; ++x) {
                const long long q = c.nreg_first[r] + x, b = c.nreg_beg[q], e = c.nreg_end[q];
                if (b + 1 > c.reg_end || e < c.reg_beg) continue;
                const int at = atomicAdd(&c.ctr[2], 1);
                if (at < c.cap) { A.st[at] = (int)b; A.en[at] = (int)e; A.label[at] = c.nreg_label[q]; }
            }
        }
        SYNC();
        nR = c.ctr[2];
        if (nR > c.cap) { if (tid == 0) { *c.status = ST_REG_CAP; *c.n_regs = 0; } return; }
    } else {
        for (int i = tid; i < c.n_cnreg; i += nt) { A.st[i] = (int)c.cn_beg[i]; A.en[i] = (int)c.cn_end[i]; A.label[i] = c.cn_label[i]; }
        SYNC();
    }
    // ---- pre_process_noisy_regs
    if (nR > 0) {
        rank_sort(A, B, nR, tid, nt);                                               // cr_index
        SYNC();
        if (c.n_low > 0) {                                                          // cr_extend_noisy_regs_with_low_comp
            for (int i = tid; i < nR; i += nt) {
                const int start = B.st[i] + 1, end = B.en[i]; int ns = start, ne = end;
                int lo = 0, hi = c.n_low;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)c.low_beg[mid] < end) lo = mid + 1; else hi = mid; }
                for (int j = lo - 1; j >= 0 && c.low_pmax[j] > start - 1; --j)
                    if (start - 1 < (int)c.low_end[j]) { if ((int)c.low_beg[j] + 1 < ns) ns = (int)c.low_beg[j] + 1; if ((int)c.low_end[j] > ne) ne = (int)c.low_end[j]; }
                A.st[i] = ns - 1; A.en[i] = ne; A.label[i] = B.label[i];
            }
            SYNC();
            rank_sort(A, B, nR, tid, nt);
            SYNC();
        }
        // B holds the sorted list
        if (tid == 0) { Ivs x = B, y = A; const int m = merge_list(x, y, nR, -1, c.tot); c.ctr[0] = m; if (x.st != A.st) for (int k = 0; k < m; ++k) { A.st[k] = x.st[k]; A.en[k] = x.en[k]; A.label[k] = x.label[k]; } }
        SYNC();
        nR = c.ctr[0];
        for (int k = tid; k < nR; k += nt) { c.tot[k] = 0; c.noi[k] = 0; }
        SYNC();
        // every kept read votes in the regions its span overlaps: noisy when one of its own noisy intervals overlaps the region
        for (int r = tid; r < c.n_reads; r += nt) {
            if (skipped(c, r)) continue;
            const int qs = (int)(c.read_beg[r] - 1), qe = (int)c.read_end[r];
            const int ub = starts_below(A.st, nR, qe);
            for (int k = ub - 1; k >= 0 && qs < A.en[k]; --k) {                      // disjoint + sorted: ends ascend too
                atomicAdd(&c.tot[k], 1);
                bool hit = false;
                for (int x = 0; x < c.n_nreg[r] && !hit; ++x) { const long long q = c.nreg_first[r] + x; if ((int)c.nreg_beg[q] < A.en[k] && A.st[k] < (int)c.nreg_end[q]) hit = true; }
                if (hit) atomicAdd(&c.noi[k], 1);
            }
        }
        SYNC();
        if (tid == 0) {
            const float min_ratio = (float)c.min_af; int o = 0;
            for (int k = 0; k < nR; ++k) {
                if (c.noi[k] < c.min_alt_dp || (float)c.noi[k] / c.tot[k] < min_ratio) continue;
                A.st[o] = A.st[k]; A.en[o] = A.en[k]; A.label[o] = A.label[k]; ++o;
            }
            c.ctr[0] = o;
        }
        SYNC();
        nR = c.ctr[0];
    }
    if (n == 0) {            // classify_cand_vars is not called for a chunk without candidate sites (src/collect_var.c:2923)
        if (tid == 0) { *c.n_regs = nR; if (nR > c.reg_cap) *c.status = ST_REG_CAP; }
        if (nR <= c.reg_cap) for (int k = tid; k < nR; k += nt) { c.out_regs[k] = A.st[k]; c.out_regs[nR + k] = A.en[k]; reinterpret_cast<int *>(c.out_regs + 2 * (size_t)nR)[k] = A.label[k]; }
        return;
    }
    // ---- classify_cand_vars after classify_var_cate.  var_pos_cr: the sites that are not LOW_COV (ONT: nor strand-biased), with a running
    // maximum of their ends in list order
    for (int i = tid; i < n; i += nt) { c.var_cate[i] = c.var_cate_in[i]; c.keep[i] = 0; }
    SYNC();
    if (tid == 0) {
        int m = INT32_MIN;
        for (int i = 0; i < n; ++i) {
            const int cate = c.var_cate[i];
            if (!(cate == LOW_COV_VAR || (c.is_ont && cate == STRAND_BIAS_VAR))) {
                const int en = (int)(c.site_type[i] == CINS ? c.site_pos[i] : c.site_pos[i] + c.site_ref_len[i] - 1);
                if (en > m) m = en;
            }
            c.vp_pmax[i] = m;                                                       // over the members up to and including i
        }
    }
    SYNC();
    for (int i = tid; i < n; i += nt) {
        const int cate = c.var_cate[i];
        if (cate == NON_VAR || cate == STRAND_BIAS_VAR) continue;
        const int qs = (int)(c.site_pos[i] - 1), qe = (int)(c.site_type[i] == CINS ? c.site_pos[i] : c.site_pos[i] + c.site_ref_len[i] - 1);
        if (nR > 0 && regs_overlap(A, nR, qs, qe)) { c.var_cate[i] = NON_VAR; continue; }
        if (cate == LOW_COV_VAR) continue;
        const bool in_reg = c.site_pos[i] >= c.reg_beg && c.site_pos[i] <= c.reg_end;
        if (cate == REP_HET_VAR) { if (in_reg) add_var_cr(c, i, false, nR); continue; }
        // does another member of var_pos_cr overlap the site?  The list is in collect_all_cand_var_sites' order: by anchor (an indel's anchor is
        // the base before it), so a member's start is its anchor or one less.  Back while an earlier member's end still passes the site's start;
        // forward while a later anchor can still start before the site's end.
        bool ovlp = false;
        for (int j = i - 1; j >= 0 && !ovlp && c.vp_pmax[j] > qs; --j) {
            const int cj = c.var_cate_in[j];
            if (cj == LOW_COV_VAR || (c.is_ont && cj == STRAND_BIAS_VAR)) continue;
            const int sj = (int)(c.site_pos[j] - 1), ej = (int)(c.site_type[j] == CINS ? c.site_pos[j] : c.site_pos[j] + c.site_ref_len[j] - 1);
            ovlp = sj < qe && qs < ej;
        }
        for (int j = i + 1; j < n && !ovlp; ++j) {
            const int sj = (int)(c.site_pos[j] - 1), aj = c.site_type[j] == CDIFF ? sj + 1 : sj;
            if (aj > qe) break;
            const int cj = c.var_cate_in[j];
            if (cj == LOW_COV_VAR || (c.is_ont && cj == STRAND_BIAS_VAR)) continue;
            const int ej = (int)(c.site_type[j] == CINS ? c.site_pos[j] : c.site_pos[j] + c.site_ref_len[j] - 1);
            ovlp = sj < qe && qs < ej;
        }
        if (ovlp && in_reg) add_var_cr(c, i, true, nR);
        if (cate == LOW_AF_VAR) c.var_cate[i] = LOW_COV_VAR;
    }
    SYNC();
    const int n_add = c.ctr[1];
    if (nR + n_add > c.cap) { if (tid == 0) { *c.status = ST_REG_CAP; *c.n_regs = 0; } return; }
    if (n_add > 0) {                                                                // cr_merge2
        rank_sort(A, B, nR + n_add, tid, nt);
        SYNC();
        if (tid == 0) { Ivs x = B, y = A; const int m = merge_list(x, y, nR + n_add, -1, c.tot); c.ctr[0] = m; if (x.st != A.st) for (int k = 0; k < m; ++k) { A.st[k] = x.st[k]; A.en[k] = x.en[k]; A.label[k] = x.label[k]; } }
        SYNC();
        nR = c.ctr[0];
    }
    // ---- post_process_noisy_regs: collect_noisy_reg_start_end's sweep (one thread), then the flank extension per region
    int *max_left = c.tot, *min_right = c.noi;
    if (tid == 0) {
        for (int k = 0; k < nR; ++k) max_left[k] = min_right[k] = -1;
        for (int k = 0, v = 0; k < nR && v < n;) {
            if (c.var_cate[v] & NOT_CAND) { v++; continue; }
            const int vs = (int)c.site_pos[v], ve = (int)(c.site_pos[v] + c.site_ref_len[v] - 1), rs = A.st[k] + 1, re = A.en[k];
            if (vs > re) { if (min_right[k] == -1) min_right[k] = v; k++; }
            else if (ve < rs) { max_left[k] = v; v++; }
            else v++;
        }
    }
    SYNC();
    for (int k = tid; k < nR; k += nt) {
        const int ml = max_left[k] == -1 ? 0 : max_left[k], mr = min_right[k] == -1 ? n - 1 : min_right[k];
        const int flank = c.flank;
        int cs = A.st[k] + 1 - flank, ce = A.en[k] + flank;
        for (int v = ml; v >= 0; --v) {
            if (c.var_cate[v] & NOT_CAND) continue;
            const int vs = (int)c.site_pos[v], ve = (int)(c.site_pos[v] + c.site_ref_len[v] - 1);
            if (ve < cs - 1) break;
            else if (vs - flank < cs) cs = vs - flank;
        }
        for (int v = mr; v < n; ++v) {
            if (c.var_cate[v] & NOT_CAND) continue;
            const int vs = (int)c.site_pos[v], ve = (int)(c.site_pos[v] + c.site_ref_len[v] - 1);
            if (vs > ce + 1) break;
            else if (ve + flank > ce) ce = ve + flank;
        }
        B.st[k] = cs; B.en[k] = ce; B.label[k] = A.label[k];
    }
    SYNC();
    rank_sort(B, A, nR, tid, nt);
    SYNC();
    if (tid == 0) { Ivs x = A, y = B; const int m = merge_list(x, y, nR, 0, c.tot); c.ctr[0] = m; if (x.st != A.st) for (int k = 0; k < m; ++k) { A.st[k] = x.st[k]; A.en[k] = x.en[k]; A.label[k] = x.label[k]; } }
    SYNC();
    nR = c.ctr[0];
    // ---- the sites that stay: not inside a region (cr_is_contained looks at the last interval starting at or before the site only)
    for (int i = tid; i < n; i += nt) {
        if (c.var_cate[i] & NOT_CAND) continue;
        const int qs = (int)(c.site_pos[i] - 1), qe = (int)(c.site_pos[i] + c.site_ref_len[i]);
        if (nR > 0) {
            int lo = 0, hi = nR;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.st[mid] <= qs) lo = mid + 1; else hi = mid; }
            bool inside = false;
            for (int k = lo - 1; k >= 0 && k < nR; ++k) { if (A.st[k] >= qe) break; if (A.st[k] <= qs && A.en[k] >= qe) { inside = true; break; } }
            if (inside) { c.var_cate[i] = NON_VAR; continue; }
        }
        c.keep[i] = 1;
    }
    if (tid == 0) { *c.n_regs = nR; if (nR > c.reg_cap) *c.status = ST_REG_CAP; }
    if (nR <= c.reg_cap) for (int k = tid; k < nR; k += nt) { c.out_regs[k] = A.st[k]; c.out_regs[nR + k] = A.en[k]; reinterpret_cast<int *>(c.out_regs + 2 * (size_t)nR)[k] = A.label[k]; }
}

} // namespace noisyreg
} // namespace lcd
