// phase_device.cuh -- device-side logic of K4: read -> haplotype assignment and variant phasing of one region
// chunk, replacing assign_hap_based_on_germline_het_vars_kmeans (reference src/assign_hap.c:473-547 and its helpers
// :16-467).  Integer arithmetic only.
//
// B200 design: one CTA per chunk (chunks are independent until stitch_var_main); all state of a chunk -- the
// read x variant allele profile in CSR form, per-variant consensus / per-haplotype allele counts -- stays in HBM/L2
// for the whole call.
//   * seed pass (:499-527): inherently sequential over reads (every assignment moves the consensus the next read is
//     scored against), so one warp walks the reads in the reference's order (variants from the seed outwards; the
//     reads covering a variant in the order cgranges returns them, precomputed by the host glue together with a
//     prefix maximum of the span ends so that the candidates of a variant are one short index range), and spreads
//     the per-read work -- scoring against both haplotypes and the profile / consensus update -- over its 32 lanes.
//   * iterations (:530-542): the agree / conflict counts of adjacent heterozygous variants are accumulated per read
//     in parallel (atomics), the flip parity / phase-set propagation over the variants is a ballot scan, and the
//     re-assignment of all reads runs one read per thread (the in-place fill-in of a half-known consensus is
//     idempotent, and the consensus itself is only rebuilt after all reads have been scored).
// The file compiles for the host as well (tests/emu: a "warp" of one lane, a CTA of one thread).
#pragma once
#include <stdint.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace phase {

#ifdef LCD_EMU
constexpr int WARP = 1;
#define LCD_PHASE_TID 0
#define LCD_PHASE_NT 1
#else
constexpr int WARP = 32;
#define LCD_PHASE_TID ((int)threadIdx.x)
#define LCD_PHASE_NT ((int)blockDim.x)
#endif

enum { CLEAN_HET_SNP = 0x004, CLEAN_HET_INDEL = 0x008, CLEAN_HOM_VAR = 0x080, NOISY_CAND_HET_VAR = 0x100, NOISY_CAND_HOM_VAR = 0x200 };
constexpr int GERMLINE_CLEAN = CLEAN_HET_SNP | CLEAN_HET_INDEL | CLEAN_HOM_VAR;
constexpr int CDIFF = 8;               // BAM_CDIFF

struct __align__(16) Chunk {
    int32_t n_reads, n_vars, target, is_ont;
    int32_t n_cr, pad;
    int64_t read_off, var_off;         // first read / variant of the chunk in the concatenated arrays
};

struct KernelArgs {
    const Chunk *chunks; int32_t n_chunks;
    // per read (concatenated over chunks)
    const int32_t *ordered_ids; const uint8_t *is_skipped; const int32_t *pstart, *pend; const int64_t *allele_off; const int8_t *alleles;
    const int32_t *cr_order, *cr_pmax_end;        // reads in cgranges order; running maximum of their span ends
    int32_t *haps; long long *phase_sets; int32_t *agree, *conflict;
    // per variant
    const int32_t *cate, *type, *hp, *nuniq, *alle_covs, *total_cov; const long long *pos;
    int32_t *cons, *prof; long long *var_ps;
    int32_t *valid, *flags, *n_agree, *n_conf, *snap;     // scratch: valid-variant list, is_het, counts, consensus snapshot
};

struct Phaser {
    // views of one chunk
    int nr, nv, target, is_ont, n_cr, n_valid;
    const int32_t *ordered, *pstart, *pend, *cr_order, *cr_pmax; const uint8_t *skipped; const int64_t *aoff; const int8_t *alleles;
    const int32_t *cate, *type, *hp, *nuniq, *covs, *tcov; const long long *pos;
    int32_t *haps, *agree, *conflict, *cons, *prof, *valid, *is_het, *n_agree, *n_conf, *snap; long long *psets, *var_ps;
    int *sh;                          // >= 8 ints of shared scratch

    __device__ __forceinline__ int allele(int r, int v) const { return alleles[aoff[r] + (v - pstart[r])]; }
    __device__ static void cta_sync() {
#ifndef LCD_EMU
        __syncthreads();
#endif
    }

    // read_to_cons_allele_score :127-147 (fills in a half-known consensus in place)
    __device__ __forceinline__ int score(int hap, int v, int ct, int al) {
        const int w = (ct == CLEAN_HET_SNP || ct == CLEAN_HET_INDEL) ? 2 : 1;
        int32_t *ca = cons + 3 * v;
        int a = ca[hap], b = ca[3 - hap];
        if (a == -1 && b == -1) return 0;
        if (a == -1) { a = 1 - b; ca[hap] = a; }
        if (b == -1) { b = 1 - a; ca[3 - hap] = b; }
        if (a == al) return w;
        if (a == -1) return 0;
        return -w;
    }
    // update_var_hap_to_cons_alle :244-268
    __device__ __forceinline__ void update_cons(int v, int hap) {
        int max_cov = 0, best = -1, total = 0;
        const int32_t *p = prof + 12 * v + 4 * hap;
        for (int i = 0; i < nuniq[v]; ++i) { const int x = p[i]; total += x; if (x > max_cov) { max_cov = x; best = i; } }
        if (is_ont && hp[v] == 1 && max_cov < total * 0.67) best = -1;
        cons[3 * v + hap] = best;
    }
    // the per-variant part of init_assign_read_hap_based_on_cons_alle :158-182: contributions of variant v to
    // acc = {score1, score2, used1, used2, agree1, agree2, conflict1, conflict2}
    __device__ __forceinline__ void contribute(int r, int v, int (&acc)[8]) {
        const int ct = cate[v];
        if ((ct & target) == 0) return;
        if (hp[v] == 1 || ct == NOISY_CAND_HOM_VAR) return;
        const int al = allele(r, v);
        if (al < 0) return;
#pragma unroll
        for (int hap = 1; hap <= 2; ++hap) {
            const int s = score(hap, v, ct, al);
            if (s != 0) {
                if (ct != CLEAN_HOM_VAR) acc[2 + hap - 1]++;
                if ((ct & GERMLINE_CLEAN) > 0 && type[v] == CDIFF) { if (s > 0) acc[4 + hap - 1]++; else acc[6 + hap - 1]++; }
            }
            if (ct != CLEAN_HOM_VAR) acc[hap - 1] += s;
        }
    }
    // the decision of init_assign_read_hap_based_on_cons_alle :183-196
    __device__ __forceinline__ int decide(int r, const int (&acc)[8], bool write) {
        int max_hap = 0, max_s = 0, min_hap = 0, min_s = 0;
#pragma unroll
        for (int hap = 1; hap <= 2; ++hap) {
            const int s = acc[hap - 1];
            if (s > max_s) { max_hap = hap; max_s = s; }
            else if (s < min_s) { min_hap = hap; min_s = s; }
        }
        int ag = 0, cf = 0, res;
        if (acc[2] == 0 && acc[3] == 0) res = -1;
        else if (max_s == 0 && min_s == 0) res = 0;
        else if (max_s > 0) { ag = acc[4 + max_hap - 1]; cf = acc[6 + max_hap - 1]; res = max_hap; }
        else res = 3 - min_hap;
        if (write) { agree[r] = ag; conflict[r] = cf; }
        return res;
    }

    // one read with the lanes of a warp dealt over its variant span (seed pass)
    __device__ int assign_read_warp(int r, int lane) {
        int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int v = pstart[r] + lane; v <= pend[r]; v += WARP) contribute(r, v, acc);
#ifndef LCD_EMU
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __reduce_add_sync(0xffffffffu, acc[i]);
#endif
        return decide(r, acc, lane == 0);
    }

    __device__ void run(const KernelArgs &a, const Chunk &c, int *shared) {
        const int tid = LCD_PHASE_TID, NT = LCD_PHASE_NT, lane = tid % WARP;
        sh = shared;
        nr = c.n_reads; nv = c.n_vars; target = c.target; is_ont = c.is_ont; n_cr = c.n_cr;
        ordered = a.ordered_ids + c.read_off; skipped = a.is_skipped + c.read_off; pstart = a.pstart + c.read_off; pend = a.pend + c.read_off;
        aoff = a.allele_off + c.read_off; alleles = a.alleles; cr_order = a.cr_order + c.read_off; cr_pmax = a.cr_pmax_end + c.read_off;
        haps = a.haps + c.read_off; psets = a.phase_sets + c.read_off; agree = a.agree + c.read_off; conflict = a.conflict + c.read_off;
        cate = a.cate + c.var_off; type = a.type + c.var_off; hp = a.hp + c.var_off; nuniq = a.nuniq + c.var_off; covs = a.alle_covs + 4 * c.var_off;
        tcov = a.total_cov + c.var_off; pos = a.pos + c.var_off;
        cons = a.cons + 3 * c.var_off; prof = a.prof + 12 * c.var_off; var_ps = a.var_ps + c.var_off;
        valid = a.valid + c.var_off; is_het = a.flags + c.var_off; n_agree = a.n_agree + c.var_off; n_conf = a.n_conf + c.var_off; snap = a.snap + 2 * c.var_off;

        // valid variants (:476-481), in order: thread 0 compacts (a few thousand entries at most)
        if (tid == 0) {
            int n = 0;
            for (int v = 0; v < nv; ++v) if (cate[v] & target) valid[n++] = v;
            sh[0] = n;
        }
        cta_sync();
        n_valid = sh[0];
        if (n_valid == 0) return;                                              // :483-486: nothing is touched
        // read_init_hap_phase_set :16-20, var_init_hap_profile_cons_allele :39-63
        for (int r = tid; r < nr; r += NT) { haps[r] = 0; psets[r] = -1; }
        for (int k = tid; k < n_valid; k += NT) {
            const int v = valid[k];
            for (int i = 0; i < 12; ++i) prof[12 * v + i] = 0;
            int best = -1;
            if (!(is_ont == 1 && hp[v])) { int mx = 0; for (int i = 0; i < nuniq[v]; ++i) if (covs[4 * v + i] > mx) { mx = covs[4 * v + i]; best = i; } }
            cons[3 * v] = best;
            cons[3 * v + 1] = cons[3 * v + 2] = (cate[v] == NOISY_CAND_HOM_VAR || cate[v] == CLEAN_HOM_VAR) ? 1 : -1;
        }
        // select_init_var :94-125 (thread 0)
        if (tid == 0) {
            int best[4] = {-1, -1, -1, -1}, depth[4] = {0, 0, 0, 0};
            for (int k = 0; k < n_valid; ++k) {
                const int v = valid[k], ct = cate[v];
                int cls = -1;
                if (ct == CLEAN_HET_SNP) cls = 0;
                else if (ct == CLEAN_HET_INDEL) cls = 1;
                else if (ct == NOISY_CAND_HET_VAR) { if (type[v] == CDIFF) cls = 2; else if (hp[v] == 0) cls = 3; }
                if (cls >= 0 && (best[cls] == -1 || depth[cls] < tcov[v])) { best[cls] = k; depth[cls] = tcov[v]; }
            }
            int init_k = -1;
            for (int cls = 0; cls < 4 && init_k < 0; ++cls) init_k = best[cls];
            sh[1] = init_k;
        }
        cta_sync();
        const int init_k = sh[1];
        // ---- seed pass :499-527 (warp 0)
        if (init_k != -1 && tid < WARP) {
            for (int step = 0; step < n_valid; ++step) {
                const int k = step == 0 ? init_k : (step <= init_k ? init_k - step : step);
                const int v = valid[k];
                if (cate[v] == NOISY_CAND_HOM_VAR || cate[v] == CLEAN_HOM_VAR) continue;
                // reads covering v sit in cr positions [xlo, xhi): start <= v (sorted by start) and running max end >= v
                int lo = 0, hi = n_cr;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (pstart[cr_order[mid]] <= v) lo = mid + 1; else hi = mid; }
                const int xhi = lo;
                lo = 0; hi = xhi;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (cr_pmax[mid] >= v) hi = mid; else lo = mid + 1; }
                for (int x = lo; x < xhi; ++x) {
                    const int r = cr_order[x];
                    if (pend[r] < v || skipped[r] || haps[r] != 0) continue;
                    int hap = assign_read_warp(r, lane);
                    if (hap == -1) hap = 1;
#ifndef LCD_EMU
                    __syncwarp();
#endif
                    if (lane == 0) haps[r] = hap;
                    // update_var_hap_profile_cons_alle_based_on_read_hap :270-290
                    for (int u = pstart[r] + lane; u <= pend[r]; u += WARP) {
                        if ((cate[u] & target) == 0) continue;
                        const int al = allele(r, u);
                        if (al < 0) continue;
                        if (hap == 0) { for (int h = 1; h <= 2; ++h) { prof[12 * u + 4 * h + al] += 1; update_cons(u, h); } }
                        else { prof[12 * u + 4 * hap + al] += 1; update_cons(u, hap); }
                    }
#ifndef LCD_EMU
                    __syncwarp();
#endif
                }
            }
        }
        cta_sync();
        // ---- iterations :530-542
        for (int iter = 0; iter < 10; ++iter) {
            // iter_update_var_hap_cons_phase_set :345-422
            for (int k = tid; k < n_valid; k += NT) {
                const int v = valid[k]; const int32_t *ca = cons + 3 * v;
                is_het[k] = (ca[1] != -1 && ca[2] != -1 && ca[1] != ca[2] && hp[v] == 0) ? 1 : 0;
                n_agree[k] = 0; n_conf[k] = 0;
            }
            // per-variant index in the valid list is needed to address the counts: snap[] doubles as the inverse map here
            for (int k = tid; k < n_valid; k += NT) snap[valid[k]] = k;
            if (tid == 0) { sh[2] = 0; sh[3] = 0; }
            cta_sync();
            // agree / conflict of every pair of adjacent heterozygous variants, accumulated read by read: a read counts
            // for a pair iff it spans both, so within a read the pairs are its consecutive het variants (check_agree_haps :307-320)
            for (int x = tid; x < n_cr; x += NT) {
                const int r = cr_order[x];
                if (skipped[r]) continue;
                const int hap = haps[r];
                if (hap == 0) continue;
                int pv = -1, pa = 0;
                for (int v = pstart[r]; v <= pend[r]; ++v) {
                    if ((cate[v] & target) == 0) continue;
                    const int k = snap[v];
                    if (!is_het[k]) continue;
                    const int al = allele(r, v);
                    if (pv >= 0 && pa >= 0 && al >= 0) {
                        const int32_t *c1 = cons + 3 * pv, *c2 = cons + 3 * v;
                        if (c1[hap] == pa && c2[hap] == al) atomicAdd(&n_agree[k], 1);
                        else if (c1[hap] == pa && c2[3 - hap] == al) atomicAdd(&n_conf[k], 1);
                    }
                    pv = v; pa = al;
                }
            }
            cta_sync();
            // flip parity / phase-set propagation over the valid variants (warp 0, ballot scans over 32-variant groups)
            if (tid < WARP) {
                int flip = 0, changed1 = 0; long long ps = -1;
                for (int base = 0; base < n_valid; base += WARP) {
                    const int k = base + lane;
                    const bool in = k < n_valid;
                    int v = 0, het = 0, newps = 0, tog = 0; long long own = -1;
                    if (in) {
                        v = valid[k]; own = type[v] == CDIFF ? pos[v] : pos[v] - 1;
                        het = k > 0 && is_het[k];
                        if (k == 0) newps = 1;
                        else if (het) { if (n_agree[k] < 2 && n_conf[k] < 2) newps = 1; else if (n_conf[k] > n_agree[k]) tog = 1; }
                    }
#ifndef LCD_EMU
                    const unsigned mt = __ballot_sync(0xffffffffu, tog), mn = __ballot_sync(0xffffffffu, newps);
                    const unsigned below = (2u << lane) - 1;                  // lanes <= this one
                    const int my_flip = flip ^ (__popc(mt & below) & 1);
                    const unsigned mnb = mn & below;
                    const int src = mnb ? 31 - __clz(mnb) : -1;
                    const long long own_src = __shfl_sync(0xffffffffu, own, src < 0 ? 0 : src);
                    const long long my_ps = src < 0 ? ps : own_src;
                    if (in) { var_ps[v] = my_ps; if (het && my_flip) changed1 = 1; }
                    flip ^= __popc(mt) & 1;
                    if (mn) ps = __shfl_sync(0xffffffffu, own, 31 - __clz(mn));
#else
                    if (in) { if (tog) flip ^= 1; if (newps) ps = own; var_ps[v] = ps; if (het && flip) changed1 = 1; }
#endif
                }
#ifndef LCD_EMU
                changed1 = __any_sync(0xffffffffu, changed1);
#endif
                if (lane == 0) sh[2] = changed1;
                // (the reference swaps hap_to_cons_alle[1] and [2] of a flipped variant twice, :409-413: a no-op)
            }
            cta_sync();
            // iter_update_var_hap_to_cons_alle :425-467
            for (int k = tid; k < n_valid; k += NT) { const int v = valid[k]; for (int i = 0; i < 12; ++i) prof[12 * v + i] = 0; }
            // snapshot of the consensus: n_agree / n_conf are free again
            for (int k = tid; k < n_valid; k += NT) { n_agree[k] = cons[3 * valid[k] + 1]; n_conf[k] = cons[3 * valid[k] + 2]; }
            cta_sync();
            for (int i = tid; i < nr; i += NT) {
                const int r = ordered[i];
                if (skipped[r]) continue;
                int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int v = pstart[r]; v <= pend[r]; ++v) contribute(r, v, acc);
                int hap = decide(r, acc, true);
                if (hap == -1) hap = 0;
                haps[r] = hap;
                for (int u = pstart[r]; u <= pend[r]; ++u) {                   // update_var_hap_profile_based_on_read_hap :292-305
                    if ((cate[u] & target) == 0) continue;
                    const int al = allele(r, u);
                    if (al < 0) continue;
                    if (hap == 0) { atomicAdd(&prof[12 * u + 4 + al], 1); atomicAdd(&prof[12 * u + 8 + al], 1); }
                    else atomicAdd(&prof[12 * u + 4 * hap + al], 1);
                }
            }
            cta_sync();
            for (int k = tid; k < n_valid; k += NT) {
                const int v = valid[k];
                update_cons(v, 1); update_cons(v, 2);
                if (cons[3 * v + 1] != n_agree[k] || cons[3 * v + 2] != n_conf[k]) sh[3] = 1;
            }
            cta_sync();
            const int changed = sh[2] | sh[3];
            cta_sync();
            if (!changed) break;
        }
        // update_read_phase_set :322-339
        for (int i = tid; i < nr; i += NT) {
            const int r = ordered[i];
            if (skipped[r] || pstart[r] == -1) continue;
            long long ps = -1;
            for (int v = pstart[r]; v <= pend[r]; ++v) {
                if ((cate[v] & target) == 0) continue;
                const int32_t *ca = cons + 3 * v;
                if (ca[1] != -1 && ca[2] != -1 && ca[1] != ca[2]) ps = var_ps[v];
                if (ps != -1) break;
            }
            psets[r] = ps;
        }
    }
};

} // namespace phase
} // namespace lcd
