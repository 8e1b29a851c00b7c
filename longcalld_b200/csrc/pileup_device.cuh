// pileup_device.cuh -- device-side logic of K2: the per-site coverage pass of the pileup scan, replacing
// collect_cand_vars (reference src/collect_var.c:238-249) = update_cand_vars_from_digar (src/bam_utils.c:287-329) over
// every kept read: a merge-join of the read's difference list (digar1_t X / I / D events, position sorted) against the
// chunk's sorted candidate sites, counting per site reference / alternative observations (by strand) and low-quality ones.
//
// B200 design: ONE THREAD PER READ over all chunks of a batch (~1 000 reads per 500 kb chunk, ~100 k per 50 Mb): a read
// touches a few hundred events and sites, the counts commute, so reads run independently and meet only in the
// atomicAdd on the 8 counters of a site (32-byte records, L2 resident).  All inputs are flat SoA arrays concatenated
// over the chunks; nothing is copied per read.
// The file compiles for the host as well (tests/emu).
#pragma once
#include <stdint.h>
#include "../../include/lcd_gpu.h"

namespace lcd {
namespace pileup {

enum { CINS = 1, CDEL = 2, CEQUAL = 7, CDIFF = 8 };

struct __align__(16) Chunk { int32_t n_sites, min_bq, min_sv_len, pad; int64_t site_off; };

struct KernelArgs {
    const Chunk *chunks; int64_t n_reads_total;
    // per read (concatenated)
    const int32_t *read_chunk; const uint8_t *read_active;      // active = listed in ordered_read_ids and not skipped
    const long long *read_beg, *read_end; const uint8_t *read_is_rev;
    const long long *digar_first; const int32_t *n_digar; const long long *qual_off; const uint8_t *qual;
    // per event
    const long long *digar_pos; const int8_t *digar_type; const int32_t *digar_len, *digar_qi; const uint8_t *digar_low_qual;
    const long long *digar_alt_off; const uint8_t *digar_alt;
    // per site
    const long long *site_pos; const int32_t *site_type, *site_ref_len, *site_alt_len; const long long *site_alt_off; const uint8_t *site_alt;
    int32_t *site_counts;                                        // [n_sites_total][8]
};

// exact_comp_var_site_ins (src/collect_var.c:1901-1935) of site s against the site made from event d
// (make_var_site_from_digar, src/collect_var.c:1113-1121)
__device__ __forceinline__ int comp_site_event(const KernelArgs &a, long long s, long long d, int min_sv_len) {
    const int st = a.site_type[s], dt = a.digar_type[d];
    const long long ps = st == CDIFF ? a.site_pos[s] : a.site_pos[s] - 1, pd = dt == CDIFF ? a.digar_pos[d] : a.digar_pos[d] - 1;
    if (ps < pd) return -1;
    if (ps > pd) return 1;
    if (st < dt) return -1;
    if (st > dt) return 1;
    const int dl = a.digar_len[d];
    const int d_ref = dt == CINS ? 0 : (dt == CDEL ? dl : 1), d_alt = dt == CDEL ? 0 : dl;
    const int s_ref = a.site_ref_len[s], s_alt = a.site_alt_len[s];
    if (s_ref < d_ref) return -1;
    if (s_ref > d_ref) return 1;
    if (st == CDIFF || (st == CINS && s_alt < min_sv_len)) {
        if (s_alt < d_alt) return -1;
        if (s_alt > d_alt) return 1;
        const uint8_t *x = a.site_alt + a.site_alt_off[s], *y = a.digar_alt + a.digar_alt_off[d];
        for (int i = 0; i < s_alt; ++i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
        return 0;
    } else if (st == CINS) {
        const int mn = s_alt < d_alt ? s_alt : d_alt, mx = s_alt > d_alt ? s_alt : d_alt;
        if (mn >= mx * 0.8) return 0;
        return s_alt - d_alt;
    }
    return 0;
}

// update_var_site_with_allele, src/bam_utils.c:234-243
__device__ __forceinline__ void count(const KernelArgs &a, long long s, bool low_qual, int strand, int allele) {
    int32_t *c = a.site_counts + 8 * s;
    if (low_qual) { atomicAdd(c + 1, 1); return; }
    atomicAdd(c, 1); atomicAdd(c + 2 + allele, 1); atomicAdd(c + 4 + 2 * strand + allele, 1);
}

// update_cand_vars_from_digar, src/bam_utils.c:287-329, for read g
__device__ void process_read(const KernelArgs &a, long long g) {
    if (!a.read_active[g]) return;
    const Chunk ch = a.chunks[a.read_chunk[g]];
    const long long s0 = ch.site_off, s_end = ch.site_off + ch.n_sites;
    const long long beg = a.read_beg[g], end = a.read_end[g];
    const int strand = a.read_is_rev[g];
    long long s;
    {   // get_var_site_start, src/bam_utils.c:229-241
        const long long target = beg > 0 ? beg - 1 : beg;
        long long left = s0, right = s_end;
        while (left < right) {
            const long long mid = left + (right - left) / 2;
            const long long mp = a.site_type[mid] == CDIFF ? a.site_pos[mid] : a.site_pos[mid] - 1;
            if (mp < target) left = mid + 1; else right = mid;
        }
        while (left < s_end && a.site_pos[left] < beg) left++;
        s = left;
    }
    long long d = a.digar_first[g];
    const long long d_end = d + a.n_digar[g];
    const uint8_t *qual = a.qual + a.qual_off[g];
    while (s < s_end && d < d_end) {
        const int dt = a.digar_type[d];
        if (dt == CEQUAL) { d++; continue; }
        const int ret = comp_site_event(a, s, d, ch.min_sv_len);
        if (ret < 0) { count(a, s, false, strand, 0); s++; }
        else if (ret == 0) {
            bool low = a.digar_low_qual[d] != 0;
            if (!low) {                                             // get_digar_ave_qual, src/bam_utils.c:258-280
                const int qi = a.digar_qi[d];
                int ave = 0;
                if (qi >= 0) {
                    int q0, q1;
                    if (dt == CDEL) { if (qi == 0) { q0 = q1 = 0; } else { q0 = qi - 1; q1 = qi; } }
                    else { q0 = qi; q1 = qi + a.digar_len[d] - 1; }
                    int sum = 0;
                    for (int i = q0; i <= q1; ++i) sum += qual[i];
                    ave = sum / (q1 - q0 + 1);
                }
                low = ave < ch.min_bq;
            }
            count(a, s, low, strand, 1); s++;
        } else d++;
    }
    for (; s < s_end; ++s) {
        if (a.site_pos[s] > end) break;
        count(a, s, false, strand, 0);
    }
}

} // namespace pileup
} // namespace lcd
