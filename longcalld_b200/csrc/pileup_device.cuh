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

struct __align__(16) Chunk { int32_t n_sites, min_bq, min_sv_len, pad; int64_t site_off; int64_t alt_base; int64_t salt_base; int64_t pad2; };   // alt_base / salt_base: first digar_alt / site_alt byte of the chunk (digar_alt_off / site_alt_off are relative to them)

struct KernelArgs {
    const Chunk *chunks; int64_t n_reads_total;
    // per read (concatenated)
    const int32_t *read_chunk; const uint8_t *read_active;      // active = listed in ordered_read_ids and not skipped
    const uint8_t *read_dropped;                                 // optional (K1 -> K2 on the device): reads K1's skip test dropped
    const long long *read_beg, *read_end; const uint8_t *read_is_rev;
    const long long *digar_first; const int32_t *n_digar; const long long *qual_off; const uint8_t *qual;
    // per event
    const long long *digar_pos; const int8_t *digar_type; const int32_t *digar_len, *digar_qi; const uint8_t *digar_low_qual;
    const long long *digar_alt_off; const uint8_t *digar_alt;
    // per site
    const long long *site_pos; const int32_t *site_type, *site_ref_len, *site_alt_len; const long long *site_alt_off; const uint8_t *site_alt;
    int32_t *site_counts;                                        // [n_sites_total][8]
    // read x variant profile (K3)
    const int32_t *var_cate;                                     // per site (= candidate variant)
    const long long *nreg_first; const int32_t *n_nreg; const long long *nreg_beg, *nreg_end;   // per-read noisy intervals [beg, end)
    const long long *row_off; const int32_t *row_cap;            // per read: first entry / capacity of its profile row
    int32_t *prof_start, *prof_end; long long *allele_off; int8_t *alleles; int32_t *alt_qi;
    int32_t *status;                                             // set to a negative code when a row overflows its capacity
};

// exact_comp_var_site_ins (src/collect_var.c:1901-1935) of site s against the site made from event d
// (make_var_site_from_digar, src/collect_var.c:1113-1121)
__device__ __forceinline__ int comp_site_event(const KernelArgs &a, long long s, long long d, int min_sv_len, long long alt_base, long long salt_base) {
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
        const uint8_t *x = a.site_alt + salt_base + a.site_alt_off[s], *y = a.digar_alt + alt_base + a.digar_alt_off[d];
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
    if (!a.read_active[g] || (a.read_dropped && a.read_dropped[g])) return;
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
        const int ret = comp_site_event(a, s, d, ch.min_sv_len, ch.alt_base, ch.salt_base);
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

// get_var_start / get_var_site_start (src/bam_utils.c:215-241): first site a read starting at `beg` can meet
__host__ __device__ inline long long first_site(const long long *site_pos, const int32_t *site_type, long long s0, long long s_end, long long beg) {
    const long long target = beg > 0 ? beg - 1 : beg;
    long long left = s0, right = s_end;
    while (left < right) {
        const long long mid = left + (right - left) / 2;
        const long long mp = site_type[mid] == CDIFF ? site_pos[mid] : site_pos[mid] - 1;
        if (mp < target) left = mid + 1; else right = mid;
    }
    while (left < s_end && site_pos[left] < beg) left++;
    return left;
}
// capacity of a read's profile row: sites from first_site() up to the first one whose sort position is right of end + 1
// (no event of the read lies further right, so the merge-join never visits a later site)
__host__ __device__ inline long long row_end_site(const long long *site_pos, const int32_t *site_type, long long v0, long long s_end, long long end) {
    long long left = v0, right = s_end;
    while (left < right) {
        const long long mid = left + (right - left) / 2;
        const long long mp = site_type[mid] == CDIFF ? site_pos[mid] : site_pos[mid] - 1;
        if (mp <= end + 1) left = mid + 1; else right = mid;
    }
    return left;
}

enum { NON_VAR = 0x800, CAND_SOMATIC_VAR = 0x040 };

// update_read_vs_all_var_profile_from_digar (src/bam_utils.c:446-552, germline categories) for read g; the row is
// stored from the read's first candidate site v0: alleles[row_off + (v - v0)], as collect_read_var_profile
// (src/collect_var.c:1389-1431) would leave it in read_var_profile_t relative to start_var_idx
__device__ void profile_read(const KernelArgs &a, long long g) {
    a.prof_start[g] = -1; a.prof_end[g] = -2; a.allele_off[g] = a.row_off[g];
    if (!a.read_active[g] || (a.read_dropped && a.read_dropped[g])) return;
    const Chunk ch = a.chunks[a.read_chunk[g]];
    const long long s0 = ch.site_off, s_end = ch.site_off + ch.n_sites;
    const long long beg = a.read_beg[g], end = a.read_end[g];
    const long long v0 = first_site(a.site_pos, a.site_type, s0, s_end, beg);
    const long long row = a.row_off[g]; const int cap = a.row_cap[g];
    for (int k = 0; k < cap; ++k) { a.alleles[row + k] = -1; a.alt_qi[row + k] = -1; }
    long long v = v0, d = a.digar_first[g];
    const long long d_end = d + a.n_digar[g];
    const uint8_t *qual = a.qual + a.qual_off[g];
    long long start = -1, last = -2;
    auto set = [&](long long vi, int al, int qi) {
        if (vi - v0 >= cap) { *a.status = -4; return; }
        if (start == -1) start = vi;
        last = vi;
        a.alleles[row + (vi - v0)] = (int8_t)al; a.alt_qi[row + (vi - v0)] = qi;
    };
    while (v < s_end && d < d_end) {
        if (a.var_cate[v] == NON_VAR) { v++; continue; }
        const int dt = a.digar_type[d];
        if (dt == CEQUAL) { d++; continue; }
        // comp_ovlp_var_site = ovlp_var_site (src/collect_var.c:79-95) + exact_comp_var_site (:1878-1898)
        const int st = a.site_type[v], dl = a.digar_len[d];
        const int s_ref = a.site_ref_len[v], s_alt = a.site_alt_len[v];
        const int d_ref = dt == CINS ? 0 : (dt == CDEL ? dl : 1), d_alt = dt == CDEL ? 0 : dl;
        const int beg1 = (int)a.site_pos[v], end1 = beg1 + s_ref, beg2 = (int)a.digar_pos[d], end2 = beg2 + d_ref;
        bool ovlp;
        if (s_ref == 0 && d_ref == 0) ovlp = beg1 == beg2;
        else if (s_ref == 0) ovlp = beg1 > beg2 && end1 < end2;
        else if (d_ref == 0) ovlp = beg2 > beg1 && end2 < end1;
        else ovlp = !(beg1 >= end2 || beg2 >= end1);
        int ret;
        {
            const long long ps = st == CDIFF ? a.site_pos[v] : a.site_pos[v] - 1, pd = dt == CDIFF ? a.digar_pos[d] : a.digar_pos[d] - 1;
            if (ps != pd) ret = ps < pd ? -1 : 1;
            else if (st != dt) ret = st < dt ? -1 : 1;
            else if (s_ref != d_ref) ret = s_ref < d_ref ? -1 : 1;
            else if (s_alt != d_alt) ret = s_alt < d_alt ? -1 : 1;
            else {
                ret = 0;
                if (st == CDIFF || st == CINS) {
                    const uint8_t *x = a.site_alt + ch.salt_base + a.site_alt_off[v], *y = a.digar_alt + ch.alt_base + a.digar_alt_off[d];
                    for (int i = 0; i < s_alt; ++i) if (x[i] != y[i]) { ret = x[i] < y[i] ? -1 : 1; break; }
                }
            }
        }
        if (!ovlp) {
            if (ret < 0) { set(v, 0, -1); v++; }
            else if (ret > 0) d++;
            else { v++; d++; }
        } else if (ret == 0) {
            int ave = 0;                                            // get_digar_ave_qual, src/bam_utils.c:258-280
            const int qi = a.digar_qi[d];
            if (!a.digar_low_qual[d] && qi >= 0) {
                int q0, q1;
                if (dt == CDEL) { if (qi == 0) { q0 = q1 = 0; } else { q0 = qi - 1; q1 = qi; } }
                else { q0 = qi; q1 = qi + dl - 1; }
                int sum = 0;
                for (int i = q0; i <= q1; ++i) sum += qual[i];
                ave = sum / (q1 - q0 + 1);
            }
            set(v, ave < ch.min_bq ? -2 : 1, qi); v++;
        } else { set(v, -1, -1); v++; }
    }
    for (; v < s_end; ++v) {
        const long long p = a.site_pos[v];
        if (p > end) break;
        bool noisy = false;                                         // is_in_noisy_reg, src/bam_utils.h:136-141
        for (long long k = a.nreg_first[g]; k < a.nreg_first[g] + a.n_nreg[g]; ++k) if (a.nreg_beg[k] < p + 1 && p < a.nreg_end[k]) { noisy = true; break; }
        if (noisy) continue;
        set(v, 0, -1);
    }
    a.prof_start[g] = start < 0 ? -1 : (int32_t)(start - s0);
    a.prof_end[g] = last < 0 ? -2 : (int32_t)(last - s0);
    if (start >= 0) a.allele_off[g] = row + (start - v0);
}

} // namespace pileup
} // namespace lcd
