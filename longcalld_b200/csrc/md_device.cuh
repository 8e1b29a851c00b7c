// md_device.cuh -- device-side logic of K1's MD-tag front end: reads with plain-M CIGARs and an MD tag (the bundled ONT data) are turned into
// the =/X CIGARs the difference-list kernels consume, by the walk collect_digar_from_MD_tag makes over (CIGAR, MD) (reference
// src/bam_utils.c:1037-1094): one '=' op per piece of a matching run the reference emits (a run may continue over an insertion), one 1-base
// 'X' op per mismatch letter, the "0" after a mismatch or a deletion skipped exactly where the reference skips it.  Everything else the MD
// variant does (quality flags, the sliding window, clips, the skip test) is what the =/X variant does on those ops.
//
// B200 design: thread per read, two passes (count, fill) around the same exclusive scan K1 uses; an MD string is a few hundred bytes read
// once per pass.  The file compiles for the host as well (tests/emu).
#pragma once
#include <stdint.h>

namespace lcd {
namespace md {

enum { CMATCH = 0, CDEL = 2, CEQUAL = 7, CDIFF = 8 };
enum { MD_OK = 0, MD_MISMATCH = 3, MD_EQX_OP = 4 };          // MD and CIGAR do not match / an =/X op next to an MD tag (the reference exits on both)

__device__ __forceinline__ bool is_digit(char c) { return c >= '0' && c <= '9'; }
__device__ __forceinline__ bool is_alpha(char c) { const char l = c | 0x20; return l >= 'a' && l <= 'z'; }

// emit(op, len) is called for every op of the =/X CIGAR, in order
template <class Emit>
__device__ __forceinline__ int walk(const uint32_t *cg, int nc, const char *md, Emit emit) {
    long long md_i = 0, last_eq = 0;
    for (int i = 0; i < nc; ++i) {
        const int op = cg[i] & 15; long long m = cg[i] >> 4;
        if (op == CMATCH) {
            for (;;) {
                if (last_eq > 0) {                                    // a run that started before this op (it continued over an insertion)
                    if (last_eq >= m) { emit(CEQUAL, m); last_eq -= m; m = 0; }
                    else { emit(CEQUAL, last_eq); m -= last_eq; last_eq = 0; }
                } else if (is_digit(md[md_i])) {
                    long long eq = 0;
                    while (is_digit(md[md_i])) { eq = eq * 10 + (md[md_i] - '0'); ++md_i; }
                    if (eq > m) { last_eq = eq - m; eq = m; }
                    else if (eq == 0) continue;
                    emit(CEQUAL, eq); m -= eq;
                } else if (is_alpha(md[md_i])) {
                    emit(CDIFF, 1); m -= 1;
                    md_i += (md[md_i + 1] == '0') ? 2 : 1;             // the 0 between two mismatches
                } else return MD_MISMATCH;
                if (m <= 0) break;
            }
        } else if (op == CDEL) {
            emit(CDEL, m);
            if (md[md_i] != '^') return MD_MISMATCH;                  // truncated / malformed tag: never step past its NUL
            ++md_i;
            while (md[md_i] && is_alpha(md[md_i])) ++md_i;
            if (md[md_i] == '0') ++md_i;                              // the 0 after a deletion
        } else if (op == CEQUAL || op == CDIFF) return MD_EQX_OP;
        else emit(op, m);
    }
    return MD_OK;
}

struct KernelArgs {
    long long n_reads_total;
    const uint8_t *read_active;
    const int32_t *n_cigar0; const long long *cigar_off0; const uint32_t *cigar0;      // the reads' own CIGARs
    const long long *md_off; const char *md;                                           // md_off[g] < 0: the read's CIGAR is =/X already
    long long *cnt;                                                                    // count pass: ops of the converted CIGAR per read
    const long long *first;                                                            // exclusive scan of cnt
    int32_t *n_cigar; long long *cigar_off; uint32_t *cigar;                           // fill pass: what K1 consumes
    int32_t *status;
};

__device__ void count_read(const KernelArgs &a, long long g) {
    long long n = 0;
    if (a.read_active[g]) {
        const uint32_t *cg = a.cigar0 + a.cigar_off0[g]; const int nc = a.n_cigar0[g];
        if (a.md_off[g] < 0) n = nc;
        else {
            const int st = walk(cg, nc, a.md + a.md_off[g], [&](int, long long) { ++n; });
            if (st) { *a.status = st; n = 0; }
        }
    }
    a.cnt[g] = n;
}

__device__ void fill_read(const KernelArgs &a, long long g) {
    const long long o = a.first[g];
    a.cigar_off[g] = o; a.n_cigar[g] = (int32_t)(a.first[g + 1] - o);
    if (!a.read_active[g] || a.first[g + 1] == o) return;
    const uint32_t *cg = a.cigar0 + a.cigar_off0[g]; const int nc = a.n_cigar0[g];
    uint32_t *out = a.cigar + o;
    if (a.md_off[g] < 0) { for (int i = 0; i < nc; ++i) out[i] = cg[i]; return; }
    long long k = 0;
    walk(cg, nc, a.md + a.md_off[g], [&](int op, long long len) { out[k++] = ((uint32_t)len << 4) | (uint32_t)op; });
}

} // namespace md
} // namespace lcd
