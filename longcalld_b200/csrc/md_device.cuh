// md_device.cuh -- device-side logic of K1's MD-tag front end: reads with plain-M CIGARs and an MD tag (the bundled ONT data) are turned into
// the =/X CIGARs the difference-list kernels consume, by the walk collect_digar_from_MD_tag makes over (CIGAR, MD) (reference
// src/bam_utils.c:1037-1094): one '=' op per piece of a matching run the reference emits (a run may continue over an insertion), one 1-base
// 'X' op per mismatch letter, the "0" after a mismatch or a deletion skipped exactly where the reference skips it.  Everything else the MD
// variant does (quality flags, the sliding window, clips, the skip test) is what the =/X variant does on those ops.
//
// B200 design: thread per read, two passes (count, fill) around the same exclusive scan K1 uses; an MD string is a few hundred bytes read
// once per pass.  The file compiles for the host as well (tests/emu).
//
// The same front end serves the two other variants of the reference's driver (src/collect_var.c:1072-1080):
//   cs tag   (collect_digar_from_cs_tag, src/bam_utils.c:844-1001): the records come from the walk over the cs string -- ":n" / "=seq" runs,
//            "*xy" mismatches, "+seq" insertions, "-seq" deletions, "~..." introns (skipped WITHOUT advancing the position) -- and clips from the
//            FIRST and the LAST CIGAR op only, with a candidate count that differs from the =/X variant's at contig ends (ops CS_SOFT / CS_HARD
//            below).  The reference takes the alt bases from the tag's letters; the kernels take them from SEQ, so the walk checks that the two
//            agree (they do in every valid BAM) and rejects the read loudly otherwise (CS_SEQ_MISMATCH).
//   no tag   (collect_digar_from_ref_seq, :1176-1290): every base of an M / = / X op is compared with the chunk's reference window; bases
//            outside the window are passed over without a record and WITHOUT closing the running match (op SKIP below, emitted before the
//            run it interrupts -- the reference places the run's record by counting back from where it ends).
#pragma once
#include <stdint.h>

namespace lcd {
namespace md {

enum { CMATCH = 0, CINS = 1, CDEL = 2, CREF_SKIP = 3, CSOFT = 4, CHARD = 5, CEQUAL = 7, CDIFF = 8,
       SKIP = 9,                // pseudo-op: pos += len, qi += len, no record (no-tag variant, bases outside the reference window)
       CS_SOFT = 10, CS_HARD = 11 };   // pseudo-ops: a clip of a cs-tagged read (counted as a candidate whenever it is long, wherever it lies)
enum { MD_OK = 0, MD_MISMATCH = 3, MD_EQX_OP = 4,            // MD and CIGAR do not match / an =/X op next to an MD tag (the reference exits on both)
       CS_BAD = 5, CS_SEQ_MISMATCH = 6, TAG_BAD_KIND = 7 };  // malformed cs string / cs letters differ from SEQ / unknown tag kind
enum { KIND_EQX = -1, KIND_MD = 0, KIND_CS = 1, KIND_REFSEQ = 2 };

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

__device__ __forceinline__ int nt4(char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }
__device__ __forceinline__ int seq_code(const uint8_t *bseq, long long qi) {       // seq_nt16_int[bam_seqi(bseq, qi)]
    const int c = (bseq[qi >> 1] >> ((~qi & 1) << 2)) & 15;
    return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
}

// cs tag -> ops (collect_digar_from_cs_tag, src/bam_utils.c:844-1001)
template <class Emit>
__device__ __forceinline__ int walk_cs(const uint32_t *cg, int nc, const char *cs, const uint8_t *bseq, int qlen, Emit emit) {
    if (nc <= 0) return CS_BAD;
    long long qi = 0;
    { const int op = cg[0] & 15; const long long len = cg[0] >> 4; if (op == CSOFT || op == CHARD) { emit(op == CSOFT ? CS_SOFT : CS_HARD, len); if (op == CSOFT) qi += len; } }
    long long i = 0;
    while (cs[i]) {
        const char c = cs[i];
        if (c == ':') {
            long long len = 0; ++i;
            if (!is_digit(cs[i])) return CS_BAD;
            while (is_digit(cs[i])) { len = len * 10 + (cs[i] - '0'); ++i; }
            emit(CEQUAL, len); qi += len;
        } else if (c == '=') {
            long long len = 0; ++i;
            while (is_alpha(cs[i])) { ++len; ++i; }
            emit(CEQUAL, len); qi += len;
        } else if (c == '*') {
            if (!cs[i + 1] || !cs[i + 2]) return CS_BAD;
            if (qi >= qlen || nt4(cs[i + 2]) != seq_code(bseq, qi)) return CS_SEQ_MISMATCH;
            emit(CDIFF, 1); ++qi; i += 3;
        } else if (c == '+') {
            long long len = 0; ++i;
            while (is_alpha(cs[i])) { if (qi + len >= qlen || nt4(cs[i]) != seq_code(bseq, qi + len)) return CS_SEQ_MISMATCH; ++len; ++i; }
            emit(CINS, len); qi += len;
        } else if (c == '-') {
            long long len = 0; ++i;
            while (is_alpha(cs[i])) { ++len; ++i; }
            emit(CDEL, len);
        } else if (c == '~') {                                        // intron: skipped, the position is not advanced (:945-947)
            ++i;
            while (is_alpha(cs[i]) || is_digit(cs[i])) ++i;
        } else return CS_BAD;
    }
    { const int op = cg[nc - 1] & 15; const long long len = cg[nc - 1] >> 4; if (op == CSOFT || op == CHARD) emit(op == CSOFT ? CS_SOFT : CS_HARD, len); }
    return MD_OK;
}

// no tag: the read against the chunk's reference window (collect_digar_from_ref_seq, src/bam_utils.c:1176-1290)
// ref: ASCII bases of positions ref_beg .. ref_end (1-based, inclusive); pos0: the read's 0-based start
template <class Emit>
__device__ __forceinline__ int walk_refseq(const uint32_t *cg, int nc, const char *ref, long long ref_beg, long long ref_end, long long pos0,
                                           const uint8_t *bseq, Emit emit) {
    long long pos = pos0 + 1, qi = 0;
    for (int k = 0; k < nc; ++k) {
        const int op = cg[k] & 15; const long long len = cg[k] >> 4;
        if (op == CMATCH || op == CDIFF || op == CEQUAL) {
            long long eq = 0, out = 0;
            for (long long j = 0; j < len; ++j, ++pos, ++qi) {
                if (pos < ref_beg || pos > ref_end) { ++out; continue; }
                if (nt4(ref[pos - ref_beg]) != seq_code(bseq, qi)) {
                    if (out) { emit(SKIP, out); out = 0; }
                    if (eq) { emit(CEQUAL, eq); eq = 0; }
                    emit(CDIFF, 1);
                } else ++eq;
            }
            if (out) emit(SKIP, out);
            if (eq) emit(CEQUAL, eq);
        } else {
            emit(op, len);
            if (op == CDEL || op == CREF_SKIP) pos += len;
            else if (op == CINS || op == CSOFT) qi += len;
        }
    }
    return MD_OK;
}

struct KernelArgs {
    long long n_reads_total;
    const uint8_t *read_active;
    const int32_t *n_cigar0; const long long *cigar_off0; const uint32_t *cigar0;      // the reads' own CIGARs
    const long long *md_off; const char *md;                                           // md_off[g] < 0: the read's CIGAR is =/X already (or it has no tag)
    const int8_t *kind;                                                                // nullptr: KIND_MD wherever md_off[g] >= 0, else the read's variant
    const int32_t *read_chunk; const long long *ref_off, *ref_beg, *ref_end; const char *ref;    // per chunk: the reference window (KIND_REFSEQ)
    const long long *read_pos0; const int32_t *l_qseq; const long long *seq_off; const uint8_t *bseq;
    long long *cnt;                                                                    // count pass: ops of the converted CIGAR per read
    const long long *first;                                                            // exclusive scan of cnt
    int32_t *n_cigar; long long *cigar_off; uint32_t *cigar;                           // fill pass: what K1 consumes
    long long *rlen;                                                                   // fill pass: reference length of the read's OWN CIGAR (bam_endpos; the op stream may skip introns)
    int32_t *status;
};

__device__ __forceinline__ int read_kind(const KernelArgs &a, long long g) { return a.kind ? a.kind[g] : (a.md_off[g] < 0 ? KIND_EQX : KIND_MD); }

template <class Emit>
__device__ __forceinline__ int walk_read(const KernelArgs &a, long long g, int kind, Emit emit) {
    const uint32_t *cg = a.cigar0 + a.cigar_off0[g]; const int nc = a.n_cigar0[g];
    if (kind == KIND_MD) return walk(cg, nc, a.md + a.md_off[g], emit);
    if (kind == KIND_CS) return walk_cs(cg, nc, a.md + a.md_off[g], a.bseq + a.seq_off[g], a.l_qseq[g], emit);
    if (kind == KIND_REFSEQ) { const int c = a.read_chunk[g]; return walk_refseq(cg, nc, a.ref + a.ref_off[c], a.ref_beg[c], a.ref_end[c], a.read_pos0[g], a.bseq + a.seq_off[g], emit); }
    return TAG_BAD_KIND;
}

__device__ void count_read(const KernelArgs &a, long long g) {
    long long n = 0;
    if (a.read_active[g]) {
        const int kind = read_kind(a, g);
        if (kind == KIND_EQX) n = a.n_cigar0[g];
        else {
            const int st = walk_read(a, g, kind, [&](int, long long) { ++n; });
            if (st) { *a.status = st; n = 0; }
        }
    }
    a.cnt[g] = n;
}

__device__ void fill_read(const KernelArgs &a, long long g) {
    const long long o = a.first[g];
    a.cigar_off[g] = o; a.n_cigar[g] = (int32_t)(a.first[g + 1] - o);
    if (a.rlen) {
        long long rl = 0;
        if (a.read_active[g]) { const uint32_t *c0 = a.cigar0 + a.cigar_off0[g]; for (int k = 0; k < a.n_cigar0[g]; ++k) { const int op = c0[k] & 15; if (op == CMATCH || op == CDEL || op == CREF_SKIP || op == CEQUAL || op == CDIFF) rl += c0[k] >> 4; } }
        a.rlen[g] = rl;
    }
    if (!a.read_active[g] || a.first[g + 1] == o) return;
    const uint32_t *cg = a.cigar0 + a.cigar_off0[g]; const int nc = a.n_cigar0[g];
    uint32_t *out = a.cigar + o;
    const int kind = read_kind(a, g);
    if (kind == KIND_EQX) { for (int i = 0; i < nc; ++i) out[i] = cg[i]; return; }
    long long k = 0;
    walk_read(a, g, kind, [&](int op, long long len) { out[k++] = ((uint32_t)len << 4) | (uint32_t)op; });
}

} // namespace md
} // namespace lcd
