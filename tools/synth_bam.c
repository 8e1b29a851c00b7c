/* tools/synth_bam.c -- MEASUREMENT INFRASTRUCTURE: synthetic long-read BAM + FASTA of the shape BASELINE.json names
 * (SURVEY.md section 8d), so that `longcallD call` itself -- the unmodified reference and the GPU build -- can be timed on the same input.
 *
 *   synth_bam <out_prefix> <ref_mb> <hifi|ont> [seed=11] [coverage=30] [n_contigs=1] [mosaic=0] [style=eqx|m|md|cs]
 *     -> <out_prefix>.fa (+ .fai), <out_prefix>.bam (+ .bai): coordinate-sorted, MAPQ 60, true alignments + NM; style picks which of the
 *        reference's four difference-list variants the reads take (src/collect_var.c:1072-1080): =/X CIGARs (default), plain-M CIGARs
 *        without tags (bases are compared with the reference), plain-M + MD tag, plain-M + cs tag (short form)
 *
 * reference : uniform ACGT; a homopolymer (6-30 bp) every ~0.8 kb; an STR / VNTR (unit 2-60 bp x 3-40 copies) every ~6 kb
 * diploid   : SNPs 1 / 1 000 bp (2/3 het), small indels 1 / 3 000 bp (70 % as copy-number changes of the planted repeats), SV insertions /
 *             deletions of 50 bp - 6 kb every ~300 kb; mosaic=1 adds TE-like inserts (300 bp - 6 kb random sequence + poly-A + 5-20 bp TSD)
 *             carried by 5 % of the reads, one per ~2 Mb
 * reads     : HiFi  length ~ N(15 kb, 3 kb), error 0.2 % (80 % homopolymer-length indels), Q20 - Q40
 *             ONT   length log-normal (N50 ~ 25 kb), error 1.5 % (subs 40 / ins 25 / del 35 %), Q8 - Q30
 *             3 % of the reads carry a soft clip of 100 - 2 000 bp; strand is random
 * Everything is seeded (xorshift64*); built against the reference's own htslib by tools/Makefile. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "htslib/sam.h"
#include "htslib/faidx.h"
#include "htslib/kstring.h"

static uint64_t rs = 88172645463325252ull;
static inline uint64_t rnd(void) { rs ^= rs >> 12; rs ^= rs << 25; rs ^= rs >> 27; return rs * 2685821657736338717ull; }
static inline double urand(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static inline int irand(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }       /* inclusive */
static double nrand(void) { double u = urand(), v = urand(); if (u < 1e-300) u = 1e-300; return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); }
static const char ACGT[] = "ACGT";

typedef struct { int64_t pos; int type; int len; char *alt; int hap; int mosaic; } var_t;    /* type: 0 SNP, 1 INS (after pos-1, before pos), 2 DEL of [pos, pos+len); hap: 1, 2 or 3 (both) */
static var_t *vars; static size_t n_vars, m_vars;
static void add_var(int64_t pos, int type, int len, char *alt, int hap, int mosaic) {
    if (n_vars == m_vars) { m_vars = m_vars ? m_vars * 2 : 1024; vars = (var_t*)realloc(vars, m_vars * sizeof(var_t)); }
    var_t v = { pos, type, len, alt, hap, mosaic }; vars[n_vars++] = v;
}
static int cmp_var(const void *a, const void *b) { const var_t *x = (const var_t*)a, *y = (const var_t*)b; return x->pos < y->pos ? -1 : x->pos > y->pos; }
static char *rand_seq(int n) { char *s = (char*)malloc(n + 1); for (int i = 0; i < n; ++i) s[i] = ACGT[rnd() & 3]; s[n] = 0; return s; }

typedef struct { int64_t pos; int unit, copies; } rep_t;

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: synth_bam <out_prefix> <ref_mb> <hifi|ont> [seed] [coverage] [n_contigs] [mosaic] [eqx|m|md|cs]\n"); return 1; }
    const char *prefix = argv[1]; const double ref_mb = atof(argv[2]); const int ont = strcmp(argv[3], "ont") == 0;
    const uint64_t seed = argc > 4 ? strtoull(argv[4], 0, 10) : 11; const double cov = argc > 5 ? atof(argv[5]) : 30.0;
    const int n_ctg = argc > 6 ? atoi(argv[6]) : 1, mosaic = argc > 7 ? atoi(argv[7]) : 0;
    const int style = argc > 8 ? (strcmp(argv[8], "m") == 0 ? 1 : strcmp(argv[8], "md") == 0 ? 2 : strcmp(argv[8], "cs") == 0 ? 3 : 0) : 0;
    kstring_t md = { 0, 0, NULL }, cs = { 0, 0, NULL }; uint32_t *mcig = NULL; size_t n_mcig = 0, m_mcig = 0;
    rs ^= seed * 0x9E3779B97F4A7C15ull; for (int i = 0; i < 8; ++i) rnd();
    const int64_t L = (int64_t)(ref_mb * 1e6 / n_ctg);
    char fn[4096];
    snprintf(fn, sizeof(fn), "%s.fa", prefix); FILE *fa = fopen(fn, "w");
    snprintf(fn, sizeof(fn), "%s.bam", prefix);
    samFile *out = sam_open(fn, "wb1");
    if (!fa || !out) { fprintf(stderr, "cannot open outputs\n"); return 1; }
    hts_set_threads(out, 4);
    sam_hdr_t *hdr = sam_hdr_init();
    sam_hdr_add_line(hdr, "HD", "VN", "1.6", "SO", "coordinate", NULL);
    for (int c = 0; c < n_ctg; ++c) { char name[32], len[32]; snprintf(name, 32, "chr%d", c + 1); snprintf(len, 32, "%lld", (long long)L); sam_hdr_add_line(hdr, "SQ", "SN", name, "LN", len, NULL); }
    if (sam_hdr_write(out, hdr) < 0) return 1;
    bam1_t *b = bam_init1();
    uint64_t n_reads = 0, n_bases = 0;
    for (int ctg = 0; ctg < n_ctg; ++ctg) {
        /* ---- reference with planted low-complexity sequence */
        char *ref = (char*)malloc(L + 1);
        for (int64_t i = 0; i < L; ++i) ref[i] = ACGT[rnd() & 3];
        ref[L] = 0;
        rep_t *reps = NULL; size_t n_rep = 0, m_rep = 0;
        for (int64_t p = 1000; p < L - 4000; p += irand(250, 750)) {
            if (rnd() % 8 == 0) {                        /* STR / VNTR: ~1 per 6 kb */
                const int unit = (rnd() & 3) ? irand(2, 6) : irand(7, 60), copies = irand(3, unit > 20 ? 12 : 40);
                for (int c = 1; c < copies; ++c) memcpy(ref + p + (int64_t)c * unit, ref + p, unit);
                if (n_rep == m_rep) { m_rep = m_rep ? 2 * m_rep : 256; reps = (rep_t*)realloc(reps, m_rep * sizeof(rep_t)); }
                rep_t r = { p, unit, copies }; reps[n_rep++] = r;
            } else { const int n = irand(6, 30); memset(ref + p, ref[p], n);
                if (n_rep == m_rep) { m_rep = m_rep ? 2 * m_rep : 256; reps = (rep_t*)realloc(reps, m_rep * sizeof(rep_t)); }
                rep_t r = { p, 1, n }; reps[n_rep++] = r; }
        }
        fprintf(fa, ">chr%d\n", ctg + 1);
        for (int64_t i = 0; i < L; i += 60) { fwrite(ref + i, 1, (size_t)(L - i < 60 ? L - i : 60), fa); fputc('\n', fa); }
        /* ---- diploid truth */
        n_vars = 0;
        for (int64_t p = 500 + irand(0, 1000); p < L - 500; p += irand(500, 1500)) {                 /* SNPs */
            char *a = (char*)malloc(2); do { a[0] = ACGT[rnd() & 3]; } while (a[0] == ref[p]); a[1] = 0;
            const int h = rnd() % 3; add_var(p, 0, 1, a, h == 0 ? 3 : h, 0);
        }
        for (int64_t p = 2000 + irand(0, 4000); p < L - 8000; p += irand(800, 2400)) {                /* small indels */
            const int hap = (rnd() % 3 == 0) ? 3 : 1 + (int)(rnd() & 1);
            if (rnd() % 10 < 7 && n_rep) {                /* a copy-number change of the nearest planted repeat */
                size_t lo = 0, hi = n_rep; while (lo < hi) { size_t m = (lo + hi) / 2; if (reps[m].pos < p) lo = m + 1; else hi = m; }
                const rep_t *r = reps + (lo < n_rep ? lo : n_rep - 1);
                const int k = r->unit > 10 ? 1 : irand(1, 3), len = r->unit * k;
                if (len * (k + 1) >= r->unit * r->copies) continue;
                if (rnd() & 1) { char *a = (char*)malloc(len + 1); memcpy(a, ref + r->pos, len); a[len] = 0; add_var(r->pos + r->unit, 1, len, a, hap, 0); }
                else add_var(r->pos + r->unit, 2, len, NULL, hap, 0);
            } else { const int len = irand(1, 12); if (rnd() & 1) add_var(p, 1, len, rand_seq(len), hap, 0); else add_var(p, 2, len, NULL, hap, 0); }
        }
        for (int64_t p = 100000 + irand(0, 100000); p < L - 50000; p += irand(200000, 400000)) {      /* SVs */
            const int len = (rnd() & 3) ? irand(50, 600) : irand(600, 6000), hap = (rnd() % 4 == 0) ? 3 : 1 + (int)(rnd() & 1);
            if (rnd() & 1) add_var(p, 1, len, rand_seq(len), hap, 0); else add_var(p, 2, len, NULL, hap, 0);
        }
        if (mosaic) for (int64_t p = 700000 + irand(0, 600000); p < L - 50000; p += irand(1500000, 2500000)) {   /* TE-like mosaic inserts */
            const int body = (rnd() & 1) ? irand(280, 320) : irand(1000, 6000), pa = irand(15, 40), tsd = irand(5, 20), len = body + pa + tsd;
            char *a = rand_seq(len); memset(a + body, 'A', pa); memcpy(a + body + pa, ref + p - tsd, tsd);
            add_var(p, 1, len, a, 1 + (int)(rnd() & 1), 1);
        }
        qsort(vars, n_vars, sizeof(var_t), cmp_var);
        /* drop variants that touch the one before them */
        { size_t k = 0; int64_t last_end = -1;
          for (size_t i = 0; i < n_vars; ++i) { const int64_t e = vars[i].pos + (vars[i].type == 2 ? vars[i].len : 1);
              if (vars[i].pos <= last_end + 1) { free(vars[i].alt); continue; } vars[k++] = vars[i]; last_end = e; }
          n_vars = k; }
        /* ---- reads, by increasing start */
        const double mean_len = ont ? 20000.0 : 15000.0, start_rate = cov / mean_len;          /* reads starting per reference base */
        const double err = ont ? 0.015 : 0.002;
        size_t v0 = 0;
        kstring_t seq = { 0, 0, NULL }, qual = { 0, 0, NULL }; uint32_t *cig = NULL; size_t n_cig = 0, m_cig = 0;
#define PUSH(op, ln) do { if ((ln) > 0) { if (n_cig && (cig[n_cig - 1] & 15) == (uint32_t)(op)) cig[n_cig - 1] += (uint32_t)(ln) << 4; else { if (n_cig == m_cig) { m_cig = m_cig ? 2 * m_cig : 1024; cig = (uint32_t*)realloc(cig, m_cig * 4); } cig[n_cig++] = ((uint32_t)(ln) << 4) | (op); } } } while (0)
#define EMIT(base, q) do { kputc((base), &seq); kputc((char)(q), &qual); } while (0)
        for (int64_t s = 0; s < L - 1000; ++s) {
            if (urand() >= start_rate) continue;
            int len = ont ? (int)exp(9.6 + 0.75 * nrand()) : (int)(15000 + 3000 * nrand());
            if (len < 1000) len = 1000; if (len > 200000) len = 200000;
            const int64_t e = s + len < L ? s + len : L;
            const int hap = 1 + (int)(rnd() & 1); const int carries_mosaic = urand() < 0.10;       /* 10 % of one haplotype's reads = 5 % allele fraction */
            seq.l = qual.l = 0; n_cig = 0; int nm = 0;
            if (rnd() % 100 < 3) { const int c = irand(100, 2000); for (int i = 0; i < c; ++i) EMIT(ACGT[rnd() & 3], ont ? irand(8, 30) : irand(20, 40)); PUSH(BAM_CSOFT_CLIP, c); }
            while (v0 < n_vars && vars[v0].pos < s) ++v0;
            size_t v = v0; int64_t p = s;
            const int qlo = ont ? 8 : 20, qhi = ont ? 30 : 40;
            while (p < e) {
                if (v < n_vars && vars[v].pos == p && (vars[v].hap & hap) && (!vars[v].mosaic || carries_mosaic) && p > s) {
                    const var_t *x = vars + v++;
                    if (x->type == 0) { EMIT(x->alt[0], irand(qlo, qhi)); PUSH(BAM_CDIFF, 1); nm++; p++; continue; }
                    if (x->type == 1) { for (int i = 0; i < x->len; ++i) EMIT(x->alt[i], irand(qlo, qhi)); PUSH(BAM_CINS, x->len); nm += x->len; continue; }
                    if (p + x->len < e) { PUSH(BAM_CDEL, x->len); nm += x->len; p += x->len; continue; }
                } else if (v < n_vars && vars[v].pos == p) ++v;
                if (urand() < err && p > s + 5 && p < e - 5) {                       /* sequencing error */
                    const double u = urand();
                    if (!ont) {                       /* HiFi: 80 % homopolymer-length indels */
                        if (u < 0.4) { EMIT(ref[p], irand(8, 20)); EMIT(ref[p], irand(8, 20)); PUSH(BAM_CEQUAL, 1); PUSH(BAM_CINS, 1); nm++; p++; continue; }
                        if (u < 0.8) { PUSH(BAM_CDEL, 1); nm++; p++; continue; }
                    } else {
                        if (u < 0.25) { EMIT(ACGT[rnd() & 3], irand(5, 15)); PUSH(BAM_CINS, 1); nm++; continue; }
                        if (u < 0.60) { PUSH(BAM_CDEL, 1); nm++; p++; continue; }
                    }
                    char c; do { c = ACGT[rnd() & 3]; } while (c == ref[p]);
                    EMIT(c, irand(5, 15)); PUSH(BAM_CDIFF, 1); nm++; p++; continue;
                }
                EMIT(ref[p], irand(qlo, qhi)); PUSH(BAM_CEQUAL, 1); p++;
            }
            if (n_cig == 0 || seq.l == 0) continue;
            { const int op = cig[n_cig - 1] & 15; if (op == BAM_CDEL) n_cig--; }                  /* an alignment does not end in a deletion */
            char name[64]; snprintf(name, sizeof(name), "r%d_%llu", ctg + 1, (unsigned long long)n_reads);
            const uint32_t *wcig = cig; size_t n_wcig = n_cig;
            if (style) {                   /* the same alignment as a plain-M CIGAR, with its MD and (short-form) cs strings */
                md.l = cs.l = 0; n_mcig = 0;
                int64_t rp = s; size_t qi = 0; long md_run = 0, cs_run = 0; uint32_t m = 0;
#define MPUSH(op, ln) do { if (n_mcig == m_mcig) { m_mcig = m_mcig ? 2 * m_mcig : 1024; mcig = (uint32_t*)realloc(mcig, m_mcig * 4); } mcig[n_mcig++] = ((uint32_t)(ln) << 4) | (op); } while (0)
#define MFLUSH() do { if (m) { MPUSH(BAM_CMATCH, m); m = 0; } } while (0)
#define CSFLUSH() do { if (cs_run) { ksprintf(&cs, ":%ld", cs_run); cs_run = 0; } } while (0)
                for (size_t k = 0; k < n_cig; ++k) {
                    const int op = cig[k] & 15; const uint32_t ln = cig[k] >> 4;
                    if (op == BAM_CEQUAL) { m += ln; md_run += ln; cs_run += ln; rp += ln; qi += ln; }
                    else if (op == BAM_CDIFF) {
                        for (uint32_t j = 0; j < ln; ++j) { ksprintf(&md, "%ld%c", md_run, ref[rp]); md_run = 0; CSFLUSH(); ksprintf(&cs, "*%c%c", ref[rp] | 0x20, seq.s[qi] | 0x20); ++rp; ++qi; }
                        m += ln;
                    } else if (op == BAM_CINS) { MFLUSH(); MPUSH(BAM_CINS, ln); CSFLUSH(); kputc('+', &cs); for (uint32_t j = 0; j < ln; ++j) kputc(seq.s[qi + j] | 0x20, &cs); qi += ln; }
                    else if (op == BAM_CDEL) {
                        MFLUSH(); MPUSH(BAM_CDEL, ln); ksprintf(&md, "%ld^", md_run); md_run = 0; CSFLUSH(); kputc('-', &cs);
                        for (uint32_t j = 0; j < ln; ++j) { kputc(ref[rp + j], &md); kputc(ref[rp + j] | 0x20, &cs); }
                        rp += ln;
                    } else { MFLUSH(); MPUSH(op, ln); if (op == BAM_CSOFT_CLIP) qi += ln; }
                }
                MFLUSH(); CSFLUSH(); ksprintf(&md, "%ld", md_run);
                wcig = mcig; n_wcig = n_mcig;
            }
            if (bam_set1(b, strlen(name), name, (rnd() & 1) ? BAM_FREVERSE : 0, ctg, s, 60, n_wcig, wcig, -1, -1, 0, seq.l, seq.s, qual.s, 16 + (style == 2 ? md.l + 8 : 0) + (style == 3 ? cs.l + 8 : 0)) < 0) return 1;
            bam_aux_update_int(b, "NM", nm);
            if (style == 2 && bam_aux_append(b, "MD", 'Z', (int)md.l + 1, (const uint8_t*)md.s) < 0) return 1;
            if (style == 3 && bam_aux_append(b, "cs", 'Z', (int)cs.l + 1, (const uint8_t*)cs.s) < 0) return 1;
            if (sam_write1(out, hdr, b) < 0) return 1;
            n_reads++; n_bases += seq.l;
        }
        free(seq.s); free(qual.s); free(cig); free(ref); free(reps);
        for (size_t i = 0; i < n_vars; ++i) free(vars[i].alt);
    }
    bam_destroy1(b); sam_hdr_destroy(hdr);
    if (sam_close(out) < 0) return 1;
    fclose(fa);
    snprintf(fn, sizeof(fn), "%s.fa", prefix); if (fai_build(fn) < 0) return 1;
    snprintf(fn, sizeof(fn), "%s.bam", prefix); if (sam_index_build(fn, 0) < 0) return 1;
    fprintf(stderr, "[synth_bam] %s: %d contig(s) x %lld bp, %llu reads, %llu bases (%.1fx)\n", prefix, n_ctg, (long long)L, (unsigned long long)n_reads,
            (unsigned long long)n_bases, (double)n_bases / ((double)L * n_ctg));
    return 0;
}
