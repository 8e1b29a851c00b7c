// tests/emu/wfa_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K6 device logic
// (longcalld_b200/csrc/wfa_device.cuh) single-lane on the host.  Same signature as the oracle's
// lcd_oracle_wfa_align so the tests can diff the two.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/wfa_device.cuh"
#include <stdlib.h>
#include <vector>

using namespace lcd::wfa;

extern "C" int emu_wfa_align(const uint8_t *pattern, int plen, const uint8_t *text, int tlen,
                             const lcd_wfa_params_t *par, char *ops, lcd_wfa_result_t *res,
                             int arena_kib /* small private arena to exercise the overflow path */) {
    const size_t pb = ((size_t)plen + 12 + 15) & ~(size_t)15, tb = ((size_t)tlen + 12 + 15) & ~(size_t)15;
    std::vector<uint8_t> seqs(pb + tb + 16);
    memcpy(seqs.data(), pattern, plen); memset(seqs.data() + plen, '!', pb - plen);
    memcpy(seqs.data() + pb, text, tlen); memset(seqs.data() + pb + tlen, '?', tb - tlen);
    Problem p; memset(&p, 0, sizeof(p));
    p.pat = 0; p.txt = pb; p.ops = 0; p.plen = plen; p.tlen = tlen; p.par = *par;
    p.s_cap = 8 * (plen + tlen) + 256;
    std::vector<char> opsbuf(2 * (plen + tlen) + 32);
    const size_t pool_words = (size_t)96 << 20;        // 384 MiB
    static int32_t *pool = (int32_t *)malloc(pool_words * 4);
    // poison the part of the arena a small problem can touch: the kernel must never depend on
    // what a previous problem left behind
    { static uint32_t x = 0x9e3779b9u; const size_t nw = std::min<size_t>(pool_words, (size_t)2 << 20);
      for (size_t i = 0; i < nw; ++i) { x = x * 1664525u + 1013904223u; pool[i] = (int32_t)(x >> 3) - (1 << 27); } }
    std::vector<WfSet> meta(p.s_cap);
    memset(meta.data(), 0x5a, sizeof(WfSet) * meta.size());
    static uint32_t bitmap[64];
    for (int i = 0; i < 64; ++i) if (bitmap[i]) return -9;   /* a previous problem leaked chunks */
    uint32_t queue = 0; int32_t order = 0;
    DevResult dr; memset(&dr, 0, sizeof(dr));
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.problems = &p; a.order = &order; a.n = 1; a.queue = &queue; a.seqs = seqs.data(); a.ops = opsbuf.data();
    a.results = &dr; a.pool = pool; a.meta = meta.data(); a.meta_cap = p.s_cap;
    a.arena_base = 0; a.arena_units = (uint32_t)arena_kib * 64;
    a.overflow_base = a.arena_units;
    a.n_chunks = (uint32_t)((pool_words / 4 - a.arena_units) / OVERFLOW_CHUNK_UNITS);
    a.chunk_bitmap = bitmap;
    WfSet ring[RING]; int red[64];
    std::vector<uint8_t> seq_smem(WARP_SEQ_SMEM);
    Aligner<1> al(a);
    al.g.lane = 0; al.g.ring = ring; al.g.red = red; al.g.phase = 0;
    al.pool = pool; al.gmeta = meta.data();
    al.align(p, &dr, seq_smem.data(), WARP_SEQ_SMEM, a.arena_base, a.arena_base + a.arena_units, 0);
    res->status = dr.status; res->score = dr.score; res->n_ops = dr.n_ops; res->end_v = dr.end_v; res->end_h = dr.end_h;
    if (ops && dr.status >= 0) { memcpy(ops, opsbuf.data() + dr.ops_begin, dr.n_ops); ops[dr.n_ops] = 0; }
    return 0;
}
