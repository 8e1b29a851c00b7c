// tests/emu/poa_simt_emu.cpp -- TEST INFRASTRUCTURE: runs the WARP policy of the product's K5 device logic
// (longcalld_b200/csrc/poa_device.cuh: strip rows, packed chain rows, ballot backtrack, parallel fusion) on the
// host under the one-warp SIMT emulator (simt_emu.h).  Same signature as the oracle's lcd_oracle_poa.
#include "simt_emu.h"
#include "../../longcalld_b200/csrc/poa_device.cuh"
#include <vector>
using namespace lcd::poa;

static int mode = 0;     /* bit1: tight workspace budgets */
extern "C" void emu_poa_mode(int m) { mode = m; }
static uint8_t *g_read_clu = nullptr; static int g_min_w = 0; static DevResult g_last;      /* the 2-consensus form (emu_poa_ncons) */
extern "C" long emu_poa_stat(int i) { return simt_stat[i]; }
static const int32_t *g_sub_beg = nullptr, *g_sub_end = nullptr;
extern "C" int emu_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                       const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
                       uint8_t *msa, int32_t *msa_len, int32_t msa_cap);
// partially covering reads: same signature as the oracle's lcd_oracle_poa_sub
extern "C" int emu_poa_sub(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const int32_t *sub_beg, const int32_t *sub_end,
                           const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    g_sub_beg = sub_beg; g_sub_end = sub_end;
    const int rc = emu_poa(n_seq, seqs, seq_off, seq_len, p, cons, cons_len, msa, msa_len, msa_cap);
    g_sub_beg = g_sub_end = nullptr;
    return rc;
}
extern "C" int emu_poa(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                       const lcd_poa_params_t *p, uint8_t *cons, int32_t *cons_len,
                       uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    Problem pb; memset(&pb, 0, sizeof(pb));
    pb.seq_base = 0; pb.read_first = 0; pb.n_reads = n_seq; pb.par = *p; pb.cons_off = 0;
    for (int i = 0; i < n_seq; ++i) { pb.sum_len += seq_len[i]; if (seq_len[i] > pb.max_len) pb.max_len = seq_len[i]; }
    const uint64_t words = (uint64_t)48 << 20;            // 192 MiB arena
    static int32_t *arena = (int32_t *)malloc(words * 4);
    { static uint32_t x = 12345; for (size_t i = 0; i < ((size_t)4 << 20); ++i) { x = x * 1664525u + 1013904223u; arena[i] = (int32_t)x; } }   // poison
    std::vector<uint8_t> msa_pool((size_t)msa_cap + 64);
    unsigned long long msa_used = 0; uint32_t queue = 0; int32_t order = 0;
    DevResult dr; memset(&dr, 0, sizeof(dr));
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.problems = &pb; a.order = &order; a.n = 1; a.queue = &queue; a.seqs = seqs; a.read_off = seq_off; a.read_len = seq_len;
    a.cons = cons; a.msa = msa_pool.data(); a.msa_cap = (unsigned long long)msa_cap; a.msa_used = &msa_used;
    a.results = &dr; a.arena = arena; a.arena_words = words; a.read_clu = g_read_clu; pb.min_w = g_min_w; a.sub_beg = g_sub_beg; a.sub_end = g_sub_end;
    pb.node_cap = pb.sum_len + 34; pb.edge_cap = 3 * (pb.sum_len + n_seq) + 64;
    if (mode & 2) { pb.node_cap = 2 * pb.max_len + 64; pb.edge_cap = 3 * pb.node_cap; }
    static __attribute__((aligned(16))) int16_t cache[POA_SMEM_HALFS];
    simt::run_warp([&] {
        Poa<WarpLanes> poa;
        poa.row_cache = cache;
        poa.run(a, pb, &dr, arena);
    });
    *cons_len = dr.cons_len; *msa_len = 0;
    if (dr.status == ST_OK && msa) { *msa_len = dr.msa_len; memcpy(msa, msa_pool.data() + dr.msa_off, (size_t)(n_seq + (dr.n_cons == 2 ? 2 : 1)) * dr.msa_len); }
    g_last = dr;
    return dr.status;
}

// abpoa_aln_msa_cons with max_n_cons consensus sequences: same signature as the oracle's lcd_oracle_poa_ncons
#include <math.h>
extern "C" int emu_poa_ncons(int n_seq, const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len, const lcd_poa_params_t *p, double min_freq,
                             uint8_t *cons, int32_t *cons_len, int32_t *n_cons, uint8_t *read_clu, uint8_t *msa, int32_t *msa_len, int32_t msa_cap) {
    const int cw = (int)ceil(n_seq * min_freq);
    g_read_clu = read_clu; g_min_w = cw > 2 ? cw : 2;
    memset(read_clu, 0, n_seq);
    const int rc = emu_poa(n_seq, seqs, seq_off, seq_len, p, cons, cons_len, msa, msa_len, msa_cap);
    g_read_clu = nullptr; g_min_w = 0;
    *n_cons = rc == 0 ? g_last.n_cons : 0; cons_len[1] = rc == 0 ? g_last.cons_len2 : 0;
    return rc;
}
