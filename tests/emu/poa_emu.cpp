// tests/emu/poa_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K5 device logic
// (longcalld_b200/csrc/poa_device.cuh) on the host.  A "warp" is a 32-element array (HostLanes), so
// the lane-level vector semantics (shuffle scans, band masks) are exercised exactly as on the GPU.
// Same signature as the oracle's lcd_oracle_poa so the tests can diff the two.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/poa_device.cuh"
#include <stdlib.h>
#include <vector>

namespace lcd { namespace poa {
struct HostLanes {
    struct vec { int v[32]; };
    static constexpr int NT = 1, NW = 1;
    static constexpr bool TWO_PHASE = false;
    static int lane() { return 0; }
    static int tid() { return 0; }
    static int warp() { return 0; }
    static void sync() {}
    static vec load(const int16_t *p) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = p[l]; return r; }
    static vec load_m1(const int16_t *p, int first) { vec r; r.v[0] = first; for (int l = 1; l < 32; ++l) r.v[l] = p[l - 1]; return r; }
    static void store(int16_t *p, vec x) { for (int l = 0; l < 32; ++l) p[l] = (int16_t)x.v[l]; }
    static vec set1(int x) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = x; return r; }
    static vec add(vec a, vec b) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = (int16_t)(a.v[l] + b.v[l]); return r; }
    static vec sub(vec a, vec b) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = (int16_t)(a.v[l] - b.v[l]); return r; }
    static vec vmax(vec a, vec b) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = a.v[l] > b.v[l] ? a.v[l] : b.v[l]; return r; }
    static vec shift_up(vec x, int n, int fill) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = l < n ? fill : x.v[l - n]; return r; }
    static int lane_value(vec x, int l) { return x.v[l]; }
    static vec keep(vec x, int lo, int hi, int fill) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = (l >= lo && l <= hi) ? x.v[l] : fill; return r; }
    template <class F> static vec map_cols(int col0, F f) { vec r; for (int l = 0; l < 32; ++l) r.v[l] = f(col0 + l); return r; }
    static bool row_max(vec x, int lo, int hi, int &m, int &first, int &last) {
        if (lo > hi) return false;
        m = INT32_MIN; for (int l = lo; l <= hi; ++l) if (x.v[l] > m) m = x.v[l];
        first = -1; for (int l = lo; l <= hi; ++l) if (x.v[l] == m) { if (first < 0) first = l; last = l; }
        return true;
    }
};
struct HostLanes2 : HostLanes { static constexpr bool TWO_PHASE = true; };   // exercises the two-phase (CTA) row code
}}
using namespace lcd::poa;

static int mode = 0;     /* bit0: thread-per-problem lane policy (packed int16x2); bit1: tight workspace budgets; bit2: on-chip previous-row cache; bit3: two-phase rows (the CTA-per-problem code path) */
extern "C" void emu_poa_mode(int m) { mode = m; }
static uint8_t *g_read_clu = nullptr; static int g_min_w = 0; static DevResult g_last;      /* the 2-consensus form (emu_poa_ncons) */
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
    a.results = &dr; a.arena = arena; a.arena_words = words; a.read_clu = g_read_clu; pb.min_w = g_min_w;
    pb.node_cap = pb.sum_len + 34; pb.edge_cap = 3 * (pb.sum_len + n_seq) + 64;
    if (mode & 2) { pb.node_cap = 2 * pb.max_len + 64; pb.edge_cap = 3 * pb.node_cap; }   /* tight first-attempt budgets */
    if (mode & 1) { Poa<ThreadLanes> poa; poa.run(a, pb, &dr, arena); }
    else if (mode & 8) { Poa<HostLanes2> poa; static int gsbuf[Poa<HostLanes2>::GS_INTS]; poa.gs = gsbuf; poa.run(a, pb, &dr, arena); }
    else { Poa<HostLanes> poa; static int16_t cache[2 * 3 * 8 * 32]; poa.row_cache = (mode & 4) ? cache : nullptr; poa.run(a, pb, &dr, arena); }
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
