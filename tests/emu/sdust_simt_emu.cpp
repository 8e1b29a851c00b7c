// tests/emu/sdust_simt_emu.cpp -- TEST INFRASTRUCTURE: the product's K0 device logic (longcalld_b200/csrc/sdust_device.cuh) under the one-warp SIMT emulator
// (32 fibers, barriers are rendezvous): the block decomposition, the racing segment appends and the per-segment replays run as 32 cooperating threads.
#include "simt_emu.h"
#include "../../longcalld_b200/csrc/sdust_device.cuh"
#include <vector>

using namespace lcd::sdust;
struct WarpSync { void operator()() const { __syncwarp(); } };

extern "C" int emu_sdust(const uint8_t *seq, int l_seq, int T, int W, int64_t *beg, int64_t *end, int64_t cap) {
    Chunk c; memset(&c, 0, sizeof(c));
    const int seg_cap = l_seq / (W + 20) + 256;
    std::vector<int> pv(l_seq + 1, 0x55555555), ss(seg_cap, 0x55555555), sc(seg_cap, 0x55555555), so(seg_cap, 0x55555555), ctr(4, 0x55555555);
    std::vector<unsigned char> tr(l_seq + 1, 0x55);
    std::vector<long long> ob(cap + 1), oe(cap + 1);
    long long n_out = 0; int status = 0;
    c.seq = (const char *)seq; c.n = l_seq; c.T = T; c.W = W; c.base = 0;
    c.out_beg = ob.data(); c.out_end = oe.data(); c.cap = cap; c.n_out = &n_out; c.status = &status;
    c.prevvalid = pv.data(); c.trig = tr.data(); c.seg_start = ss.data(); c.seg_cnt = sc.data(); c.seg_off = so.data(); c.seg_cap = seg_cap; c.ctr = ctr.data();
    std::vector<long long> sb((size_t)l_seq / 4 + (size_t)seg_cap * (W / 4 + 3) + 16, 0x5555555555555555LL), se(sb.size(), 0x5555555555555555LL);
    c.stage_beg = sb.data(); c.stage_end = se.data(); c.stage_cap = (long long)sb.size();
    simt::run_warp([&] { run_chunk(c, (int)threadIdx.x, 32, WarpSync()); });
    if (status) return status;
    for (long long k = 0; k < n_out; ++k) { beg[k] = ob[k]; end[k] = oe[k]; }
    return (int)n_out;
}
