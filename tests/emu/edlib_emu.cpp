// tests/emu/edlib_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K7 device logic
// (longcalld_b200/csrc/edlib_device.cuh, one thread per problem) on the host.  Same signature as the
// oracle's lcd_oracle_edlib_align so the tests can diff the two.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/edlib_device.cuh"
#include <stdlib.h>
#include <vector>

using namespace lcd::edlib;

extern "C" int emu_edlib_align(const uint8_t *query, int qlen, const uint8_t *target, int tlen,
                               int mode, int want_path, uint8_t *aln, lcd_edlib_result_t *res) {
    std::vector<uint8_t> seqs((size_t)qlen + tlen + 16);
    memcpy(seqs.data(), query, qlen); memcpy(seqs.data() + qlen, target, tlen);
    Problem p; memset(&p, 0, sizeof(p));
    p.q_off = 0; p.t_off = (uint64_t)qlen; p.qlen = qlen; p.tlen = tlen; p.mode = mode; p.want_path = want_path;
    p.ws_words = workspace_words(qlen, tlen, want_path);
    // exactly the words the host plan hands out, poisoned, with a guard zone behind them
    std::vector<Word> ws(p.ws_words + 64);
    for (size_t i = 0; i < ws.size(); ++i) ws[i] = 0xa5a5a5a5deadbeefull ^ (i * 0x9e3779b97f4a7c15ull);
    std::vector<Word> guard(ws.end() - 64, ws.end());
    std::vector<uint8_t> out((size_t)qlen + tlen + 16, 0xee);
    DevResult r; memset(&r, 0, sizeof(r));
    Aligner al;
    al.align(p, seqs.data(), out.data(), ws.data(), &r);
    for (int i = 0; i < 64; ++i) if (ws[p.ws_words + i] != guard[i]) return -9;      // wrote past its workspace
    res->status = r.status; res->edit_distance = r.edit_distance; res->start_loc = r.start_loc; res->end_loc = r.end_loc; res->aln_len = r.aln_len;
    if (aln && r.aln_len > 0) memcpy(aln, out.data(), r.aln_len);
    return 0;
}
