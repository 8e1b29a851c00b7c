// tests/emu/classify_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K2b device logic (longcalld_b200/csrc/classify_device.cuh, one thread
// per site) on the host over one chunk.  Same signature as the oracle's lcd_oracle_classify_sites.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/classify_device.cuh"
#include <vector>
#include <math.h>

using namespace lcd::classify;

extern "C" int emu_classify_sites(const lcd_classify_input_t *in, int32_t *var_cate) {
    static double lg[LGAMMA_MAX_I + 1]; static bool lg_ready = false;
    if (!lg_ready) { for (int i = 0; i <= LGAMMA_MAX_I; ++i) lg[i] = lgamma((double)i); lg_ready = true; }
    Chunk c; memset(&c, 0, sizeof(c));
    c.min_dp = in->min_dp; c.min_alt_dp = in->min_alt_dp; c.max_xgaps = in->max_xgaps; c.is_ont = in->is_ont; c.min_af = in->min_af; c.max_af = in->max_af;
    c.ref_beg = in->ref_beg; c.ref_end = in->ref_end; c.ref_off = 0; c.alt_base = 0;
    std::vector<int32_t> site_chunk(in->n_sites + 1, 0);
    std::vector<int32_t> counts((size_t)8 * (in->n_sites + 1));                 // (16-byte aligned copy: the kernel reads int4 records)
    memcpy(counts.data(), in->site_counts, sizeof(int32_t) * 8 * (size_t)in->n_sites);
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_sites_total = in->n_sites; a.site_chunk = site_chunk.data();
    a.site_pos = (const long long *)in->site_pos; a.site_type = in->site_type; a.site_ref_len = in->site_ref_len; a.site_alt_len = in->site_alt_len;
    a.site_alt_off = (const long long *)in->site_alt_off; a.site_alt = in->site_alt; a.site_counts = counts.data(); a.ref = in->ref_seq; a.var_cate = var_cate; a.lgamma_cache = lg;
    for (long long s = 0; s < in->n_sites; ++s) classify_site(a, s);
    return 0;
}
