// tests/emu/noisyreg_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K2c device logic (longcalld_b200/csrc/noisyreg_device.cuh, one CTA per
// chunk) on the host with a "CTA" of one thread over one chunk.  Same signature as the oracle's lcd_oracle_noisy_regs.
#include "cuda_emu.h"
#include "../../include/lcd_gpu.h"
#include "../../longcalld_b200/csrc/noisyreg_device.cuh"
#include <vector>

using namespace lcd::noisyreg;
struct NoSync { void operator()() const {} };

// chained: 0 = the interval count in the chunk struct; 1 = read through n_low_dev as when K2c is chained to a K0 plan (the struct's own count poisoned);
// 2 = the same with K0 reporting a failure for the chunk
static int emu_run(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out, int chained) {
    Chunk c; memset(&c, 0, sizeof(c));
    const size_t cap = (size_t)in->n_cnreg + in->n_sites + 8;
    c.reg_beg = in->reg_beg; c.reg_end = in->reg_end; c.min_af = in->min_af; c.min_alt_dp = in->min_alt_dp; c.flank = in->noisy_reg_flank_len; c.is_ont = in->is_ont;
    c.n_sites = in->n_sites; c.n_reads = in->n_reads; c.n_cnreg = (int)in->n_cnreg; c.n_low = (int)in->n_low; c.cap = (int)cap;
    c.site_pos = (const long long *)in->site_pos; c.site_type = in->site_type; c.site_ref_len = in->site_ref_len; c.var_cate_in = in->var_cate;
    c.cn_beg = (const long long *)in->cnreg_beg; c.cn_end = (const long long *)in->cnreg_end; c.cn_label = in->cnreg_label;
    c.low_beg = (const long long *)in->low_beg; c.low_end = (const long long *)in->low_end;
    c.is_skipped = in->is_skipped; c.read_beg = (const long long *)in->read_beg; c.read_end = (const long long *)in->read_end; c.digar_first = (const long long *)in->digar_first;
    c.n_digar = in->n_digar; c.digar_pos = (const long long *)in->digar_pos; c.digar_type = (const signed char *)in->digar_type; c.digar_len = in->digar_len;
    c.nreg_first = (const long long *)in->nreg_first; c.n_nreg = in->n_nreg; c.nreg_beg = (const long long *)in->nreg_beg; c.nreg_end = (const long long *)in->nreg_end;
    std::vector<int> sc(6 * cap + 4 * cap + (size_t)in->n_low + in->n_sites + 64, 0x55555555);      // poisoned scratch
    std::vector<long long> ob(3 * cap + 2);
    long long nregs = 0; int status = 0;
    c.var_cate = out->var_cate; c.keep = out->keep; c.out_regs = ob.data(); c.reg_cap = (long long)cap; c.n_regs = &nregs; c.status = &status;
    int *p = sc.data();
    c.A.st = p; c.A.en = p + cap; c.A.label = p + 2 * cap; c.B.st = p + 3 * cap; c.B.en = p + 4 * cap; c.B.label = p + 5 * cap; p += 6 * cap;
    c.tot = p; c.noi = p + cap; p += 2 * cap; c.low_pmax = p; p += in->n_low + 1; c.vp_pmax = p; p += in->n_sites + 1; c.ctr = p;
    const long long n_low_dev = in->n_low; const int low_status = chained == 2 ? -5 : 0;
    if (chained) { c.n_low = -12345; c.n_low_dev = &n_low_dev; c.low_status = &low_status; }
    run_chunk(c, 0, 1, NoSync());
    if (status) return status;
    out->n_regs = nregs;
    if (nregs > out->reg_cap) return -5;
    for (long long k = 0; k < nregs; ++k) { out->reg_beg[k] = ob[k]; out->reg_end[k] = ob[nregs + k]; out->reg_label[k] = reinterpret_cast<const int *>(ob.data() + 2 * nregs)[k]; }
    return 0;
}
extern "C" int emu_noisy_regs(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) { return emu_run(in, out, 0); }
extern "C" int emu_noisy_regs_chained(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) { return emu_run(in, out, 1); }
extern "C" int emu_noisy_regs_k0_failed(const lcd_noisyreg_input_t *in, lcd_noisyreg_output_t *out) { return emu_run(in, out, 2); }
