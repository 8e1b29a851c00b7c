// tests/emu/phase_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K4 device logic (longcalld_b200/csrc/phase_device.cuh)
// on the host as a one-thread CTA, fed by the same host marshalling as the GPU plan would do (cgranges order, running
// maximum of the span ends).  Same signature as the oracle's lcd_oracle_assign_hap so the tests can diff the two.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/phase_device.cuh"
#include <vector>
#include <algorithm>

using namespace lcd::phase;

extern "C" void lcd_oracle_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out);   // checker side: the order only

extern "C" int emu_assign_hap(const lcd_phase_input_t *in, lcd_phase_output_t *out) {
    const int nr = in->n_reads, nv = in->n_vars;
    Chunk c; memset(&c, 0, sizeof(c));
    c.n_reads = nr; c.n_vars = nv; c.target = in->target_var_cate; c.is_ont = in->is_ont; c.read_off = 0; c.var_off = 0;
    std::vector<int32_t> st, lb, order(nr + 1), pmax(nr + 1);
    for (int i = 0; i < nr; ++i) {
        const int r = in->ordered_read_ids[i];
        if (in->is_skipped[r] || in->prof_start[r] < 0 || in->prof_end[r] < 0) continue;
        st.push_back(in->prof_start[r]); lb.push_back(r);
    }
    c.n_cr = (int)st.size();
    lcd_oracle_cr_order(c.n_cr, st.data(), lb.data(), order.data());
    int run_max = INT32_MIN;
    for (int x = 0; x < c.n_cr; ++x) { run_max = std::max(run_max, in->prof_end[order[x]]); pmax[x] = run_max; }
    std::vector<long long> psets(nr + 1), var_ps(nv + 1), pos(nv + 1);
    for (int r = 0; r < nr; ++r) psets[r] = out->phase_sets[r];
    for (int v = 0; v < nv; ++v) { var_ps[v] = out->var_phase_set[v]; pos[v] = in->pos[v]; }
    std::vector<int32_t> valid(nv + 1), flags(nv + 1), nag(nv + 1), ncf(nv + 1), snap(2 * nv + 2), nuniq(nv + 1);
    for (int v = 0; v < nv; ++v) nuniq[v] = std::min(4, in->n_uniq_alles[v]);
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_chunks = 1;
    a.ordered_ids = in->ordered_read_ids; a.is_skipped = in->is_skipped; a.pstart = in->prof_start; a.pend = in->prof_end;
    a.allele_off = in->allele_off; a.alleles = in->alleles; a.cr_order = order.data(); a.cr_pmax_end = pmax.data();
    a.haps = out->haps; a.phase_sets = psets.data(); a.agree = out->n_clean_agree_snps; a.conflict = out->n_clean_conflict_snps;
    a.cate = in->var_cate; a.type = in->var_type; a.hp = in->is_hp_indel; a.nuniq = nuniq.data(); a.alle_covs = in->alle_covs; a.total_cov = in->total_cov;
    a.pos = pos.data(); a.cons = out->hap_to_cons_alle; a.prof = out->hap_to_alle_profile; a.var_ps = var_ps.data();
    a.valid = valid.data(); a.flags = flags.data(); a.n_agree = nag.data(); a.n_conf = ncf.data(); a.snap = snap.data();
    int sh[8] = {0};
    Phaser p;
    p.run(a, c, sh);
    for (int r = 0; r < nr; ++r) out->phase_sets[r] = psets[r];
    for (int v = 0; v < nv; ++v) out->var_phase_set[v] = var_ps[v];
    return 0;
}
