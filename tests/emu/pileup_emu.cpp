// tests/emu/pileup_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K2 device logic (longcalld_b200/csrc/pileup_device.cuh,
// one thread per read) on the host over one chunk.  Same signature as the oracle's lcd_oracle_collect_cand_vars.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/pileup_device.cuh"
#include <vector>

using namespace lcd::pileup;

extern "C" int emu_collect_cand_vars(const lcd_pileup_input_t *in, lcd_pileup_output_t *out) {
    Chunk c; c.n_sites = in->n_sites; c.min_bq = in->min_bq; c.min_sv_len = in->min_sv_len; c.pad = 0; c.site_off = 0; c.alt_base = 0; c.salt_base = 0; c.pad2 = 0;
    std::vector<int32_t> read_chunk(in->n_reads + 1, 0);
    std::vector<uint8_t> active(in->n_reads + 1, 0);
    for (int i = 0; i < in->n_reads; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) active[r] = 1; }
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_reads_total = in->n_reads; a.read_chunk = read_chunk.data(); a.read_active = active.data();
    a.read_beg = (const long long *)in->read_beg; a.read_end = (const long long *)in->read_end; a.read_is_rev = in->read_is_rev;
    a.digar_first = (const long long *)in->digar_first; a.n_digar = in->n_digar; a.qual_off = (const long long *)in->qual_off; a.qual = in->qual;
    a.digar_pos = (const long long *)in->digar_pos; a.digar_type = in->digar_type; a.digar_len = in->digar_len; a.digar_qi = in->digar_qi;
    a.digar_low_qual = in->digar_low_qual; a.digar_alt_off = (const long long *)in->digar_alt_off; a.digar_alt = in->digar_alt;
    a.site_pos = (const long long *)in->site_pos; a.site_type = in->site_type; a.site_ref_len = in->site_ref_len; a.site_alt_len = in->site_alt_len;
    a.site_alt_off = (const long long *)in->site_alt_off; a.site_alt = in->site_alt; a.site_counts = out->site_counts;
    memset(out->site_counts, 0, sizeof(int32_t) * 8 * (size_t)in->n_sites);
    for (long long g = 0; g < in->n_reads; ++g) process_read(a, g);
    return 0;
}

// K3: profile_read over one chunk, rows laid out as the host plan does (first_site / row_end_site)
extern "C" int emu_read_var_profile(const lcd_pileup_input_t *in, const lcd_profile_extra_t *ex, lcd_profile_output_t *out) {
    Chunk c; c.n_sites = in->n_sites; c.min_bq = in->min_bq; c.min_sv_len = in->min_sv_len; c.pad = 0; c.site_off = 0; c.alt_base = 0; c.salt_base = 0; c.pad2 = 0;
    const int nr = in->n_reads;
    std::vector<int32_t> read_chunk(nr + 1, 0), row_cap(nr + 1, 0);
    std::vector<uint8_t> active(nr + 1, 0);
    std::vector<long long> row_off(nr + 1, 0);
    for (int i = 0; i < nr; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) active[r] = 1; }
    long long tot = 0;
    for (int r = 0; r < nr; ++r) {
        const long long v0 = first_site((const long long *)in->site_pos, in->site_type, 0, in->n_sites, in->read_beg[r]);
        const long long v1 = row_end_site((const long long *)in->site_pos, in->site_type, v0, in->n_sites, in->read_end[r]);
        row_off[r] = tot; row_cap[r] = (int32_t)(v1 - v0); tot += v1 - v0;
    }
    if (tot > out->alleles_cap) return -3;
    int32_t status = 0;
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_reads_total = nr; a.read_chunk = read_chunk.data(); a.read_active = active.data();
    a.read_beg = (const long long *)in->read_beg; a.read_end = (const long long *)in->read_end; a.read_is_rev = in->read_is_rev;
    a.digar_first = (const long long *)in->digar_first; a.n_digar = in->n_digar; a.qual_off = (const long long *)in->qual_off; a.qual = in->qual;
    a.digar_pos = (const long long *)in->digar_pos; a.digar_type = in->digar_type; a.digar_len = in->digar_len; a.digar_qi = in->digar_qi;
    a.digar_low_qual = in->digar_low_qual; a.digar_alt_off = (const long long *)in->digar_alt_off; a.digar_alt = in->digar_alt;
    a.site_pos = (const long long *)in->site_pos; a.site_type = in->site_type; a.site_ref_len = in->site_ref_len; a.site_alt_len = in->site_alt_len;
    a.site_alt_off = (const long long *)in->site_alt_off; a.site_alt = in->site_alt;
    a.var_cate = ex->var_cate; a.nreg_first = (const long long *)ex->nreg_first; a.n_nreg = ex->n_nreg; a.nreg_beg = (const long long *)ex->nreg_beg; a.nreg_end = (const long long *)ex->nreg_end;
    a.row_off = row_off.data(); a.row_cap = row_cap.data(); a.prof_start = out->prof_start; a.prof_end = out->prof_end; a.allele_off = (long long *)out->allele_off;
    a.alleles = out->alleles; a.alt_qi = out->alt_qi; a.status = &status;
    for (long long g = 0; g < nr; ++g) profile_read(a, g);
    out->n_alleles = tot;
    return status;
}
