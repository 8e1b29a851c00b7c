// tests/emu/sites_emu.cpp -- TEST INFRASTRUCTURE: runs the product's candidate-site device logic (longcalld_b200/csrc/sites_device.cuh:
// count / scatter per read, group / emit per position bin) on the host over one chunk.  Same signature as the oracle's
// lcd_oracle_collect_sites; the scatter runs over the reads in reverse so that the bin contents arrive in another order than the oracle's.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/sites_device.cuh"
#include <vector>

using namespace lcd::sites;

extern "C" int emu_collect_sites(const lcd_pileup_input_t *in, int64_t reg_beg, int64_t reg_end, lcd_sites_output_t *out) {
    const int nr = in->n_reads;
    std::vector<int32_t> read_chunk(nr + 1, 0);
    std::vector<uint8_t> active(nr + 1, 0);
    for (int i = 0; i < nr; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) active[r] = 1; }
    long long lo = 0, hi = 0; bool any = false;
    for (int r = 0; r < nr; ++r) if (active[r]) {
        if (!any || in->read_beg[r] - 1 < lo) lo = in->read_beg[r] - 1;
        if (!any || in->read_end[r] + 1 > hi) hi = in->read_end[r] + 1;
        any = true;
    }
    if (reg_beg != -1 && reg_beg - 1 > lo) lo = reg_beg - 1;
    if (reg_end != -1 && reg_end < hi) hi = reg_end;
    Chunk c; c.reg_beg = reg_beg; c.reg_end = reg_end; c.lo = lo; c.bin0 = 0; c.n_bins = hi >= lo ? ((hi - lo) >> BIN_SHIFT) + 1 : 1; c.alt_base = 0; c.min_sv_len = in->min_sv_len; c.pad = 0;
    const long long nb = c.n_bins;
    std::vector<int32_t> bin_count(nb + 1, 0), bin_cursor(nb + 1, 0), bin_keep(nb + 1, 0);
    std::vector<long long> bin_first(nb + 1, 0), keep_first(nb + 1, 0);
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_reads_total = nr; a.n_bins_total = nb; a.read_chunk = read_chunk.data(); a.read_active = active.data();
    a.digar_first = (const long long *)in->digar_first; a.n_digar = in->n_digar; a.digar_pos = (const long long *)in->digar_pos; a.digar_type = in->digar_type;
    a.digar_len = in->digar_len; a.digar_low_qual = in->digar_low_qual; a.digar_alt_off = (const long long *)in->digar_alt_off; a.digar_alt = in->digar_alt;
    a.bin_count = bin_count.data(); a.bin_first = bin_first.data(); a.bin_cursor = bin_cursor.data(); a.bin_keep = bin_keep.data(); a.keep_first = keep_first.data();
    for (long long g = 0; g < nr; ++g) count_read(a, g);
    for (long long b = 0; b < nb; ++b) bin_first[b + 1] = bin_first[b] + bin_count[b];
    std::vector<long long> cand(bin_first[nb] + 1, 0);
    a.cand = cand.data();
    for (long long g = nr - 1; g >= 0; --g) scatter_read(a, g);
    for (long long b = 0; b < nb; ++b) group_bin(a, 0, b);
    for (long long b = 0; b < nb; ++b) keep_first[b + 1] = keep_first[b] + bin_keep[b];
    const long long ns = keep_first[nb];
    out->n_sites = ns;
    if (ns > out->cap) return -3;
    std::vector<long long> aoff(ns + 1, 0);
    a.site_pos = (long long *)out->site_pos; a.site_type = out->site_type; a.site_ref_len = out->site_ref_len; a.site_alt_len = out->site_alt_len;
    a.site_src = (long long *)out->site_src; a.site_alt_off = aoff.data();
    for (long long b = 0; b < nb; ++b) emit_bin(a, 0, b);
    return 0;
}
