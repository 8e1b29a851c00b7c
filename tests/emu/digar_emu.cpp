// tests/emu/digar_emu.cpp -- TEST INFRASTRUCTURE: runs the product's K1 device logic (longcalld_b200/csrc/digar_device.cuh:
// count_read / fill_read per read as one thread, hist_read as a single lane) on the host over one chunk, with the host plan's
// exclusive scan and interval compaction restated.  Same signature as the oracle's lcd_oracle_collect_digar_eqx.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/digar_device.cuh"
#include <vector>

using namespace lcd::digar;
extern "C" void lcd_oracle_cr_order(int n, const int32_t *start, const int32_t *label, int32_t *order_out);   // cgranges order (> 64 intervals)

static const long long *g_rlen = nullptr;       // set by emu_collect_digar_tags: the reads' own reference lengths
extern "C" int emu_collect_digar_eqx(const lcd_digar_input_t *in, lcd_digar_output_t *out) {
    const long long nr = in->n_reads, stride = nr + 1;
    Chunk c; memset(&c, 0, sizeof(c));
    c.min_bq = in->min_bq; c.max_xgaps = in->noisy_reg_max_xgaps; c.win = in->noisy_reg_slide_win; c.end_clip_reg = in->end_clip_reg; c.flank_win = in->end_clip_reg_flank_win;
    c.max_noisy_frac = in->max_noisy_frac_per_read; c.max_var_ratio = in->max_var_ratio_per_read; c.whole_ref_len = in->whole_ref_len; c.read0 = 0;
    std::vector<int32_t> read_chunk(nr + 1, 0); std::vector<uint8_t> active(nr + 1, 0);
    for (int i = 0; i < nr; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) active[r] = 1; }
    long long nq = 0;
    for (int r = 0; r < nr; ++r) if (active[r]) nq = std::max<long long>(nq, in->qual_off[r] + in->l_qseq[r]);
    uint8_t *qual = (uint8_t*)aligned_alloc(16, (size_t)((nq + 47) & ~15ll)); memset(qual, 0, (size_t)((nq + 47) & ~15ll)); memcpy(qual, in->qual, (size_t)nq);
    std::vector<long long> cnt(3 * stride, 0), first(3 * stride, 0);
    std::vector<unsigned long long> qc(256, 0); int32_t status = 0;
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.chunks = &c; a.n_reads_total = nr; a.read_chunk = read_chunk.data(); a.read_active = active.data();
    a.read_pos0 = (const long long *)in->read_pos0; a.read_is_rev = in->read_is_rev; a.is_palindrome = in->is_palindrome;
    a.n_cigar = in->n_cigar; a.cigar_off = (const long long *)in->cigar_off; a.cigar = in->cigar; a.rlen = g_rlen;
    a.l_qseq = in->l_qseq; a.seq_off = (const long long *)in->seq_off; a.bseq = in->bseq; a.qual_off = (const long long *)in->qual_off; a.qual = qual;
    std::vector<int32_t> ndig(nr + 1, 0); a.n_digar = ndig.data();
    a.cnt = cnt.data(); a.first = first.data(); a.stride = stride; a.qual_counts = qc.data(); a.status = &status;
    for (long long g = 0; g < nr; ++g) count_read(a, g);
    for (int j = 0; j < 3; ++j) { long long run = 0; for (long long g = 0; g <= nr; ++g) { first[j * stride + g] = run; run += g < nr ? cnt[j * stride + g] : 0; } }
    const long long nd = first[nr], na = first[stride + nr], ncap = first[2 * stride + nr];
    int rc = 0;
    if (nd > out->digar_cap || na > out->alt_cap) { free(qual); return -3; }
    std::vector<long long> nb(ncap + 1), ne(ncap + 1); std::vector<int32_t> nl(ncap + 1), nn(nr + 1, 0);
    a.skip = out->skip; a.read_beg = (long long *)out->read_beg; a.read_end = (long long *)out->read_end;
    a.digar_pos = (long long *)out->digar_pos; a.digar_type = out->digar_type; a.digar_len = out->digar_len; a.digar_qi = out->digar_qi;
    a.digar_low_qual = out->digar_low_qual; a.digar_alt_off = (long long *)out->digar_alt_off; a.digar_alt = out->digar_alt;
    a.n_nreg = nn.data(); a.nreg_beg = nb.data(); a.nreg_end = ne.data(); a.nreg_label = nl.data();
    for (long long g = 0; g < nr; ++g) fill_read(a, g);
    if (status) { free(qual); return -10 - status; }
    unsigned hist[256]; memset(hist, 0, sizeof(hist));
    for (long long g = 0; g < nr; ++g) if (active[g]) hist_read(a, g, 0, 1, hist);
    for (int b = 0; b < 256; ++b) out->qual_counts[b] = hist[b];
    long long top = 0; out->n_cnreg = 0;
    for (long long r = 0; r < nr; ++r) {
        out->digar_first[r] = first[r]; out->n_digar[r] = (int32_t)cnt[r]; out->nreg_first[r] = top; out->n_nreg[r] = nn[r];
        if (top + nn[r] > out->nreg_cap) { rc = -4; break; }
        const long long f = first[2 * stride + r];
        std::vector<int32_t> ord(nn[r]);
        for (int x = 0; x < nn[r]; ++x) ord[x] = x;
        if (nn[r] > 64) {
            std::vector<int32_t> st(nn[r]), id(nn[r]);
            for (int x = 0; x < nn[r]; ++x) { st[x] = (int32_t)nb[f + x]; id[x] = x; }
            lcd_oracle_cr_order(nn[r], st.data(), id.data(), ord.data());
        }
        for (int x = 0; x < nn[r]; ++x) { out->nreg_beg[top + x] = nb[f + ord[x]]; out->nreg_end[top + x] = ne[f + ord[x]]; out->nreg_label[top + x] = nl[f + ord[x]]; }
        top += nn[r];
    }
    for (long long i = 0; i < nr && !rc; ++i) {
        const int r = in->ordered_read_ids[i];
        if (!active[r] || out->skip[r]) continue;
        for (long long y = out->nreg_first[r]; y < out->nreg_first[r] + out->n_nreg[r]; ++y)
            if (!(out->nreg_beg[y] + 1 > in->reg_end || out->nreg_end[y] < in->reg_beg)) {
                if (out->n_cnreg >= out->cnreg_cap) { rc = -4; break; }
                out->cnreg_beg[out->n_cnreg] = out->nreg_beg[y]; out->cnreg_end[out->n_cnreg] = out->nreg_end[y]; out->cnreg_label[out->n_cnreg] = out->nreg_label[y]; out->n_cnreg++;
            }
    }
    out->n_digar_total = nd; out->n_alt_total = na; out->n_nreg_total = top;
    free(qual);
    return rc;
}

// The tag front end (md_device.cuh: count, scan, fill, thread per read) followed by the pass above: what lcd_digar_tags_batch runs.
// kind / off / text as in lcd_read_tags_t; returns -20 - status when the front end rejects a read.
#include "../../longcalld_b200/csrc/md_device.cuh"
extern "C" int emu_collect_digar_tags(const lcd_digar_input_t *in, const int8_t *kind, const int64_t *off, const char *text, const char *ref, int64_t ref_beg, int64_t ref_end,
                                      lcd_digar_output_t *out) {
    namespace M = lcd::md;
    const long long nr = in->n_reads;
    std::vector<uint8_t> active(nr + 1, 0);
    for (int i = 0; i < nr; ++i) { const int r = in->ordered_read_ids[i]; if (!in->is_skipped[r]) active[r] = 1; }
    std::vector<long long> cnt(nr + 2, 0), first(nr + 2, 0), coff(nr + 1, 0), toff(nr + 1, -1); std::vector<int32_t> ncig(nr + 1, 0), read_chunk(nr + 1, 0); int32_t status = 0;
    for (long long r = 0; r < nr; ++r) toff[r] = (kind[r] == M::KIND_MD || kind[r] == M::KIND_CS) ? off[r] : -1;
    long long roff = 0, rb = ref_beg, re = ref_end;
    M::KernelArgs a; memset(&a, 0, sizeof(a));
    a.n_reads_total = nr; a.read_active = active.data(); a.n_cigar0 = in->n_cigar; a.cigar_off0 = (const long long *)in->cigar_off; a.cigar0 = in->cigar;
    a.md_off = toff.data(); a.md = text; a.kind = kind; a.read_chunk = read_chunk.data(); a.ref_off = &roff; a.ref_beg = &rb; a.ref_end = &re; a.ref = ref;
    a.read_pos0 = (const long long *)in->read_pos0; a.l_qseq = in->l_qseq; a.seq_off = (const long long *)in->seq_off; a.bseq = in->bseq;
    a.cnt = cnt.data(); a.first = first.data(); a.n_cigar = ncig.data(); a.cigar_off = coff.data(); a.status = &status;
    for (long long g = 0; g < nr; ++g) M::count_read(a, g);
    if (status) return -20 - status;
    long long run = 0; for (long long g = 0; g <= nr; ++g) { first[g] = run; run += g < nr ? cnt[g] : 0; }
    std::vector<uint32_t> cig(run + 4, 0); a.cigar = cig.data();
    std::vector<long long> rlen(nr + 1, 0); a.rlen = rlen.data();
    for (long long g = 0; g < nr; ++g) M::fill_read(a, g);
    lcd_digar_input_t x = *in;
    x.cigar = cig.data(); x.cigar_off = (const int64_t *)coff.data(); x.n_cigar = ncig.data();
    g_rlen = rlen.data();
    const int rc = emu_collect_digar_eqx(&x, out);
    g_rlen = nullptr;
    return rc;
}
