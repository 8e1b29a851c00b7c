// digar_kernel.cu -- K1 launchers and host plan: batched =/X difference-list pass over the reads of many chunks.
// Device logic and design notes: digar_device.cuh.
#include "lcd_common.cuh"
#include "digar_device.cuh"
#include "md_device.cuh"
#include <algorithm>

namespace lcd {
namespace digar {

constexpr int THREADS = 128;
constexpr int HIST_WARPS = 8;
constexpr int SCAN_THREADS = 256;      // small CTAs: the scan has to fit next to a resident persistent DP grid (lcd_gpu_reserve_sms)

__global__ void __launch_bounds__(THREADS)
digar_count_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        count_read(a, g);
}

// Exclusive scan of cnt[j][0 .. n] (the entry at n counts as 0, so first[j][n] is the total); one CTA per array j:
// every thread sums a contiguous segment, the partial sums are scanned through shared memory, the segment is rewritten.
__global__ void __launch_bounds__(SCAN_THREADS)
digar_scan_kernel(const long long *cnt, long long *first, long long n, long long stride) {
    __shared__ long long warp_sum[SCAN_THREADS / 32];
    const long long *in = cnt + blockIdx.x * stride; long long *out = first + blockIdx.x * stride;
    const long long seg = (n + 1 + SCAN_THREADS - 1) / SCAN_THREADS;
    const long long i0 = min(n + 1, seg * (long long)threadIdx.x), i1 = min(n + 1, i0 + seg);
    long long s = 0;
    for (long long i = i0; i < i1; ++i) s += i < n ? in[i] : 0;
    long long incl = s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < SCAN_THREADS / 32 ? warp_sum[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        if (lane < SCAN_THREADS / 32) warp_sum[lane] = wi - w;
    }
    __syncthreads();
    long long run = warp_sum[warp] + incl - s;
    for (long long i = i0; i < i1; ++i) { out[i] = run; run += i < n ? in[i] : 0; }
}

// MD-tag front end (md_device.cuh): (CIGAR with M, MD) -> the =/X CIGAR the kernels above and below consume, thread per read
__global__ void __launch_bounds__(THREADS)
md_count_kernel(const md::KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        md::count_read(a, g);
}
__global__ void __launch_bounds__(THREADS)
md_fill_kernel(const md::KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        md::fill_read(a, g);
}

__global__ void __launch_bounds__(THREADS)
digar_fill_kernel(const KernelArgs a) {
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < a.n_reads_total; g += (long long)gridDim.x * THREADS)
        fill_read(a, g);
}

// One warp per contiguous range of reads; per-warp shared-memory histogram, flushed when the range crosses into another chunk.
__global__ void __launch_bounds__(HIST_WARPS * 32)
digar_hist_kernel(const KernelArgs a, long long reads_per_warp) {
    __shared__ unsigned hist[HIST_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *h = hist[warp];
    for (int b = lane; b < 256; b += 32) h[b] = 0;
    __syncwarp();
    const long long gw = (long long)blockIdx.x * HIST_WARPS + warp;
    const long long r0 = gw * reads_per_warp, r1 = min(a.n_reads_total, r0 + reads_per_warp);
    int cur = -1;
    for (long long g = r0; g < r1; ++g) {
        if (!a.read_active[g]) continue;
        const int c = a.read_chunk[g];
        if (c != cur) {
            if (cur >= 0) {
                __syncwarp();
                for (int b = lane; b < 256; b += 32) { const unsigned v = h[b]; if (v) { atomicAdd(a.qual_counts + 256ll * cur + b, (unsigned long long)v); h[b] = 0; } }
                __syncwarp();
            }
            cur = c;
        }
        hist_read(a, g, lane, 32, h);
    }
    if (cur >= 0) {
        __syncwarp();
        for (int b = lane; b < 256; b += 32) { const unsigned v = h[b]; if (v) atomicAdd(a.qual_counts + 256ll * cur + b, (unsigned long long)v); }
    }
}

// cgranges' cr_index order for more than 64 intervals (src/cgranges.c:13-64): in-place MSD radix sort on the start, 8 bits a
// pass from the top byte of the 64-bit key, buckets of at most 64 finished by insertion sort.  Not stable -- reproduced literally.
struct Intv { uint64_t key; long long beg, end; int32_t label; };
static void intv_insertion_sort(Intv *beg, Intv *end) {
    for (Intv *i = beg + 1; i < end; ++i)
        if (i->key < (i - 1)->key) {
            Intv tmp = *i, *j;
            for (j = i; j > beg && tmp.key < (j - 1)->key; --j) *j = *(j - 1);
            *j = tmp;
        }
}
static void intv_radix_sort(Intv *beg, Intv *end, int shift) {
    Intv *bb[256], *be[256]; size_t cnt[256] = {0};
    for (Intv *i = beg; i != end; ++i) cnt[(i->key >> shift) & 255]++;
    Intv *p = beg;
    for (int k = 0; k < 256; ++k) { bb[k] = p; p += cnt[k]; be[k] = p; }
    for (int k = 0; k < 256;) {
        if (bb[k] != be[k]) {
            int l = (int)((bb[k]->key >> shift) & 255);
            if (l != k) {
                Intv tmp = *bb[k], swap;
                do { swap = tmp; tmp = *bb[l]; *bb[l]++ = swap; l = (int)((tmp.key >> shift) & 255); } while (l != k);
                *bb[k]++ = tmp;
            } else ++bb[k];
        } else ++k;
    }
    if (shift) {
        const int next = shift > 8 ? shift - 8 : 0;
        Intv *b0 = beg;
        for (int k = 0; k < 256; ++k) {
            Intv *e0 = be[k];
            if (e0 - b0 > 64) intv_radix_sort(b0, e0, next);
            else if (e0 - b0 > 1) intv_insertion_sort(b0, e0);
            b0 = e0;
        }
    }
}

template <typename T, typename U> static void append(std::vector<T> &dst, const U *src, size_t n, long long add = 0) {
    const size_t o = dst.size(); dst.resize(o + n);
    for (size_t i = 0; i < n; ++i) dst[o + i] = (T)(src[i] + (U)add);
}

struct DigarPlan : Plan {
    bool uses_pool() const override { return false; }
    std::vector<Chunk> chunks; std::vector<long long> read_off, reg_beg, reg_end; std::vector<int32_t> ordered;
    std::vector<uint8_t> h_active;
    long long tot_reads = 0, tot_cigar = 0, tot_seq = 0, tot_qual = 0, stride = 1;
    long long tot_digar = -1, tot_alt = 0, tot_ncap = 0;
    DevBuf<Chunk> d_chunks; DevBuf<int32_t> d_read_chunk, d_ncig, d_lq; DevBuf<uint8_t> d_active, d_rev, d_pal, d_bseq, d_qual; DevBuf<uint32_t> d_cigar;
    DevBuf<long long> d_pos0, d_coff, d_soff, d_qoff, d_cnt, d_first, d_rlen;
    DevBuf<uint8_t> d_skip, d_dlow, d_dalt; DevBuf<long long> d_beg, d_end, d_dpos, d_daoff, d_nbeg, d_nend; DevBuf<int8_t> d_dtype;
    DevBuf<int32_t> d_dlen, d_dqi, d_nnreg, d_nlabel, d_status, d_ndig; DevBuf<unsigned long long> d_qc;
    std::vector<long long> h_first; std::vector<int32_t> h_nnreg; bool have_index = false;

    // tags: MD strings only (lcd_digar_md_*); rtags: the general form, a variant per read (lcd_digar_tags_*)
    int build(int n_, const lcd_digar_input_t *in, const lcd_md_tags_t *tags = nullptr, const lcd_read_tags_t *rtags = nullptr) {
        n = n_;
        Context &c = ctx();
        if (n == 0) return 0;
        std::vector<int32_t> read_chunk, ncig, lq; std::vector<uint8_t> rev, pal; std::vector<long long> pos0, coff, soff, qoff;
        std::vector<long long> cig_base(n), seq_base(n), qual_base(n), cig_n(n), seq_n(n), qual_n(n), md_base(n, 0), md_n(n, 0), md_off; long long tot_md = 0;
        std::vector<int8_t> kinds; std::vector<long long> ref_off(n, 0), ref_beg(n, 0), ref_end(n, -1), ref_n(n, 0); long long tot_ref = 0;
        chunks.resize(n); read_off.resize(n + 1); reg_beg.resize(n); reg_end.resize(n);
        for (int i = 0; i < n; ++i) {
            const lcd_digar_input_t &x = in[i];
            if (x.n_reads < 0) { set_error("lcd_digar: chunk %d has a negative read count", i); return -1; }
            Chunk &k = chunks[i]; read_off[i] = tot_reads;
            k.min_bq = x.min_bq; k.max_xgaps = x.noisy_reg_max_xgaps; k.win = x.noisy_reg_slide_win; k.end_clip_reg = x.end_clip_reg; k.flank_win = x.end_clip_reg_flank_win;
            k.pad = 0; k.max_noisy_frac = x.max_noisy_frac_per_read; k.max_var_ratio = x.max_var_ratio_per_read; k.whole_ref_len = x.whole_ref_len; k.read0 = tot_reads;
            if (k.win < 2) { set_error("lcd_digar: chunk %d has a sliding window of %d (the reference needs >= 2)", i, k.win); return -1; }
            reg_beg[i] = x.reg_beg; reg_end[i] = x.reg_end;
            long long nc = 0, ns = 0, nq = 0;
            std::vector<uint8_t> listed(x.n_reads, 0);
            for (int r = 0; r < x.n_reads; ++r) { const int id = x.ordered_read_ids[r]; if (id >= 0 && id < x.n_reads) listed[id] = 1; }
            for (int r = 0; r < x.n_reads; ++r) {
                if (x.n_cigar[r] < 0 || x.l_qseq[r] < 0 || x.cigar_off[r] < 0 || x.seq_off[r] < 0 || x.qual_off[r] < 0) { set_error("lcd_digar: chunk %d read %d has invalid sizes", i, r); return -1; }
                const bool act = listed[r] && !x.is_skipped[r];
                h_active.push_back(act); read_chunk.push_back(i);
                if (!act) continue;
                nc = std::max<long long>(nc, x.cigar_off[r] + x.n_cigar[r]); ns = std::max<long long>(ns, x.seq_off[r] + (x.l_qseq[r] + 1) / 2);
                nq = std::max<long long>(nq, x.qual_off[r] + x.l_qseq[r]);
            }
            cig_base[i] = tot_cigar; seq_base[i] = tot_seq; qual_base[i] = tot_qual; cig_n[i] = nc; seq_n[i] = ns; qual_n[i] = nq;
            if (tags || rtags) {      // tags: NUL-terminated strings at text + off[r] (< 0: none)
                long long top = 0; bool need_ref = false;
                const int64_t *t_off = tags ? tags[i].md_off : rtags[i].off; const char *t_text = tags ? tags[i].md : rtags[i].text;
                for (int r = 0; r < x.n_reads; ++r) {
                    const bool act = listed[r] && !x.is_skipped[r];
                    int kind = !act ? md::KIND_EQX : tags ? (t_off[r] < 0 ? md::KIND_EQX : md::KIND_MD) : rtags[i].kind[r];
                    if (kind < md::KIND_EQX || kind > md::KIND_REFSEQ) { set_error("lcd_digar: chunk %d read %d has an unknown tag kind %d", i, r, kind); return -1; }
                    const bool has_text = kind == md::KIND_MD || kind == md::KIND_CS;
                    if (has_text && (!t_off || !t_text || t_off[r] < 0)) { set_error("lcd_digar: chunk %d read %d is tagged but has no tag text", i, r); return -1; }
                    const long long o = has_text ? t_off[r] : -1;
                    md_off.push_back(o < 0 ? -1 : tot_md + o);
                    if (o >= 0) top = std::max<long long>(top, o + (long long)strlen(t_text + o) + 1);
                    kinds.push_back((int8_t)kind);
                    if (kind == md::KIND_REFSEQ) need_ref = true;
                }
                md_base[i] = tot_md; md_n[i] = top; tot_md += top;
                if (need_ref) {
                    if (!rtags[i].ref_seq || rtags[i].ref_end < rtags[i].ref_beg) { set_error("lcd_digar: chunk %d has untagged plain-M reads but no reference window", i); return -1; }
                    ref_off[i] = tot_ref; ref_beg[i] = rtags[i].ref_beg; ref_end[i] = rtags[i].ref_end; ref_n[i] = rtags[i].ref_end - rtags[i].ref_beg + 1; tot_ref += (ref_n[i] + 15) & ~15ll;
                }
            }
            append(ordered, x.ordered_read_ids, x.n_reads);
            append(pos0, x.read_pos0, x.n_reads); append(rev, x.read_is_rev, x.n_reads); append(pal, x.is_palindrome, x.n_reads);
            append(ncig, x.n_cigar, x.n_reads); append(lq, x.l_qseq, x.n_reads);
            append(coff, x.cigar_off, x.n_reads, tot_cigar); append(soff, x.seq_off, x.n_reads, tot_seq); append(qoff, x.qual_off, x.n_reads, tot_qual);
            tot_reads += x.n_reads; tot_cigar += nc; tot_seq += (ns + 15) & ~15ll; tot_qual += (nq + 15) & ~15ll;   // 16-byte aligned chunk bases
        }
        read_off[n] = tot_reads; stride = tot_reads + 1;
        auto pad = [](auto &v) { v.push_back(0); };
        pad(read_chunk); pad(h_active); pad(pos0); pad(rev); pad(pal); pad(ncig); pad(lq); pad(coff); pad(soff); pad(qoff);
        if (tags || rtags) { md_off.push_back(-1); kinds.push_back((int8_t)md::KIND_EQX); }
        cudaStream_t s = cur_stream();
        if (d_chunks.upload(chunks.data(), n, s) || d_read_chunk.upload(read_chunk.data(), read_chunk.size(), s) || d_active.upload(h_active.data(), h_active.size(), s) ||
            d_pos0.upload(pos0.data(), pos0.size(), s) || d_rev.upload(rev.data(), rev.size(), s) || d_pal.upload(pal.data(), pal.size(), s) ||
            d_ncig.upload(ncig.data(), ncig.size(), s) || d_lq.upload(lq.data(), lq.size(), s) || d_coff.upload(coff.data(), coff.size(), s) ||
            d_soff.upload(soff.data(), soff.size(), s) || d_qoff.upload(qoff.data(), qoff.size(), s)) return -1;
        // the bulk arrays go straight from the caller's buffers into their slice of the device arrays (no host staging copy)
        if (d_cigar.alloc(tot_cigar + 4) || d_bseq.alloc(tot_seq + 32) || d_qual.alloc(tot_qual + 32)) return -1;
        LCD_CUDA_OK(cudaMemsetAsync(d_qual.p + tot_qual, 0, 32, s));
        for (int i = 0; i < n; ++i) {
            if (cig_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_cigar.p + cig_base[i], in[i].cigar, sizeof(uint32_t) * cig_n[i], cudaMemcpyHostToDevice, s));
            if (seq_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_bseq.p + seq_base[i], in[i].bseq, seq_n[i], cudaMemcpyHostToDevice, s));
            if (qual_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_qual.p + qual_base[i], in[i].qual, qual_n[i], cudaMemcpyHostToDevice, s));
        }
        if (d_cnt.alloc(3 * stride) || d_first.alloc(3 * stride) || d_skip.alloc(stride) || d_beg.alloc(stride) || d_end.alloc(stride) || d_nnreg.alloc(stride) || d_ndig.alloc(stride) ||
            d_qc.alloc(256 * (size_t)n) || d_status.alloc(1)) return -1;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        if ((tags || rtags) && tot_reads) {
            std::vector<const char *> text(n), ref(n, nullptr);
            for (int i = 0; i < n; ++i) { text[i] = tags ? tags[i].md : rtags[i].text; if (rtags && ref_n[i]) ref[i] = rtags[i].ref_seq; }
            if (convert_md(s, md_off, text, md_base, md_n, tot_md, rtags ? &kinds : nullptr, ref, ref_off, ref_beg, ref_end, ref_n, tot_ref)) return -1;
        }
        return 0;
    }

    // MD-tagged reads: d_cigar / d_coff / d_ncig so far hold the reads' own CIGARs; replace them by the =/X CIGARs of the reference's MD walk
    // (the same for cs-tagged and untagged plain-M reads: kinds says which walk a read takes)
    int convert_md(cudaStream_t s, const std::vector<long long> &md_off, const std::vector<const char *> &text, const std::vector<long long> &md_base,
                   const std::vector<long long> &md_n, long long tot_md, const std::vector<int8_t> *kinds, const std::vector<const char *> &ref,
                   const std::vector<long long> &ref_off, const std::vector<long long> &ref_beg, const std::vector<long long> &ref_end,
                   const std::vector<long long> &ref_n, long long tot_ref) {
        Context &c = ctx();
        DevBuf<long long> d_mdoff, d_cnt1, d_first1; DevBuf<char> d_md; DevBuf<uint32_t> d_cig2; DevBuf<long long> d_coff2; DevBuf<int32_t> d_ncig2, d_st;
        DevBuf<int8_t> d_kind; DevBuf<long long> d_roff, d_rbeg, d_rend; DevBuf<char> d_ref;
        if (kinds && (d_kind.upload(kinds->data(), kinds->size(), s) || d_roff.upload(ref_off.data(), ref_off.size(), s) || d_rbeg.upload(ref_beg.data(), ref_beg.size(), s) ||
                      d_rend.upload(ref_end.data(), ref_end.size(), s) || d_ref.alloc(tot_ref + 16))) return -1;
        if (kinds) for (int i = 0; i < n; ++i) if (ref_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_ref.p + ref_off[i], ref[i], (size_t)ref_n[i], cudaMemcpyHostToDevice, s));
        if (d_mdoff.upload(md_off.data(), md_off.size(), s) || d_md.alloc(tot_md + 16) || d_cnt1.alloc(stride) || d_first1.alloc(stride) || d_coff2.alloc(stride) ||
            d_ncig2.alloc(stride) || d_st.alloc(1)) return -1;
        LCD_CUDA_OK(cudaMemsetAsync(d_md.p + tot_md, 0, 16, d_md.st));          // NUL slack behind the last tag: a truncated tag ends the walk, it is never read past
        for (int i = 0; i < n; ++i) if (md_n[i]) LCD_CUDA_OK(cudaMemcpyAsync(d_md.p + md_base[i], text[i], (size_t)md_n[i], cudaMemcpyHostToDevice, s));
        LCD_CUDA_OK(cudaMemsetAsync(d_st.p, 0, sizeof(int32_t), s));
        md::KernelArgs a; memset(&a, 0, sizeof(a));
        a.n_reads_total = tot_reads; a.read_active = d_active.p; a.n_cigar0 = d_ncig.p; a.cigar_off0 = d_coff.p; a.cigar0 = d_cigar.p; a.md_off = d_mdoff.p; a.md = d_md.p;
        a.cnt = d_cnt1.p; a.first = d_first1.p; a.n_cigar = d_ncig2.p; a.cigar_off = d_coff2.p; a.status = d_st.p;
        if (kinds) { a.kind = d_kind.p; a.read_chunk = d_read_chunk.p; a.ref_off = d_roff.p; a.ref_beg = d_rbeg.p; a.ref_end = d_rend.p; a.ref = d_ref.p; }
        a.read_pos0 = d_pos0.p; a.l_qseq = d_lq.p; a.seq_off = d_soff.p; a.bseq = d_bseq.p;
        const int grid = (int)std::min<long long>((tot_reads + THREADS - 1) / THREADS, (long long)c.sm_count * 16);
        md_count_kernel<<<grid, THREADS, 0, s>>>(a);
        digar_scan_kernel<<<1, SCAN_THREADS, 0, s>>>(d_cnt1.p, d_first1.p, tot_reads, stride);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches += 2;
        long long total = 0; int32_t st = 0;
        LCD_DRAIN(s);
        LCD_CUDA_OK(cudaMemcpyAsync(&total, d_first1.p + tot_reads, sizeof(long long), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(&st, d_st.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        if (st == md::MD_MISMATCH) { set_error("lcd_digar: a read's MD tag and CIGAR do not match (the reference stops as well: src/bam_utils.c:1083)"); return -2; }
        if (st == md::MD_EQX_OP) { set_error("lcd_digar: a read with an MD tag has =/X CIGAR ops (the reference stops as well: src/bam_utils.c:1139); pass md_off < 0 for such reads"); return -2; }
        if (st == md::CS_BAD) { set_error("lcd_digar: a read's cs tag is malformed (the reference stops as well: src/bam_utils.c:949)"); return -2; }
        if (st == md::CS_SEQ_MISMATCH) { set_error("lcd_digar: a read's cs tag spells bases that differ from its SEQ (the reference takes the alt bases from the tag; such reads are not handled on the GPU)"); return -2; }
        if (st) { set_error("lcd_digar: tag front end failed with status %d", st); return -2; }
        if (d_cig2.alloc(total + 4) || d_rlen.alloc(stride)) return -1;
        a.cigar = d_cig2.p; a.rlen = d_rlen.p;
        md_fill_kernel<<<grid, THREADS, 0, s>>>(a);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches++;
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        std::swap(d_cigar.p, d_cig2.p); std::swap(d_cigar.n, d_cig2.n); std::swap(d_cigar.st, d_cig2.st);
        std::swap(d_coff.p, d_coff2.p); std::swap(d_coff.n, d_coff2.n); std::swap(d_coff.st, d_coff2.st);
        std::swap(d_ncig.p, d_ncig2.p); std::swap(d_ncig.n, d_ncig2.n); std::swap(d_ncig.st, d_ncig2.st);
        tot_cigar = total;
        return 0;
    }

    void args(KernelArgs &a) {
        memset(&a, 0, sizeof(a));
        a.chunks = d_chunks.p; a.n_reads_total = tot_reads; a.read_chunk = d_read_chunk.p; a.read_active = d_active.p;
        a.read_pos0 = d_pos0.p; a.read_is_rev = d_rev.p; a.is_palindrome = d_pal.p; a.n_cigar = d_ncig.p; a.cigar_off = d_coff.p; a.cigar = d_cigar.p; a.rlen = d_rlen.p;
        a.l_qseq = d_lq.p; a.seq_off = d_soff.p; a.bseq = d_bseq.p; a.qual_off = d_qoff.p; a.qual = d_qual.p;
        a.cnt = d_cnt.p; a.first = d_first.p; a.stride = stride;
        a.skip = d_skip.p; a.read_beg = d_beg.p; a.read_end = d_end.p; a.n_digar = d_ndig.p;
        a.digar_pos = d_dpos.p; a.digar_type = d_dtype.p; a.digar_len = d_dlen.p; a.digar_qi = d_dqi.p; a.digar_low_qual = d_dlow.p; a.digar_alt_off = d_daoff.p; a.digar_alt = d_dalt.p;
        a.n_nreg = d_nnreg.p; a.nreg_beg = d_nbeg.p; a.nreg_end = d_nend.p; a.nreg_label = d_nlabel.p; a.qual_counts = d_qc.p; a.status = d_status.p;
    }

    int run(cudaStream_t s) override {
        Context &c = ctx();
        have_index = false;
        if (n == 0) return 0;
        LCD_CUDA_OK(cudaMemsetAsync(d_qc.p, 0, sizeof(unsigned long long) * 256 * (size_t)n, s));
        LCD_CUDA_OK(cudaMemsetAsync(d_status.p, 0, sizeof(int32_t), s));
        if (tot_reads == 0) return 0;
        KernelArgs a; args(a);
        const int grid = (int)std::min<long long>((tot_reads + THREADS - 1) / THREADS, (long long)c.sm_count * 16);
        digar_count_kernel<<<grid, THREADS, 0, s>>>(a);
        digar_scan_kernel<<<3, SCAN_THREADS, 0, s>>>(d_cnt.p, d_first.p, tot_reads, stride);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches += 2;
        if (tot_digar < 0) {        // first run: the output arrays are sized from the scan totals (the sizes do not change between runs)
            long long t[3];
            LCD_DRAIN(s);
            for (int j = 0; j < 3; ++j) LCD_CUDA_OK(cudaMemcpyAsync(t + j, d_first.p + j * stride + tot_reads, sizeof(long long), cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
            tot_digar = t[0]; tot_alt = t[1]; tot_ncap = t[2];
            if (d_dpos.alloc(tot_digar + 1) || d_dtype.alloc(tot_digar + 1) || d_dlen.alloc(tot_digar + 1) || d_dqi.alloc(tot_digar + 1) || d_dlow.alloc(tot_digar + 1) ||
                d_daoff.alloc(tot_digar + 1) || d_dalt.alloc(tot_alt + 1) || d_nbeg.alloc(tot_ncap + 1) || d_nend.alloc(tot_ncap + 1) || d_nlabel.alloc(tot_ncap + 1)) return -1;
            args(a);
        }
        digar_fill_kernel<<<grid, THREADS, 0, s>>>(a);
        const long long warps = (long long)c.sm_count * 8 * HIST_WARPS;
        const long long rpw = std::max<long long>(1, (tot_reads + warps - 1) / warps);
        const int hgrid = (int)((tot_reads + rpw * HIST_WARPS - 1) / (rpw * HIST_WARPS));
        digar_hist_kernel<<<hgrid, HIST_WARPS * 32, 0, s>>>(a, rpw);
        LCD_CUDA_OK(cudaGetLastError());
        c.launches += 2;
        return 0;
    }

    int work_units(cudaStream_t, uint64_t *units) override { *units = (uint64_t)tot_qual; return 0; }   // read bases scanned

    int index(cudaStream_t s) {
        if (have_index) return 0;
        if (tot_digar < 0) { set_error("lcd_digar: the plan has not been run"); return -1; }
        h_first.assign(3 * stride, 0); h_nnreg.assign(stride, 0);
        LCD_DRAIN(s);
        int32_t status = 0;
        if (tot_reads) {
            LCD_CUDA_OK(cudaMemcpyAsync(h_first.data(), d_first.p, sizeof(long long) * 3 * stride, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(h_nnreg.data(), d_nnreg.p, sizeof(int32_t) * tot_reads, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaMemcpyAsync(&status, d_status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        if (status == ST_BAD_OP) { set_error("lcd_digar: a read's CIGAR holds an 'M' op; only =/X CIGARs are implemented on the GPU (the reference stops as well: src/bam_utils.c:766)"); return -2; }
        if (status) { set_error("lcd_digar: an output slice overflowed on the device (status %d)", status); return -3; }
        have_index = true;
        return 0;
    }

    int view(cudaStream_t s, DigarView *v) {
        if (tot_reads && index(s)) return -1;
        v->n_chunks = n; v->n_reads_total = tot_reads; v->tot_events = tot_digar < 0 ? 0 : tot_digar;
        v->read_off = read_off; v->alt_base.assign(n, 0); v->min_bq.assign(n, 0); v->h_active = h_active;
        for (int i = 0; i < n; ++i) { v->alt_base[i] = tot_reads ? h_first[stride + read_off[i]] : 0; v->min_bq[i] = chunks[i].min_bq; }
        v->h_beg.assign(stride, 0); v->h_end.assign(stride, 0);
        if (tot_reads && v->want_host_spans) {
            LCD_CUDA_OK(cudaMemcpyAsync(v->h_beg.data(), d_beg.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(v->h_end.data(), d_end.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaStreamSynchronize(s));
        }
        v->active = d_active.p; v->dropped = d_skip.p; v->rev = d_rev.p; v->qual = d_qual.p; v->dlow = d_dlow.p; v->dalt = d_dalt.p;
        v->beg = d_beg.p; v->end = d_end.p; v->dfirst = d_first.p; v->qoff = d_qoff.p; v->dpos = d_dpos.p; v->daoff = d_daoff.p;
        v->nfirst = d_first.p + 2 * stride; v->nbeg = d_nbeg.p; v->nend = d_nend.p;
        v->ndig = d_ndig.p; v->dlen = d_dlen.p; v->dqi = d_dqi.p; v->nnreg = d_nnreg.p; v->dtype = d_dtype.p; v->nlabel = d_nlabel.p;
        v->nreg_total.assign(n, 0); v->reg_beg.assign(n, 0); v->reg_end.assign(n, 0);
        for (int i = 0; i < n; ++i) { long long t = 0; for (long long g = read_off[i]; g < read_off[i + 1]; ++g) t += h_nnreg[g]; v->nreg_total[i] = t; v->reg_beg[i] = reg_beg[i]; v->reg_end[i] = reg_end[i]; }
        return 0;
    }

    int sizes(cudaStream_t s, int i, int64_t *nd, int64_t *na, int64_t *nr) {
        if (i < 0 || i >= n) { set_error("lcd_digar_plan_sizes: chunk %d out of range", i); return -1; }
        if (tot_reads == 0 || tot_digar < 0) { if (tot_reads) { set_error("lcd_digar: the plan has not been run"); return -1; } *nd = *na = *nr = 0; return 0; }
        if (index(s)) return -1;
        const long long g0 = read_off[i], g1 = read_off[i + 1];
        *nd = h_first[g1] - h_first[g0]; *na = h_first[stride + g1] - h_first[stride + g0];
        long long t = 0;
        for (long long g = g0; g < g1; ++g) t += h_nnreg[g];
        *nr = t;
        return 0;
    }

    int fetch(cudaStream_t s, lcd_digar_output_t *out) {
        if (n == 0) return 0;
        if (tot_reads == 0) {
            for (int i = 0; i < n; ++i) { memset(out[i].qual_counts, 0, sizeof(int64_t) * 256); out[i].n_cnreg = out[i].n_digar_total = out[i].n_alt_total = out[i].n_nreg_total = 0; }
            return 0;
        }
        if (index(s)) return -1;
        LCD_DRAIN(s);
        std::vector<uint8_t> skip(stride); std::vector<long long> beg(stride), end(stride), nb(tot_ncap + 1), ne(tot_ncap + 1); std::vector<int32_t> nl(tot_ncap + 1);
        std::vector<unsigned long long> qc(256 * (size_t)n);
        LCD_CUDA_OK(cudaMemcpyAsync(skip.data(), d_skip.p, tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(beg.data(), d_beg.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
        LCD_CUDA_OK(cudaMemcpyAsync(end.data(), d_end.p, sizeof(long long) * tot_reads, cudaMemcpyDeviceToHost, s));
        if (tot_ncap) {
            LCD_CUDA_OK(cudaMemcpyAsync(nb.data(), d_nbeg.p, sizeof(long long) * tot_ncap, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(ne.data(), d_nend.p, sizeof(long long) * tot_ncap, cudaMemcpyDeviceToHost, s));
            LCD_CUDA_OK(cudaMemcpyAsync(nl.data(), d_nlabel.p, sizeof(int32_t) * tot_ncap, cudaMemcpyDeviceToHost, s));
        }
        LCD_CUDA_OK(cudaMemcpyAsync(qc.data(), d_qc.p, sizeof(unsigned long long) * 256 * (size_t)n, cudaMemcpyDeviceToHost, s));
        // the difference lists go straight into the caller's arrays, chunk by chunk (the device layout is chunk-contiguous)
        for (int i = 0; i < n; ++i) {
            const long long g0 = read_off[i], g1 = read_off[i + 1];
            const long long d0 = h_first[g0], nd = h_first[g1] - d0, a0 = h_first[stride + g0], na = h_first[stride + g1] - a0;
            if (nd > out[i].digar_cap || na > out[i].alt_cap) {
                cudaStreamSynchronize(s);
                set_error("lcd_digar: chunk %d needs %lld records / %lld alt bases, the caller provided %lld / %lld (lcd_digar_capacity, lcd_digar_plan_sizes)", i, nd, na, (long long)out[i].digar_cap, (long long)out[i].alt_cap);
                return -3;
            }
            if (nd) {
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_pos, d_dpos.p + d0, sizeof(long long) * nd, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_type, d_dtype.p + d0, nd, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_len, d_dlen.p + d0, sizeof(int32_t) * nd, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_qi, d_dqi.p + d0, sizeof(int32_t) * nd, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_low_qual, d_dlow.p + d0, nd, cudaMemcpyDeviceToHost, s));
                LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_alt_off, d_daoff.p + d0, sizeof(long long) * nd, cudaMemcpyDeviceToHost, s));
            }
            if (na) LCD_CUDA_OK(cudaMemcpyAsync(out[i].digar_alt, d_dalt.p + a0, na, cudaMemcpyDeviceToHost, s));
            out[i].n_digar_total = nd; out[i].n_alt_total = na;
        }
        LCD_CUDA_OK(cudaStreamSynchronize(s));
        std::vector<Intv> tmp;
        for (int i = 0; i < n; ++i) {
            lcd_digar_output_t &o = out[i];
            const long long g0 = read_off[i], nr = read_off[i + 1] - g0, d0 = h_first[g0];
            long long top = 0;
            for (long long r = 0; r < nr; ++r) {
                const long long g = g0 + r; const int k = h_nnreg[g];
                o.skip[r] = skip[g]; o.read_beg[r] = h_active[g] ? beg[g] : 0; o.read_end[r] = h_active[g] ? end[g] : 0;
                o.digar_first[r] = h_first[g] - d0; o.n_digar[r] = (int32_t)(h_first[g + 1] - h_first[g]);
                o.nreg_first[r] = top; o.n_nreg[r] = k;
                if (top + k > o.nreg_cap) { set_error("lcd_digar: chunk %d needs more than %lld interval slots", i, (long long)o.nreg_cap); return -3; }
                const long long f = h_first[2 * stride + g];
                if (k > 64) {       // cr_index of a large interval set: cgranges' own (unstable) radix sort on the cr_add order the device kept
                    tmp.resize(k);
                    for (int x = 0; x < k; ++x) { tmp[x].key = (uint64_t)(uint32_t)(int32_t)nb[f + x]; tmp[x].beg = nb[f + x]; tmp[x].end = ne[f + x]; tmp[x].label = nl[f + x]; }
                    intv_radix_sort(tmp.data(), tmp.data() + k, 56);
                    for (int x = 0; x < k; ++x) { o.nreg_beg[top + x] = tmp[x].beg; o.nreg_end[top + x] = tmp[x].end; o.nreg_label[top + x] = tmp[x].label; }
                } else
                    for (int x = 0; x < k; ++x) { o.nreg_beg[top + x] = nb[f + x]; o.nreg_end[top + x] = ne[f + x]; o.nreg_label[top + x] = nl[f + x]; }
                top += k;
            }
            o.n_nreg_total = top;
            // what collect_digar_from_eqx_cigar adds to chunk->chunk_noisy_regs (src/bam_utils.c:819-832): kept reads in ordered_read_ids order
            o.n_cnreg = 0;
            for (long long x = 0; x < nr; ++x) {
                const int r = ordered[g0 + x];
                if (r < 0 || r >= nr || !h_active[g0 + r] || skip[g0 + r]) continue;
                for (long long y = o.nreg_first[r]; y < o.nreg_first[r] + o.n_nreg[r]; ++y)
                    if (!(o.nreg_beg[y] + 1 > reg_end[i] || o.nreg_end[y] < reg_beg[i])) {
                        if (o.n_cnreg >= o.cnreg_cap) { set_error("lcd_digar: chunk %d needs more than %lld chunk interval slots", i, (long long)o.cnreg_cap); return -3; }
                        o.cnreg_beg[o.n_cnreg] = o.nreg_beg[y]; o.cnreg_end[o.n_cnreg] = o.nreg_end[y]; o.cnreg_label[o.n_cnreg] = o.nreg_label[y]; o.n_cnreg++;
                    }
            }
            for (int b = 0; b < 256; ++b) o.qual_counts[b] = (int64_t)qc[256 * (size_t)i + b];
        }
        return 0;
    }
};

} // namespace digar

int digar_plan_view(Plan *plan, cudaStream_t s, DigarView *v) {
    digar::DigarPlan *p = dynamic_cast<digar::DigarPlan *>(plan);
    if (!p) { set_error("not a digar plan"); return -1; }
    if (p->n && p->tot_reads && p->tot_digar < 0) { set_error("the digar plan has not been run"); return -1; }
    return p->view(s, v);
}
} // namespace lcd

using namespace lcd;

extern "C" {

int lcd_digar_capacity(const lcd_digar_input_t *in, int64_t *digar_cap, int64_t *alt_cap, int64_t *nreg_cap) {
    if (!in || !digar_cap || !alt_cap || !nreg_cap) { set_error("lcd_digar_capacity: null arguments"); return -1; }
    long long nd = 0, na = 0, ni = 0;
    for (int r = 0; r < in->n_reads; ++r) {
        const uint32_t *cg = in->cigar + in->cigar_off[r];
        for (int k = 0; k < in->n_cigar[r]; ++k) {
            const int op = cg[k] & 15; const long long len = cg[k] >> 4;
            if (op == digar::CDIFF) { nd += len; na += len; ni += len; }
            else if (op == digar::CINS) { nd++; na += len; ni++; }
            else if (op == digar::CDEL) { nd++; ni++; }
            else if (op == digar::CEQUAL || op == digar::CSOFT || op == digar::CHARD) nd++;
        }
        ni += 2;
    }
    *digar_cap = nd + 1; *alt_cap = na + 1; *nreg_cap = ni + 1;
    return 0;
}

lcd_plan_t *lcd_digar_md_plan_create(int n_chunks, const lcd_digar_input_t *in, const lcd_md_tags_t *tags) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && (!in || !tags))) { set_error("lcd_digar_md_plan_create: invalid arguments"); return nullptr; }
    digar::DigarPlan *p = new digar::DigarPlan();
    if (p->build(n_chunks, in, tags)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_digar_md_batch(int n_chunks, const lcd_digar_input_t *in, const lcd_md_tags_t *tags, lcd_digar_output_t *out) {
    lcd_plan_t *plan = lcd_digar_md_plan_create(n_chunks, in, tags);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_digar_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

lcd_plan_t *lcd_digar_tags_plan_create(int n_chunks, const lcd_digar_input_t *in, const lcd_read_tags_t *tags) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && (!in || !tags))) { set_error("lcd_digar_tags_plan_create: invalid arguments"); return nullptr; }
    for (int i = 0; i < n_chunks; ++i) if (in[i].n_reads > 0 && !tags[i].kind) { set_error("lcd_digar_tags_plan_create: chunk %d has no kind array", i); return nullptr; }
    digar::DigarPlan *p = new digar::DigarPlan();
    if (p->build(n_chunks, in, nullptr, tags)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_digar_tags_batch(int n_chunks, const lcd_digar_input_t *in, const lcd_read_tags_t *tags, lcd_digar_output_t *out) {
    lcd_plan_t *plan = lcd_digar_tags_plan_create(n_chunks, in, tags);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_digar_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

lcd_plan_t *lcd_digar_plan_create(int n_chunks, const lcd_digar_input_t *in) {
    if (ensure_ready()) return nullptr;
    if (n_chunks < 0 || (n_chunks > 0 && !in)) { set_error("lcd_digar_plan_create: invalid arguments"); return nullptr; }
    digar::DigarPlan *p = new digar::DigarPlan();
    if (p->build(n_chunks, in)) { delete p; return nullptr; }
    return reinterpret_cast<lcd_plan_t *>(p);
}

int lcd_digar_plan_sizes(lcd_plan_t *plan, void *stream, int chunk, int64_t *n_digar, int64_t *n_alt, int64_t *n_nreg) {
    digar::DigarPlan *p = dynamic_cast<digar::DigarPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !n_digar || !n_alt || !n_nreg) { set_error("lcd_digar_plan_sizes: not a digar plan / null outputs"); return -1; }
    return p->sizes(pick_stream(stream), chunk, n_digar, n_alt, n_nreg);
}

int lcd_digar_plan_fetch(lcd_plan_t *plan, void *stream, lcd_digar_output_t *out) {
    digar::DigarPlan *p = dynamic_cast<digar::DigarPlan *>(reinterpret_cast<Plan *>(plan));
    if (!p || !out) { set_error("lcd_digar_plan_fetch: not a digar plan / null outputs"); return -1; }
    return p->fetch(pick_stream(stream), out);
}

int lcd_digar_batch(int n_chunks, const lcd_digar_input_t *in, lcd_digar_output_t *out) {
    lcd_plan_t *plan = lcd_digar_plan_create(n_chunks, in);
    if (!plan) return -1;
    int rc = lcd_plan_run(plan, nullptr);
    if (!rc) rc = lcd_digar_plan_fetch(plan, nullptr, out);
    lcd_plan_destroy(plan);
    return rc;
}

}
