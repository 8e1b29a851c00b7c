// scan.cuh -- device-wide exclusive scan of int32 counts into int64 offsets (out[0 .. n], out[n] = total), used for the
// counting-sort bins of the candidate-site list.  Three launches: per-tile sums, one CTA over the tile sums, per-tile rescan.
// Every thread owns 8 consecutive elements (32-byte loads, 64-byte stores), tiles of 2048.
#pragma once
#include "lcd_common.cuh"

namespace lcd {
namespace scan {

constexpr int THREADS = 256, PER_THREAD = 8, TILE = THREADS * PER_THREAD;
constexpr int TS = 256;          // threads of the one CTA that scans the tile sums (small: it has to fit next to a resident persistent DP grid)

__device__ __forceinline__ long long block_exclusive(long long v, long long *smem, long long *total) {     // smem: THREADS / 32 + 1 entries
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < THREADS / 32 ? smem[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        if (lane < THREADS / 32) smem[lane] = wi - w;
        if (lane == THREADS / 32 - 1) smem[THREADS / 32] = wi;
    }
    __syncthreads();
    const long long r = smem[warp] + incl - v;
    if (total) *total = smem[THREADS / 32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(THREADS)
tile_sum_kernel(const int32_t *in, long long n, long long *tile_sum) {
    __shared__ long long sm[THREADS / 32 + 1];
    const long long i0 = (long long)blockIdx.x * TILE + (long long)threadIdx.x * PER_THREAD;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) if (i0 + k < n) s += in[i0 + k];
    long long tot;
    block_exclusive(s, sm, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(TS)
tile_scan_kernel(long long *tile_sum, long long n_tiles) {      // in place: tile_sum[t] <- sum of the tiles before t; tile_sum[n_tiles] <- total
    __shared__ long long warp_sum[TS / 32];
    const long long seg = (n_tiles + TS - 1) / TS;
    const long long i0 = min(n_tiles, seg * (long long)threadIdx.x), i1 = min(n_tiles, i0 + seg);
    long long s = 0;
    for (long long i = i0; i < i1; ++i) s += tile_sum[i];
    long long incl = s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < TS / 32 ? warp_sum[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        if (lane < TS / 32) warp_sum[lane] = wi - w;
    }
    __syncthreads();
    long long run = warp_sum[warp] + incl - s;
    for (long long i = i0; i < i1; ++i) { const long long v = tile_sum[i]; tile_sum[i] = run; run += v; }
    if (threadIdx.x == TS - 1) tile_sum[n_tiles] = run;
}

__global__ void __launch_bounds__(THREADS)
tile_rescan_kernel(const int32_t *in, long long n, const long long *tile_off, long long *out) {
    __shared__ long long sm[THREADS / 32 + 1];
    const long long i0 = (long long)blockIdx.x * TILE + (long long)threadIdx.x * PER_THREAD;
    int v[PER_THREAD]; long long s = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) { v[k] = i0 + k < n ? in[i0 + k] : 0; s += v[k]; }
    long long run = tile_off[blockIdx.x] + block_exclusive(s, sm, nullptr);
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_off[gridDim.x];
}

// out must hold n + 1 entries, tmp at least n / TILE + 2
inline int exclusive_scan(const int32_t *in, long long n, long long *out, DevBuf<long long> &tmp, cudaStream_t s) {
    const long long n_tiles = (n + TILE - 1) / TILE;
    if (n_tiles == 0) { LCD_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(long long), s)); return 0; }
    if (tmp.n < (size_t)n_tiles + 2 && tmp.alloc((size_t)n_tiles + 2)) return -1;
    tile_sum_kernel<<<(unsigned)n_tiles, THREADS, 0, s>>>(in, n, tmp.p);
    tile_scan_kernel<<<1, TS, 0, s>>>(tmp.p, n_tiles);
    tile_rescan_kernel<<<(unsigned)n_tiles, THREADS, 0, s>>>(in, n, tmp.p, out);
    LCD_CUDA_OK(cudaGetLastError());
    ctx().launches += 3;
    return 0;
}

} // namespace scan
} // namespace lcd
