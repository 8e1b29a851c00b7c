// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE: single-lane host stand-ins for the CUDA intrinsics
// used by longcalld_b200/csrc/*_device.cuh, so that the product's device logic can be compiled with
// g++ (group size G = 1) and checked against the oracle on the CPU box (and under ASAN/valgrind).
// It checks LOGIC only -- races, shuffles and barriers are exercised by the -m gpu tests.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#define LCD_EMU 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
using std::max;
using std::min;
static inline unsigned __vadd2(unsigned a, unsigned b) { return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return ((a - b) & 0xffffu) | (((a >> 16) - (b >> 16)) << 16); }
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
    const int16_t al = (int16_t)a, bl = (int16_t)b, ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
    return (uint16_t)(al > bl ? al : bl) | ((unsigned)(uint16_t)(ah > bh ? ah : bh) << 16);
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __ffsll(unsigned long long x) { return __builtin_ffsll((long long)x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __clzll(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline int __reduce_max_sync(unsigned, int v) { return v; }
static inline int __reduce_min_sync(unsigned, int v) { return v; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline void __threadfence_block() {}
static inline void __threadfence() {}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p += v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p |= v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p &= v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
