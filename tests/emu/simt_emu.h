// tests/emu/simt_emu.h -- TEST INFRASTRUCTURE: a one-warp SIMT emulator for the host.
//
// cuda_emu.h runs the product's device logic with ONE lane (collectives are identities), which cannot exercise
// the warp-cooperative code (strip rows, ballot backtrack, parallel fusion, shuffles of the packed DP).  This
// header runs the real warp code: 32 lanes are 32 fibers (hand-rolled x86-64 context switch, one OS thread),
// every *_sync collective is a rendezvous of the live lanes, and memory is plain host memory.  The scheduling
// is deterministic (round robin, lanes only switch at collectives), which matches the guarantees the device
// code may rely on: values written before a collective are visible after it.
//
// Use: include this header INSTEAD of cuda_emu.h (LCD_EMU stays undefined, so the GPU-only sections of the
// *_device.cuh headers are compiled), then   simt::run_warp([&] { ...code using threadIdx.x... });
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <algorithm>
#include <functional>
#define LCD_SIMT_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __restrict__
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
using std::max;
using std::min;
struct simt_dim3 { unsigned x, y, z; };
static simt_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};

namespace simt {
constexpr int NL = 32;
constexpr size_t STACK = 1 << 20;
struct Fiber { void *sp; char *stack; bool alive; };
static Fiber fib[NL];
static void *main_sp;
static int cur = -1, alive = 0, arrived = 0;
static unsigned gen = 0;
static uint64_t xbuf[2][NL];
static std::function<void()> *body;

extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
)");

static void switch_to(int next) {            // from a lane (or main, cur == -1) to lane `next` (or main, -1)
    void **save = cur < 0 ? &main_sp : &fib[cur].sp;
    void *load = next < 0 ? main_sp : fib[next].sp;
    cur = next;
    threadIdx.x = next < 0 ? 0 : (unsigned)next;
    simt_switch(save, load);
    // back in the saved context: cur / threadIdx were set by whoever switched to us
}
static int next_alive(int from) {
    for (int k = 1; k <= NL; ++k) { const int l = (from + k) % NL; if (fib[l].alive) return l; }
    return -1;
}
static void yield() { const int n = next_alive(cur); if (n >= 0 && n != cur) switch_to(n); }
static void barrier() {
    if (++arrived >= alive) { arrived = 0; ++gen; return; }
    const unsigned my = gen;
    while (gen == my) yield();
}
static void entry() {
    (*body)();
    fib[cur].alive = false; --alive;
    if (alive > 0 && arrived >= alive) { arrived = 0; ++gen; }      // the lanes still waiting need not wait for an exited one
    const int n = next_alive(cur);
    switch_to(n);                                                      // never returns here
    abort();
}
static void run_warp(std::function<void()> f) {
    body = &f; alive = NL; arrived = 0; gen = 0;
    for (int l = 0; l < NL; ++l) {
        if (!fib[l].stack) fib[l].stack = (char *)aligned_alloc(64, STACK);
        uintptr_t top = ((uintptr_t)fib[l].stack + STACK) & ~(uintptr_t)63;
        void **sp = (void **)top;
        *--sp = nullptr;                       // fake return address of entry (keeps rsp % 16 == 8 at entry, as after a call)
        *--sp = (void *)entry;                 // ret target of simt_switch
        for (int k = 0; k < 6; ++k) *--sp = nullptr;   // r15 r14 r13 r12 rbx rbp
        fib[l].sp = sp; fib[l].alive = true;
    }
    cur = -1;
    switch_to(0);
    threadIdx.x = 0;
}
template <typename T> static T exch(T v, int src) {
    static_assert(sizeof(T) <= 8, "exchange of at most 8 bytes");
    const int g = gen & 1;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); xbuf[g][cur] = raw;
    barrier();
    T r; memcpy(&r, &xbuf[g][src & 31], sizeof(T));
    return r;
}
template <typename T, typename F> static T reduce_all(T v, F f) {
    const int g = gen & 1;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); xbuf[g][cur] = raw;
    barrier();
    bool first = true; T acc = v;
    for (int l = 0; l < NL; ++l) if (fib[l].alive) { T x; memcpy(&x, &xbuf[g][l], sizeof(T)); acc = first ? x : f(acc, x); first = false; }
    return acc;
}
} // namespace simt

template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return simt::exch(v, src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int l = simt::cur; return simt::exch(v, l - d < 0 ? l : l - d); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { const int l = simt::cur; return simt::exch(v, l + d > 31 ? l : l + d); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exch(v, simt::cur ^ m); }
static inline unsigned __ballot_sync(unsigned, int p) {
    const int g = simt::gen & 1;
    simt::xbuf[g][simt::cur] = p ? 1 : 0;
    simt::barrier();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (simt::fib[l].alive && simt::xbuf[g][l]) m |= 1u << l;
    return m;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
static inline int __reduce_max_sync(unsigned, int v) { return simt::reduce_all(v, [](int a, int b) { return a > b ? a : b; }); }
static inline int __reduce_min_sync(unsigned, int v) { return simt::reduce_all(v, [](int a, int b) { return a < b ? a : b; }); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return simt::reduce_all(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return simt::reduce_all(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return simt::reduce_all(v, [](unsigned a, unsigned b) { return a | b; }); }
static inline unsigned __reduce_and_sync(unsigned, unsigned v) { return simt::reduce_all(v, [](unsigned a, unsigned b) { return a & b; }); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return simt::reduce_all(v, [](unsigned a, unsigned b) { return a + b; }); }
static inline int __reduce_add_sync(unsigned, int v) { return simt::reduce_all(v, [](int a, int b) { return a + b; }); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier(); }
static inline void __syncthreads() { simt::barrier(); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline long long clock64() { return 0; }
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p += v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p |= v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p &= v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- scalar bit / SIMD-in-word intrinsics (semantics of the CUDA math API) ----
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __ffsll(unsigned long long x) { return __builtin_ffsll((long long)x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __clzll(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) { if (sh > 32) sh = 32; const unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << sh) >> 32); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) { if (sh > 32) sh = 32; const unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> sh); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) { const unsigned sel = (s >> (4 * i)) & 0xf; unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff; if (sel & 8) byte = (byte & 0x80) ? 0xff : 0; r |= byte << (8 * i); }
    return r;
}
#define SIMT_H2(expr_lo, expr_hi) ((unsigned)(uint16_t)(expr_lo) | ((unsigned)(uint16_t)(expr_hi) << 16))
#define SIMT_LO(x) ((int)(int16_t)((x) & 0xffffu))
#define SIMT_HI(x) ((int)(int16_t)((x) >> 16))
static inline unsigned __vadd2(unsigned a, unsigned b) { return SIMT_H2(SIMT_LO(a) + SIMT_LO(b), SIMT_HI(a) + SIMT_HI(b)); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return SIMT_H2(SIMT_LO(a) - SIMT_LO(b), SIMT_HI(a) - SIMT_HI(b)); }
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return SIMT_H2(std::max(SIMT_LO(a), SIMT_LO(b)), std::max(SIMT_HI(a), SIMT_HI(b))); }
static inline unsigned __vmins2(unsigned a, unsigned b) { return SIMT_H2(std::min(SIMT_LO(a), SIMT_LO(b)), std::min(SIMT_HI(a), SIMT_HI(b))); }
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool *ph, bool *pl) { *pl = SIMT_LO(a) >= SIMT_LO(b); *ph = SIMT_HI(a) >= SIMT_HI(b); return __vmaxs2(a, b); }
static inline unsigned __vcmpeq2(unsigned a, unsigned b) { return ((a & 0xffffu) == (b & 0xffffu) ? 0xffffu : 0u) | ((a >> 16) == (b >> 16) ? 0xffff0000u : 0u); }
static inline unsigned __vcmpgts2(unsigned a, unsigned b) { return (SIMT_LO(a) > SIMT_LO(b) ? 0xffffu : 0u) | (SIMT_HI(a) > SIMT_HI(b) ? 0xffff0000u : 0u); }
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
static inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
