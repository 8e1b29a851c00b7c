// lcd_common.cuh -- shared host/device plumbing of liblcd_gpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>
#include "../../include/lcd_gpu.h"

namespace lcd {

void set_error(const char *fmt, ...);
#define LCD_CUDA_OK(call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            lcd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)
#define LCD_CUDA_OK_PTR(call)                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            lcd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return nullptr;                                                                    \
        }                                                                                      \
    } while (0)

// One context per process: device, library stream, SM count, and the workspace pool that the DP
// kernels carve wavefronts / DP planes from.  The pool is one allocation: [private arenas | overflow].
struct Context {
    std::atomic<bool> ready{false};
    int device = 0;
    int sm_count = 148;
    int reserved_sms = 0;             // CTA slots of this many SMs are left free by the persistent DP grids (lcd_gpu_reserve_sms)
    int dp_sms() const { return sm_count - reserved_sms > 8 ? sm_count - reserved_sms : 8; }
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;  // created on first lcd_gpu_aux_stream()
    std::vector<cudaStream_t> extra_streams;   // lcd_gpu_new_stream: one per host thread of a multi-threaded caller
    int32_t *pool = nullptr;          // int32 words
    size_t pool_words = 0;
    // The pool is cut into windows; a plan of the DP engines carves its workspace from ONE window for the time of a run.  Plans come in
    // two classes -- 0: POA (K5), 1: WFA / edlib (K6, K7) -- and every class has its list of windows (lcd_gpu_init: one window for both;
    // lcd_gpu_split_pool: one each; lcd_gpu_pool_windows: several each, so that several batches of one engine are in flight together).
    // Every window orders its own users: the event is recorded behind the last launch that carves from the window and the next run() on
    // that window -- whatever stream it was given -- waits for it first.  (The mutex only covers the enqueue; the persistent grids keep
    // using the window after run() has returned.)
    static constexpr int BITMAP_WORDS = 4096;     // overflow chunks in use (bit set), device memory, one bitmap per window
    struct Window { size_t off = 0, words = 0; std::mutex mu; cudaEvent_t done = nullptr; uint32_t *bitmap = nullptr; };
    std::vector<Window*> windows;
    std::vector<int> cls_win[2];
    std::atomic<unsigned> rr[2];
    size_t class_words(int cls) const { return windows[cls_win[cls][0]]->words; }       // windows of a class have one size
    Window *pick_window(int cls);     // an idle window of the class (its last user's event has fired), else the next one in turn
    int set_windows(int n0, int n1, size_t lower_words);
    std::mutex mu;                    // guards the context's own lists (streams, windows)
    size_t requested_pool_bytes = 0;  // what lcd_gpu_init was first called with (0: default)
    std::atomic<unsigned long long> launches{0};
};
Context &ctx();
int ensure_ready();
// The stream a host thread's plans upload, run and fetch on when it passes no stream: the library stream, unless the thread chose
// another one with lcd_gpu_set_thread_stream (e.g. the auxiliary stream, to overlap one stage's copies with another's kernels).
cudaStream_t &thread_stream();
// A host thread other than the one that called lcd_gpu_init starts on device 0: every entry point binds the calling thread to the
// library's device first (once per thread).
void bind_thread();
inline cudaStream_t cur_stream() { bind_thread(); cudaStream_t t = thread_stream(); return t ? t : ctx().stream; }

struct Plan {
    virtual ~Plan() {}
    virtual int run(cudaStream_t s) = 0;
    virtual int work_units(cudaStream_t s, uint64_t *units) = 0;
    // plans whose kernels carve workspace from the context's pool are serialised by lcd_plan_run; the others (K1, K1b, K2, K3, K4:
    // their buffers are their own) may run from another host thread / stream while a pool plan is in flight
    virtual bool uses_pool() const { return true; }
    virtual int pool_window() const { return 0; }       // the plan's class: 0 POA, 1 WFA / edlib
    Context::Window *win = nullptr;                     // the window lcd_plan_run chose for the current run
    // completes what run() left pending on the stream (default: nothing beyond draining it); called by lcd_plan_sync and the fetches
    virtual int finish(cudaStream_t s) { (void)s; return 0; }       // which window of a split pool the plan's kernels carve from
    int n = 0;
};

// Device buffer from the stream-ordered pool of the library stream (cudaMallocAsync; the pool keeps freed
// memory, so the per-call plans of the *_batch entry points do not pay cudaMalloc / cudaFree every time).
// Every plan synchronises the library stream after its uploads, which orders the allocation before any use on
// another stream; buffers are released when the plan is destroyed, after its last fetch has synchronised.
template <typename T>
struct DevBuf {
    T *p = nullptr; size_t n = 0; cudaStream_t st = nullptr;     // st: the stream the buffer was allocated on (and is freed on)
    int alloc(size_t count) {
        free_(); n = count;
        if (count == 0) return 0;
        st = cur_stream();
        cudaError_t e = cudaMallocAsync((void**)&p, count * sizeof(T), st);
        if (e != cudaSuccess) { set_error("cudaMallocAsync(%zu) -> %s", count * sizeof(T), cudaGetErrorString(e)); p = nullptr; return -1; }
        return 0;
    }
    int upload(const T *h, size_t count, cudaStream_t s) {
        if (alloc(count)) return -1;
        if (count == 0) return 0;
        cudaError_t e = cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st);
        (void)s;
        if (e != cudaSuccess) { set_error("H2D -> %s", cudaGetErrorString(e)); return -1; }
        return 0;
    }
    void free_() { if (p) cudaFreeAsync(p, st); p = nullptr; n = 0; }
    ~DevBuf() { free_(); }
};

// K1's results as K2 / K3 consume them in place (device pointers into a digar plan that has been run; valid while it lives).
struct DigarView {
    int n_chunks = 0; long long n_reads_total = 0, tot_events = 0;
    std::vector<long long> read_off, alt_base, h_beg, h_end; std::vector<int32_t> min_bq; std::vector<uint8_t> h_active;
    const uint8_t *active = nullptr, *dropped = nullptr, *rev = nullptr, *qual = nullptr, *dlow = nullptr, *dalt = nullptr;
    const long long *beg = nullptr, *end = nullptr, *dfirst = nullptr, *qoff = nullptr, *dpos = nullptr, *daoff = nullptr, *nfirst = nullptr, *nbeg = nullptr, *nend = nullptr;
    const int32_t *ndig = nullptr, *dlen = nullptr, *dqi = nullptr, *nnreg = nullptr, *nlabel = nullptr; const int8_t *dtype = nullptr;
    bool want_host_spans = true;              // h_beg / h_end (a D2H copy of every read's span): consumers that work on the device only switch it off
    std::vector<long long> nreg_total;        // per chunk: noisy intervals of its reads
    std::vector<long long> reg_beg, reg_end;  // per chunk: the region
};
int digar_plan_view(Plan *plan, cudaStream_t s, DigarView *v);      // digar_kernel.cu

// K1b's site lists as K2 consumes them in place (device pointers into a sites plan that has been run; valid while it lives).
struct SitesView {
    int n_chunks = 0; std::vector<long long> site_off;      // n_chunks + 1 entries: chunk i's sites are [site_off[i], site_off[i + 1])
    std::vector<int32_t> min_sv_len;
    const long long *spos = nullptr, *saoff = nullptr; const int32_t *stype = nullptr, *sref = nullptr, *salt = nullptr;
};
int sites_plan_view(Plan *plan, Plan *digar, cudaStream_t s, SitesView *v);   // sites_kernel.cu

// Host <-> device copies of pageable memory go through one staging path of the driver: a copy queued BEHIND a long kernel holds that path
// until the kernel ends, and every other thread's copies wait with it (measured on B200: a second host thread's 2 MB D2H took 250 ms
// while a POA launch with its status copy queued behind it was in flight).  So results are read back only after the stream has drained.
#define LCD_DRAIN(s) LCD_CUDA_OK(cudaStreamSynchronize(s))

// K2's sites and counters as K2b consumes them in place (device pointers into a pileup plan; valid while it -- and the plans it was
// created on -- live).  site_alt_off is relative to salt_base[chunk] in site_alt.
struct PileupView {
    int n_chunks = 0; std::vector<long long> site_off, salt_base;
    const long long *spos = nullptr, *saoff = nullptr; const int32_t *stype = nullptr, *sref = nullptr, *salt = nullptr, *counts = nullptr; const uint8_t *site_alt = nullptr;
};
int pileup_plan_view(Plan *plan, PileupView *v);      // pileup_kernel.cu

// K2b's sites and categories as K2c consumes them in place (device pointers into a classify plan; valid while it -- and the plans it was
// created on -- live; the categories are there once the plan has run on the stream K2c runs on, or an earlier one it waits for).
struct ClassifyView {
    int n_chunks = 0; std::vector<long long> site_off;
    const long long *spos = nullptr; const int32_t *stype = nullptr, *sref = nullptr, *cate = nullptr;
};
int classify_plan_view(Plan *plan, ClassifyView *v);  // classify_kernel.cu

// K0's intervals as K2c consumes them in place (per chunk: device arrays of starts / ends in ascending order and the device word that holds their number).
struct SdustView {
    int n_chunks = 0; std::vector<const long long *> beg, end, n_out; std::vector<const int *> status; std::vector<long long> cap;
};
int sdust_plan_view(Plan *plan, SdustView *v);        // sdust_kernel.cu

inline cudaStream_t pick_stream(void *s) { bind_thread(); return s ? (cudaStream_t)s : cur_stream(); }

} // namespace lcd
