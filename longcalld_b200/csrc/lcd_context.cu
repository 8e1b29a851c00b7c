// lcd_context.cu -- context, error reporting and plan plumbing of liblcd_gpu.so.
#include <stdarg.h>
#include "lcd_common.cuh"

namespace lcd {

static thread_local char g_err[512] = "";
static char g_err_global[512] = "";
static std::mutex g_err_mu;

void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(g_err_mu);
    strncpy(g_err_global, g_err, sizeof(g_err_global) - 1);
}

Context &ctx() { static Context c; return c; }

// n0 windows for class 0 in [0, lower_words), n1 for class 1 in the rest; n1 == 0: ONE window over the whole pool for both classes.
// Called with c.mu held, before any plan runs (the windows of plans in flight would be pulled from under them).
int Context::set_windows(int n0, int n1, size_t lower_words) {
    for (Window *w : windows) { if (w->done) cudaEventDestroy(w->done); if (w->bitmap) cudaFree(w->bitmap); delete w; }
    windows.clear(); cls_win[0].clear(); cls_win[1].clear(); rr[0] = 0; rr[1] = 0;
    auto add = [&](size_t off, size_t words) -> int {
        Window *w = new Window(); w->off = off; w->words = words;
        if (cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming) != cudaSuccess || cudaMalloc((void**)&w->bitmap, sizeof(uint32_t) * BITMAP_WORDS) != cudaSuccess ||
            cudaMemset(w->bitmap, 0, sizeof(uint32_t) * BITMAP_WORDS) != cudaSuccess) { set_error("pool window: CUDA allocation failed"); delete w; return -1; }
        windows.push_back(w); return (int)windows.size() - 1;
    };
    if (n1 == 0) { const int i = add(0, pool_words); if (i < 0) return -1; cls_win[0].push_back(i); cls_win[1].push_back(i); return 0; }
    const size_t w0 = (lower_words / n0) & ~(size_t)63, w1 = ((pool_words - lower_words) / n1) & ~(size_t)63;
    for (int k = 0; k < n0; ++k) { const int i = add((size_t)k * w0, w0); if (i < 0) return -1; cls_win[0].push_back(i); }
    for (int k = 0; k < n1; ++k) { const int i = add(lower_words + (size_t)k * w1, w1); if (i < 0) return -1; cls_win[1].push_back(i); }
    return 0;
}
Context::Window *Context::pick_window(int cls) {
    const std::vector<int> &v = cls_win[cls];
    const unsigned start = rr[cls]++;
    for (size_t k = 0; k < v.size(); ++k) {
        Window *w = windows[v[(start + k) % v.size()]];
        if (cudaEventQuery(w->done) == cudaSuccess) return w;
    }
    return windows[v[start % v.size()]];
}
cudaStream_t &thread_stream() { static thread_local cudaStream_t s = nullptr; return s; }

void bind_thread() {
    static thread_local int bound = -1;
    Context &c = ctx();
    if (!c.ready || bound == c.device) return;
    if (cudaSetDevice(c.device) == cudaSuccess) bound = c.device;
}

int ensure_ready() {
    if (ctx().ready) { bind_thread(); return 0; }
    return lcd_gpu_init(0, 0);
}

} // namespace lcd

using namespace lcd;

extern "C" {

int lcd_gpu_abi_version(void) { return LCD_GPU_ABI_VERSION; }

// Grows the device's stream-ordered memory pool (what every plan's buffers come from) by `bytes` once: a block freed on one stream is not handed to another stream
// before that stream has synchronised, so host threads that create plans side by side now and then find no reusable block and make the pool grow -- and growing it
// while a persistent grid is resident stalled every allocating thread until the grid retired (measured: one e2e step in ~15 of bench.py at 450 ms instead of 320 ms,
// all four host threads blocked for the length of the POA launch).  With a reserve mapped beforehand the pool hands out cached memory instead.
int lcd_gpu_reserve_plan_memory(size_t bytes) {
    if (ensure_ready()) return -1;
    Context &c = ctx();
    size_t free_b = 0, total_b = 0;
    LCD_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    if (bytes > free_b / 2) bytes = free_b / 2;
    if (bytes < ((size_t)1 << 20)) return 0;
    void *p = nullptr;
    LCD_CUDA_OK(cudaMallocAsync(&p, bytes, c.stream));
    LCD_CUDA_OK(cudaFreeAsync(p, c.stream));
    LCD_CUDA_OK(cudaStreamSynchronize(c.stream));
    return 0;
}
const char *lcd_gpu_last_error(void) {
    if (g_err[0]) return g_err;
    std::lock_guard<std::mutex> lk(g_err_mu);                 // another thread's last message, copied into this thread's buffer
    strncpy(g_err, g_err_global, sizeof(g_err) - 1);
    return g_err;
}
uint64_t lcd_gpu_launch_count(void) { return ctx().launches; }
void *lcd_gpu_stream(void) { return ctx().ready ? (void*)ctx().stream : nullptr; }
void *lcd_gpu_aux_stream(void) {
    Context &c = ctx();
    if (!c.ready) return nullptr;
    bind_thread();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.aux_stream && cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("lcd_gpu_aux_stream: cudaStreamCreate failed"); return nullptr; }
    return (void*)c.aux_stream;
}
int lcd_gpu_pool_windows(int n_poa, int n_aln, size_t lower_bytes) {
    Context &c = ctx();
    if (!c.ready) { set_error("lcd_gpu_pool_windows: call lcd_gpu_init first"); return -1; }
    std::lock_guard<std::mutex> lk(c.mu);
    const size_t w = (lower_bytes / 4) & ~(size_t)63;
    if (lower_bytes == 0) return c.set_windows(1, 0, 0);
    if (n_poa < 1 || n_aln < 1 || n_poa > 64 || n_aln > 64 || w / n_poa < (1u << 20) || w + (size_t)n_aln * (1u << 20) > c.pool_words) {
        set_error("lcd_gpu_pool_windows: %d + %d windows with %zu bytes below do not fit a pool of %zu bytes", n_poa, n_aln, lower_bytes, c.pool_words * 4); return -1;
    }
    return c.set_windows(n_poa, n_aln, w);
}
int lcd_gpu_split_pool(size_t lower_bytes) { return lcd_gpu_pool_windows(1, 1, lower_bytes); }
int lcd_gpu_reserve_sms(int n_sms) {
    Context &c = ctx();
    if (n_sms < 0 || n_sms >= c.sm_count) { set_error("lcd_gpu_reserve_sms: %d out of range", n_sms); return -1; }
    c.reserved_sms = n_sms;
    return 0;
}
void lcd_gpu_set_thread_stream(void *stream) { thread_stream() = (cudaStream_t)stream; }
void *lcd_gpu_new_stream(void) {
    if (ensure_ready()) return nullptr;
    Context &c = ctx();
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { set_error("lcd_gpu_new_stream: cudaStreamCreate failed"); return nullptr; }
    std::lock_guard<std::mutex> lk(c.mu);
    c.extra_streams.push_back(s);
    return (void*)s;
}

int lcd_gpu_init(int device, size_t pool_bytes) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready) {
        // a second init must describe the context that is up: the library is one context per process
        if (device != c.device && !(device == 0 && pool_bytes == 0)) { set_error("lcd_gpu_init: already initialised on device %d (asked for %d)", c.device, device); return -1; }
        if (pool_bytes && pool_bytes != c.requested_pool_bytes) { set_error("lcd_gpu_init: already initialised with a pool of %zu bytes (asked for %zu)", c.pool_words * 4, pool_bytes); return -1; }
        return 0;
    }
    c.requested_pool_bytes = pool_bytes;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("lcd_gpu_init: no CUDA device (%s); liblcd_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return -1;
    }
    if (device < 0 || device >= ndev) { set_error("lcd_gpu_init: device %d out of range (0..%d)", device, ndev - 1); return -1; }
    LCD_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    LCD_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("lcd_gpu_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return -1;
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    LCD_CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    {   // plans allocate from the device's stream-ordered pool; keep what they free for the next plan
        cudaMemPool_t mp;
        LCD_CUDA_OK(cudaDeviceGetDefaultMemPool(&mp, device));
        unsigned long long keep = ~0ull;
        LCD_CUDA_OK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    if (pool_bytes == 0) {
        size_t free_b = 0, total_b = 0;
        LCD_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
        pool_bytes = free_b / 4;                       // a quarter of free HBM ...
        const size_t cap = (size_t)32 << 30;           // ... at most 32 GiB
        if (pool_bytes > cap) pool_bytes = cap;
    }
    pool_bytes &= ~(size_t)255;
    LCD_CUDA_OK(cudaMalloc((void**)&c.pool, pool_bytes));
    c.pool_words = pool_bytes / 4;
    if (c.set_windows(1, 0, 0)) return -1;
    c.ready = true;
    return 0;
}

void lcd_gpu_shutdown(void) {
    Context &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) return;
    cudaDeviceSynchronize();
    if (c.pool) cudaFree(c.pool);
    for (Context::Window *w : c.windows) { if (w->done) cudaEventDestroy(w->done); if (w->bitmap) cudaFree(w->bitmap); delete w; }
    c.windows.clear(); c.cls_win[0].clear(); c.cls_win[1].clear();
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.aux_stream) cudaStreamDestroy(c.aux_stream);
    for (cudaStream_t s : c.extra_streams) cudaStreamDestroy(s);
    c.extra_streams.clear();
    c.aux_stream = nullptr;
    c.pool = nullptr; c.stream = nullptr; c.ready = false;
}

int lcd_plan_run(lcd_plan_t *plan, void *stream) {
    if (!plan) { set_error("lcd_plan_run: null plan"); return -1; }
    if (ensure_ready()) return -1;
    Context &c = ctx();
    Plan *p = reinterpret_cast<Plan*>(plan);
    if (!p->uses_pool()) return p->run(pick_stream(stream));
    cudaStream_t s = pick_stream(stream);
    Context::Window *w = c.pick_window(p->pool_window() == 1 ? 1 : 0);
    std::lock_guard<std::mutex> lk(w->mu);
    p->win = w;
    LCD_CUDA_OK(cudaStreamWaitEvent(s, w->done, 0));        // the window's previous user, on whatever stream it ran
    const int rc = p->run(s);
    LCD_CUDA_OK(cudaEventRecord(w->done, s));
    return rc;
}

int lcd_plan_sync(lcd_plan_t *plan, void *stream) {
    LCD_CUDA_OK(cudaStreamSynchronize(pick_stream(stream)));
    if (plan && reinterpret_cast<Plan*>(plan)->finish(pick_stream(stream))) return -1;
    LCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int lcd_plan_work_units(lcd_plan_t *plan, void *stream, uint64_t *units) {
    if (!plan) { set_error("lcd_plan_work_units: null plan"); return -1; }
    return reinterpret_cast<Plan*>(plan)->work_units(pick_stream(stream), units);
}

void lcd_plan_destroy(lcd_plan_t *plan) {
    bind_thread();
    if (plan) delete reinterpret_cast<Plan*>(plan);
}

}
