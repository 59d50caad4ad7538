// runtime.cu -- process-wide state of liblapack_b200: CUDA error latch, stream-ordered workspace
// pool, look-ahead streams/events, device properties, and the DMMA peak micro-benchmark used as the
// measured FP64 tensor-core roofline denominator.
#include "lb_internal.h"
#include <atomic>
#include <mutex>

namespace lb {

// Two latches: `pending` is consumed by the call that reports it (so one failed call does not poison the return codes of
// later, successful calls); `last` stays readable through lb200_last_cuda_error() until lb200_clear_cuda_error().
static std::atomic<int> g_cuda_err{0}, g_cuda_err_last{0};
void record_cuda_error(cudaError_t e) {
    int expected = 0;
    g_cuda_err.compare_exchange_strong(expected, (int)e);
    g_cuda_err_last.store((int)e);
}
int last_cuda_error() { int e = g_cuda_err.load(); return e ? e : g_cuda_err_last.load(); }
int pending_cuda_error() { return g_cuda_err.load(); }
int take_cuda_error() {
    const int e = g_cuda_err.exchange(0);
    if (e) (void)cudaGetLastError();
    return e;
}
void clear_cuda_error() {
    g_cuda_err.store(0);
    g_cuda_err_last.store(0);
    (void)cudaGetLastError();
}

// Device word that turns the Level-3 kernels launched while it is set into no-ops once it becomes non-zero: DPOTRF aborts at
// the first non-positive leading minor (dpotrf.f:219-220,239-240), so everything queued behind the failing leaf must not run
// on the unfactored block (it would spread Inf/NaN over the trailing matrix).  Set / cleared by the Cholesky drivers under
// the library mutex.
static thread_local const int* g_kernel_guard = nullptr;   // per host thread: concurrent callers must not see each other's guard
const int* kernel_guard() { return g_kernel_guard; }
void set_kernel_guard(const int* p) { g_kernel_guard = p; }

std::recursive_mutex& driver_mutex() {
    static std::recursive_mutex m;
    return m;
}

StreamOut*& stream_out() {
    static StreamOut* so = nullptr;
    return so;
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

void* ws_alloc(cudaStream_t s, size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    static bool pool_cfg = false;
    if (!pool_cfg) {
        // keep freed blocks cached in the default pool instead of returning them to the driver
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            // The panel, update and side streams all take scratch from this pool.  With the default policy the allocator
            // may hand a block freed on one stream to another stream by INSERTING a dependency between the two streams
            // (cudaMemPoolReuseAllowInternalDependencies), which silently serialises the high-priority panel behind a
            // long low-priority update.  LAPACK_B200_POOL_INTERNAL_DEPS=0 restricts reuse across streams to frees that have
            // already completed (A/B knob).
            const char* e = getenv("LAPACK_B200_POOL_INTERNAL_DEPS");
            int off = (e && e[0] == '0') ? 0 : 1;      // default: CUDA's default (allowed); =0 was measured neutral on one GPU
            cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
        }
        pool_cfg = true;
    }
    LB_CUDA_CHECK(cudaMallocAsync(&p, bytes, s));
    return p;
}
void ws_free(cudaStream_t s, void* p) {
    if (p) LB_CUDA_CHECK(cudaFreeAsync(p, s));
}

Aux& aux(int level) {
    static Aux a[2];
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Aux& x = a[level ? 1 : 0];
    if (!x.ready) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = highest priority
        auto pr = [&](int steps_below_top) { const int p = hi + steps_below_top; return p > lo ? lo : p; };
        if (level == 0) {
            // level 0: panel on top, preparation next, updates / side work two steps above the lowest priority when the device has
            // that many levels, so that a level-0 factorization running as the "panel" of level 1 outranks level 1's updates
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.panel_stream, cudaStreamNonBlocking, pr(0)));
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.prep_stream, cudaStreamNonBlocking, pr(1)));
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.update_stream, cudaStreamNonBlocking, pr(3)));
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.side_stream, cudaStreamNonBlocking, pr(3)));
        } else {
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.panel_stream, cudaStreamNonBlocking, pr(2)));   // carries the level-0 driver
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.prep_stream, cudaStreamNonBlocking, pr(4)));
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.update_stream, cudaStreamNonBlocking, lo));
            LB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.side_stream, cudaStreamNonBlocking, lo));
        }
        for (int i = 0; i < 32; ++i) LB_CUDA_CHECK(cudaEventCreateWithFlags(&x.ev[i], cudaEventDisableTiming));
        x.ready = true;
    }
    return x;
}

// ----------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmarks: `iters` x 8 independent DMMA.8x8x4 (or 16 DFMA) chains per warp.
__global__ void dmma_peak_kernel(int iters, double* out) {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i * 1e-9; }
    double x = 1.0 + threadIdx.x * 1e-12, y = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1])
                         : "d"(x), "d"(y));
    }
    double sacc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc += d[i][0] + d[i][1];
    if (sacc == 123.456) out[0] = sacc;
}
__global__ void dfma_peak_kernel(int iters, double* out) {
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = threadIdx.x * 1e-9 + i;
    double x = 1.0 + threadIdx.x * 1e-12, y = 1e-13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = fma(d[i], x, y);
    }
    double sacc = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) sacc += d[i];
    if (sacc == 123.456) out[0] = sacc;
}

// returns TFLOP/s; kind 0 = DMMA, 1 = DFMA
double fp64_peak(cudaStream_t s, int kind, int warps_per_cta, int ctas_per_sm, int iters) {
    double* out = (double*)ws_alloc(s, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int grid = num_sms() * ctas_per_sm, block = warps_per_cta * 32;
    for (int rep = 0; rep < 2; ++rep) {
        if (rep == 1) cudaEventRecord(e0, s);
        if (kind == 0) dmma_peak_kernel<<<grid, block, 0, s>>>(iters, out);
        else dfma_peak_kernel<<<grid, block, 0, s>>>(iters, out);
        if (rep == 1) cudaEventRecord(e1, s);
    }
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ws_free(s, out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double flops_per_warp = kind == 0 ? (double)iters * 8 * 512.0 : (double)iters * 16 * 64.0;
    double total = flops_per_warp * warps_per_cta * (double)grid;
    return total / (ms * 1e-3) * 1e-12;
}

}  // namespace lb
