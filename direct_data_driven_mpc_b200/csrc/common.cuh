// Shared host/device helpers for libddmpc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/ddmpc.h"

#define DDMPC_ADMM_RELAX 1.8   // over-relaxation of the box-row ADMM (all three implementations share it)

namespace ddmpc {

extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

#define DDMPC_CUDA(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::ddmpc::fail(DDMPC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,          \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
    } while (0)

#define DDMPC_LAUNCH_CHECK()                                                              \
    do {                                                                                  \
        ::ddmpc::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return ::ddmpc::fail(DDMPC_ERR_CUDA, "kernel launch failed: %s (%s:%d)",      \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
    } while (0)

#define DDMPC_TRY(expr)                                                                   \
    do {                                                                                  \
        int _s = (expr);                                                                  \
        if (_s != DDMPC_OK) return _s;                                                    \
    } while (0)

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// True the first time it is called with `mask` on the current device (function attributes are per device).
inline bool first_time_on_device(std::atomic<unsigned long long> &mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}

// RAII device buffer (setup scratch + plan storage).  Memory comes from the device's stream-ordered pool
// (cudaMallocAsync, ordered on the stream of the calling entry point) that keeps what it has been given (release threshold = max;
// ddmpc_trim_memory() hands it back): a batched controller setup
// allocates and frees a dozen buffers of up to hundreds of MB, and with cudaMalloc/cudaFree (each a device-wide
// synchronisation plus page-table work) that cost 80-250 ms of wall time per call against 13 ms of kernels.
// Scratch is freed in the order of the stream its kernels ran on (ScratchStreamScope below), and ddmpc_set_destroy
// synchronises the device, so returning memory to the pool never races with work using it.
inline cudaError_t pool_ready() {
    static std::atomic<unsigned long long> done{0};
    if (!first_time_on_device(done)) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    uint64_t keep = ~0ull;   // a finite threshold makes every large setup re-map gigabytes (0.4-3 s instead of 0.13 s)
    return cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}

// Stream that DevBuf allocations made by the calling thread are ordered on.  Entry points that enqueue work on a
// caller stream open a ScratchStreamScope for it, so scratch is allocated AND freed in that stream's order: an early
// error return then cannot hand memory that kernels on the stream are still using back to the pool.
extern thread_local cudaStream_t g_scratch_stream;
struct ScratchStreamScope {
    cudaStream_t prev;
    explicit ScratchStreamScope(cudaStream_t s) : prev(g_scratch_stream) { g_scratch_stream = s; }
    ~ScratchStreamScope() { g_scratch_stream = prev; }
};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;   // the stream the allocation is ordered on
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr;
        bytes = 0;
    }
    // long-lived buffers (plan storage) outlive the stream they were created on: ddmpc_set_destroy synchronises the
    // device and frees them on the legacy stream
    void detach_stream() { stream = nullptr; }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 8;
        cudaError_t e = pool_ready();
        if (e != cudaSuccess) return e;
        stream = g_scratch_stream;
        e = cudaMallocAsync(&p, n, stream);
        if (e == cudaSuccess) bytes = n;
        else p = nullptr;
        return e;
    }
    double *d() const { return (double *)p; }
    int *i() const { return (int *)p; }
};

}  // namespace ddmpc
