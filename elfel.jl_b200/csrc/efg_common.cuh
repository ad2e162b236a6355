// efg_common.cuh -- context, device-buffer bookkeeping and error plumbing of libelfelgpu.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <new>

#include "../../include/elfel_gpu.h"

struct EfgError {
    int code;
    std::string msg;
};

[[noreturn]] inline void efg_throw(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw EfgError{code, buf};
}

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            efg_throw(err__ == cudaErrorMemoryAllocation ? EFG_ERR_OOM : EFG_ERR_CUDA,            \
                      "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

// Device allocations owned by a ctx, stream-ordered (cudaMallocAsync / cudaFreeAsync on the ctx stream with a
// never-trimmed pool): the symbolic phase allocates and frees many multi-GB temporaries and plain
// cudaMalloc/cudaFree (device-synchronising, unmapping) made its time vary by 10x.
struct DevPool {
    int64_t bytes = 0;
    cudaStream_t stream = nullptr;
    void *alloc(size_t n)
    {
        void *p = nullptr;
        if (n == 0) n = 8;
        cudaError_t e = cudaMallocAsync(&p, n, stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            efg_throw(EFG_ERR_OOM, "cudaMallocAsync of %zu bytes failed: %s", n, cudaGetErrorString(e));
        }
        bytes += (int64_t)n;
        return p;
    }
    void free(void *p) { cudaFreeAsync(p, stream); }
};

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int64_t held = 0;
    DevPool *pool = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void alloc(DevPool &pl, size_t count)
    {
        release();
        pool = &pl;
        const int64_t before = pl.bytes;
        p = (T *)pl.alloc(count * sizeof(T));
        held = pl.bytes - before;
        n = count;
    }
    void release()
    {
        if (p) {
            if (pool) { pool->free(p); pool->bytes -= held; } else cudaFree(p);
        }
        p = nullptr; n = 0; held = 0;
    }
    size_t size_bytes() const { return n * sizeof(T); }
};

static inline unsigned grid_for(int64_t n, int block, int64_t cap = (int64_t)148 * 64)
{
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}
