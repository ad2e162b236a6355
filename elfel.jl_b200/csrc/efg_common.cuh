// efg_common.cuh -- context, device-buffer bookkeeping and error plumbing of libelfelgpu.so
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <new>
#include <cstdlib>
#include <algorithm>
#include <iterator>

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

// Device memory of a ctx: a private arena.  Slabs come from plain cudaMalloc (measured on B200: 16 GB in 26 ms, against
// 635 ms for the first cudaMallocAsync of that size -- growing a stream-ordered pool made the FIRST symbolic phase of a
// process take seconds) and are handed out by a host-side best-fit free list with coalescing, so the steady state makes
// no driver call at all and the time of the symbolic phase no longer depends on allocator state.  Every block is used
// on the ctx's own stream only (or on its copy stream behind an event), so re-using a freed block for later work of the
// same stream needs no synchronisation.  Nothing is shared with other users of the device (no process-global pool
// attributes are touched); efg_destroy returns everything.
#include <map>
#include <unordered_map>
struct DevPool {
    int64_t bytes = 0;                    // handed out
    int64_t reserved = 0;                 // held in slabs
    int64_t peak = 0;                     // high-water mark of `bytes`
    cudaStream_t stream = nullptr;
    size_t next_slab = (size_t)256 << 20; // growth hint for the next slab (raised by reserve())
    struct Slab { char *base; size_t size; };
    std::vector<Slab> slabs;
    std::map<char *, size_t> free_blocks;             // address -> size, coalesced
    std::unordered_map<void *, size_t> live;

    static size_t round_up(size_t n) { return (n + 511) & ~(size_t)511; }
    void add_free(char *p, size_t n)
    {
        auto it = free_blocks.lower_bound(p);
        if (it != free_blocks.begin()) {        // merge with the block before, when adjacent and in the same slab
            auto pr = std::prev(it);
            if (pr->first + pr->second == p && same_slab(pr->first, p)) { p = pr->first; n += pr->second; free_blocks.erase(pr); }
        }
        if (it != free_blocks.end() && p + n == it->first && same_slab(p, it->first)) { n += it->second; free_blocks.erase(it); }
        free_blocks[p] = n;
    }
    bool same_slab(const char *a, const char *b) const
    {
        for (const Slab &s : slabs)
            if (a >= s.base && a < s.base + s.size) return b >= s.base && b < s.base + s.size;
        return false;
    }
    bool grow(size_t need)
    {
        size_t want = need > next_slab ? need : next_slab;
        if (want < (size_t)reserved / 8) want = round_up((size_t)reserved / 8);     // geometric growth: few slabs, bounded waste
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && want > need) { cudaGetLastError(); want = need; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) {                 // return completely free slabs to the driver and try once more
            cudaGetLastError();
            if (trim() == 0) return false;
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) { cudaGetLastError(); return false; }
        }
        slabs.push_back(Slab{(char *)p, want});
        reserved += (int64_t)want;
        next_slab = (size_t)256 << 20;          // a reserve() hint covers one slab
        if (getenv("EFG_TRACE")) fprintf(stderr, "[efg trace] arena: new slab of %.2f GB (asked %.2f GB), %zu slabs, %.2f GB reserved\n", want / 1e9, need / 1e9, slabs.size(), reserved / 1e9);
        add_free((char *)p, want);
        return true;
    }
    // the caller knows roughly how much the coming phase needs: make the next slab that large (one cudaMalloc instead of many)
    void reserve(size_t total)
    {
        const size_t have = (size_t)(reserved - bytes);
        if (total > have) next_slab = std::max(next_slab, round_up(total - have));
    }
    void *alloc(size_t n)
    {
        n = round_up(n ? n : 8);
        for (int attempt = 0; attempt < 2; attempt++) {
            auto best = free_blocks.end();
            for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it)
                if (it->second >= n && (best == free_blocks.end() || it->second < best->second)) best = it;
            if (best != free_blocks.end()) {
                char *p = best->first;
                const size_t sz = best->second;
                free_blocks.erase(best);
                if (sz > n) free_blocks[p + n] = sz - n;
                live[p] = n;
                bytes += (int64_t)n;
                if (bytes > peak) peak = bytes;
                return p;
            }
            if (attempt == 0 && !grow(n)) break;
        }
        efg_throw(EFG_ERR_OOM, "device allocation of %zu bytes failed (%lld bytes held by this ctx)", n, (long long)reserved);
    }
    void free(void *p)
    {
        auto it = live.find(p);
        if (it == live.end()) return;
        const size_t n = it->second;
        live.erase(it);
        bytes -= (int64_t)n;
        add_free((char *)p, n);
    }
    // give completely unused slabs back to the driver; returns the number of bytes released
    size_t trim()
    {
        size_t released = 0;
        for (size_t k = 0; k < slabs.size();) {
            auto it = free_blocks.find(slabs[k].base);
            if (it != free_blocks.end() && it->second == slabs[k].size) {
                if (stream) cudaStreamSynchronize(stream);
                cudaFree(slabs[k].base);
                released += slabs[k].size;
                reserved -= (int64_t)slabs[k].size;
                free_blocks.erase(it);
                slabs.erase(slabs.begin() + (long)k);
            } else k++;
        }
        return released;
    }
    void destroy()
    {
        for (const Slab &s : slabs) cudaFree(s.base);
        slabs.clear(); free_blocks.clear(); live.clear();
        bytes = reserved = 0;
    }
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
        p = (T *)pl.alloc(count * sizeof(T));
        n = count;
    }
    void release()
    {
        if (p) {
            if (pool) pool->free(p); else cudaFree(p);
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
