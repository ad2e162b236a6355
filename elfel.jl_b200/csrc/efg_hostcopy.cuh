// efg_hostcopy.cuh -- the row indices' way to a HOST array (efg_fetch_pattern_async / efg_fetch_csc).
//
// The reference's SparseMatrixCSC holds rowval as Int64, 1-based (src/Assemblers.jl:121-123); the device holds Int32,
// 0-based.  Widening on the device doubles the PCIe traffic of the structure (config 2: 5.9 GB instead of 2.9 GB at
// ~55 GB/s).  Here the Int32 array crosses the link as it is, chunk by chunk, into the UPPER HALF of the caller's Int64
// array, and a few host threads widen it in place (out[i] = in32[i] + 1, ascending) while the next chunks and then the
// values are still on the link.
//
// Why in place is safe: input i sits at byte 4*nnz + 4*i of the destination, output i at byte 8*i.
//   - writing output i touches bytes [8i, 8i+8): input index 2i - nnz <= i, already consumed (ascending order);
//   - a chunk [a, b) processed by several threads at once destroys the inputs [2a - nnz, 2b - nnz): all of them must lie in
//     finished chunks, i.e. 2b - nnz <= a.  Chunks therefore shrink geometrically towards the end (b <= (a + nnz) / 2);
//   - the DMA of a later chunk writes at 4*nnz + 4*j, j > i, which is beyond every output written so far (8i < 4nnz + 4j).
#pragma once
#include <thread>
#include <atomic>
#include <chrono>
#include <emmintrin.h>

struct HostWiden {
    std::vector<std::thread> threads;
    std::vector<cudaEvent_t> events;          // events[c]: chunk c has arrived (owned by the ctx, reused by every fetch:
                                              // creating ~100 blocking-sync events per call cost 35 ms)
    std::vector<int64_t> cuts;                // chunk c = [cuts[c], cuts[c+1])
    std::atomic<int> arrived{0};              // barrier between chunks
    std::atomic<int> failed{0};
    int nthreads = 0;
    int64_t *dst = nullptr;
    int64_t nnz = 0;
    int device = 0;
    void join()
    {
        for (auto &t : threads) if (t.joinable()) t.join();
        threads.clear();
    }
    ~HostWiden() { join(); }
};

// out[i] = in[i] + 1 for i in [a, b); out 8-byte aligned; streaming (non-temporal) stores: the output is not read again here
static inline void widen_range(const int32_t *in, int64_t *out, int64_t a, int64_t b)
{
    int64_t i = a;
    while (i < b && (reinterpret_cast<uintptr_t>(out + i) & 15)) { out[i] = (int64_t)in[i] + 1; i++; }
    const __m128i zero = _mm_setzero_si128(), one = _mm_set1_epi64x(1);
    for (; i + 4 <= b; i += 4) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(in + i));      // 4 non-negative Int32
        _mm_stream_si128(reinterpret_cast<__m128i *>(out + i), _mm_add_epi64(_mm_unpacklo_epi32(v, zero), one));
        _mm_stream_si128(reinterpret_cast<__m128i *>(out + i + 2), _mm_add_epi64(_mm_unpackhi_epi32(v, zero), one));
    }
    for (; i < b; i++) out[i] = (int64_t)in[i] + 1;
    _mm_sfence();
}

static void widen_worker(HostWiden *w, int tid)
{
    cudaSetDevice(w->device);
    const int32_t *in = reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(w->dst) + 4 * w->nnz);
    const int nch = (int)w->cuts.size() - 1;
    for (int c = 0; c < nch; c++) {
        // blocking-sync events: a polling loop with short sleeps measured 40 ms slower per fetch (timer slack), host-function
        // callbacks 20 ms slower (they stall the copy stream); their creation (~0.4 ms each) is paid once per ctx
        if (cudaEventSynchronize(w->events[(size_t)c]) != cudaSuccess) { w->failed = 1; cudaGetLastError(); }
        const int64_t a = w->cuts[(size_t)c], b = w->cuts[(size_t)c + 1], len = b - a;
        const int64_t lo = a + len * tid / w->nthreads, hi = a + len * (tid + 1) / w->nthreads;
        if (w->failed) { /* skip the work, keep the barrier protocol */ }
        else if (2 * b - w->nnz > a) { if (tid == 0) widen_range(in, w->dst, a, b); }     // short tail: one thread, ascending
        else widen_range(in, w->dst, lo, hi);
        // every thread has finished chunk c before anyone starts c + 1 (whose outputs may overwrite chunk c's inputs)
        w->arrived.fetch_add(1, std::memory_order_acq_rel);
        const int want = (c + 1) * w->nthreads;
        while (w->arrived.load(std::memory_order_acquire) < want) std::this_thread::yield();
    }
}

static int host_widen_threads(int device_count)
{
    if (const char *e = getenv("EFG_HOST_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) return v; }
    const int hw = (int)std::thread::hardware_concurrency();
    int t = hw / (device_count > 0 ? device_count : 1);      // one process per GPU shares the host cores
    if (t > 8) t = 8;
    if (t < 2) t = 2;
    return t;
}

// chunk boundaries: `ch` entries each while the parallel-safety bound b <= (a + nnz) / 2 allows, then halving
static std::vector<int64_t> widen_cuts(int64_t nnz, int64_t ch)
{
    std::vector<int64_t> cuts{0};
    int64_t a = 0;
    while (a < nnz) {
        int64_t b = a + ch;
        const int64_t safe = (a + nnz) / 2;
        if (b > safe) b = safe;
        if (b <= a || nnz - a <= 4096) b = nnz;          // short tail: one chunk, widened by one thread (widen_worker)
        cuts.push_back(b);
        a = b;
    }
    return cuts;
}
