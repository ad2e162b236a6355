// CPU test of efg_hostcopy.cuh (in-place Int32 -> Int64 widening by several threads); the CUDA calls are stubbed.
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <cstring>
typedef int cudaEvent_t; typedef int cudaError_t; const int cudaSuccess = 0;
static int cudaEventSynchronize(cudaEvent_t) { return 0; }
static int cudaEventQuery(cudaEvent_t) { return 0; }
const int cudaErrorNotReady = 600;
static int cudaEventDestroy(cudaEvent_t) { return 0; }
static int cudaGetLastError() { return 0; }
static int cudaSetDevice(int) { return 0; }
#include "efg_hostcopy.cuh"
int main() {
    for (int64_t nnz : {1LL, 3LL, 17LL, 4097LL, 100000LL, 5000001LL}) {
      for (int64_t ch : {1024LL, 1LL << 20}) {
        std::vector<int64_t> buf((size_t)nnz + 2);
        int64_t *dst = buf.data() + 1;   // misaligned by 8 on purpose
        int32_t *in = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(dst) + 4 * nnz);
        for (int64_t i = 0; i < nnz; i++) in[i] = (int32_t)((i * 2654435761u) & 0x7fffffff);
        std::vector<int32_t> ref(in, in + nnz);
        HostWiden w; w.dst = dst; w.nnz = nnz; w.nthreads = 5; w.cuts = widen_cuts(nnz, ch);
        w.events.assign(w.cuts.size() - 1, 0);
        for (int t = 0; t < w.nthreads; t++) w.threads.emplace_back(widen_worker, &w, t);
        for (auto &t : w.threads) t.join();
        w.threads.clear(); w.events.clear();
        int64_t bad = 0;
        for (int64_t i = 0; i < nnz; i++) bad += dst[i] != (int64_t)ref[i] + 1;
        if (bad) { printf("nnz %lld ch %lld: %lld wrong entries\n", (long long)nnz, (long long)ch, (long long)bad); return 1; }
      }
    }
}
