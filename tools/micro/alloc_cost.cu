// How expensive are large device allocations on this box?  (cold-call cost of the symbolic phase)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main()
{
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t GB = (size_t)1 << 30;
    for (size_t g : {1, 4, 16, 32}) {
        void *p;
        double t0 = now(); CK(cudaMalloc(&p, g * GB)); double t1 = now();
        CK(cudaMemsetAsync(p, 0, g * GB, st)); CK(cudaStreamSynchronize(st)); double t2 = now();
        CK(cudaMemsetAsync(p, 0, g * GB, st)); CK(cudaStreamSynchronize(st)); double t3 = now();
        CK(cudaFree(p)); double t4 = now();
        printf("cudaMalloc %2zu GB: malloc %8.2f ms, first memset %7.2f, second memset %7.2f, free %7.2f\n", g, t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    }
    {   // default pool, release threshold max
        cudaMemPool_t mp; CK(cudaDeviceGetDefaultMemPool(&mp, 0));
        uint64_t thr = UINT64_MAX; CK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr));
        for (size_t g : {1, 4, 16, 32}) {
            void *p;
            double t0 = now(); CK(cudaMallocAsync(&p, g * GB, st)); CK(cudaStreamSynchronize(st)); double t1 = now();
            CK(cudaMemsetAsync(p, 0, g * GB, st)); CK(cudaStreamSynchronize(st)); double t2 = now();
            CK(cudaFreeAsync(p, st)); CK(cudaStreamSynchronize(st)); double t3 = now();
            CK(cudaMallocAsync(&p, g * GB, st)); CK(cudaStreamSynchronize(st)); double t4 = now();
            CK(cudaFreeAsync(p, st)); CK(cudaStreamSynchronize(st));
            printf("cudaMallocAsync(default pool) %2zu GB: cold %8.2f ms, memset %7.2f, free %6.2f, warm re-alloc %6.2f\n", g, t1 - t0, t2 - t1, t3 - t2, t4 - t3);
        }
        CK(cudaMemPoolTrimTo(mp, 0));
    }
    {   // many medium allocations, like the symbolic phase: 64 x 256 MB cold, then again warm
        std::vector<void *> v(64);
        double t0 = now();
        for (auto &p : v) CK(cudaMallocAsync(&p, 256 << 20, st));
        CK(cudaStreamSynchronize(st)); double t1 = now();
        for (auto &p : v) CK(cudaFreeAsync(p, st));
        CK(cudaStreamSynchronize(st)); double t2 = now();
        for (auto &p : v) CK(cudaMallocAsync(&p, 256 << 20, st));
        CK(cudaStreamSynchronize(st)); double t3 = now();
        for (auto &p : v) CK(cudaFreeAsync(p, st));
        CK(cudaStreamSynchronize(st));
        printf("64 x 256 MB cudaMallocAsync: cold %8.2f ms, free %6.2f, warm %6.2f\n", t1 - t0, t2 - t1, t3 - t2);
    }
    {   // private pool
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned; props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice; props.location.id = 0;
        cudaMemPool_t mp; CK(cudaMemPoolCreate(&mp, &props));
        uint64_t thr = UINT64_MAX; CK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr));
        void *p;
        double t0 = now(); CK(cudaMallocFromPoolAsync(&p, 16 * GB, mp, st)); CK(cudaStreamSynchronize(st)); double t1 = now();
        CK(cudaFreeAsync(p, st)); CK(cudaStreamSynchronize(st));
        double t2 = now(); CK(cudaMemPoolDestroy(mp)); double t3 = now();
        printf("private pool: 16 GB cold %8.2f ms, destroy %7.2f ms\n", t1 - t0, t3 - t2);
    }
    return 0;
}
