// efg_twopass.cuh -- the general two-pass path.
//
// Symbolic: every COO triplet the reference would append gets the key (col, row); a stable radix
// sort (input enumerated element-major in the reference's append order) groups equal keys in
// append order, the group heads are the CSC nonzeros (rows ascending inside a column, explicit
// zeros kept), exactly what sparse(I,J,V,m,n) produces (src/Assemblers.jl:121-123).
// Numeric: kernel 1 writes all element matrices to HBM, kernel 2 sums each nonzero's
// contributions left to right in append order (deterministic, no atomics).
// ~2.5x the algorithmic traffic: this path is the cross-check and the fallback for meshes whose
// node valence exceeds the tiled path's limits, not the fast path.
#pragma once
#include <cub/cub.cuh>
#include "efg_ctx.cuh"

template <class F>
__global__ void k_tp_keys(DofSrc src, int64_t nel, int64_t nrow, int64_t ncol, ColMap cm, int64_t ncl,
                          uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int *__restrict__ errflag)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nel; e += stride) {
        int32_t d[F::ND];
        F::edofs(src, e, d);
        bool bad = false;
#pragma unroll
        for (int a = 0; a < F::ND; a++) bad |= (d[a] < 0) || (d[a] >= nrow && d[a] >= ncol);
#pragma unroll
        for (int j = 0; j < F::ND; j++)
#pragma unroll
            for (int i = 0; i < F::ND; i++)
                if (F::mask(i, j)) {
                    bad |= (d[i] >= nrow) || (d[j] >= ncol);
                    const int64_t t = e * F::NT + F::kidx(i, j);
                    const int64_t lc = (d[j] >= 0) ? cm.local(d[j]) : -1;
                    const uint64_t cl = lc >= 0 ? (uint64_t)lc : (uint64_t)ncl;
                    keys[t] = (cl << 32) | (uint32_t)d[i];
                    vals[t] = (uint32_t)t;
                }
        if (bad) *errflag = 1;
    }
}

#define TP_CHUNK 2048
// heads per chunk of the sorted key array
__global__ void k_tp_count_heads(const uint64_t *__restrict__ keys, int64_t n, uint64_t ncl,
                                 int64_t *__restrict__ chunk_heads, int64_t *__restrict__ chunk_valid)
{
    const int64_t base = (int64_t)blockIdx.x * TP_CHUNK;
    int heads = 0, valid = 0;
    for (int o = threadIdx.x; o < TP_CHUNK; o += blockDim.x) {
        const int64_t p = base + o;
        if (p < n) {
            const uint64_t k = keys[p];
            if ((k >> 32) < ncl) {
                valid++;
                if (p == 0 || keys[p - 1] != k) heads++;
            }
        }
    }
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int h = BR(tmp).Sum(heads);
    __syncthreads();
    const int v = BR(tmp).Sum(valid);
    if (threadIdx.x == 0) { chunk_heads[blockIdx.x] = h; chunk_valid[blockIdx.x] = v; }
}

__global__ void k_tp_fill_heads(const uint64_t *__restrict__ keys, int64_t n, uint64_t ncl,
                                const int64_t *__restrict__ chunk_off, int64_t *__restrict__ seg_start,
                                int32_t *__restrict__ rowval, unsigned long long *__restrict__ colcnt)
{
    typedef cub::BlockScan<int, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * TP_CHUNK;
    const int64_t off = chunk_off[blockIdx.x];
    for (int o0 = 0; o0 < TP_CHUNK; o0 += 256) {
        const int64_t p = base + o0 + threadIdx.x;
        int flag = 0;
        uint64_t k = 0;
        if (p < n) {
            k = keys[p];
            flag = ((k >> 32) < ncl) && (p == 0 || keys[p - 1] != k);
        }
        int rank, total;
        BS(tmp).ExclusiveSum(flag, rank, total);
        const int c = carry;
        if (flag) {
            const int64_t s = off + c + rank;
            seg_start[s] = p;
            rowval[s] = (int32_t)(uint32_t)(k & 0xffffffffu);
            atomicAdd(&colcnt[k >> 32], 1ull);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

__global__ void k_tp_colptr_from_counts(const int64_t *__restrict__ excl, int64_t ncl, int64_t nnz, int64_t *__restrict__ colptr)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= ncl; j += stride)
        colptr[j] = (j < ncl ? excl[j] : nnz) + 1;
}

template <class F> void twopass_symbolic(efg_ctx *ctx)
{
    const MeshDev &m0 = ctx->mesh[0];
    const int64_t nel = m0.nel;
    const int64_t ntrip = nel * F::NT;
    if (ntrip >= (int64_t)1 << 32)
        efg_throw(EFG_ERR_LIMIT, "two-pass path: %lld triplets exceed the 2^32 limit; shard the mesh (efg_set_column_range)", (long long)ntrip);
    const int64_t ncl = ctx->ncl;
    DofSrc src{m0.conn.p, ctx->mesh[1].conn.p, ctx->space[0].dof.p, ctx->space[1].dof.p, ctx->space[2].dof.p, ctx->space[0].cdof.p, ctx->space[1].cdof.p, ctx->space[2].cdof.p};

    DevBuf<uint64_t> keys, keys2;
    DevBuf<uint32_t> vals;
    DevBuf<int> errflag;
    keys.alloc(ctx->pool, ntrip); keys2.alloc(ctx->pool, ntrip);
    vals.alloc(ctx->pool, ntrip); ctx->tp.perm.alloc(ctx->pool, ntrip);
    errflag.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemsetAsync(errflag.p, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_tp_keys<F>, grid_for(nel, 256), 256, 0, src, nel, ctx->nrow, ctx->ncol, COLMAP(ctx), ncl, keys.p, vals.p, errflag.p);
    int herr = 0;
    CUDA_CHECK(cudaMemcpyAsync(&herr, errflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (herr)
        efg_throw(EFG_ERR_INDEX, "ArgumentError: a dof number is < 1 or exceeds nrow/ncol (was every space numbered, incl. data dofs?)");

    int bits = 1;
    while (((int64_t)1 << bits) <= ncl) bits++;
    const int end_bit = 32 + bits;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, ctx->tp.perm.p, ntrip, 0, end_bit, ctx->stream);
    DevBuf<char> tmp;
    tmp.alloc(ctx->pool, tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, vals.p, ctx->tp.perm.p, ntrip, 0, end_bit, ctx->stream));
    ctx->launches += 8;
    keys.release(); vals.release(); tmp.release();

    const int64_t nchunk = (ntrip + TP_CHUNK - 1) / TP_CHUNK;
    DevBuf<int64_t> ch_heads, ch_valid, ch_off;
    ch_heads.alloc(ctx->pool, nchunk + 1); ch_valid.alloc(ctx->pool, nchunk + 1); ch_off.alloc(ctx->pool, nchunk + 1);
    CUDA_CHECK(cudaMemsetAsync(ch_heads.p, 0, (nchunk + 1) * sizeof(int64_t), ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(ch_valid.p, 0, (nchunk + 1) * sizeof(int64_t), ctx->stream));
    LAUNCH(ctx, k_tp_count_heads, (unsigned)nchunk, 256, 0, keys2.p, ntrip, (uint64_t)ncl, ch_heads.p, ch_valid.p);
    tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ch_heads.p, ch_off.p, nchunk + 1, ctx->stream);
    tmp.alloc(ctx->pool, tmp_bytes);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, ch_heads.p, ch_off.p, nchunk + 1, ctx->stream));
    int64_t nnz = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nnz, ch_off.p + nchunk, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    // valid triplets = sum of ch_valid
    DevBuf<int64_t> vsum;
    vsum.alloc(ctx->pool, 1);
    size_t tb2 = 0;
    cub::DeviceReduce::Sum(nullptr, tb2, ch_valid.p, vsum.p, nchunk, ctx->stream);
    DevBuf<char> tmp2;
    tmp2.alloc(ctx->pool, tb2);
    CUDA_CHECK(cub::DeviceReduce::Sum(tmp2.p, tb2, ch_valid.p, vsum.p, nchunk, ctx->stream));
    int64_t nvalid = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nvalid, vsum.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 4;

    ctx->nnz = nnz;
    ctx->tp.ntrip_valid = nvalid;
    ctx->tp.seg_start.alloc(ctx->pool, nnz + 1);
    ctx->rowval.alloc(ctx->pool, nnz > 0 ? nnz : 1);
    ctx->colptr.alloc(ctx->pool, ncl + 1);
    DevBuf<unsigned long long> colcnt;
    DevBuf<int64_t> colex;
    colcnt.alloc(ctx->pool, ncl + 1); colex.alloc(ctx->pool, ncl + 1);
    CUDA_CHECK(cudaMemsetAsync(colcnt.p, 0, (ncl + 1) * sizeof(unsigned long long), ctx->stream));
    LAUNCH(ctx, k_tp_fill_heads, (unsigned)nchunk, 256, 0, keys2.p, ntrip, (uint64_t)ncl, ch_off.p, ctx->tp.seg_start.p, ctx->rowval.p, colcnt.p);
    CUDA_CHECK(cudaMemcpyAsync(ctx->tp.seg_start.p + nnz, &nvalid, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (int64_t *)colcnt.p, colex.p, ncl + 1, ctx->stream);
    tmp.alloc(ctx->pool, tmp_bytes);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, (int64_t *)colcnt.p, colex.p, ncl + 1, ctx->stream));
    LAUNCH(ctx, k_tp_colptr_from_counts, grid_for(ncl + 1, 256), 256, 0, colex.p, ncl, nnz, ctx->colptr.p);
    ctx->launches += 2;
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->tp.Ke.alloc(ctx->pool, ntrip);
    ctx->nzval.alloc(ctx->pool, nnz > 0 ? nnz : 1);
}

// numeric kernel 1: all element matrices, Ke[k*nel + e] (coalesced across elements)
template <class F> struct KeEmit {
    static constexpr bool TRI = false;
    double *__restrict__ Ke;
    int64_t e, nel;
    template <int J> __device__ __forceinline__ void col(const double (&out)[F::ND]) {
#pragma unroll
        for (int i = 0; i < F::ND; i++)
            if (F::mask(i, J)) Ke[(int64_t)F::kidx(i, J) * nel + e] = out[i];
    }
};

template <class F, bool S>
__global__ void __launch_bounds__(128) k_tp_elem_matrices(const int32_t *__restrict__ gconn, const double2 *__restrict__ gxy, const double *__restrict__ gz,
                                                          int64_t nel, double *__restrict__ Ke)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nel; e += stride) {
        double X[F::GK], Y[F::GK];
        load_xy<F::GK>(gconn, gxy, e, X, Y);
        KeEmit<F> emit{Ke, e, nel};
        if constexpr (form_dim3<F>::value) {        // FEH1_T4: third coordinate plane
            double Z[F::GK];
#pragma unroll
            for (int a = 0; a < F::GK; a++) Z[a] = __ldg(&gz[gconn[e * F::GK + a]]);
            F::template element3<S>(X, Y, Z, 0xffffffffu, emit);
        } else {
            F::template element<S>(X, Y, 0xffffffffu, emit);
        }
    }
}

// numeric kernel 2: nonzero s = left-to-right sum of its contributions (append order)
template <int NT>
__global__ void k_tp_gather(const uint32_t *__restrict__ perm, const int64_t *__restrict__ seg_start,
                            const double *__restrict__ Ke, int64_t nel, int64_t nnz, double *__restrict__ nzval)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz; s += stride) {
        const int64_t p0 = seg_start[s], p1 = seg_start[s + 1];
        uint32_t t = perm[p0];
        double acc = Ke[(int64_t)(t % NT) * nel + (t / NT)];
        for (int64_t p = p0 + 1; p < p1; p++) {
            t = perm[p];
            acc = __dadd_rn(acc, Ke[(int64_t)(t % NT) * nel + (t / NT)]);
        }
        nzval[s] = acc;
    }
}

template <class F> void twopass_numeric(efg_ctx *ctx)
{
    const MeshDev &gm = ctx->mesh[F::GMESH];
    const int64_t nel = ctx->mesh[0].nel;
    if (ctx->opt_strict)
        LAUNCH(ctx, (k_tp_elem_matrices<F, true>), grid_for(nel, 128, 148 * 32), 128, 0, gm.conn.p, gm.xy.p, gm.z.p, nel, ctx->tp.Ke.p);
    else
        LAUNCH(ctx, (k_tp_elem_matrices<F, false>), grid_for(nel, 128, 148 * 32), 128, 0, gm.conn.p, gm.xy.p, gm.z.p, nel, ctx->tp.Ke.p);
    LAUNCH(ctx, k_tp_gather<F::NT>, grid_for(ctx->nnz, 256, 148 * 32), 256, 0, ctx->tp.perm.p, ctx->tp.seg_start.p,
           ctx->tp.Ke.p, nel, ctx->nnz, ctx->nzval.p);
    ctx->numeric_launches += 2;
}
