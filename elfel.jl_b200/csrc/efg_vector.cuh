// efg_vector.cuh -- the callers either side of the matrix path, all on the device:
//   f1  system VECTOR assembly: SysvecAssembler start!/assemble!/finish! fed by a LocalVectorAssembler
//       (src/Assemblers.jl:196-232, src/LocalAssemblers.jl:95-152) for the load term of the heat examples,
//       `fe[j] += N[j]*Q*JxW` (examples/heat/poisson/t3.jl:57, q4.jl:47);
//   f3  the post-processing integrators of the Stokes examples: location(el, qp) (src/FEIterators.jl:227-235) and
//       evaluate_pressure_error / evaluate_velocity_error (examples/stokes/colliding_flow/ht_p2_p1.jl:120-178);
//   f2  what the examples do with K right after finish!: `KT = K*T` (examples/heat/poisson/t3.jl:78) and the
//       partition K[1:nu,1:nu], K[1:nu,nu+1:end] (t3.jl:79, examples/stokes/colliding_flow/ht_p2_p1_gen.jl solve!),
//       so a 10 GB matrix is not copied to the host just to be sliced.
//
// Determinism / parity: the reference adds contributions to val[gi] element by element (ascending element, local
// index ascending), and SparseArrays' K*x accumulates y[r] column by column.  Both orders are reproduced exactly:
// a stable radix sort groups contributions by destination keeping their original order, one thread then sums a
// destination left to right.  No atomics.  All arithmetic here uses explicitly rounded operations (no FMA
// contraction) -- these kernels are memory-bound, so the results are bit-identical to the CPU oracle in every mode.
#pragma once
#include "efg_tiled.cuh"

struct VecData {
    // f1: dof -> contributions (t = e*NEN + local index), grouped by local row, original order kept
    DevBuf<uint32_t> adjptr, adj;
    DevBuf<double> fe;          // NEN x nel, SoA: fe[j*nel + e]
    DevBuf<double> val;         // owned rows
    int64_t nrl = 0;            // owned rows (= nrow, or the ctx's column ranges when sharded)
    int64_t nrow = 0;
    bool have_sym = false, have_val = false;
    // f2: row-major view of the CSC result
    DevBuf<int64_t> rowptr;     // nrow+1
    DevBuf<uint32_t> tperm;     // row-major position -> CSC position
    DevBuf<int32_t> tcol;       // row-major position -> local column
    bool have_csr = false;
    // f2: last extracted block
    DevBuf<int64_t> bcolptr;
    DevBuf<int64_t> bfirst;     // first CSC position of each block column
    int64_t bnnz = 0, br0 = 0, bc0 = 0, bncol = 0;
    bool have_block = false;
    double vec_ms = 0, spmv_ms = 0;
};

static inline VecData *&vec_data(efg_ctx *ctx) { return *reinterpret_cast<VecData **>(&ctx->vec_opaque); }
inline void vec_release(efg_ctx *ctx)
{
    VecData *&d = vec_data(ctx);
    delete d;
    d = nullptr;
}
static inline VecData *vec_get(efg_ctx *ctx)
{
    VecData *&d = vec_data(ctx);
    if (!d) d = new VecData();
    return d;
}
// ---- f1: vector assembly ------------------------------------------------------------------------------
template <int NEN>
__global__ void k_vec_keys(const int32_t *__restrict__ conn, const int32_t *__restrict__ dof, int64_t nel, int64_t nrow, ColMap cm,
                           uint32_t nrl, uint32_t *__restrict__ cnt, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int *__restrict__ err)
{
    GRID_STRIDE(t, nel * NEN) {
        const int32_t d = dof[conn[t]];
        if (d < 0 || d >= nrow) *err = 1;                 // Julia: BoundsError on val[gi]
        const int64_t l = (d < 0 || d >= nrow) ? -1 : cm.local(d);
        if (l >= 0) { atomicAdd(&cnt[l], 1u); keys[t] = (uint32_t)l; } else keys[t] = nrl;
        vals[t] = (uint32_t)t;
    }
}

// one thread per element: fe[j] = sum_q (N[j]*Q)*JxW, quadrature points in order, starting from 0.0
template <int NEN, int NQ>
__global__ void __launch_bounds__(256) k_vec_heat_load(const int32_t *__restrict__ conn, const double2 *__restrict__ xy, int64_t nel,
                                                       double Q, double *__restrict__ fe)
{
    const QTab &tg = c_tab[kind_slot(NEN)];
    const int nq = NQ ? NQ : tg.npts;          // NQ = 0: the rule's size is read at run time (the less common rules)
    GRID_STRIDE(e, nel) {
        double X[NEN], Y[NEN];
        load_xy<NEN>(conn, xy, e, X, Y);
        double f[NEN];
#pragma unroll
        for (int j = 0; j < NEN; j++) f[j] = 0.0;
#pragma unroll
        for (int q = 0; q < nq; q++) {
            // _jac + Jacobian(Val{2}): src/FElements.jl:148-156,120-129 (node-order sum, first term assigned)
            double J00 = __dmul_rn(X[0], tg.gp[q][0][0]), J01 = __dmul_rn(X[0], tg.gp[q][0][1]);
            double J10 = __dmul_rn(Y[0], tg.gp[q][0][0]), J11 = __dmul_rn(Y[0], tg.gp[q][0][1]);
#pragma unroll
            for (int n = 1; n < NEN; n++) {
                J00 = __dadd_rn(J00, __dmul_rn(X[n], tg.gp[q][n][0])); J01 = __dadd_rn(J01, __dmul_rn(X[n], tg.gp[q][n][1]));
                J10 = __dadd_rn(J10, __dmul_rn(Y[n], tg.gp[q][n][0])); J11 = __dadd_rn(J11, __dmul_rn(Y[n], tg.gp[q][n][1]));
            }
            const double JxW = __dmul_rn(__dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01)), tg.w[q]);
#pragma unroll
            for (int j = 0; j < NEN; j++) f[j] = __dadd_rn(f[j], __dmul_rn(__dmul_rn(tg.N[q][j], Q), JxW));
        }
#pragma unroll
        for (int j = 0; j < NEN; j++) fe[(int64_t)j * nel + e] = f[j];
    }
}

// the same on FEH1_T4 (examples/heat/poisson/t4.jl:41-55): 3x3 Jacobian, Jacobian(Val{3}) of src/FElements.jl:138-146
template <int NQ>
__global__ void __launch_bounds__(256) k_vec_heat_load_t4(const int32_t *__restrict__ conn, const double2 *__restrict__ xy, const double *__restrict__ z,
                                                          int64_t nel, double Q, double *__restrict__ fe)
{
    const QTab &tg = c_tab[EFG_TAB_T4];
    GRID_STRIDE(e, nel) {
        double X[4], Y[4], Z[4];
        load_xy<4>(conn, xy, e, X, Y);
#pragma unroll
        for (int a = 0; a < 4; a++) Z[a] = __ldg(&z[conn[e * 4 + a]]);
        double g[4][3], det;
        HeatFormT4<NQ>::template geometry<true>(X, Y, Z, g, det);
        double f[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double JxW = __dmul_rn(det, tg.w[q]);
#pragma unroll
            for (int j = 0; j < 4; j++) f[j] = __dadd_rn(f[j], __dmul_rn(__dmul_rn(tg.N[q][j], Q), JxW));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) fe[(int64_t)j * nel + e] = f[j];
    }
}

// one thread per owned row: val = ((0.0 + c1) + c2) + ... in the reference's order
template <int NEN>
__global__ void k_vec_gather(const uint32_t *__restrict__ adjptr, const uint32_t *__restrict__ adj, const double *__restrict__ fe,
                             int64_t nel, int64_t nrl, double *__restrict__ val)
{
    GRID_STRIDE(r, nrl) {
        double acc = 0.0;
        const uint32_t a1 = adjptr[r + 1];
        for (uint32_t p = adjptr[r]; p < a1; p++) {
            const uint32_t t = adj[p];
            acc = __dadd_rn(acc, fe[(int64_t)(t % NEN) * nel + (t / NEN)]);
        }
        val[r] = acc;
    }
}

template <int NEN> static void vec_symbolic(efg_ctx *ctx, VecData *vd, int64_t nrow)
{
    const MeshDev &m = ctx->mesh[0];
    const int64_t nel = m.nel, np = nel * NEN;
    if (np >= ((int64_t)1 << 32)) efg_throw(EFG_ERR_LIMIT, "vector assembly: nel*nen exceeds 2^32; shard the mesh");
    // owned rows: the ctx's column ranges when the owner-computes sharding is active for an nrow x nrow system
    const bool sharded = ctx->started && ctx->have_range && ctx->ncol == nrow;
    ColMap cm = sharded ? COLMAP(ctx) : ColMap{0, nrow, 0, nullptr, nullptr, nullptr};
    const int64_t nrl = sharded ? ctx->ncl : nrow;
    DevBuf<int> err;
    err.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    DevBuf<uint32_t> cnt;
    cnt.alloc(ctx->pool, (size_t)nrl + 2);
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, ((size_t)nrl + 2) * sizeof(uint32_t), ctx->stream));
    vd->adjptr.alloc(ctx->pool, (size_t)nrl + 2);
    vd->adj.alloc(ctx->pool, (size_t)(np > 0 ? np : 1));
    {
        const size_t n = (size_t)(np > 0 ? np : 1);
        uint32_t *k1 = reinterpret_cast<uint32_t *>(tl_scratch(ctx, 3 * n * sizeof(uint32_t))), *k2 = k1 + n, *v1 = k2 + n;
        LAUNCH(ctx, k_vec_keys<NEN>, grid_for(np, 256), 256, 0, m.conn.p, ctx->space[0].dof.p, nel, nrow, cm, (uint32_t)nrl, cnt.p, k1, v1, err.p);
        if (tl_read(ctx, err.p))
            efg_throw(EFG_ERR_INDEX, "BoundsError: a dof number is < 1 or exceeds nrow (was the space numbered, incl. data dofs?)");
        if (np > 0) tl_sort_pairs(ctx, k1, k2, v1, vd->adj.p, np, bits_for(nrl));
    }
    tl_excl_scan(ctx, cnt.p, vd->adjptr.p, nrl + 1);
    vd->fe.alloc(ctx->pool, (size_t)(np > 0 ? np : 1));
    vd->val.alloc(ctx->pool, (size_t)(nrl > 0 ? nrl : 1));
    vd->nrl = nrl; vd->nrow = nrow;
    vd->have_sym = true;
    ctx->scratch.release();
}

template <int NEN, int NQ> static void vec_numeric_heat(efg_ctx *ctx, VecData *vd, double Q)
{
    const MeshDev &m = ctx->mesh[0];
    LAUNCH(ctx, (k_vec_heat_load<NEN, NQ>), grid_for(m.nel, 256, (int64_t)148 * 32), 256, 0, m.conn.p, m.xy.p, m.nel, Q, vd->fe.p);
    LAUNCH(ctx, k_vec_gather<NEN>, grid_for(vd->nrl, 256, (int64_t)148 * 32), 256, 0, vd->adjptr.p, vd->adj.p, vd->fe.p, m.nel, vd->nrl, vd->val.p);
}

template <int NQ> static void vec_numeric_heat_t4(efg_ctx *ctx, VecData *vd, double Q)
{
    const MeshDev &m = ctx->mesh[0];
    LAUNCH(ctx, (k_vec_heat_load_t4<NQ>), grid_for(m.nel, 256, (int64_t)148 * 32), 256, 0, m.conn.p, m.xy.p, m.z.p, m.nel, Q, vd->fe.p);
    LAUNCH(ctx, k_vec_gather<4>, grid_for(vd->nrl, 256, (int64_t)148 * 32), 256, 0, vd->adjptr.p, vd->adj.p, vd->fe.p, m.nel, vd->nrl, vd->val.p);
}

// ---- f2: K*x in SparseArrays' accumulation order, and sub-blocks of K -------------------------------------
__global__ void k_csr_keys(const int64_t *__restrict__ colptr, int64_t ncl, const int32_t *__restrict__ rowval,
                           uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, int32_t *__restrict__ colof, uint32_t *__restrict__ rowcnt)
{
    GRID_STRIDE(c, ncl) {
        const int64_t p1 = colptr[c + 1] - 1;
        for (int64_t p = colptr[c] - 1; p < p1; p++) {
            const int32_t r = rowval[p];
            keys[p] = (uint32_t)r; vals[p] = (uint32_t)p; colof[p] = (int32_t)c;
            atomicAdd(&rowcnt[r], 1u);
        }
    }
}
__global__ void k_csr_cols(const uint32_t *__restrict__ tperm, const int32_t *__restrict__ colof, int64_t nnz, int32_t *__restrict__ tcol)
{
    GRID_STRIDE(q, nnz) tcol[q] = colof[tperm[q]];
}
// y[r] = ((0.0 + a_rc1*x_c1) + a_rc2*x_c2) + ... , columns ascending, product and sum rounded separately
__global__ void k_spmv_rows(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ tperm, const int32_t *__restrict__ tcol,
                            const double *__restrict__ nzval, const double *__restrict__ x, int64_t nrow, double *__restrict__ y)
{
    GRID_STRIDE(r, nrow) {
        double acc = 0.0;
        const int64_t q1 = rowptr[r + 1];
        for (int64_t q = rowptr[r]; q < q1; q++) acc = __dadd_rn(acc, __dmul_rn(nzval[tperm[q]], __ldg(&x[tcol[q]])));
        y[r] = acc;
    }
}

// Heat forms: K is BITWISE symmetric (the element matrix is built from its upper triangle and both (r,c) and (c,r) sum
// the same element contributions in the same order), so row r of K is column r of the CSC arrays read in place:
// ascending rows of the column = ascending columns of the row = SparseArrays' accumulation order, with contiguous
// reads and no row-major view at all.
__global__ void k_spmv_cols_sym(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowval, const double *__restrict__ nzval,
                                const double *__restrict__ x, int64_t n, double *__restrict__ y)
{
    GRID_STRIDE(c, n) {
        double acc = 0.0;
        const int64_t p1 = colptr[c + 1] - 1;
        for (int64_t p = colptr[c] - 1; p < p1; p++) acc = __dadd_rn(acc, __dmul_rn(nzval[p], __ldg(&x[rowval[p]])));
        y[c] = acc;
    }
}

static void vec_build_csr(efg_ctx *ctx, VecData *vd)
{
    const int64_t nnz = ctx->nnz, nrow = ctx->nrow, ncl = ctx->ncl;
    if (nnz >= ((int64_t)1 << 32)) efg_throw(EFG_ERR_LIMIT, "efg_spmv: nnz exceeds 2^32");
    const size_t n = (size_t)(nnz > 0 ? nnz : 1);
    DevBuf<uint32_t> rowcnt;
    rowcnt.alloc(ctx->pool, (size_t)nrow + 2);
    CUDA_CHECK(cudaMemsetAsync(rowcnt.p, 0, ((size_t)nrow + 2) * sizeof(uint32_t), ctx->stream));
    vd->rowptr.alloc(ctx->pool, (size_t)nrow + 2);
    vd->tperm.alloc(ctx->pool, n);
    vd->tcol.alloc(ctx->pool, n);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(tl_scratch(ctx, 4 * n * sizeof(uint32_t))), *k2 = k1 + n, *v1 = k2 + n;
    int32_t *colof = reinterpret_cast<int32_t *>(v1 + n);
    LAUNCH(ctx, k_csr_keys, grid_for(ncl, 128), 128, 0, ctx->colptr.p, ncl, ctx->rowval.p, k1, v1, colof, rowcnt.p);
    if (nnz > 0) tl_sort_pairs(ctx, k1, k2, v1, vd->tperm.p, nnz, bits_for(nrow));     // stable: columns stay ascending inside a row
    LAUNCH(ctx, k_csr_cols, grid_for(nnz, 256), 256, 0, vd->tperm.p, colof, nnz, vd->tcol.p);
    {
        cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const uint32_t *> it(rowcnt.p, cub::CastOp<int64_t>());
        tl_excl_scan(ctx, it, vd->rowptr.p, nrow + 1);
    }
    vd->have_csr = true;
    ctx->scratch.release();
}

// entries of column c (local) with r0 <= row < r1: rows are ascending inside a column
__global__ void k_block_count(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowval, int64_t bc0, int64_t bncol,
                              int32_t r0, int32_t r1, int64_t *__restrict__ cnt, int64_t *__restrict__ first)
{
    GRID_STRIDE(j, bncol) {
        const int64_t p0 = colptr[bc0 + j] - 1, p1 = colptr[bc0 + j + 1] - 1;
        int64_t lo = p0, hi = p1;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < r0) lo = mid + 1; else hi = mid; }
        const int64_t a = lo;
        hi = p1;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < r1) lo = mid + 1; else hi = mid; }
        cnt[j] = lo - a;
        first[j] = a;
    }
}
__global__ void k_block_colptr_out(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) { GRID_STRIDE(i, n) out[i] = in[i] + 1; }
// one warp per block column: copy its entries (rows rebased to the block, 1-based Int64)
__global__ void k_block_copy(const int64_t *__restrict__ bcolptr, const int64_t *__restrict__ first, int64_t bncol, const int32_t *__restrict__ rowval,
                             const double *__restrict__ nzval, int64_t r0, int64_t *__restrict__ orow, double *__restrict__ oval)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = w; j < bncol; j += nw) {
        const int64_t o = bcolptr[j], n = bcolptr[j + 1] - o, a = first[j];
        for (int64_t k = lane; k < n; k += 32) {
            if (orow) orow[o + k] = (int64_t)rowval[a + k] - r0 + 1;
            if (oval) oval[o + k] = nzval[a + k];
        }
    }
}

// ---- f3: location(el, qp) and the L2 error integrators ---------------------------------------------------
// out[(e*NQ + q)*2 + {0,1}] = sum_i x_i * N_i(q): first term assigned, node order, individually rounded operations
template <int NEN, int NQ>
__global__ void __launch_bounds__(256) k_qp_locations(const int32_t *__restrict__ conn, const double2 *__restrict__ xy, int64_t nel, double *__restrict__ out)
{
    const QTab &tg = c_tab[kind_slot(NEN)];
    const int nq = NQ ? NQ : tg.npts;
    GRID_STRIDE(e, nel) {
        double X[NEN], Y[NEN];
        load_xy<NEN>(conn, xy, e, X, Y);
#pragma unroll
        for (int q = 0; q < nq; q++) {
            double lx = __dmul_rn(X[0], tg.N[q][0]), ly = __dmul_rn(Y[0], tg.N[q][0]);
#pragma unroll
            for (int i = 1; i < NEN; i++) { lx = __dadd_rn(lx, __dmul_rn(X[i], tg.N[q][i])); ly = __dadd_rn(ly, __dmul_rn(Y[i], tg.N[q][i])); }
            reinterpret_cast<double2 *>(out)[e * nq + q] = make_double2(lx, ly);
        }
    }
}

struct ErrComp { const int32_t *dof; int ncs, comp; const int32_t *cdof = nullptr; };    // cdof: cell dofs of an FEH1_T3_BUBBLE space (row f5)

// one thread per element: sum over its quadrature points of JxW * sum_c (field_c - truth_c)^2, in the reference's
// operation order; the element sums are then added by a fixed-shape tree (cub::DeviceReduce), so the result is
// reproducible run to run and agrees with the CPU loop's single running sum to rounding (tests: 1e-12 relative)
template <int NEN, int NQ, int NC>
__global__ void __launch_bounds__(256) k_l2_error_elem(const int32_t *__restrict__ conn, const double2 *__restrict__ xy, int64_t nel, ErrComp c0, ErrComp c1,
                                                       const double *__restrict__ U, int64_t nU, const double *__restrict__ truth,
                                                       double *__restrict__ eout, int *__restrict__ err)
{
    const QTab &tg = c_tab[kind_slot(NEN)];
    const int nq = NQ ? NQ : tg.npts;
    GRID_STRIDE(e, nel) {
        double X[NEN], Y[NEN], v0[NEN], v1[NEN];
#pragma unroll
        for (int a = 0; a < NEN; a++) {
            const int32_t n = conn[e * NEN + a];
            const double2 p = __ldg(&xy[n]);
            X[a] = p.x; Y[a] = p.y;
            const int32_t d0 = c0.dof[(int64_t)n * c0.ncs + c0.comp];
            const int32_t d1 = NC > 1 ? c1.dof[(int64_t)n * c1.ncs + c1.comp] : 0;
            if (d0 < 0 || d0 >= nU || d1 < 0 || d1 >= nU) { *err = 1; v0[a] = v1[a] = 0.0; continue; }
            v0[a] = U[d0];
            v1[a] = NC > 1 ? U[d1] : 0.0;
        }
        // FEH1_T3_BUBBLE: eldofvals = vertex values, then the bubble's (examples/stokes/colliding_flow/p1b_p1.jl:150-160)
        const bool bub = NEN == 3 && c0.cdof != nullptr;
        double b0 = 0.0, b1 = 0.0;
        if (bub) {
            const int32_t d0 = c0.cdof[e], d1 = NC > 1 ? c1.cdof[e] : 0;
            if (d0 < 0 || d0 >= nU || d1 < 0 || d1 >= nU) *err = 1;
            else { b0 = U[d0]; b1 = NC > 1 ? U[d1] : 0.0; }
        }
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < nq; q++) {
            double J00 = __dmul_rn(X[0], tg.gp[q][0][0]), J01 = __dmul_rn(X[0], tg.gp[q][0][1]);
            double J10 = __dmul_rn(Y[0], tg.gp[q][0][0]), J11 = __dmul_rn(Y[0], tg.gp[q][0][1]);
#pragma unroll
            for (int n = 1; n < NEN; n++) {
                J00 = __dadd_rn(J00, __dmul_rn(X[n], tg.gp[q][n][0])); J01 = __dadd_rn(J01, __dmul_rn(X[n], tg.gp[q][n][1]));
                J10 = __dadd_rn(J10, __dmul_rn(Y[n], tg.gp[q][n][0])); J11 = __dadd_rn(J11, __dmul_rn(Y[n], tg.gp[q][n][1]));
            }
            const double JxW = __dmul_rn(__dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01)), tg.w[q]);
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < NEN; j++) {
                a0 = __dadd_rn(a0, __dmul_rn(v0[j], tg.N[q][j]));
                if (NC > 1) a1 = __dadd_rn(a1, __dmul_rn(v1[j], tg.N[q][j]));
            }
            if (bub) {
                const double Nb = c_tab[3].N[q][3];
                a0 = __dadd_rn(a0, __dmul_rn(b0, Nb));
                if (NC > 1) a1 = __dadd_rn(a1, __dmul_rn(b1, Nb));
            }
            const double *t = truth + (e * nq + q) * NC;
            const double d0 = __dsub_rn(a0, t[0]);
            double sq = __dmul_rn(d0, d0);
            if (NC > 1) { const double d1 = __dsub_rn(a1, t[1]); sq = __dadd_rn(sq, __dmul_rn(d1, d1)); }
            acc = __dadd_rn(acc, __dmul_rn(JxW, sq));
        }
        eout[e] = acc;
    }
}

template <int NEN, int NQ> static void vec_locations(efg_ctx *ctx, const MeshDev &m, double *out)
{
    LAUNCH(ctx, (k_qp_locations<NEN, NQ>), grid_for(m.nel, 256), 256, 0, m.conn.p, m.xy.p, m.nel, out);
}
template <int NEN, int NQ> static void vec_l2_elem(efg_ctx *ctx, const MeshDev &m, int nc, ErrComp c0, ErrComp c1, const double *U, int64_t nU,
                                                   const double *truth, double *eout, int *err)
{
    if (nc == 1) LAUNCH(ctx, (k_l2_error_elem<NEN, NQ, 1>), grid_for(m.nel, 256), 256, 0, m.conn.p, m.xy.p, m.nel, c0, c1, U, nU, truth, eout, err);
    else LAUNCH(ctx, (k_l2_error_elem<NEN, NQ, 2>), grid_for(m.nel, 256), 256, 0, m.conn.p, m.xy.p, m.nel, c0, c1, U, nU, truth, eout, err);
}
// (element kind, number of quadrature points) -> instantiation
template <class Fn> static bool vec_dispatch_kq(int kind, int npts, Fn &&fn)
{
    switch (kind * 100 + npts) {
    case 301: fn(std::integral_constant<int, 3>{}, std::integral_constant<int, 1>{}); return true;
    case 303: fn(std::integral_constant<int, 3>{}, std::integral_constant<int, 3>{}); return true;
    case 601: fn(std::integral_constant<int, 6>{}, std::integral_constant<int, 1>{}); return true;
    case 603: fn(std::integral_constant<int, 6>{}, std::integral_constant<int, 3>{}); return true;
    case 401: fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{}); return true;
    case 404: fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{}); return true;
    case 409: fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 9>{}); return true;
    }
    // the less common rules: the number of points is read from the table at run time (NQ = 0)
    const bool tri_rule = npts == 4 || npts == 6 || npts == 7 || npts == 9 || npts == 12 || npts == 13;
    if (kind == 3 && tri_rule) { fn(std::integral_constant<int, 3>{}, std::integral_constant<int, 0>{}); return true; }
    if (kind == 6 && tri_rule) { fn(std::integral_constant<int, 6>{}, std::integral_constant<int, 0>{}); return true; }
    if (kind == 4 && (npts == 16 || npts == 25)) { fn(std::integral_constant<int, 4>{}, std::integral_constant<int, 0>{}); return true; }
    return false;
}
