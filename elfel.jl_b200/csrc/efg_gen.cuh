// efg_gen.cuh -- SURVEY 8f row f4: the inputs of the path made ON the device, behind the C ABI.
//   efg_gen_mesh      T3block / Q4block / T6block of MeshSteward as the reference's examples call them
//                     (examples/heat/poisson/t3.jl:34, q4.jl:24, examples/elasticity/stretch/t6.jl:34): nodes x-fastest,
//                     elements i outer / j inner, T3 orientation :a (test/qmesh-conn.dat), T6 corner nodes before mid-side
//                     nodes, mid-side nodes in first-encounter order (the host mirror's T6block_fast, closed form)
//   efg_gen_space     FESpace on a generated (or uploaded) mesh: ncomp dofs per node, nothing prescribed yet
//   efg_setebc_box    setebc!(fesp, 0, i, comp, 0.0) for every node i of vselect(geom; box = [x0 x1 y0 y1])
//   efg_setebc_nodes  the same for an explicit node list (src/FESpaces.jl:141-173, src/FEFields.jl:124-131)
//   efg_number_dofs   numberdofs!(spaces): numberfreedofs! then numberdatadofs! over the spaces in order, node-major /
//                     component-minor, free dofs first (src/FEFields.jl:137-177, src/FESpaces.jl:250-259)
// Nothing crosses PCIe: config 2's mesh and numbering (5.6 GB of Int64/Float64 on the host) exist only as the Int32 /
// Float64 device arrays the symbolic phase reads.  efg_fetch_mesh / efg_fetch_dofnums copy them out (Int64, 1-based) for a
// caller that wants FEField.dofnums or the mesh on the host.
#pragma once
#include "efg_ctx.cuh"

__device__ __forceinline__ double gen_coord(int64_t i, double len, int64_t n) { return __ddiv_rn(__dmul_rn((double)i, len), (double)n); }   // (i * L) / n

__global__ void k_gen_grid_xy(int64_t nL, int64_t nW, double L, double W, double2 *__restrict__ xy)
{
    GRID_STRIDE(k, (nL + 1) * (nW + 1)) xy[k] = make_double2(gen_coord(k % (nL + 1), L, nL), gen_coord(k / (nL + 1), W, nW));
}
__global__ void k_gen_t3(int64_t nL, int64_t nW, int32_t *__restrict__ conn)
{
    GRID_STRIDE(c, nL * nW) {
        const int64_t i = c / nW, j = c % nW;
        const int32_t f = (int32_t)(j * (nL + 1) + i);
        int32_t *e = conn + 6 * c;
        e[0] = f; e[1] = f + 1; e[2] = f + (int32_t)nL + 2;
        e[3] = f; e[4] = f + (int32_t)nL + 2; e[5] = f + (int32_t)nL + 1;
    }
}
__global__ void k_gen_q4(int64_t nL, int64_t nW, int32_t *__restrict__ conn)
{
    GRID_STRIDE(c, nL * nW) {
        const int64_t i = c / nW, j = c % nW;
        const int32_t f = (int32_t)(j * (nL + 1) + i);
        int32_t *e = conn + 4 * c;
        e[0] = f; e[1] = f + 1; e[2] = f + (int32_t)nL + 2; e[3] = f + (int32_t)nL + 1;
    }
}
// T6block: index (among the mid-side nodes) of the "right" edge of cell (i, j); the cell's new edges are numbered
// [bottom (only j == 0)], right, diagonal, top, [left (only i == 0)] in that order, cells walked i outer / j inner
__device__ __forceinline__ int64_t t6_right(int64_t i, int64_t j, int64_t nW)
{
    const int64_t col0 = 4 * nW + 1, coln = 3 * nW + 1;
    const int64_t base = i == 0 ? 0 : col0 + (i - 1) * coln;
    const int64_t per = i == 0 ? 4 : 3;
    return base + (j == 0 ? 1 : 1 + per * j);
}
__global__ void k_gen_t6(int64_t nL, int64_t nW, double L, double W, int32_t *__restrict__ conn, double2 *__restrict__ xy)
{
    const int64_t nv = (nL + 1) * (nW + 1);
    GRID_STRIDE(c, nL * nW) {
        const int64_t i = c / nW, j = c % nW;
        const int64_t right = t6_right(i, j, nW), diag = right + 1, top = right + 2;
        const int64_t bottom = j == 0 ? right - 1 : t6_right(i, j - 1, nW) + 2;
        const int64_t left = i == 0 ? right + 3 : t6_right(i - 1, j, nW);
        const int32_t f = (int32_t)(j * (nL + 1) + i), n = (int32_t)nL;
        int32_t *e = conn + 12 * c;
        e[0] = f; e[1] = f + 1; e[2] = f + n + 2; e[3] = (int32_t)(nv + bottom); e[4] = (int32_t)(nv + right); e[5] = (int32_t)(nv + diag);
        e[6] = f; e[7] = f + n + 2; e[8] = f + n + 1; e[9] = (int32_t)(nv + diag); e[10] = (int32_t)(nv + top); e[11] = (int32_t)(nv + left);
        // coordinates of the edges this cell numbers: 0.5 * (x_a + x_b) of the corner coordinates
        const double x0 = gen_coord(i, L, nL), x1 = gen_coord(i + 1, L, nL), y0 = gen_coord(j, W, nW), y1 = gen_coord(j + 1, W, nW);
        const double xm = __dmul_rn(0.5, __dadd_rn(x0, x1)), ym = __dmul_rn(0.5, __dadd_rn(y0, y1));
        if (j == 0) xy[nv + bottom] = make_double2(xm, __dmul_rn(0.5, __dadd_rn(y0, y0)));
        xy[nv + right] = make_double2(__dmul_rn(0.5, __dadd_rn(x1, x1)), ym);
        xy[nv + diag] = make_double2(xm, ym);
        xy[nv + top] = make_double2(xm, __dmul_rn(0.5, __dadd_rn(y1, y1)));
        if (i == 0) xy[nv + left] = make_double2(__dmul_rn(0.5, __dadd_rn(x0, x0)), ym);
    }
}
__global__ void k_gen_ebc_box(const double2 *__restrict__ xy, int64_t nnodes, int ncomp, int comp /* -1: all */, double x0, double x1, double y0, double y1,
                              uint8_t *__restrict__ isdatum)
{
    GRID_STRIDE(k, nnodes) {
        const double2 p = xy[k];
        if (p.x >= x0 && p.x <= x1 && p.y >= y0 && p.y <= y1)
            for (int c = 0; c < ncomp; c++) if (comp < 0 || c == comp) isdatum[k * ncomp + c] = 1;
    }
}
__global__ void k_gen_ebc_nodes(const int64_t *__restrict__ ids /* 1-based */, int64_t n, int64_t nnodes, int ncomp, int comp, uint8_t *__restrict__ isdatum, int *__restrict__ err)
{
    GRID_STRIDE(t, n) {
        const int64_t k = ids[t] - 1;
        if (k < 0 || k >= nnodes) { *err = 1; continue; }
        for (int c = 0; c < ncomp; c++) if (comp < 0 || c == comp) isdatum[k * ncomp + c] = 1;
    }
}
struct NotU8 { __device__ __forceinline__ int64_t operator()(uint8_t v) const { return v ? 0 : 1; } };
struct IsU8 { __device__ __forceinline__ int64_t operator()(uint8_t v) const { return v ? 1 : 0; } };
// dof numbers (0-based) from the exclusive scans of the free / datum flags
__global__ void k_gen_numbers(const uint8_t *__restrict__ isdatum, const int64_t *__restrict__ freepos, const int64_t *__restrict__ datapos, int64_t n,
                              int64_t free0, int64_t data0, int32_t *__restrict__ dof)
{
    GRID_STRIDE(t, n) dof[t] = (int32_t)(isdatum[t] ? data0 + datapos[t] : free0 + freepos[t]);
}
__global__ void k_gen_index_out(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) { GRID_STRIDE(i, n) out[i] = (int64_t)in[i] + 1; }
__global__ void k_gen_shift(double2 *__restrict__ xy, int64_t n, double dx, double dy)
{
    GRID_STRIDE(k, n) { double2 p = xy[k]; p.x = __dadd_rn(p.x, dx); p.y = __dadd_rn(p.y, dy); xy[k] = p; }
}
__global__ void k_gen_corners(const int32_t *__restrict__ c6, int64_t nel, int32_t *__restrict__ c3, int32_t *__restrict__ maxid)
{
    GRID_STRIDE(e, nel) {
        int32_t mx = 0;
        for (int a = 0; a < 3; a++) { const int32_t v = c6[e * 6 + a]; c3[e * 3 + a] = v; mx = v > mx ? v : mx; }
        atomicMax(maxid, mx);
    }
}
