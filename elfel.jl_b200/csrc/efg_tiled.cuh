// efg_tiled.cuh -- the fused path: owner-computes tiles, element matrices never leave the SM.
//
// Data layout in HBM (built once by the symbolic phase, all on the GPU):
//   tiles[T]   TileDesc: one CTA work item = a compact group of elements (contiguous in a space-
//              filling-curve order of the mesh) + the halo elements that touch the matrix columns
//              the tile OWNS (owner of a column = lowest tile among the elements containing its dof).
//   tconn      geometry connectivity of every tile element (incl. halo), tile order, int32
//   tmask      which local columns of that element the tile owns, uint16
//   pk         for every owned nonzero ("slot"), in (column,row) order: one 32-bit word naming where its first
//              and second contribution sit in the CTA's shared-memory stage (reference append order: ascending
//              element, column-major inside an element); hidx/heavy: lists for nonzeros with more contributions
//   runs       maximal groups of owned columns that are contiguous in nzval
// Numeric kernel, per tile: (1) one thread per tile element: coordinates -> Jacobian/gradients ->
// the owned columns of the element matrix, in registers -> shared-memory stage, (2) one thread per
// owned nonzero: left-to-right sum of its contributions from the stage (deterministic, no atomics),
// (3) coalesced store of nzval.  Every nonzero is written exactly once; nothing else is written.
#pragma once
#include <cub/cub.cuh>
#include <climits>
#include "efg_ctx.cuh"

#define TL_CAP 96          // max distinct rows in one matrix column (tiled path)
#define TL_MAXND 16

struct TileDescFull {
    int64_t slot0;          // first tile-order slot            (symbolic phase only)
    int64_t gidx0;          // first tile-order heavy contribution (symbolic phase only)
    int64_t elem0;          // first tile element (index into tconn/tmask)
    int64_t meta0;          // byte offset of the tile's metadata block (16-byte aligned)
    int32_t nelem;          // tile elements incl. halo
    int32_t nslot;          // owned nonzeros
    int32_t nrun;           // runs of nonzeros contiguous in nzval
    int32_t nq;             // staged columns (= stage row stride)
    int32_t ncontrib;       // contributions to the HEAVY owned nonzeros (entries of hidx)
    int32_t meta_bytes;     // size of the metadata block (multiple of 16)
    int32_t run0;           // first run (symbolic phase only)
    int32_t nheavy;         // owned nonzeros with more than TL_LIGHT contributions
    int64_t heavy0;         // first heavy entry in tile order (symbolic phase only)
    uint16_t qbase[TL_MAXND + 1]; // qbase[r] = first staged column of "r-th owned column of an element"
    uint16_t nqs;           // stage row stride: nq rounded up to an odd number (rows i of one staged column fall in distinct banks)
    uint16_t pad2_[2];
    int64_t geo0;           // byte offset of the tile's geometry block (TL_GEO)
    int32_t geo_bytes;      // its size (multiple of 16); 0 = the form reads tconn/xy from global memory
    int32_t nnode;          // unique nodes of the tile's elements
};
static_assert(sizeof(TileDescFull) % 8 == 0, "TileDescFull is copied as 8-byte words");

// metadata block of a tile, each part 16-byte aligned:
//   [pk: u32 x nslot]      per owned nonzero: stage index of its 1st | 2nd << 16 contribution (0xFFFF = none);
//                          0xFFFEFFFE marks a "heavy" nonzero (> TL_LIGHT contributions), 0xFFFFFFFF a padding slot
//   [hidx: u16 x ncontrib] stage indices of the contributions of the heavy nonzeros, append order
//   [rowbase: i64 x nslot/32] index in nzval of the first slot of every row of 32 slots   [heavy: TileHeavy x nheavy]
__host__ __device__ static inline int tl_align16(int b) { return (b + 15) & ~15; }
// SPLIT forms keep per-element geometry (gradients + JxW at every quadrature point, SoA) and the column masks in
// shared memory between the two steps of phase 1
template <class F> __host__ __device__ constexpr int tl_gsz() { return F::SPLIT ? F::NQ * (2 * F::BK + 1) : 0; }
__host__ __device__ static inline int tl_geo_bytes(int gsz, int nelem) { return gsz ? tl_align16(nelem * (gsz * 8 + 2)) : 0; }
__host__ __device__ static inline int tl_meta_goff_bytes(int nslot) { return tl_align16(4 * nslot); }          // pk
__host__ __device__ static inline int tl_meta_gidx_bytes(int nc) { return tl_align16(2 * nc); }              // hidx
struct TileHeavy { uint16_t s, o, c, pad; };
// row table: the owned nonzeros of a tile are laid out in ROWS of 32 slots that never straddle a run (every run of nonzeros
// contiguous in nzval starts at a multiple of 32 slots; the tail of its last row is padding, pk = 0xFFFFFFFF), so the
// destination of lane l of row r is simply rowbase[r] + l
__host__ __device__ static inline int tl_meta_rows_bytes(int nslot) { return tl_align16(8 * (nslot >> 5)); }
// geometry block of a tile, each part 16-byte aligned: [txy: double2 x nnode] [conn16: u16 x GK x nelem] [mask16: u16 x nelem]
__host__ __device__ static inline int tl_geo_xy_bytes(int nnode) { return 16 * nnode; }
__host__ __device__ static inline int tl_geo_conn_bytes(int gk, int nelem) { return tl_align16(2 * gk * nelem); }
// 3-D meshes (FEH1_T4) append [tz: double x nnode]
__host__ __device__ static inline int tl_geo_z_off(int gk, int nelem, int nnode) { return tl_geo_xy_bytes(nnode) + tl_geo_conn_bytes(gk, nelem) + tl_align16(2 * nelem); }
__host__ __device__ static inline int tl_geo_block_bytes(int gk, int nelem, int nnode, int dim3 = 0)
{
    return tl_geo_z_off(gk, nelem, nnode) + (dim3 ? tl_align16(8 * nnode) : 0);
}
#define TL_GEO_CAP 4096   // nelem*GK entries a tile's local numbering is built from (256 threads x 16)   // tile slot, first entry in hidx, number of contributions
// Gather variants that were A/B-measured SLOWER on B200 and removed (profiles/README.md): a branch-free light gather
// through a neutral -0.0 stage entry (34 % fewer instructions in the loop, one more LDS per single-contribution
// nonzero: +10 %), warp-uniform two-run destination tracking (+2 %), 8 nonzeros in flight per lane (+13 %).  The
// kernel is bound by the LSU data pipe and by latency, not by instruction issue.
#ifndef TL_U
#define TL_U 4          // light nonzeros in flight per lane
#endif
#ifndef TL_MINB
#define TL_MINB 2       // CTAs per SM of the one-thread-per-element kernels with 5-8 local dofs (T6 heat); see TL_BLOCK_NS
#endif
#ifndef TL_GEO
#define TL_GEO 1        // one-thread-per-element forms: the tile's unique node coordinates + 16-bit local connectivity +
                        // column masks form a "geometry block" fetched by ONE TMA bulk copy at tile start (no dependent
                        // global loads in phase 1); 0: per-element global loads of tconn -> xy
#endif
// Persistent CTAs (one per CTA slot) that fetch the NEXT tile's geometry block while the current tile is in its gather
// phase, and its gather metadata while the next tile computes.  Measured (profiles/README.md): no gain for the
// one-thread-per-element forms with large tiles (T6 heat 3.96 vs 3.86 ms, Q4 2.11 vs 2.06 ms; T3 0.78 vs 0.80 ms).
#ifndef TL_PERSIST_NS
#define TL_PERSIST_NS 0
#endif
#ifndef TL_PERSIST_SPLIT
#define TL_PERSIST_SPLIT 0
#endif
template <class F> constexpr bool tl_persist() { return TL_GEO && (F::SPLIT ? TL_PERSIST_SPLIT : TL_PERSIST_NS); }
#ifndef TL_GATHER_SELECT
#define TL_GATHER_SELECT 0
#endif
#ifndef TL_U_SPLIT
#define TL_U_SPLIT 3     // rows of slots in flight per warp in the light walk of the vector forms (elasticity: 4 / 3 / 2 rows -> 3.290 / 3.255 /
#endif                   // 3.269 ms; their CTAs are at the register cap, fewer rows in flight leave more registers to phase 1)
#ifndef TL_U_SPLIT15
#define TL_U_SPLIT15 1   // the same for the 15-column Stokes forms (Stokes gen, 4 / 3 / 2 / 1 rows: 1.390 / 1.380 / 1.375 / 1.369 ms)
#endif
#ifndef TL_TAIL_PASS
#define TL_TAIL_PASS 1   // Stokes forms: pressure columns from their own item space (see tl_phase1b_pairs)
#endif
#ifndef TL_GEO_QSPLIT
#define TL_GEO_QSPLIT 0  // vector forms, phase 1a: one thread per (element, quadrature point) instead of one per element: SLOWER
                         // (elasticity 3.29 -> 3.52 ms, Stokes gen 1.444 -> 1.470): idle threads cost nothing while the co-resident CTA
                         // runs, the repeated coordinate loads and run-time table indices do (profiles/r2_ab_cta_shapes_and_layouts.txt)
#endif
#ifndef TL_GEO_SPLIT
#define TL_GEO_SPLIT 1  // vector forms: phase 1a reads the tile's geometry block too (the block sits behind the shared
                        // geometry/metadata area)
#endif
#ifndef TL_GS_ALIAS
#define TL_GS_ALIAS 1   // vector forms: 1 = the per-element geometry Gs of phase 1 shares its shared-memory area with the gather
                        // metadata (whose TMA copy can then only be issued after phase 1); 0 = separate areas, the copy is
                        // issued at tile start like for the scalar forms, at the price of smaller tiles
#endif
#ifndef TL_GS_ALIAS15
#define TL_GS_ALIAS15 1 // the same switch for the 15-dof Stokes forms (their tiles are at the 32-element floor already)
#endif
#ifndef TL_PAIRS
#define TL_PAIRS 1      // vector forms, phase 1b: one thread per node (both dofs' columns) instead of one per column
#endif
#ifndef TL_SYM_STAGE
#define TL_SYM_STAGE 1  // forms with a bitwise-symmetric element matrix (heat) stage its upper triangle only: ND(ND+1)/2 rows of
                        // one entry per tile element instead of ND rows of one entry per (element, owned column)
#endif
template <class F> __host__ __device__ constexpr bool tl_gs_alias() { return F::ND >= 15 ? TL_GS_ALIAS15 : TL_GS_ALIAS; }
template <class F> __host__ __device__ constexpr bool tl_sym() { return TL_SYM_STAGE && F::SYM && !F::SPLIT; }
template <class F> __host__ __device__ constexpr int tl_srows() { return tl_sym<F>() ? F::ND * (F::ND + 1) / 2 : F::ND; }
// stage row of entry (a, b), a <= b, of an nd x nd upper triangle: the DIAGONAL entries take rows 0 .. nd-1 (a contribution
// to a diagonal nonzero of the matrix is recognised by its stage index alone: fused load vector), the others follow column by column
__host__ __device__ constexpr int tl_tri(int a, int b, int nd) { return a == b ? a : nd + b * (b - 1) / 2 + a; }
// fused load vector (examples/heat/poisson/t3.jl:57: fe[j] += N[j]*Q*JxW in the loop that builds ke):
// forms that can stage fe next to their element matrix
template <class F> __host__ __device__ constexpr bool tl_can_fuse() { return TL_SYM_STAGE && F::SYM && !F::SPLIT && F::GK == F::ND && !form_dim3<F>::value; }
#ifndef TL_VEC_CONN
#define TL_VEC_CONN 1   // tile connectivity read with 8/16-byte loads (T6: 3 x int2, Q4: 1 x int4) instead of 4-byte loads
#endif
#define TL_PK_NONE 0xFFFFu
#define TL_PK_HEAVY 0xFFFEFFFEu
// bytes of the stage: ND rows of nqs staged columns
__host__ __device__ static inline int tl_stage_bytes(int nd, int nqs) { return tl_align16(nd * nqs * 8); }

// what the pattern phase (A) leaves for the tile phase (B); released once the tiles exist
struct TiledSym {
    DevBuf<int32_t> edof;            // combined element dof vectors, ND x nel
    DevBuf<uint32_t> adj, adjptr;    // column -> (element, local column) pairs, ascending element
    DevBuf<uint32_t> eorder;         // elements along the space-filling curve (empty: natural order)
    DevBuf<uint8_t> colcnt, hcnt;    // rows per column, heavy nonzeros per column
    DevBuf<uint16_t> ccnt;           // contributions to the heavy nonzeros of a column
    DevBuf<int64_t> colptr0;         // 0-based column offsets
    int64_t npairs = 0;
    bool have_order = false;
};

struct TiledData {
    DevBuf<TileDescFull> tiles;
    DevBuf<int32_t> tconn;
    DevBuf<uint16_t> tmask;
    DevBuf<unsigned char> meta;  // per-tile metadata blocks (bulk-copied to shared memory by the numeric kernel)
    DevBuf<unsigned char> geo;   // per-tile geometry blocks (TL_GEO)
    int64_t geo_total = 0;
    int64_t ntelem = 0, ncontrib = 0, nruns = 0, meta_bytes = 0;
    int smem_bytes = 0, stage_bytes = 0, meta_max = 0;
    int off_meta = 0, off_geo = 0, off_gs = 0;   // persistent kernel: fixed shared-memory offsets (maxima over tiles)
    bool persist = false;
    bool complete = false;       // false: the tile phase stopped early (shared-memory footprint too large for this tile size)
    int block = 256;
};

static inline TiledData *&tiled_data(efg_ctx *ctx)
{
    static_assert(sizeof(void *) == 8, "");
    return *reinterpret_cast<TiledData **>(&ctx->tl_opaque);
}
static inline TiledSym *&tiled_sym(efg_ctx *ctx) { return *reinterpret_cast<TiledSym **>(&ctx->tl_sym_opaque); }

inline void tiled_release_tiles(efg_ctx *ctx)
{
    TiledData *&d = tiled_data(ctx);
    delete d;
    d = nullptr;
    ctx->tl.ntiles = 0;
}
inline void tiled_release(efg_ctx *ctx)
{
    tiled_release_tiles(ctx);
    TiledSym *&y = tiled_sym(ctx);
    delete y;
    y = nullptr;
}

// ---- CUB helpers ---------------------------------------------------------------------------------
template <class K, class V>
static void tl_sort_pairs(efg_ctx *ctx, const K *kin, K *kout, const V *vin, V *vout, int64_t n, int end_bit)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, end_bit, ctx->stream);
    DevBuf<char> tmp;
    tmp.alloc(ctx->pool, tb);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin, kout, vin, vout, n, 0, end_bit, ctx->stream));
    ctx->launches += (end_bit + 7) / 8 + 2;
}   // temporaries are freed stream-ordered: no synchronisation needed
template <class K> static void tl_sort_keys(efg_ctx *ctx, const K *kin, K *kout, int64_t n, int end_bit)
{
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, kin, kout, n, 0, end_bit, ctx->stream);
    DevBuf<char> tmp;
    tmp.alloc(ctx->pool, tb);
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, kin, kout, n, 0, end_bit, ctx->stream));
    ctx->launches += (end_bit + 7) / 8 + 2;
}   // temporaries are freed stream-ordered: no synchronisation needed
template <class In, class Out> static void tl_excl_scan(efg_ctx *ctx, In in, Out out, int64_t n)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, ctx->stream);
    DevBuf<char> tmp;
    tmp.alloc(ctx->pool, tb);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, ctx->stream));
    ctx->launches += 2;
}
static inline char *tl_scratch(efg_ctx *ctx, size_t bytes)
{
    if (ctx->scratch.n < bytes) ctx->scratch.alloc(ctx->pool, bytes + bytes / 8);
    return ctx->scratch.p;
}
static inline int bits_for(int64_t maxval)
{
    int b = 1;
    while (((int64_t)1 << b) <= maxval) b++;
    return b;
}
// Small device -> host read-backs (counts, maxima, error flags) go through a page-locked, device-mapped "mailbox": a
// one-thread kernel stores the words there and the host reads them after the stream has drained.  No copy engine is
// involved, so a read-back of the symbolic phase never queues behind the chunks of a pattern fetch that is using the
// device -> host engine at that moment (measured: the tile phase of config 2 took 83 ms instead of 51 ms next to it).
__global__ void k_mailbox(const unsigned char *__restrict__ src, int nbytes, unsigned char *__restrict__ dst)
{
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
static void tl_read_bytes(efg_ctx *ctx, const void *dptr, void *out, int nbytes)
{
    if (!ctx->mailbox) {
        CUDA_CHECK(cudaHostAlloc(&ctx->mailbox, 256, cudaHostAllocMapped));
        CUDA_CHECK(cudaHostGetDevicePointer(&ctx->mailbox_dev, ctx->mailbox, 0));
    }
    if (nbytes > 256) efg_throw(EFG_ERR_INVALID, "internal: mailbox read of %d bytes", nbytes);
    k_mailbox<<<1, 32, 0, ctx->stream>>>(static_cast<const unsigned char *>(dptr), nbytes, static_cast<unsigned char *>(ctx->mailbox_dev));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->mailbox, (size_t)nbytes);
}
template <class T> static T tl_read(efg_ctx *ctx, const T *dptr)
{
    T h;
    tl_read_bytes(ctx, dptr, &h, (int)sizeof(T));
    return h;
}

// EFG_TRACE=1: wall-clock per symbolic step on stderr (synchronises; for diagnosing host-side stalls)
#include <cstring>
#include <chrono>
#include <cstdlib>
struct TlTrace {
    bool on; std::chrono::steady_clock::time_point t; efg_ctx *ctx;
    explicit TlTrace(efg_ctx *c) : on(getenv("EFG_TRACE") != nullptr), t(std::chrono::steady_clock::now()), ctx(c) {}
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(ctx->stream);
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[efg trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

#define GRID_STRIDE(i, n) \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride__ = (int64_t)gridDim.x * blockDim.x; i < (n); i += stride__)

// ---- symbolic kernels ------------------------------------------------------------------------------
template <class F>
__global__ void k_tl_edofs(DofSrc src, int64_t nel, int64_t nrow, int64_t ncol, int32_t *__restrict__ edof, int *__restrict__ err)
{
    GRID_STRIDE(e, nel) {
        int32_t d[F::ND];
        F::edofs(src, e, d);
        bool bad = false;
#pragma unroll
        for (int a = 0; a < F::ND; a++) {
            bad |= d[a] < 0 || d[a] >= nrow || d[a] >= ncol;
            edof[e * F::ND + a] = d[a];
        }
        if (bad) *err = 1;
    }
}

struct BBox { double x0, y0, x1, y1; };
struct BBoxOp {
    __device__ __forceinline__ BBox operator()(const BBox &a, const BBox &b) const
    {
        return BBox{fmin(a.x0, b.x0), fmin(a.y0, b.y0), fmax(a.x1, b.x1), fmax(a.y1, b.y1)};
    }
};
struct XYToBBox {
    __device__ __forceinline__ BBox operator()(const double2 &p) const { return BBox{p.x, p.y, p.x, p.y}; }
};

__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Morton key of the element centroid (first 3 nodes), 16 bits per axis
__global__ void k_tl_morton(const int32_t *__restrict__ gconn, int gk, const double2 *__restrict__ xy, int64_t nel,
                            const BBox *__restrict__ bb, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const BBox b = *bb;
    const double sx = (b.x1 > b.x0) ? 65535.0 / (b.x1 - b.x0) : 0.0, sy = (b.y1 > b.y0) ? 65535.0 / (b.y1 - b.y0) : 0.0;
    GRID_STRIDE(e, nel) {
        double cx = 0, cy = 0;
        for (int a = 0; a < 3; a++) { const double2 p = xy[gconn[e * gk + a]]; cx += p.x; cy += p.y; }
        cx *= (1.0 / 3); cy *= (1.0 / 3);
        const uint32_t qx = (uint32_t)fmin(fmax((cx - b.x0) * sx, 0.0), 65535.0);
        const uint32_t qy = (uint32_t)fmin(fmax((cy - b.y0) * sy, 0.0), 65535.0);
        keys[e] = spread16(qx) | (spread16(qy) << 1);
        vals[e] = (uint32_t)e;
    }
}
// 3-D meshes: 10 bits per axis
__device__ __forceinline__ uint32_t spread10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_tl_morton3(const int32_t *__restrict__ gconn, int gk, const double2 *__restrict__ xy, const double *__restrict__ z, int64_t nel,
                             const BBox *__restrict__ bb, const double *__restrict__ zmm, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const BBox b = *bb;
    const double z0 = zmm[0], z1 = zmm[1];
    const double sx = (b.x1 > b.x0) ? 1023.0 / (b.x1 - b.x0) : 0.0, sy = (b.y1 > b.y0) ? 1023.0 / (b.y1 - b.y0) : 0.0, sz = (z1 > z0) ? 1023.0 / (z1 - z0) : 0.0;
    GRID_STRIDE(e, nel) {
        double cx = 0, cy = 0, cz = 0;
        for (int a = 0; a < gk; a++) { const int32_t n = gconn[e * gk + a]; const double2 p = xy[n]; cx += p.x; cy += p.y; cz += z[n]; }
        cx /= gk; cy /= gk; cz /= gk;
        const uint32_t qx = (uint32_t)fmin(fmax((cx - b.x0) * sx, 0.0), 1023.0);
        const uint32_t qy = (uint32_t)fmin(fmax((cy - b.y0) * sy, 0.0), 1023.0);
        const uint32_t qz = (uint32_t)fmin(fmax((cz - z0) * sz, 0.0), 1023.0);
        keys[e] = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
        vals[e] = (uint32_t)e;
    }
}
__global__ void k_tl_etile_from_order(const uint32_t *__restrict__ eorder, int64_t nel, int te, int32_t *__restrict__ etile)
{
    GRID_STRIDE(p, nel) etile[eorder[p]] = (int32_t)(p / te);
}
__global__ void k_tl_etile_identity(int64_t nel, int te, int32_t *__restrict__ etile)
{
    GRID_STRIDE(e, nel) etile[e] = (int32_t)(e / te);
}
__global__ void k_tl_fill_i32(int32_t *p, int64_t n, int32_t v) { GRID_STRIDE(i, n) p[i] = v; }

// adjacency pair generation: (local column, pair id) for every (element, local dof) whose column this ctx owns
template <int ND>
__global__ void k_tl_pairs(const int32_t *__restrict__ edof, int64_t nel, ColMap cm, uint32_t ncl, uint32_t *__restrict__ adjcnt,
                           uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    GRID_STRIDE(t, nel * ND) {
        const int64_t lc = cm.local(edof[t]);
        if (lc >= 0) {
            atomicAdd(&adjcnt[(uint32_t)lc], 1u);
            keys[t] = (uint32_t)lc;
        } else {
            keys[t] = ncl;
        }
        vals[t] = (uint32_t)t;
    }
}
// owner tile of every owned column = the lowest tile among the elements that contain its dof
template <int ND>
__global__ void k_tl_owner(const int32_t *__restrict__ edof, const int32_t *__restrict__ etile, int64_t nel, ColMap cm, int32_t *__restrict__ owner)
{
    GRID_STRIDE(t, nel * ND) {
        const int64_t lc = cm.local(edof[t]);
        if (lc >= 0) atomicMin(&owner[lc], etile[t / ND]);
    }
}

// sorted, de-duplicated row list of one column (candidates enumerated from the adjacency);
// COUNT additionally keeps the number of contributions of every row
template <class F, bool COUNT>
__device__ __forceinline__ int tl_column_rows(const uint32_t *__restrict__ adj, uint32_t a0, uint32_t a1,
                                              const int32_t *__restrict__ edof, int32_t (&rows)[TL_CAP], uint16_t *cnt, int &ncontrib)
{
    int n = 0;
    ncontrib = 0;
    for (uint32_t p = a0; p < a1; p++) {
        const uint32_t t = adj[p];
        const uint32_t e = t / F::ND;
        const int lj = (int)(t % F::ND);
        for (int i = 0; i < F::ND; i++) {
            if (!F::mask(i, lj)) continue;
            ncontrib++;
            const int32_t r = edof[(int64_t)e * F::ND + i];
            int lo = 0, hi = n;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows[mid] < r) lo = mid + 1; else hi = mid; }
            if (lo < n && rows[lo] == r) { if (COUNT) cnt[lo]++; continue; }
            if (n == TL_CAP) return -1;
            for (int k = n; k > lo; k--) { rows[k] = rows[k - 1]; if (COUNT) cnt[k] = cnt[k - 1]; }
            rows[lo] = r;
            if (COUNT) cnt[lo] = 1;
            n++;
        }
    }
    return n;
}

#define TL_LIGHT 2   // nonzeros with at most this many contributions take the straight-line gather path

template <class F>
__global__ void k_tl_col_count(const uint32_t *__restrict__ adjptr, const uint32_t *__restrict__ adj, const int32_t *__restrict__ edof,
                               int64_t ncl, uint8_t *__restrict__ colcnt, uint16_t *__restrict__ ccnt, uint8_t *__restrict__ hcnt,
                               int *__restrict__ err, int diag_heavy /* fused load vector: every diagonal nonzero takes the heavy list */)
{
    GRID_STRIDE(cl, ncl) {
        int32_t rows[TL_CAP];
        uint16_t cnt[TL_CAP];
        int nc;
        const int n = tl_column_rows<F, true>(adj, adjptr[cl], adjptr[cl + 1], edof, rows, cnt, nc);
        if (n < 0 || nc > 65535) { *err = 2; colcnt[cl] = 0; ccnt[cl] = 0; hcnt[cl] = 0; continue; }
        int h = 0, hc = 0;
        for (int t = 0; t < n; t++) if (cnt[t] > TL_LIGHT || (diag_heavy && rows[t] == (int32_t)cl)) { h++; hc += cnt[t] + ((diag_heavy && rows[t] == (int32_t)cl) ? 2 : 0); }      // (+2: the dof, see k_tl_gather_build)
        colcnt[cl] = (uint8_t)n;
        ccnt[cl] = (uint16_t)hc;      // contributions to heavy nonzeros only
        hcnt[cl] = (uint8_t)h;
    }
}
template <class F>
__global__ void k_tl_col_fill(const uint32_t *__restrict__ adjptr, const uint32_t *__restrict__ adj, const int32_t *__restrict__ edof,
                              int64_t ncl, const int64_t *__restrict__ colptr0, int32_t *__restrict__ rowval, int64_t *__restrict__ colptr1)
{
    GRID_STRIDE(cl, ncl + 1) {
        colptr1[cl] = colptr0[cl] + 1;
        if (cl == ncl) continue;
        int32_t rows[TL_CAP];
        int nc;
        const int n = tl_column_rows<F, false>(adj, adjptr[cl], adjptr[cl + 1], edof, rows, nullptr, nc);
        const int64_t o = colptr0[cl];
        for (int k = 0; k < n; k++) rowval[o + k] = rows[k];
    }
}

__global__ void k_tl_tilecol_keys(const int32_t *__restrict__ owner, int64_t ncl, int ntiles, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    GRID_STRIDE(cl, ncl) {
        const int32_t o = owner[cl];
        keys[cl] = (o == INT_MAX) ? (uint32_t)ntiles : (uint32_t)o;
        vals[cl] = (uint32_t)cl;
    }
}
// lower_bound of every tile id in a sorted key array
template <class K> __global__ void k_tl_lower_bounds(const K *__restrict__ keys, int64_t n, int ntiles, int shift, int64_t *__restrict__ ptr)
{
    GRID_STRIDE(T, (int64_t)ntiles + 1) {
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)(keys[mid] >> shift) < T) lo = mid + 1; else hi = mid; }
        ptr[T] = lo;
    }
}
struct GatherU8 {
    const uint8_t *v; const uint32_t *idx; int64_t n;
    __device__ __forceinline__ int64_t operator()(int64_t k) const { return k < n ? (int64_t)v[idx[k]] : 0; }
};
struct GatherU16 {
    const uint16_t *v; const uint32_t *idx; int64_t n;
    __device__ __forceinline__ int64_t operator()(int64_t k) const { return k < n ? (int64_t)v[idx[k]] : 0; }
};
// run heads among the owned tile columns
__global__ void k_tl_run_flags(const uint32_t *__restrict__ tkeys, const uint32_t *__restrict__ tcols, int64_t nowned, int32_t *__restrict__ flag)
{
    GRID_STRIDE(k, nowned) flag[k] = (k == 0 || tkeys[k] != tkeys[k - 1] || tcols[k] != tcols[k - 1] + 1) ? 1 : 0;
}
__global__ void k_tl_run_fill(const int32_t *__restrict__ flag, const int64_t *__restrict__ runidx, const uint32_t *__restrict__ tkeys,
                              const uint32_t *__restrict__ tcols, int64_t nowned, const int64_t *__restrict__ tcol_slot,
                              const int64_t *__restrict__ tcol_ptr, const int64_t *__restrict__ colptr0, TileRun *__restrict__ runs,
                              int64_t *__restrict__ run_firstk)
{
    GRID_STRIDE(k, nowned) if (flag[k]) {
        const int64_t r = runidx[k];
        const uint32_t T = tkeys[k];
        runs[r].nz0 = colptr0[tcols[k]];
        runs[r].s0 = (int32_t)(tcol_slot[k] - tcol_slot[tcol_ptr[T]]);
        run_firstk[r] = k;
    }
}
__global__ void k_tl_run_len(const int64_t *__restrict__ run_firstk, int64_t nruns, int64_t nowned, const int64_t *__restrict__ tcol_slot, TileRun *__restrict__ runs)
{
    GRID_STRIDE(r, nruns) {
        const int64_t k0 = run_firstk[r], k1 = (r + 1 < nruns) ? run_firstk[r + 1] : nowned;
        runs[r].len = (int32_t)(tcol_slot[k1] - tcol_slot[k0]);
    }
}

// padded slot layout: every run starts at a multiple of 32 slots
__global__ void k_tl_run_padlen(const TileRun *__restrict__ runs, int64_t nruns, int64_t *__restrict__ padlen)
{
    GRID_STRIDE(r, nruns + 1) padlen[r] = r < nruns ? (((int64_t)runs[r].len + 31) & ~(int64_t)31) : 0;
}
// pslot[k] = padded position of the first slot of owned tile column k (pslot[nowned] = total)
__global__ void k_tl_pslot(const int32_t *__restrict__ flag, const int64_t *__restrict__ runidx, const int64_t *__restrict__ run_firstk,
                           const int64_t *__restrict__ tcol_slot, const int64_t *__restrict__ run_pad0, int64_t nowned, int64_t nruns,
                           int64_t *__restrict__ pslot)
{
    GRID_STRIDE(k, nowned + 1) {
        if (k == nowned) { pslot[k] = run_pad0[nruns]; continue; }
        const int64_t r = flag[k] ? runidx[k] : runidx[k] - 1;
        pslot[k] = run_pad0[r] + (tcol_slot[k] - tcol_slot[run_firstk[r]]);
    }
}
__global__ void k_tl_run_s0(const int64_t *__restrict__ run_firstk, int64_t nruns, const uint32_t *__restrict__ tkeys, const int64_t *__restrict__ tcol_ptr,
                            const int64_t *__restrict__ pslot, TileRun *__restrict__ runs)
{
    GRID_STRIDE(r, nruns) {
        const int64_t k = run_firstk[r];
        runs[r].s0 = (int32_t)(pslot[k] - pslot[tcol_ptr[tkeys[k]]]);
    }
}

// tile-element keys: (T << 36) | (e << 4) | lj for every (e, lj) whose column is owned
template <int ND>
__global__ void k_tl_telem_keys(const int32_t *__restrict__ edof, int64_t nel, ColMap cm, const int32_t *__restrict__ owner,
                                int ntiles, uint64_t *__restrict__ keys)
{
    GRID_STRIDE(t, nel * ND) {
        const int64_t e = t / ND;
        const int lj = (int)(t % ND);
        const int64_t lc = cm.local(edof[t]);
        const uint64_t T = lc >= 0 ? (uint64_t)(uint32_t)owner[lc] : (uint64_t)ntiles;
        keys[t] = (T << 36) | ((uint64_t)e << 4) | (uint64_t)lj;
    }
}
__global__ void k_tl_head_flags64(const uint64_t *__restrict__ keys, int64_t n, int shift, int32_t *__restrict__ flag)
{
    GRID_STRIDE(p, n) flag[p] = (p == 0 || (keys[p] >> shift) != (keys[p - 1] >> shift)) ? 1 : 0;
}
// unique (T,e) entries with their owned-column masks
__global__ void k_tl_telem_fill(const uint64_t *__restrict__ keys, int64_t n, const int32_t *__restrict__ flag, const int64_t *__restrict__ gidx_of,
                                uint64_t *__restrict__ telem_key, uint16_t *__restrict__ mask, uint32_t *__restrict__ key2, uint32_t *__restrict__ val2,
                                uint32_t *__restrict__ pc_hist /* ntiles x 17 */)
{
    GRID_STRIDE(p, n) if (flag[p]) {
        const uint64_t ke = keys[p] >> 4;
        uint32_t m = 0;
        for (int64_t q = p; q < n && (keys[q] >> 4) == ke; q++) m |= 1u << (uint32_t)(keys[q] & 15u);
        const int64_t g = gidx_of[p];
        telem_key[g] = ke;
        mask[g] = (uint16_t)m;
        const uint32_t T = (uint32_t)(ke >> 32);
        const int pc = __popc(m);
        key2[g] = (T << 5) | (uint32_t)(16 - pc);   // popcount descending inside a tile
        val2[g] = (uint32_t)g;
        atomicAdd(&pc_hist[(int64_t)T * 17 + pc], 1u);
    }
}
__global__ void k_tl_invert(const uint32_t *__restrict__ order, int64_t n, uint32_t *__restrict__ newpos)
{
    GRID_STRIDE(p, n) newpos[order[p]] = (uint32_t)p;
}
template <int GK>
__global__ void k_tl_tconn(const uint32_t *__restrict__ order, int64_t n, const uint64_t *__restrict__ telem_key, const uint16_t *__restrict__ mask,
                           const int32_t *__restrict__ gconn, int32_t *__restrict__ tconn, uint16_t *__restrict__ tmask)
{
    GRID_STRIDE(p, n) {
        const uint32_t g = order[p];
        const int64_t e = (int64_t)(telem_key[g] & 0xffffffffu);
#pragma unroll
        for (int a = 0; a < GK; a++) tconn[p * GK + a] = gconn[e * GK + a];
        tmask[p] = mask[g];
    }
}

// TL_GEO: tile-local node numbering.  One CTA per tile: block-wide radix sort of the tile's (node id, position) pairs,
// head flags -> local ids.  tlocal[elem0*GK + pos] = local id of that connectivity entry, tnodes[elem0*GK + k] = k-th
// unique node, tnn[T] = number of unique nodes.
template <int GK>
__global__ void __launch_bounds__(256) k_tl_geo_local(const int64_t *__restrict__ telem_ptr, const int32_t *__restrict__ tconn,
                                                      uint16_t *__restrict__ tlocal, int32_t *__restrict__ tnodes, int32_t *__restrict__ tnn,
                                                      int *__restrict__ err)
{
    constexpr int IPT = TL_GEO_CAP / 256;
    using Sort = cub::BlockRadixSort<uint32_t, 256, IPT, uint16_t>;
    using Scan = cub::BlockScan<int, 256>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ uint32_t last_key[256];
    const int T = blockIdx.x, tid = threadIdx.x;
    const int64_t e0 = telem_ptr[T];
    const int n = (int)(telem_ptr[T + 1] - e0) * GK;
    if (n > TL_GEO_CAP) { if (tid == 0) { *err = 5; tnn[T] = 0; } return; }
    uint32_t key[IPT];
    uint16_t pos[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = tid * IPT + k;
        key[k] = i < n ? (uint32_t)tconn[e0 * GK + i] : 0xFFFFFFFFu;
        pos[k] = (uint16_t)i;
    }
    Sort(tmp.sort).Sort(key, pos);
    __syncthreads();
    last_key[tid] = key[IPT - 1];
    __syncthreads();
    uint32_t prev = tid > 0 ? last_key[tid - 1] : 0xFFFFFFFFu;   // 0xFFFFFFFF is never a node id: the first key is a head
    int heads = 0;
    bool head[IPT];
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        head[k] = key[k] != 0xFFFFFFFFu && (key[k] != prev || (tid == 0 && k == 0));
        heads += head[k] ? 1 : 0;
        prev = key[k];
    }
    int before = 0, total = 0;
    Scan(tmp.scan).ExclusiveSum(heads, before, total);
    int id = before - 1;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        if (key[k] == 0xFFFFFFFFu) continue;
        if (head[k]) { id++; tnodes[e0 * GK + id] = (int32_t)key[k]; }
        tlocal[e0 * GK + pos[k]] = (uint16_t)id;
    }
    if (tid == 0) tnn[T] = total;
}
// fills the geometry block of every tile: coordinates of its unique nodes, 16-bit local connectivity, column masks
template <int GK>
__global__ void k_tl_geo_fill(int ntiles, const TileDescFull *__restrict__ tiles, const uint16_t *__restrict__ tlocal,
                              const int32_t *__restrict__ tnodes, const uint16_t *__restrict__ tmask, const double2 *__restrict__ xy,
                              unsigned char *__restrict__ geo, const double *__restrict__ z = nullptr)
{
    for (int T = blockIdx.x; T < ntiles; T += gridDim.x) {
        const TileDescFull &td = tiles[T];
        if (td.geo_bytes == 0) continue;
        unsigned char *gb = geo + td.geo0;
        if (z) {
            double *tz = reinterpret_cast<double *>(gb + tl_geo_z_off(GK, td.nelem, td.nnode));
            for (int k = threadIdx.x; k < td.nnode; k += blockDim.x) tz[k] = z[tnodes[td.elem0 * GK + k]];
        }
        double2 *txy = reinterpret_cast<double2 *>(gb);
        uint16_t *c16 = reinterpret_cast<uint16_t *>(gb + tl_geo_xy_bytes(td.nnode));
        uint16_t *m16 = reinterpret_cast<uint16_t *>(gb + tl_geo_xy_bytes(td.nnode) + tl_geo_conn_bytes(GK, td.nelem));
        for (int k = threadIdx.x; k < td.nnode; k += blockDim.x) txy[k] = xy[tnodes[td.elem0 * GK + k]];
        for (int k = threadIdx.x; k < td.nelem * GK; k += blockDim.x) c16[k] = tlocal[td.elem0 * GK + k];
        for (int k = threadIdx.x; k < td.nelem; k += blockDim.x) m16[k] = tmask[td.elem0 + k];
    }
}
__global__ void k_tl_tiles_geo0(int ntiles, const int64_t *__restrict__ geo_off, TileDescFull *__restrict__ tiles)
{
    GRID_STRIDE(T, ntiles) tiles[T].geo0 = geo_off[T];
}

// per owned tile column: start offset of every slot's contributions (goff) + the stage index of
// every contribution in append order (gidx), written into the owning tile's metadata block
template <class F>
__global__ void k_tl_gather_build(const uint32_t *__restrict__ tkeys, const uint32_t *__restrict__ tcols, int64_t nowned,
                                  const uint32_t *__restrict__ adjptr, const uint32_t *__restrict__ adj, const int32_t *__restrict__ edof,
                                  const int64_t *__restrict__ colptr0, const int32_t *__restrict__ rowval,
                                  const uint64_t *__restrict__ telem_key, const int64_t *__restrict__ telem_ptr,
                                  const uint16_t *__restrict__ mask, const uint32_t *__restrict__ newpos,
                                  const TileDescFull *__restrict__ tiles, const int64_t *__restrict__ tcol_slot, const int64_t *__restrict__ tcol_gidx,
                                  const int64_t *__restrict__ tcol_heavy, unsigned char *__restrict__ meta, int *__restrict__ err, int diag_heavy)
{
    GRID_STRIDE(k, nowned) {
        const uint32_t T = tkeys[k], cl = tcols[k];
        const int64_t r0 = colptr0[cl];
        const int nr = (int)(colptr0[cl + 1] - r0);
        const TileDescFull &td = tiles[T];
        const int64_t g0 = telem_ptr[T], g1 = telem_ptr[T + 1];
        unsigned char *__restrict__ mb = meta + td.meta0;
        const uint32_t sloc = (uint32_t)(tcol_slot[k] - td.slot0);
        uint32_t *__restrict__ pk = reinterpret_cast<uint32_t *>(mb) + sloc;
        uint16_t *__restrict__ hidx = reinterpret_cast<uint16_t *>(mb + tl_meta_goff_bytes(td.nslot));
        TileHeavy *__restrict__ heavy = reinterpret_cast<TileHeavy *>(mb + tl_meta_goff_bytes(td.nslot) + tl_meta_gidx_bytes(td.ncontrib) +
                                                                     tl_meta_rows_bytes(td.nslot)) + (tcol_heavy[k] - td.heavy0);
        uint16_t cnt[TL_CAP], off[TL_CAP], first[TL_CAP], second[TL_CAP];
        for (int t = 0; t < nr; t++) cnt[t] = 0;
        const uint32_t a0 = adjptr[cl], a1 = adjptr[cl + 1];
        // pass A: contributions per slot
        for (uint32_t p = a0; p < a1; p++) {
            const uint32_t t = adj[p];
            const uint32_t e = t / F::ND;
            const int lj = (int)(t % F::ND);
            for (int i = 0; i < F::ND; i++) {
                if (!F::mask(i, lj)) continue;
                const int32_t r = edof[(int64_t)e * F::ND + i];
                int lo = 0, hi = nr;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (rowval[r0 + mid] < r) lo = mid + 1; else hi = mid; }
                cnt[lo]++;
            }
        }
        uint32_t run = (uint32_t)(tcol_gidx[k] - td.gidx0);      // heavy contributions of the tile before this column
        int nh = 0;
        for (int t = 0; t < nr; t++) {
            first[t] = second[t] = TL_PK_NONE;
            off[t] = 0;
            if (diag_heavy && rowval[r0 + t] == (int32_t)cl) cnt[t] |= 0x8000u;      // forced heavy (the count itself stays below 2^15)
            if (cnt[t] > TL_LIGHT) {
                if (run + (cnt[t] & 0x7FFFu) + 2u > 65535u) *err = 3;
                heavy[nh++] = TileHeavy{(uint16_t)(sloc + t), (uint16_t)run, (uint16_t)(cnt[t] & 0x7FFFu), 0};
                off[t] = (uint16_t)run;
                run += (cnt[t] & 0x7FFFu) + ((cnt[t] & 0x8000u) ? 2u : 0u);      // a diagonal entry's list ends with its dof (two 16-bit halves)
            }
        }
        // pass B: stage indices, append order (ascending element, column-major inside an element)
        for (uint32_t p = a0; p < a1; p++) {
            const uint32_t t = adj[p];
            const uint32_t e = t / F::ND;
            const int lj = (int)(t % F::ND);
            const uint64_t want = ((uint64_t)T << 32) | e;
            int64_t lo = g0, hi = g1;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (telem_key[mid] < want) lo = mid + 1; else hi = mid; }
            const uint32_t m = mask[lo];
            const int rnk = __popc(m & ((1u << lj) - 1u));
            const uint32_t le = newpos[lo] - (uint32_t)td.elem0;
            const uint32_t q = td.qbase[rnk] + le;
            for (int i = 0; i < F::ND; i++) {
                if (!F::mask(i, lj)) continue;
                const int32_t r = edof[(int64_t)e * F::ND + i];
                int l2 = 0, h2 = nr;
                while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (rowval[r0 + mid] < r) l2 = mid + 1; else h2 = mid; }
                const uint32_t sidx = tl_sym<F>() ? (uint32_t)tl_tri(i < lj ? i : lj, i < lj ? lj : i, F::ND) * (uint32_t)td.nqs + le
                                                  : (uint32_t)i * (uint32_t)td.nqs + q;
                if (sidx >= 0xFFFEu) *err = 4;
                if (cnt[l2] > TL_LIGHT) { hidx[off[l2]] = (uint16_t)sidx; off[l2]++; }
                else if (off[l2]++ == 0) first[l2] = (uint16_t)sidx;
                else second[l2] = (uint16_t)sidx;
            }
        }
        for (int t = 0; t < nr; t++) {
            pk[t] = cnt[t] > TL_LIGHT ? TL_PK_HEAVY : ((uint32_t)first[t] | ((uint32_t)second[t] << 16));
            if (cnt[t] & 0x8000u) {      // fused load vector: the dof (= row index) of the diagonal nonzero rides behind its contributions
                const uint32_t d = (uint32_t)rowval[r0 + t];
                hidx[off[t]] = (uint16_t)(d & 0xFFFFu); hidx[off[t] + 1] = (uint16_t)(d >> 16);
            }
        }
    }
}

// row table + padding slots of every run into its tile's metadata block
__global__ void k_tl_meta_rows(int64_t nruns, const TileRun *__restrict__ runs, const int64_t *__restrict__ run_firstk, const uint32_t *__restrict__ tkeys,
                               const TileDescFull *__restrict__ tiles, unsigned char *__restrict__ meta)
{
    GRID_STRIDE(r, nruns) {
        const TileDescFull &td = tiles[tkeys[run_firstk[r]]];
        if (td.meta_bytes == 0) continue;
        unsigned char *mb = meta + td.meta0;
        uint32_t *pk = reinterpret_cast<uint32_t *>(mb);
        int64_t *rowbase = reinterpret_cast<int64_t *>(mb + tl_meta_goff_bytes(td.nslot) + tl_meta_gidx_bytes(td.ncontrib));
        const TileRun run = runs[r];
        const int padlen = (run.len + 31) & ~31;
        for (int j = 0; j < (padlen >> 5); j++) rowbase[(run.s0 >> 5) + j] = run.nz0 + 32 * (int64_t)j;
        for (int t = run.len; t < padlen; t++) pk[run.s0 + t] = 0xFFFFFFFFu;
    }
}
__global__ void k_tl_tiles_meta0(int ntiles, const int64_t *__restrict__ meta_off, TileDescFull *__restrict__ tiles)
{
    GRID_STRIDE(T, ntiles) tiles[T].meta0 = meta_off[T];
}

__global__ void k_tl_tiles_fill(int ntiles, int nd /* stage rows */, bool sym, int gsz, bool gs_alias, int gk, int dim3, const int32_t *__restrict__ tnn /* null: no geometry blocks */,
                                int64_t *__restrict__ geo_bytes, const int64_t *__restrict__ tcol_ptr, const int64_t *__restrict__ tcol_slot,
                                const int64_t *__restrict__ tcol_gidx, const int64_t *__restrict__ tcol_heavy, const int64_t *__restrict__ telem_ptr,
                                const int64_t *__restrict__ run_of_k /* exclusive scan of run-head flags */,
                                int64_t nowned, int64_t nruns, const uint32_t *__restrict__ pc_hist, TileDescFull *__restrict__ tiles,
                                int64_t *__restrict__ meta_bytes, int32_t *__restrict__ maxima /* stage+meta bytes, nq*nd, ncontrib, nelem */)
{
    GRID_STRIDE(T, ntiles) {
        TileDescFull d;
        const int64_t k0 = tcol_ptr[T], k1 = tcol_ptr[T + 1];
        d.slot0 = tcol_slot[k0];
        d.nslot = (int32_t)(tcol_slot[k1] - tcol_slot[k0]);
        d.gidx0 = tcol_gidx[k0];
        const int64_t nc = tcol_gidx[k1] - tcol_gidx[k0];
        d.ncontrib = (int32_t)(nc > 0x7fffffff ? 0x7fffffff : nc);
        d.heavy0 = tcol_heavy[k0];
        d.nheavy = (int32_t)(tcol_heavy[k1] - tcol_heavy[k0]);
        d.elem0 = telem_ptr[T];
        d.nelem = (int32_t)(telem_ptr[T + 1] - telem_ptr[T]);
        const int64_t r0 = (k0 < nowned) ? run_of_k[k0] : nruns, r1 = (k1 < nowned) ? run_of_k[k1] : nruns;
        d.run0 = (int32_t)r0;
        d.nrun = (int32_t)(r1 - r0);
        // qbase[r] = sum_{r' < r} (#elements with popcount > r')
        uint32_t above[TL_MAXND + 2];
        uint32_t acc = 0;
        for (int pc = 16; pc >= 0; pc--) { above[pc] = acc; acc += pc_hist[(int64_t)T * 17 + pc]; } // above[pc] = #elements with popcount > pc
        uint32_t q = 0;
        for (int r = 0; r <= TL_MAXND; r++) { d.qbase[r] = (uint16_t)(q > 65535u ? 65535u : q); if (r < TL_MAXND) q += above[r]; }
        d.nq = (int32_t)q;
        if (sym) q = (uint32_t)d.nelem;          // symmetric stage: one column per tile element
        d.nqs = (uint16_t)((q | 1u) > 65535u ? 65535u : (q | 1u));
        d.meta_bytes = d.nslot > 0 ? tl_meta_goff_bytes(d.nslot) + tl_meta_gidx_bytes(d.ncontrib) + tl_meta_rows_bytes(d.nslot) + tl_align16((int)sizeof(TileHeavy) * d.nheavy) : 0;
        d.meta0 = 0;
        d.pad2_[0] = d.pad2_[1] = 0;
        d.geo0 = 0;
        d.nnode = tnn ? tnn[T] : 0;
        d.geo_bytes = (tnn && d.nslot > 0) ? tl_geo_block_bytes(gk, d.nelem, d.nnode, dim3) : 0;
        tiles[T] = d;
        meta_bytes[T] = d.meta_bytes;
        geo_bytes[T] = d.geo_bytes;
        atomicMax(&maxima[0], tl_stage_bytes(nd, d.nqs) + (gs_alias ? max(tl_geo_bytes(gsz, d.nelem), d.meta_bytes) : tl_geo_bytes(gsz, d.nelem) + d.meta_bytes) + d.geo_bytes);
        atomicMax(&maxima[1], (int32_t)(((int64_t)q | 1) * nd > 0x7fffffff ? 0x7fffffff : ((int64_t)q | 1) * nd));
        if (d.nslot > 65535) atomicMax(&maxima[1], 0x7fffffff);      // TileHeavy addresses slots with 16 bits
        atomicMax(&maxima[2], d.ncontrib); atomicMax(&maxima[3], d.nelem);
        atomicMax(&maxima[4], tl_stage_bytes(nd, d.nqs)); atomicMax(&maxima[5], d.meta_bytes); atomicMax(&maxima[6], d.geo_bytes);
        atomicMax(&maxima[7], tl_geo_bytes(gsz, d.nelem));
    }
}

// ---- numeric kernel ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tl_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class F, bool WF = false> struct StageEmit {
    static constexpr bool TRI = tl_sym<F>();
    static constexpr bool WITH_F = WF;      // the worker also stages the element load vector (rows srows .. srows+ND-1)
    __device__ __forceinline__ void fvec(const double (&f)[F::ND], uint32_t mm) {
#pragma unroll
        for (int j = 0; j < F::ND; j++)
            if (mm & (1u << j)) stage[(tl_srows<F>() + j) * nq + le] = f[j];
    }
    // upper triangle of a bitwise-symmetric element matrix: entry (a, b), a <= b, is staged if column a or column b is owned
    __device__ __forceinline__ void tri(const double (&K)[F::ND][F::ND], uint32_t mm) {
#pragma unroll
        for (int b = 0; b < F::ND; b++)
#pragma unroll
            for (int a = 0; a <= b; a++)
                if (mm & ((1u << a) | (1u << b))) stage[tl_tri(a, b, F::ND) * nq + le] = K[a][b];
    }
    double *__restrict__ stage;
    const uint16_t *__restrict__ qbase;
    uint32_t m, le;
    int nq;
    template <int J> __device__ __forceinline__ void col(const double (&out)[F::ND]) {
        const uint32_t q = qbase[__popc(m & ((1u << J) - 1u))] + le;
#pragma unroll
        for (int i = 0; i < F::ND; i++)
            if (F::mask(i, J)) stage[i * nq + q] = out[i];
    }
};

__device__ __forceinline__ uint32_t tl_mbar_wait(uint32_t barA, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(barA), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void tl_bulk_load(uint32_t dstA, const void *src, uint32_t bytes, uint32_t barA)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dstA), "l"(src), "r"(bytes), "r"(barA) : "memory");
}

// One CTA per tile.  Per tile: (0) a TMA bulk copy brings the tile's gather metadata into shared memory
// while (1) one thread per tile element computes the owned columns of its element matrix into the stage,
// (2) one thread per owned nonzero sums its contributions from the stage left to right and stores nzval.
// Overlap between the compute phase of one tile and the gather/store phase of another comes from the
// CTAs co-resident on an SM.
// Phase-1 worker of the one-thread-per-element forms, kept out of line so that its register allocation does not
// depend on the code of the gather phase (the kernel-wide allocation otherwise shifts spills into this hot code).
template <class F, bool S>
__device__ __noinline__ void tl_element_to_stage(int64_t g, uint32_t le, const int32_t *__restrict__ tconn, const uint16_t *__restrict__ tmask,
                                                 const double2 *__restrict__ xy, double *__restrict__ stage, const uint16_t *__restrict__ qbase, int nq)
{
    constexpr int GK = F::GK;
    double X[GK], Y[GK];
    int32_t nd[GK];
#if TL_VEC_CONN
    if constexpr (GK == 6) {          // 24 bytes per element: three aligned 8-byte loads
        const int2 *c2 = reinterpret_cast<const int2 *>(tconn + g * 6);
        const int2 a = __ldg(c2), b = __ldg(c2 + 1), c = __ldg(c2 + 2);
        nd[0] = a.x; nd[1] = a.y; nd[2] = b.x; nd[3] = b.y; nd[4] = c.x; nd[5] = c.y;
    } else if constexpr (GK == 4) {   // 16 bytes per element: one aligned 16-byte load
        const int4 a = __ldg(reinterpret_cast<const int4 *>(tconn + g * 4));
        nd[0] = a.x; nd[1] = a.y; nd[2] = a.z; nd[3] = a.w;
    } else
#endif
    {
#pragma unroll
        for (int a = 0; a < GK; a++) nd[a] = tconn[g * GK + a];
    }
#pragma unroll
    for (int a = 0; a < GK; a++) { const double2 p = __ldg(&xy[nd[a]]); X[a] = p.x; Y[a] = p.y; }
    StageEmit<F> emit{stage, qbase, (uint32_t)tmask[g], le, nq};
    F::template element<S>(X, Y, emit.m, emit);
}

// Same worker fed from the tile's geometry block in shared memory (TL_GEO)
template <class F, bool S, bool WF = false>
__device__ __noinline__ void tl_element_to_stage_local(uint32_t le, const uint16_t *__restrict__ c16, const uint16_t *__restrict__ m16,
                                                       const double2 *__restrict__ sxy, double *__restrict__ stage, const uint16_t *__restrict__ qbase, int nq)
{
    constexpr int GK = F::GK;
    uint32_t nd[GK];
    if constexpr (GK == 6) {          // 12 bytes per element: three aligned 4-byte loads
        const uint32_t *c = reinterpret_cast<const uint32_t *>(c16 + le * 6);
        const uint32_t a = c[0], b = c[1], d = c[2];
        nd[0] = a & 0xFFFFu; nd[1] = a >> 16; nd[2] = b & 0xFFFFu; nd[3] = b >> 16; nd[4] = d & 0xFFFFu; nd[5] = d >> 16;
    } else if constexpr (GK == 4) {   // 8 bytes per element: one aligned 8-byte load
        const uint2 a = *reinterpret_cast<const uint2 *>(c16 + le * 4);
        nd[0] = a.x & 0xFFFFu; nd[1] = a.x >> 16; nd[2] = a.y & 0xFFFFu; nd[3] = a.y >> 16;
    } else {
#pragma unroll
        for (int a = 0; a < GK; a++) nd[a] = c16[le * GK + a];
    }
    double X[GK], Y[GK];
#pragma unroll
    for (int a = 0; a < GK; a++) { const double2 p = sxy[nd[a]]; X[a] = p.x; Y[a] = p.y; }
    StageEmit<F, WF> emit{stage, qbase, (uint32_t)m16[le], le, nq};
    F::template element<S>(X, Y, emit.m, emit);
}

// the same for a 3-D form (FEH1_T4): third coordinate plane behind the mask array of the geometry block
template <class F, bool S>
__device__ __noinline__ void tl_element_to_stage_local3(uint32_t le, const uint16_t *__restrict__ c16, const uint16_t *__restrict__ m16,
                                                        const double2 *__restrict__ sxy, const double *__restrict__ sz, double *__restrict__ stage,
                                                        const uint16_t *__restrict__ qbase, int nq)
{
    constexpr int GK = F::GK;
    static_assert(GK == 4, "");
    const uint2 a = *reinterpret_cast<const uint2 *>(c16 + le * 4);
    const uint32_t nd[4] = {a.x & 0xFFFFu, a.x >> 16, a.y & 0xFFFFu, a.y >> 16};
    double X[GK], Y[GK], Z[GK];
#pragma unroll
    for (int k = 0; k < GK; k++) { const double2 p = sxy[nd[k]]; X[k] = p.x; Y[k] = p.y; Z[k] = sz[nd[k]]; }
    StageEmit<F> emit{stage, qbase, (uint32_t)m16[le], le, nq};
    F::template element3<S>(X, Y, Z, emit.m, emit);
}

// SPLIT forms, phase 1a / 1b workers (inlined: out of line measured 4% slower here)
template <class F, bool S>
__device__ __forceinline__ void tl_geometry_to_smem(int64_t g, int le, int ne, const int32_t *__restrict__ tconn, const uint16_t *__restrict__ tmask,
                                                 const double2 *__restrict__ xy, double *__restrict__ Gs, uint16_t *__restrict__ smask)
{
    constexpr int GK = F::GK, NQ = F::NQ, BK = F::BK;
    double X[GK], Y[GK];
#pragma unroll
    for (int a = 0; a < GK; a++) { const double2 p = __ldg(&xy[tconn[g * GK + a]]); X[a] = p.x; Y[a] = p.y; }
    Geo<BK, NQ> G;
    geo_compute<S, GK, BK, NQ>(X, Y, G);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
#pragma unroll
        for (int n = 0; n < BK; n++) {
            Gs[(q * BK + n) * ne + le] = G.gx[q][n];
            Gs[(NQ * BK + q * BK + n) * ne + le] = G.gy[q][n];
        }
        Gs[(2 * NQ * BK + q) * ne + le] = G.JxW[q];
    }
    smask[le] = tmask[g];
}
// same, fed from the tile's geometry block in shared memory; one quadrature point at a time (small register footprint)
template <class F, bool S>
__device__ __forceinline__ void tl_geometry_to_smem_local(int le, int ne, const uint16_t *__restrict__ c16, const double2 *__restrict__ sxy,
                                                          double *__restrict__ Gs)
{
    constexpr int GK = F::GK, NQ = F::NQ, BK = F::BK;
    double X[GK], Y[GK];
#pragma unroll
    for (int a = 0; a < GK; a++) { const double2 p = sxy[c16[le * GK + a]]; X[a] = p.x; Y[a] = p.y; }
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        double gx[BK], gy[BK], JxW;
        geo_qp<S, GK, BK>(X, Y, q, gx, gy, JxW);
#pragma unroll
        for (int n = 0; n < BK; n++) {
            Gs[(q * BK + n) * ne + le] = gx[n];
            Gs[(NQ * BK + q * BK + n) * ne + le] = gy[n];
        }
        Gs[(2 * NQ * BK + q) * ne + le] = JxW;
    }
}
// same, ONE quadrature point of one element per thread (TL_GEO_QSPLIT): a tile of a vector form has fewer elements than the CTA
// has threads (32 owned + halo = 55-95 against 256 / 384), so phase 1a by element leaves three threads out of four idle
// while the others run NQ Jacobians in a row.  Same arithmetic (geo_qp, bit for bit), items q-major so that all but two warps
// read the reference-element tables at one q.  Measured slower than the per-element loop (see TL_GEO_QSPLIT): kept as the record.
template <class F, bool S>
__device__ __forceinline__ void tl_geometry_qp_to_smem_local(int le, int q, int ne, const uint16_t *__restrict__ c16, const double2 *__restrict__ sxy,
                                                             double *__restrict__ Gs)
{
    constexpr int GK = F::GK, NQ = F::NQ, BK = F::BK;
    double X[GK], Y[GK];
#pragma unroll
    for (int a = 0; a < GK; a++) { const double2 p = sxy[c16[le * GK + a]]; X[a] = p.x; Y[a] = p.y; }
    double gx[BK], gy[BK], JxW;
    geo_qp<S, GK, BK>(X, Y, q, gx, gy, JxW);
#pragma unroll
    for (int n = 0; n < BK; n++) {
        Gs[(q * BK + n) * ne + le] = gx[n];
        Gs[(NQ * BK + q * BK + n) * ne + le] = gy[n];
    }
    Gs[(2 * NQ * BK + q) * ne + le] = JxW;
}
template <class F, bool S>
__device__ __forceinline__ void tl_column_to_stage(int qc, int le, int J, int ne, int nq, const double *__restrict__ Gs, double *__restrict__ stage)
{
    const double *ge = Gs + le;       // geometry record of the element: value k at ge[k * ne]
    F::template column_single_rt<S>([&](int k) { return ge[k * ne]; }, J, [&](int i, double v) { stage[i * nq + qc] = v; });
}

// both columns of one node (local dofs J0, J0+1) of tile element le -> staged columns qc0, qc1
template <class F, bool S>
__device__ __forceinline__ void tl_column_pair_to_stage(int qc0, int qc1, int le, int J0, int ne, int nq, const double *__restrict__ Gs, double *__restrict__ stage)
{
    const double *ge = Gs + le;
    F::template column_pair_rt<S>([&](int k) { return ge[k * ne]; }, F::colnode(J0),
                                  [&](int i, double v0, double v1) { stage[i * nq + qc0] = v0; stage[i * nq + qc1] = v1; });
}
// Phase 1b of the SPLIT forms: one thread per PAIR of staged columns of a tile element (the two dofs of a node; the row
// gradients are read once for both).  Elements are ordered by their number of owned columns (descending), qbase[r] = first
// staged column of "r-th owned column of an element", so pair rank p of element le covers staged columns qbase[2p] + le and
// qbase[2p+1] + le.  Columns that do not pair up (a pressure dof, a node with one owned dof) take the single-column worker.
// Forms whose trailing columns never pair up (the pressure dofs of the Stokes forms: F::PAIR_COLS leading columns come in node pairs)
// get those columns from a second item space, (tail column k, tile element), walked after the pair items and from the LAST thread
// down: a warp of pair items used to run the pair sweep AND the single-column code whenever one of its elements had a pressure dof
// at that rank (partly owned halo elements) -- 9% of the Stokes gen kernel (profiles/r2_ab_cta_shapes_and_layouts.txt).
template <class F, class = void> struct form_pair_cols { static constexpr int value = F::ND; };
template <class F> struct form_pair_cols<F, std::void_t<decltype(F::PAIR_COLS)>> { static constexpr int value = TL_TAIL_PASS ? F::PAIR_COLS : F::ND; };
template <class F, bool S, int BLOCK>
__device__ __forceinline__ void tl_phase1b_pairs(const TileDescFull &td, const uint16_t *__restrict__ smask, int ne, int nq, const double *__restrict__ Gs,
                                                 double *__restrict__ stage, int tid)
{
    constexpr int PC = form_pair_cols<F>::value;      // columns below PC pair up by node; ranks >= PC hold tail columns only
    constexpr int NP = (PC + 1) / 2;
    int total = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) total += (int)td.qbase[2 * p + 1] - (int)td.qbase[2 * p];
    for (int t = tid; t < total; t += BLOCK) {
        int p = 0, le = t;
#pragma unroll
        for (int pp = 0; pp < NP - 1; pp++) {
            const int c = (int)td.qbase[2 * p + 1] - (int)td.qbase[2 * p];
            if (le >= c) { le -= c; p++; }
        }
        uint32_t mm = smask[le];
        for (int k = 0; k < 2 * p; k++) mm &= mm - 1;
        const int J0 = __ffs(mm) - 1;
        if (PC < F::ND && J0 >= PC) continue;      // a tail column: second item space
        mm &= mm - 1;
        int J1 = mm ? __ffs(mm) - 1 : -1;
        if (PC < F::ND && J1 >= PC) J1 = -1;
        const int qc0 = (int)td.qbase[2 * p] + le, qc1 = (int)td.qbase[2 * p + 1] + le;
        if (J1 == J0 + 1 && F::pairable(J0)) {
            tl_column_pair_to_stage<F, S>(qc0, qc1, le, J0, ne, nq, Gs, stage);
        } else {
            tl_column_to_stage<F, S>(qc0, le, J0, ne, nq, Gs, stage);
            if (J1 >= 0) tl_column_to_stage<F, S>(qc1, le, J1, ne, nq, Gs, stage);
        }
    }
    if constexpr (PC < F::ND) {
        // Measured alternatives (Stokes gen, 1.389 ms as written; 1.444 without the second item space): tail items confined to
        // the warps without pair items, in rounds: 1.419; half a column per item: 1.406; one item per element sweeping all its
        // tail columns: 1.452 (profiles/r2_ab_cta_shapes_and_layouts.txt).
        for (int t = BLOCK - 1 - tid; t < (F::ND - PC) * ne; t += BLOCK) {
            const int k = t / ne, le = t - k * ne, J = PC + k;
            const uint32_t m = smask[le];
            if (m & (1u << J)) tl_column_to_stage<F, S>((int)td.qbase[__popc(m & ((1u << J) - 1u))] + le, le, J, ne, nq, Gs, stage);
        }
    }
}

// ---- phase 2 of the numeric kernels ---------------------------------------------------------------------
// Light nonzeros (1 or 2 contributions: all but the matrix diagonals of node patches): one packed word per nonzero
// names both stage entries.  The slots are walked in rows of 32 (one per lane) that never straddle a run, TL_U rows in
// flight per warp: destination = rowbase[row] + lane, no per-nonzero run tracking.
template <int BLOCK, int U = TL_U>
__device__ __forceinline__ void tl_gather_light(const TileDescFull &td, const double *__restrict__ stage, const unsigned char *__restrict__ smeta,
                                                double *__restrict__ nzval, int lane, int warp)
{
    constexpr int NW = BLOCK / 32;
    const uint32_t *spk = reinterpret_cast<const uint32_t *>(smeta);
    const int64_t *srow = reinterpret_cast<const int64_t *>(smeta + tl_meta_goff_bytes(td.nslot) + tl_meta_gidx_bytes(td.ncontrib));
    const int nrows = td.nslot >> 5;
    for (int r0 = warp; r0 < nrows; r0 += NW * U) {
        uint32_t pk[U];
        int64_t base[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = r0 + u * NW;
            const bool in = r < nrows;
            pk[u] = in ? spk[r * 32 + lane] : 0xFFFFFFFFu;
            base[u] = in ? srow[r] : 0;
        }
        double acc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i0 = pk[u] & 0xFFFFu, i1 = pk[u] >> 16;
#if TL_GATHER_SELECT   // loads always issued (index 0 when there is no contribution), selects instead of predicated loads
            const bool h0 = i0 < 0xFFFEu, h1 = i1 < 0xFFFEu;
            const double v0 = stage[h0 ? i0 : 0u], v1 = stage[h1 ? i1 : 0u];
            acc[u] = h0 ? v0 : 0.0;
            acc[u] = h1 ? __dadd_rn(acc[u], v1) : acc[u];
#else
            acc[u] = 0.0;
            if (i0 < 0xFFFEu) acc[u] = stage[i0];
            if (i1 < 0xFFFEu) acc[u] = __dadd_rn(acc[u], stage[i1]);
#endif
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if ((pk[u] & 0xFFFFu) < 0xFFFEu) nzval[base[u] + lane] = acc[u];      // (heavy and padding slots carry 0xFFFE / 0xFFFF there)
    }
}
// Heavy nonzeros: listed per tile, one per lane (all lanes loop about equally long), left-to-right sum.
// WF (fused load vector): a tiling built with EFG_OPT_FUSE_LOAD lists EVERY diagonal nonzero here.  A heavy entry whose first
// contribution is a diagonal stage entry (index < fdiag = ND * nqs) is a diagonal nonzero of the matrix; the load-vector entry of
// its dof (= its row index) sums the fe entries of the same elements in the same order (ascending element = the order
// assemble!(av, fe) adds them, src/Assemblers.jl:217-223), found foff = srows * nqs further on in the stage.  The light walk
// above stays exactly the unfused one (a first version detected diagonal slots there and paid a dependent rowval load per row
// of slots: 5.4 ms instead of 2.8 for T6 heat).
template <int BLOCK, bool WF = false>
__device__ __forceinline__ void tl_gather_heavy(const TileDescFull &td, const double *__restrict__ stage, const unsigned char *__restrict__ smeta,
                                                double *__restrict__ nzval, int tid, uint32_t fdiag = 0, uint32_t foff = 0,
                                                const int32_t *__restrict__ rowval = nullptr, double *__restrict__ fout = nullptr)
{
    const uint16_t *hidx = reinterpret_cast<const uint16_t *>(smeta + tl_meta_goff_bytes(td.nslot));
    const int64_t *srow = reinterpret_cast<const int64_t *>(reinterpret_cast<const unsigned char *>(hidx) + tl_meta_gidx_bytes(td.ncontrib));
    const TileHeavy *heavy = reinterpret_cast<const TileHeavy *>(reinterpret_cast<const unsigned char *>(srow) + tl_meta_rows_bytes(td.nslot));
    for (int h = tid; h < td.nheavy; h += BLOCK) {
        const TileHeavy e = heavy[h];
        double acc = stage[hidx[e.o]];
        for (int k = 1; k < (int)e.c; k++) acc = __dadd_rn(acc, stage[hidx[e.o + k]]);
        const int64_t pos = srow[e.s >> 5] + (e.s & 31);
        nzval[pos] = acc;
        int32_t dof = -1;
        if constexpr (WF) { if ((uint32_t)hidx[e.o] < fdiag) dof = (int32_t)((uint32_t)hidx[e.o + e.c] | ((uint32_t)hidx[e.o + e.c + 1] << 16)); }
        if constexpr (WF) {
            if (dof >= 0) {
                double f = stage[hidx[e.o] + foff];
                for (int k = 1; k < (int)e.c; k++) f = __dadd_rn(f, stage[hidx[e.o + k] + foff]);
                fout[dof] = f;
            }
        }
    }
}

template <class F, bool S, int BLOCK, int MINB, bool WF = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_tl_numeric(const TileDescFull *__restrict__ tiles, int ntiles,
                                                            const int32_t *__restrict__ tconn, const uint16_t *__restrict__ tmask,
                                                            const double2 *__restrict__ xy, const unsigned char *__restrict__ meta,
                                                            double *__restrict__ nzval, int pf_dist, const unsigned char *__restrict__ geo,
                                                            const int32_t *__restrict__ rowval = nullptr, double *__restrict__ fout = nullptr)
{
    static_assert(!WF || (tl_can_fuse<F>() && TL_GEO), "fused load vector: symmetric-stage scalar forms on the geometry-block path");
    constexpr int SROWS = tl_srows<F>() + (WF ? F::ND : 0);      // stage rows (WF: + one row per local dof for fe)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDescFull td;
    __shared__ __align__(8) unsigned long long bar, bar2;
    constexpr bool GEO = TL_GEO && (!F::SPLIT || TL_GEO_SPLIT);
    constexpr int DW = (int)(sizeof(TileDescFull) / 8);
    constexpr int GK = F::GK;
    constexpr int NW = BLOCK / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < DW) reinterpret_cast<int64_t *>(&td)[tid] = reinterpret_cast<const int64_t *>(&tiles[blockIdx.x])[tid];
    // the tile that will occupy a CTA slot about one wave later (pf_dist = CTAs resident on the device): its
    // connectivity is pulled into L2 at the end of this CTA, so that tile's first dependent load is an L2 hit
    const uint32_t barA = tl_smem_addr(&bar), barB = tl_smem_addr(&bar2);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barB));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (td.nslot == 0) return;
    double *stage = reinterpret_cast<double *>(smem_raw);
    constexpr int GSZ = tl_gsz<F>();
    // SPLIT forms: the per-element geometry of phase 1 and the gather metadata of phase 2 share the same shared-memory
    // area, so the TMA copy is issued after phase 1; otherwise it is issued now and lands while phase 1 computes
    unsigned char *smeta = smem_raw + tl_stage_bytes(SROWS, td.nqs);
    // geometry block (GEO): txy | conn16 | mask16, behind the metadata area (vector forms: behind the area the metadata shares with Gs)
    constexpr bool LATE_META = F::SPLIT && tl_gs_alias<F>();      // metadata copied in after phase 1 (its area doubles as Gs)
    unsigned char *sgeo = smeta + (F::SPLIT ? (tl_gs_alias<F>() ? max(td.meta_bytes, tl_geo_bytes(GSZ, td.nelem)) : td.meta_bytes + tl_geo_bytes(GSZ, td.nelem)) : td.meta_bytes);
    if (GEO && tid == 0) tl_bulk_load(tl_smem_addr(sgeo), geo + td.geo0, (uint32_t)td.geo_bytes, barB);   // needed first
    if (!LATE_META && tid == 0) tl_bulk_load(tl_smem_addr(smeta), meta + td.meta0, (uint32_t)td.meta_bytes, barA);
    if (LATE_META) {   // the TMA copy is issued after phase 1 (shared area): pull the block into L2 meanwhile
        const char *mp = reinterpret_cast<const char *>(meta + td.meta0);
        for (int o = tid * 128; o < td.meta_bytes; o += BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(mp + o));
    }
    const int nq = td.nqs;      // stage row stride (odd)
    const int ncols = td.nq;    // staged columns

    if constexpr (!F::SPLIT && GEO) {
        // phase 1 from the tile's geometry block: no global loads at all -- local connectivity and coordinates come
        // from shared memory as soon as the bulk copy has landed
        tl_mbar_wait(barB, 0);
        const double2 *sxy = reinterpret_cast<const double2 *>(sgeo);
        const uint16_t *sc16 = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode));
        const uint16_t *sm16 = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode) + tl_geo_conn_bytes(GK, td.nelem));
        if constexpr (form_dim3<F>::value) {
            const double *sz = reinterpret_cast<const double *>(sgeo + tl_geo_z_off(GK, td.nelem, td.nnode));
            for (int le = tid; le < td.nelem; le += BLOCK) tl_element_to_stage_local3<F, S>((uint32_t)le, sc16, sm16, sxy, sz, stage, td.qbase, nq);
        } else
        for (int le = tid; le < td.nelem; le += BLOCK) tl_element_to_stage_local<F, S, WF>((uint32_t)le, sc16, sm16, sxy, stage, td.qbase, nq);
    } else if constexpr (!F::SPLIT) {
        // phase 1: one thread per tile element: owned columns of the element matrix -> stage
        for (int le = tid; le < td.nelem; le += BLOCK) {
            if constexpr (F::GK >= 4) {    // out of line: measured 4.47 -> 3.96 ms on config 2 (no spills in the hot code)
                tl_element_to_stage<F, S>(td.elem0 + le, (uint32_t)le, tconn, tmask, xy, stage, td.qbase, nq);
            } else {
                const int64_t g = td.elem0 + le;
                double X[GK], Y[GK];
#pragma unroll
                for (int a = 0; a < GK; a++) { const double2 p = __ldg(&xy[tconn[g * GK + a]]); X[a] = p.x; Y[a] = p.y; }
                StageEmit<F> emit{stage, td.qbase, (uint32_t)tmask[g], (uint32_t)le, nq};
                F::template element<S>(X, Y, emit.m, emit);
            }
        }
    } else {
        // phase 1a: one thread per tile element: Jacobian / JxW / gradients at every quadrature point -> shared memory
        const int ne = td.nelem;
        double *Gs = reinterpret_cast<double *>(smeta + (tl_gs_alias<F>() ? 0 : td.meta_bytes));   // SoA: Gs[k * ne + le]
        const uint16_t *smask;
        if constexpr (GEO) {
            tl_mbar_wait(barB, 0);
            const double2 *sxy = reinterpret_cast<const double2 *>(sgeo);
            const uint16_t *sc16 = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode));
            smask = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode) + tl_geo_conn_bytes(GK, ne));
#if TL_GEO_QSPLIT
            for (int t = tid; t < F::NQ * ne; t += BLOCK) { const int q = t / ne; tl_geometry_qp_to_smem_local<F, S>(t - q * ne, q, ne, sc16, sxy, Gs); }
#else
            for (int le = tid; le < ne; le += BLOCK) tl_geometry_to_smem_local<F, S>(le, ne, sc16, sxy, Gs);
#endif
        } else {
            uint16_t *sm = reinterpret_cast<uint16_t *>(Gs + GSZ * ne);
            for (int le = tid; le < ne; le += BLOCK)
                tl_geometry_to_smem<F, S>(td.elem0 + le, le, ne, tconn, tmask, xy, Gs, sm);
            smask = sm;
        }
        __syncthreads();
#if TL_PAIRS
        tl_phase1b_pairs<F, S, BLOCK>(td, smask, ne, nq, Gs, stage, tid);
#else
        // phase 1b: one thread per staged column (tile element, owned local column) -> stage
        for (int qc = tid; qc < ncols; qc += BLOCK) {
            int r = 0;
            while (r + 1 < F::ND && (int)td.qbase[r + 1] <= qc) r++;
            const int le = qc - (int)td.qbase[r];
            uint32_t mm = smask[le];
            for (int t = 0; t < r; t++) mm &= mm - 1;
            const int J = __ffs(mm) - 1;
            tl_column_to_stage<F, S>(qc, le, J, ne, nq, Gs, stage);
        }
#endif
    }
    __syncthreads();
    if (LATE_META && tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the area are done (barrier above)
        tl_bulk_load(tl_smem_addr(smeta), meta + td.meta0, (uint32_t)td.meta_bytes, barA);
    }
    tl_mbar_wait(barA, 0);

    // phase 2: every owned nonzero = left-to-right sum of its contributions (append order), stored once
    const uint32_t fdiag = (uint32_t)F::ND * (uint32_t)nq, foff = (uint32_t)tl_srows<F>() * (uint32_t)nq;
    tl_gather_light<BLOCK, F::SPLIT ? (F::ND >= 15 ? TL_U_SPLIT15 : TL_U_SPLIT) : TL_U>(td, stage, smeta, nzval, lane, warp);
    // The tile that will occupy a CTA slot about one wave later (pf_dist = CTAs resident on the device): its geometry
    // block (or connectivity) is pulled into L2 at the end of this CTA, so that tile's first load is an L2 hit.  The
    // descriptor is read here, not at kernel start, so nothing stays live across the two phases (no spills).
    const int tpf = blockIdx.x + pf_dist;
    int64_t pf_elem0 = 0; int pf_nelem = 0;
    if (tpf < ntiles) {
        if constexpr (GEO) { pf_elem0 = tiles[tpf].geo0; pf_nelem = tiles[tpf].geo_bytes; }
        else { pf_elem0 = tiles[tpf].elem0; pf_nelem = tiles[tpf].nelem; }
    }
    tl_gather_heavy<BLOCK, WF>(td, stage, smeta, nzval, tid, fdiag, foff, rowval, fout);
    if constexpr (GEO) {   // pull the geometry block of the tile one wave ahead into L2
        const char *p0 = reinterpret_cast<const char *>(geo + pf_elem0);
        for (int o = tid * 128; o < pf_nelem; o += BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
    } else {   // pull the connectivity of the tile one wave ahead into L2
        const char *p0 = reinterpret_cast<const char *>(tconn + pf_elem0 * GK);
        const int nbytes = pf_nelem * GK * 4;
        for (int o = tid * 128; o < nbytes; o += BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
        const char *p1 = reinterpret_cast<const char *>(tmask + pf_elem0);
        for (int o = tid * 128; o < pf_nelem * 2; o += BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + o));
    }
}

// Persistent variant for the TL_GEO forms: gridDim.x CTAs (one per CTA slot of the device) walk the tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  Shared memory holds three fixed areas -- stage | gather metadata | geometry
// block -- and two mbarriers.  While tile k is in its gather phase (stage + metadata in use) the geometry block of tile
// k+1 is already in flight into the geometry area; while tile k+1 computes (geometry + stage in use) its metadata lands
// in the metadata area: no phase waits for a global-memory round trip except in the first iteration.
template <class F, bool S, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_tl_numeric_p(const TileDescFull *__restrict__ tiles, int ntiles, const unsigned char *__restrict__ meta,
                                                              const unsigned char *__restrict__ geo, double *__restrict__ nzval, int off_meta, int off_geo,
                                                              int off_gs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDescFull tdb[2];
    __shared__ __align__(8) unsigned long long barG, barM;
    constexpr int DW = (int)(sizeof(TileDescFull) / 8);
    constexpr int GK = F::GK;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *stage = reinterpret_cast<double *>(smem_raw);
    unsigned char *smeta = smem_raw + off_meta, *sgeo = smem_raw + off_geo;
    const uint32_t bG = tl_smem_addr(&barG), bM = tl_smem_addr(&barM);
    int t = blockIdx.x;
    if (t >= ntiles) return;
    if (tid < DW) reinterpret_cast<int64_t *>(&tdb[0])[tid] = reinterpret_cast<const int64_t *>(&tiles[t])[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bG));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bM));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && tdb[0].nslot != 0) {
        tl_bulk_load(tl_smem_addr(sgeo), geo + tdb[0].geo0, (uint32_t)tdb[0].geo_bytes, bG);
        tl_bulk_load(tl_smem_addr(smeta), meta + tdb[0].meta0, (uint32_t)tdb[0].meta_bytes, bM);
    }
    uint32_t pg = 0, pm = 0;     // phase parities of the two barriers
    int buf = 0;
    for (; t < ntiles; t += gridDim.x, buf ^= 1) {
        const TileDescFull &td = tdb[buf];
        const TileDescFull &tdn = tdb[buf ^ 1];
        const int tn = t + gridDim.x;
        // descriptor of the next tile: written now, read after the barrier between the two phases
        if (tn < ntiles && tid < DW) reinterpret_cast<int64_t *>(&tdb[buf ^ 1])[tid] = reinterpret_cast<const int64_t *>(&tiles[tn])[tid];
        const bool work = td.nslot != 0;      // CTA-uniform
        if (work) {
            // phase 1: one thread per tile element, everything from the geometry block in shared memory
            tl_mbar_wait(bG, pg);
            pg ^= 1;
            const int nq = td.nqs;
            const double2 *sxy = reinterpret_cast<const double2 *>(sgeo);
            const uint16_t *sc16 = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode));
            const uint16_t *sm16 = reinterpret_cast<const uint16_t *>(sgeo + tl_geo_xy_bytes(td.nnode) + tl_geo_conn_bytes(GK, td.nelem));
            if constexpr (!F::SPLIT) {
                for (int le = tid; le < td.nelem; le += BLOCK) tl_element_to_stage_local<F, S>((uint32_t)le, sc16, sm16, sxy, stage, td.qbase, nq);
            } else {
                // phase 1a: one thread per tile element: Jacobian / JxW / gradients at every quadrature point -> Gs (SoA)
                const int ne = td.nelem, ncols = td.nq;
                double *Gs = reinterpret_cast<double *>(smem_raw + off_gs);
                for (int le = tid; le < ne; le += BLOCK) tl_geometry_to_smem_local<F, S>(le, ne, sc16, sxy, Gs);
                __syncthreads();
#if TL_PAIRS
                (void)ncols;
                tl_phase1b_pairs<F, S, BLOCK>(td, sm16, ne, nq, Gs, stage, tid);
#else
                // phase 1b: one thread per staged column (tile element, owned local column) -> stage
                for (int qc = tid; qc < ncols; qc += BLOCK) {
                    int r = 0;
                    while (r + 1 < F::ND && (int)td.qbase[r + 1] <= qc) r++;
                    const int le = qc - (int)td.qbase[r];
                    uint32_t mm = sm16[le];
                    for (int k = 0; k < r; k++) mm &= mm - 1;
                    const int J = __ffs(mm) - 1;
                    tl_column_to_stage<F, S>(qc, le, J, ne, nq, Gs, stage);
                }
#endif
            }
        }
        __syncthreads();      // stage complete; geometry area free; next descriptor visible
        if (tid == 0 && tn < ntiles && tdn.nslot != 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the area are done (barrier above)
            tl_bulk_load(tl_smem_addr(sgeo), geo + tdn.geo0, (uint32_t)tdn.geo_bytes, bG);
        }
        if (work) {
            tl_mbar_wait(bM, pm);
            pm ^= 1;
            tl_gather_light<BLOCK>(td, stage, smeta, nzval, lane, warp);
            tl_gather_heavy<BLOCK>(td, stage, smeta, nzval, tid);
        }
        __syncthreads();      // stage + metadata free; tdb[buf] may be overwritten in the next iteration
        if (tid == 0 && tn < ntiles && tdn.nslot != 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tl_bulk_load(tl_smem_addr(smeta), meta + tdn.meta0, (uint32_t)tdn.meta_bytes, bM);
        }
    }
}

// ---- host: symbolic ---------------------------------------------------------------------------------
// Tile sizes tried in turn (largest first): powers of two and 3*2^k keep space-filling-curve tiles compact
// (a 256-element T6 tile is a 16 x 8 block of cells).  The first size whose shared-memory footprint lets two
// CTAs share an SM is used.
// Threads per CTA x CTAs per SM of the one-thread-per-element forms, measured (profiles/r2_ab_cta_shapes_and_layouts.txt), always 20
// warps per SM at <= 96 registers (five warps per scheduler partition -- six would cap the kernel at 80 registers):
//   T3 / Q4 heat (<= 4 local dofs): 5 x 128.  2 x 320 -> 4 x 160 -> 5 x 128: Q4 1.548 -> 1.417 -> 1.378 ms, T3 0.627 -> 0.604
//     -> 0.599 ms: more CTAs in different phases overlap the compute-bound phase 1 of one with the LSU-bound gather of
//     another, and their tiles stay large enough (200-370 elements);
//   T6 heat (6 local dofs): 2 x 320.  4 x 160 gives the same kernel time (2.787 vs 2.809 ms) with tiles of 104 instead of
//     232 elements, which makes the tile phase of the symbolic part 12 ms slower (more tiles): not worth it end to end.
// SPLIT forms: 2 CTAs of 384 threads (elasticity: 256 / 288 / 320 / 352 / 384 / 416 -> 3.72 / 3.70 / 3.59 / 3.39 / 3.29 / 4.51 ms;
// 24 warps = six per scheduler partition at the 80 registers 352 threads already had) / 256 (Stokes), measured.
#ifndef TL_BLOCK_NS
#define TL_BLOCK_NS 320
#endif
#ifndef TL_BLOCK_NS4
#define TL_BLOCK_NS4 128
#endif
#ifndef TL_MINB4
#define TL_MINB4 5
#endif
#ifndef TL_BLOCK_SPLIT
#define TL_BLOCK_SPLIT (TL_PAIRS ? 384 : 256)
#endif
#ifndef TL_BLOCK_SPLIT15
#define TL_BLOCK_SPLIT15 256    // Stokes gen / veclap_alt (15 columns, tiles of 32 elements: 8 pair items per element -> 256 work items)
#endif
template <class F> __host__ __device__ constexpr int tl_block() { return F::SPLIT ? (F::ND >= 15 ? TL_BLOCK_SPLIT15 : TL_BLOCK_SPLIT) : (F::ND > 8 ? 256 : (F::ND <= 4 ? TL_BLOCK_NS4 : TL_BLOCK_NS)); }

#ifndef TL_MINB_SPLIT
#define TL_MINB_SPLIT 2
#endif
#ifndef TL_MINB_SPLIT15
#define TL_MINB_SPLIT15 TL_MINB_SPLIT
#endif
template <class F> constexpr int tl_minb() { return F::SPLIT ? (F::ND >= 15 ? TL_MINB_SPLIT15 : TL_MINB_SPLIT) : (F::ND > 8 ? 2 : (F::ND <= 4 ? TL_MINB4 : TL_MINB)); }

static const int TL_TILE_SIZES[] = {512, 384, 256, 192, 128, 96, 64, 56, 48, 40, 32};      // (smaller tiles only through EFG_OPT_TILE_ELEMS)
template <class F> __host__ __device__ constexpr int tl_items_1b() { return TL_PAIRS ? (F::ND + 1) / 2 : F::ND; }
template <class F> static int tl_default_tile_elems()
{
    // shared memory per owned element-equivalent ~ stage ND*ND*8 + gather metadata NT*2 + ~2.7*ND*ND
    // (SPLIT forms: the geometry area aliases the metadata area; symmetric stage: upper triangle of every tile element incl. halo)
    const double per_elem = tl_sym<F>() ? tl_srows<F>() * 8.0 * 1.3 + F::ND * F::ND * 2.7 + F::NT * 2.0 : F::ND * F::ND * 10.7 + F::NT * 2.0;
    int te = 32;
    for (int c : TL_TILE_SIZES)
        if (c * per_elem <= 118.0 * 1024 * 2 / tl_minb<F>()) { te = c; break; }
    if (F::SPLIT) {
        // phase 1b runs in rounds of tl_block() work items (te*ND staged columns, or te*ceil(ND/2) column pairs): avoid a
        // nearly empty last round
        for (int c : TL_TILE_SIZES) {
            if (c > te) continue;
            const double rounds = (double)c * tl_items_1b<F>() / tl_block<F>();
            if (rounds / ceil(rounds) >= 0.85) { te = c; break; }
        }
    }
    if (!F::SPLIT) {
        // phase 1 runs in rounds of tl_block() elements; a tile whose element count (own + halo) barely exceeds a
        // whole number of rounds leaves the last round almost empty.  Halo model for compact tiles: 1 + x/sqrt(te).
        const double x = F::GK == 6 ? 4.6 : (F::GK == 3 ? 4.0 : 2.9);
        const double halo = 1.0 + x / sqrt((double)te);
        const double rounds = te * halo / tl_block<F>();
        const double whole = floor(rounds);
        if (whole >= 1.0 && rounds - whole > 0.0 && rounds - whole < 0.5) {
            const int t2 = (int)(whole * tl_block<F>() * 0.95 / halo) & ~7;
            if (t2 >= 32) te = t2;
        }
    }
    return te;
}


// The dynamic-shared-memory ceiling of a kernel is a per-FUNCTION attribute shared by every ctx / host thread of the process:
// it is always raised to the device's opt-in maximum (never to the size one tiling needs), so that two ctx with different
// tile sizes cannot lower it under each other's launches (efgm_*: one host thread per device; a symbolic phase of one ctx
// used to race with the numeric launch of another).
static cudaError_t tl_raise_smem_limit(const efg_ctx *ctx, const void *kern)
{
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

// stage rows of the tiling being built: the fused load vector (EFG_OPT_FUSE_LOAD at symbolic time) adds one row per local dof
template <class F> static int tl_stage_rows(const efg_ctx *ctx)
{
    if constexpr (tl_can_fuse<F>()) return tl_srows<F>() + (ctx->opt_fuse_load ? F::ND : 0);
    else return tl_srows<F>();
}
template <class F> static const void *tl_numeric_kernel(const efg_ctx *ctx)
{
    if constexpr (tl_persist<F>()) return (const void *)k_tl_numeric_p<F, false, tl_block<F>(), tl_minb<F>()>;
    else {
        if constexpr (tl_can_fuse<F>() && TL_GEO)
            if (ctx->opt_fuse_load) return (const void *)k_tl_numeric<F, false, tl_block<F>(), tl_minb<F>(), true>;
        return (const void *)k_tl_numeric<F, false, tl_block<F>(), tl_minb<F>()>;
    }
}
// CTAs of the numeric kernel that fit an SM with the shared memory the last tile phase asked for
template <class F> static int tl_ctas_per_sm(efg_ctx *ctx)
{
    TiledData *td = tiled_data(ctx);
    const void *kern = tl_numeric_kernel<F>(ctx);
    if (tl_raise_smem_limit(ctx, kern) != cudaSuccess) { cudaGetLastError(); return 0; }
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tl_block<F>(), (size_t)td->smem_bytes));
    return per_sm;
}
// dynamic shared memory a CTA may use if two CTAs are to share an SM: half of the SM's shared memory minus the per-CTA
// reservation and the kernel's static shared memory (the occupancy query stays the final arbiter)
template <class F> static int tl_smem_budget(efg_ctx *ctx)
{
    int per_sm = 0, reserved = 0, optin = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    CUDA_CHECK(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, ctx->device));
    CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    cudaFuncAttributes fa;
    CUDA_CHECK(cudaFuncGetAttributes(&fa, tl_numeric_kernel<F>(ctx)));
    int b = per_sm / tl_minb<F>() - reserved - (int)fa.sharedSizeBytes;
    if (b > optin - (int)fa.sharedSizeBytes) b = optin - (int)fa.sharedSizeBytes;
    return b > 0 ? b : 0;
}

// ---- phase A: the CSC pattern (independent of the tile size) --------------------------------------------------------
// T0 combined element dof vectors -> T2/T3 adjacency (column -> (element, local column), ascending element: one stable radix
// sort) -> T4 per column: sorted unique rows = colptr / rowval.  When this returns (ev_pattern recorded) the pattern can
// already travel to the host while phase B and the numeric kernel run.
template <class F> static void tiled_pattern(efg_ctx *ctx)
{
    static_assert(F::ND <= TL_MAXND, "");
    constexpr int ND = F::ND;
    const MeshDev &m0 = ctx->mesh[0];
    const int64_t nel = m0.nel;
    const int64_t ncl = ctx->ncl;
    if (nel * ND >= ((int64_t)1 << 32)) efg_throw(EFG_ERR_LIMIT, "tiled path: nel*ND exceeds 2^32; shard the mesh (efg_set_column_range)");
    cudaStream_t st = ctx->stream;
    DevPool &pool = ctx->pool;
    tiled_release(ctx);
    TiledSym *sy = new TiledSym();
    tiled_sym(ctx) = sy;
    TlTrace trace(ctx);
    // one slab for everything the two phases allocate before nnz is known (pairs: edof + adj + 24 B of sort scratch, per
    // column: counts, offsets, owners, tile column lists), a second one below once nnz is known
    pool.reserve((size_t)(nel * ND) * 32 + (size_t)ncl * 24 + ((size_t)64 << 20));

    DevBuf<int> err;
    err.alloc(pool, 1);
    CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), st));
    trace.mark("start");
    // T0: combined element dof table
    sy->edof.alloc(pool, (size_t)(nel * ND));
    DofSrc src{m0.conn.p, ctx->mesh[1].conn.p, ctx->space[0].dof.p, ctx->space[1].dof.p, ctx->space[2].dof.p, ctx->space[0].cdof.p, ctx->space[1].cdof.p, ctx->space[2].cdof.p};
    LAUNCH(ctx, k_tl_edofs<F>, grid_for(nel, 256), 256, 0, src, nel, ctx->nrow, ctx->ncol, sy->edof.p, err.p);
    if (tl_read(ctx, err.p))
        efg_throw(EFG_ERR_INDEX, "ArgumentError: a dof number is < 1 or exceeds nrow/ncol (was every space numbered, incl. data dofs?)");
    trace.mark("T0 edofs");

    // T2/T3: adjacency (column -> (element, local column)), ascending element
    DevBuf<uint32_t> adjcnt;
    adjcnt.alloc(pool, (size_t)ncl + 2); sy->adjptr.alloc(pool, (size_t)ncl + 2);
    CUDA_CHECK(cudaMemsetAsync(adjcnt.p, 0, (size_t)(ncl + 2) * sizeof(uint32_t), st));
    sy->adj.alloc(pool, (size_t)(nel * ND));
    {
        const size_t np = (size_t)(nel * ND);
        uint32_t *k1 = reinterpret_cast<uint32_t *>(tl_scratch(ctx, 3 * (np + 2) * sizeof(uint64_t))), *k2 = k1 + np, *v1 = k2 + np;   // (sized for phase B's 64-bit sort)
        LAUNCH(ctx, k_tl_pairs<ND>, grid_for(nel * ND, 256), 256, 0, sy->edof.p, nel, COLMAP(ctx), (uint32_t)ncl, adjcnt.p, k1, v1);
        tl_sort_pairs(ctx, k1, k2, v1, sy->adj.p, nel * ND, bits_for(ncl));
    }
    tl_excl_scan(ctx, adjcnt.p, sy->adjptr.p, ncl + 1);
    sy->npairs = (int64_t)tl_read(ctx, sy->adjptr.p + ncl);
    trace.mark("T2/T3 adjacency");

    // T4: CSC pattern
    sy->colcnt.alloc(pool, (size_t)ncl + 1); sy->hcnt.alloc(pool, (size_t)ncl + 1); sy->ccnt.alloc(pool, (size_t)ncl + 1); sy->colptr0.alloc(pool, (size_t)ncl + 1);
    CUDA_CHECK(cudaMemsetAsync(sy->colcnt.p, 0, (size_t)ncl + 1, st));
    CUDA_CHECK(cudaMemsetAsync(sy->hcnt.p, 0, (size_t)ncl + 1, st));
    CUDA_CHECK(cudaMemsetAsync(sy->ccnt.p, 0, ((size_t)ncl + 1) * 2, st));
    LAUNCH(ctx, k_tl_col_count<F>, grid_for(ncl, 128), 128, 0, sy->adjptr.p, sy->adj.p, sy->edof.p, ncl, sy->colcnt.p, sy->ccnt.p, sy->hcnt.p, err.p, (int)(ctx->opt_fuse_load && !ctx->have_range && tl_can_fuse<F>()));
    if (tl_read(ctx, err.p))
        efg_throw(EFG_ERR_LIMIT, "tiled path: a matrix column has more than %d distinct rows (node valence too high); use EFG_OPT_PATH=1", TL_CAP);
    {
        cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const uint8_t *> it(sy->colcnt.p, cub::CastOp<int64_t>());
        tl_excl_scan(ctx, it, sy->colptr0.p, ncl + 1);
    }
    const int64_t nnz = tl_read(ctx, sy->colptr0.p + ncl);
    ctx->nnz = nnz;
    // everything proportional to nnz and to the tile elements: rowval 4 + nzval 8 + gather words 4 (+ heavy lists) per
    // nonzero, geometry blocks and tile-element lists per pair
    pool.reserve((size_t)nnz * 9 + (size_t)(nel * ND) * 8 + ((size_t)64 << 20));
    ctx->rowval.alloc(pool, (size_t)(nnz > 0 ? nnz : 1));
    ctx->colptr.alloc(pool, (size_t)ncl + 1);
    LAUNCH(ctx, k_tl_col_fill<F>, grid_for(ncl + 1, 128), 128, 0, sy->adjptr.p, sy->adj.p, sy->edof.p, ncl, sy->colptr0.p, ctx->rowval.p, ctx->colptr.p);
    CUDA_CHECK(cudaEventRecord(ctx->ev_pattern, st));
    ctx->have_pattern = true;
    trace.mark("T4 pattern");
}

// ---- phase B0: element order along a space-filling curve (independent of the tile size) -------------------------------
template <class F> static void tiled_order(efg_ctx *ctx)
{
    TiledSym *sy = tiled_sym(ctx);
    const MeshDev &gm = ctx->mesh[F::GMESH];
    const int64_t nel = ctx->mesh[0].nel;
    cudaStream_t st = ctx->stream;
    DevPool &pool = ctx->pool;
    TlTrace trace(ctx);
    sy->have_order = false;
    if (ctx->opt_sfc && nel > 32) {
        DevBuf<BBox> bb;
        bb.alloc(pool, 1);
        cub::TransformInputIterator<BBox, XYToBBox, const double2 *> it(gm.xy.p, XYToBBox());
        const BBox init{1e300, 1e300, -1e300, -1e300};
        size_t tb = 0;
        cub::DeviceReduce::Reduce(nullptr, tb, it, bb.p, gm.nnodes, BBoxOp(), init, st);
        DevBuf<char> tmp;
        tmp.alloc(pool, tb);
        CUDA_CHECK(cub::DeviceReduce::Reduce(tmp.p, tb, it, bb.p, gm.nnodes, BBoxOp(), init, st));
        ctx->launches += 2;
        DevBuf<uint32_t> k1, k2, v1;
        k1.alloc(pool, (size_t)nel); k2.alloc(pool, (size_t)nel); v1.alloc(pool, (size_t)nel);
        sy->eorder.alloc(pool, (size_t)nel);
        if constexpr (form_dim3<F>::value) {
            DevBuf<double> zmm;
            zmm.alloc(pool, 2);
            size_t t2 = 0;
            cub::DeviceReduce::Min(nullptr, t2, gm.z.p, zmm.p, gm.nnodes, st);
            DevBuf<char> tmp2;
            tmp2.alloc(pool, t2);
            CUDA_CHECK(cub::DeviceReduce::Min(tmp2.p, t2, gm.z.p, zmm.p, gm.nnodes, st));
            CUDA_CHECK(cub::DeviceReduce::Max(tmp2.p, t2, gm.z.p, zmm.p + 1, gm.nnodes, st));
            ctx->launches += 2;
            LAUNCH(ctx, k_tl_morton3, grid_for(nel, 256), 256, 0, gm.conn.p, (int)F::GK, gm.xy.p, gm.z.p, nel, bb.p, zmm.p, k1.p, v1.p);
        } else
        LAUNCH(ctx, k_tl_morton, grid_for(nel, 256), 256, 0, gm.conn.p, gm.kind, gm.xy.p, nel, bb.p, k1.p, v1.p);
        tl_sort_pairs(ctx, k1.p, k2.p, v1.p, sy->eorder.p, nel, 32);
        sy->have_order = true;
    }
    trace.mark("T1 morton order");
}

// ---- phase B: tiles of `te` elements -- owners, tile column lists, tile elements (own + halo), geometry blocks, gather words --
template <class F> static void tiled_tiles(efg_ctx *ctx, int te)
{
    constexpr int ND = F::ND;
    const MeshDev &m0 = ctx->mesh[0];
    const MeshDev &gm = ctx->mesh[F::GMESH];
    const int64_t nel = m0.nel;
    const int64_t ncl = ctx->ncl;
    const int64_t nnz = ctx->nnz;
    cudaStream_t st = ctx->stream;
    DevPool &pool = ctx->pool;
    TiledSym *sy = tiled_sym(ctx);
    const int64_t npairs = sy->npairs;
    tiled_release_tiles(ctx);
    TiledData *td = new TiledData();
    tiled_data(ctx) = td;
    TlTrace trace(ctx);

    DevBuf<int> err;
    err.alloc(pool, 1);
    CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), st));

    // T1: tile of each element
    const int ntiles = (int)((nel + te - 1) / te);
    DevBuf<int32_t> etile;
    etile.alloc(pool, (size_t)nel);
    if (sy->have_order && nel > te) LAUNCH(ctx, k_tl_etile_from_order, grid_for(nel, 256), 256, 0, sy->eorder.p, nel, te, etile.p);
    else LAUNCH(ctx, k_tl_etile_identity, grid_for(nel, 256), 256, 0, nel, te, etile.p);
    // T2: column owners
    DevBuf<int32_t> owner;
    owner.alloc(pool, (size_t)ncl + 1);
    LAUNCH(ctx, k_tl_fill_i32, grid_for(ncl + 1, 256), 256, 0, owner.p, ncl + 1, INT_MAX);
    LAUNCH(ctx, k_tl_owner<ND>, grid_for(nel * ND, 256), 256, 0, sy->edof.p, etile.p, nel, COLMAP(ctx), owner.p);
    etile.release();
    trace.mark("T1/T2 tiles+owners");

    const int32_t *edof_p = sy->edof.p;
    const uint32_t *adj_p = sy->adj.p, *adjptr_p = sy->adjptr.p;
    const int64_t *colptr0_p = sy->colptr0.p;
    // T5: tile column lists, tile-order slot / gather offsets, runs
    DevBuf<uint32_t> tkeys, tcols;
    tkeys.alloc(pool, (size_t)ncl + 1); tcols.alloc(pool, (size_t)ncl + 1);
    {
        DevBuf<uint32_t> k1, v1;
        k1.alloc(pool, (size_t)ncl + 1); v1.alloc(pool, (size_t)ncl + 1);
        LAUNCH(ctx, k_tl_tilecol_keys, grid_for(ncl, 256), 256, 0, owner.p, ncl, ntiles, k1.p, v1.p);
        tl_sort_pairs(ctx, k1.p, tkeys.p, v1.p, tcols.p, ncl, bits_for(ntiles));
    }
    DevBuf<int64_t> tcol_ptr, tcol_slot, tcol_gidx, tcol_heavy;
    tcol_ptr.alloc(pool, (size_t)ntiles + 1); tcol_slot.alloc(pool, (size_t)ncl + 2); tcol_gidx.alloc(pool, (size_t)ncl + 2); tcol_heavy.alloc(pool, (size_t)ncl + 2);
    LAUNCH(ctx, k_tl_lower_bounds<uint32_t>, grid_for(ntiles + 1, 256), 256, 0, tkeys.p, ncl, ntiles, 0, tcol_ptr.p);
    const int64_t nowned = tl_read(ctx, tcol_ptr.p + ntiles);
    {
        cub::CountingInputIterator<int64_t> cnt_it(0);
        cub::TransformInputIterator<int64_t, GatherU8, cub::CountingInputIterator<int64_t>> it8(cnt_it, GatherU8{sy->colcnt.p, tcols.p, nowned});
        tl_excl_scan(ctx, it8, tcol_slot.p, nowned + 1);
        cub::TransformInputIterator<int64_t, GatherU16, cub::CountingInputIterator<int64_t>> it16(cnt_it, GatherU16{sy->ccnt.p, tcols.p, nowned});
        tl_excl_scan(ctx, it16, tcol_gidx.p, nowned + 1);
        cub::TransformInputIterator<int64_t, GatherU8, cub::CountingInputIterator<int64_t>> ith(cnt_it, GatherU8{sy->hcnt.p, tcols.p, nowned});
        tl_excl_scan(ctx, ith, tcol_heavy.p, nowned + 1);
    }
    const int64_t ncontrib = tl_read(ctx, tcol_gidx.p + nowned);
    td->ncontrib = ncontrib;

    DevBuf<int32_t> rflag;
    DevBuf<int64_t> runidx, run_firstk;
    rflag.alloc(pool, (size_t)nowned + 1); runidx.alloc(pool, (size_t)nowned + 1);
    CUDA_CHECK(cudaMemsetAsync(rflag.p, 0, ((size_t)nowned + 1) * sizeof(int32_t), st));
    LAUNCH(ctx, k_tl_run_flags, grid_for(nowned, 256), 256, 0, tkeys.p, tcols.p, nowned, rflag.p);
    {
        cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *> it(rflag.p, cub::CastOp<int64_t>());
        tl_excl_scan(ctx, it, runidx.p, nowned + 1);
    }
    const int64_t nruns = tl_read(ctx, runidx.p + nowned);
    td->nruns = nruns;
    DevBuf<TileRun> runs;
    runs.alloc(pool, (size_t)(nruns > 0 ? nruns : 1));
    run_firstk.alloc(pool, (size_t)(nruns > 0 ? nruns : 1));
    LAUNCH(ctx, k_tl_run_fill, grid_for(nowned, 256), 256, 0, rflag.p, runidx.p, tkeys.p, tcols.p, nowned, tcol_slot.p, tcol_ptr.p, colptr0_p, runs.p, run_firstk.p);
    LAUNCH(ctx, k_tl_run_len, grid_for(nruns, 256), 256, 0, run_firstk.p, nruns, nowned, tcol_slot.p, runs.p);
    // padded slot positions (rows of 32 slots never straddle a run): pslot replaces tcol_slot wherever a POSITION in the
    // tile's metadata is meant
    DevBuf<int64_t> padlen, run_pad0, pslot;
    padlen.alloc(pool, (size_t)nruns + 2); run_pad0.alloc(pool, (size_t)nruns + 2); pslot.alloc(pool, (size_t)nowned + 2);
    LAUNCH(ctx, k_tl_run_padlen, grid_for(nruns + 1, 256), 256, 0, runs.p, nruns, padlen.p);
    tl_excl_scan(ctx, padlen.p, run_pad0.p, nruns + 1);
    LAUNCH(ctx, k_tl_pslot, grid_for(nowned + 1, 256), 256, 0, rflag.p, runidx.p, run_firstk.p, tcol_slot.p, run_pad0.p, nowned, nruns, pslot.p);
    LAUNCH(ctx, k_tl_run_s0, grid_for(nruns, 256), 256, 0, run_firstk.p, nruns, tkeys.p, tcol_ptr.p, pslot.p, runs.p);

    trace.mark("T5 tile columns/runs");
    // T6: tile element lists (own + halo), masks, popcount-descending order inside a tile
    // scratch layout: [ek1 (later: eflag) | ek2 | eidx], each nel*ND 8-byte words
    const size_t npk = (size_t)(nel * ND) + 2;
    uint64_t *ek1 = reinterpret_cast<uint64_t *>(tl_scratch(ctx, 3 * npk * sizeof(uint64_t)));
    uint64_t *ek2 = ek1 + npk;
    int64_t *eidx = reinterpret_cast<int64_t *>(ek2 + npk);
    int32_t *eflag = reinterpret_cast<int32_t *>(ek1);       // ek1 is dead once sorted into ek2
    LAUNCH(ctx, k_tl_telem_keys<ND>, grid_for(nel * ND, 256), 256, 0, edof_p, nel, COLMAP(ctx), owner.p, ntiles, ek1);
    tl_sort_keys(ctx, ek1, ek2, nel * ND, 36 + bits_for(ntiles));   // out-of-range pairs carry tile id ntiles: they sort last
    trace.mark("  T6a keys+sort64");
    CUDA_CHECK(cudaMemsetAsync(eflag, 0, ((size_t)npairs + 1) * sizeof(int32_t), st));
    LAUNCH(ctx, k_tl_head_flags64, grid_for(npairs, 256), 256, 0, ek2, npairs, 4, eflag);
    {
        cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *> it(eflag, cub::CastOp<int64_t>());
        tl_excl_scan(ctx, it, eidx, npairs + 1);
    }
    const int64_t ntelem = tl_read(ctx, eidx + npairs);
    trace.mark("  T6b flags+scan");
    td->ntelem = ntelem;
    if (ntelem >= ((int64_t)1 << 32)) efg_throw(EFG_ERR_LIMIT, "tiled path: too many tile elements");
    DevBuf<uint64_t> telem_key;
    DevBuf<uint16_t> emask;
    DevBuf<uint32_t> key2, val2, key2s, order2, newpos, pc_hist;
    telem_key.alloc(pool, (size_t)ntelem + 1); emask.alloc(pool, (size_t)ntelem + 1);
    key2.alloc(pool, (size_t)ntelem + 1); val2.alloc(pool, (size_t)ntelem + 1);
    key2s.alloc(pool, (size_t)ntelem + 1); order2.alloc(pool, (size_t)ntelem + 1); newpos.alloc(pool, (size_t)ntelem + 1);
    pc_hist.alloc(pool, (size_t)ntiles * 17);
    CUDA_CHECK(cudaMemsetAsync(pc_hist.p, 0, (size_t)ntiles * 17 * sizeof(uint32_t), st));
    LAUNCH(ctx, k_tl_telem_fill, grid_for(npairs, 256), 256, 0, ek2, npairs, eflag, eidx, telem_key.p, emask.p, key2.p, val2.p, pc_hist.p);
    tl_sort_pairs(ctx, key2.p, key2s.p, val2.p, order2.p, ntelem, 5 + bits_for(ntiles));
    trace.mark("  T6e sort popcount");
    LAUNCH(ctx, k_tl_invert, grid_for(ntelem, 256), 256, 0, order2.p, ntelem, newpos.p);
    DevBuf<int64_t> telem_ptr;
    telem_ptr.alloc(pool, (size_t)ntiles + 1);
    LAUNCH(ctx, k_tl_lower_bounds<uint64_t>, grid_for(ntiles + 1, 256), 256, 0, telem_key.p, ntelem, ntiles, 32, telem_ptr.p);
    td->tconn.alloc(pool, (size_t)(ntelem * F::GK + 1));
    td->tmask.alloc(pool, (size_t)ntelem + 1);
    LAUNCH(ctx, k_tl_tconn<F::GK>, grid_for(ntelem, 256), 256, 0, order2.p, ntelem, telem_key.p, emask.p, gm.conn.p, td->tconn.p, td->tmask.p);
    // TL_GEO (one-thread-per-element forms): tile-local node numbering for the geometry blocks
    constexpr bool GEO = TL_GEO && (!F::SPLIT || TL_GEO_SPLIT || tl_persist<F>());
    // (tlocal / tnodes live in the persistent scratch: the sort buffers it held are dead by now)
    uint16_t *tlocal = nullptr;
    int32_t *tnodes = nullptr;
    DevBuf<int32_t> tnn;
    if (GEO) {
        const size_t nent = (size_t)(ntelem * F::GK + 8);
        char *sc = tl_scratch(ctx, nent * 4 + ((nent * 2 + 15) & ~(size_t)15));
        tnodes = reinterpret_cast<int32_t *>(sc);
        tlocal = reinterpret_cast<uint16_t *>(sc + nent * 4);
        tnn.alloc(pool, (size_t)ntiles + 1);
        LAUNCH(ctx, k_tl_geo_local<F::GK>, (unsigned)ntiles, 256, 0, telem_ptr.p, td->tconn.p, tlocal, tnodes, tnn.p, err.p);
        if (tl_read(ctx, err.p))
            efg_throw(EFG_ERR_LIMIT, "tiled path: a tile has more than %d connectivity entries; lower EFG_OPT_TILE_ELEMS", TL_GEO_CAP);
    }

    trace.mark("T6 tile elements");
    // tile descriptors + metadata block layout
    DevBuf<int32_t> maxima;
    DevBuf<int64_t> mbytes, moff, gbytes, goffs;
    maxima.alloc(pool, 8);
    mbytes.alloc(pool, (size_t)ntiles + 1); moff.alloc(pool, (size_t)ntiles + 1);
    gbytes.alloc(pool, (size_t)ntiles + 1); goffs.alloc(pool, (size_t)ntiles + 1);
    CUDA_CHECK(cudaMemsetAsync(maxima.p, 0, 8 * sizeof(int32_t), st));
    CUDA_CHECK(cudaMemsetAsync(mbytes.p, 0, ((size_t)ntiles + 1) * sizeof(int64_t), st));
    CUDA_CHECK(cudaMemsetAsync(gbytes.p, 0, ((size_t)ntiles + 1) * sizeof(int64_t), st));
    td->tiles.alloc(pool, (size_t)ntiles);
    LAUNCH(ctx, k_tl_tiles_fill, grid_for(ntiles, 128), 128, 0, ntiles, tl_stage_rows<F>(ctx), tl_sym<F>(), tl_gsz<F>(), tl_gs_alias<F>(), (int)F::GK, (int)form_dim3<F>::value, GEO ? tnn.p : (const int32_t *)nullptr, gbytes.p, tcol_ptr.p, pslot.p, tcol_gidx.p, tcol_heavy.p, telem_ptr.p, runidx.p,
           nowned, nruns, pc_hist.p, td->tiles.p, mbytes.p, maxima.p);
    int32_t hmax[8];
    tl_read_bytes(ctx, maxima.p, hmax, (int)sizeof hmax);
    ctx->tl.max_nq = hmax[1] / tl_stage_rows<F>(ctx);
    ctx->tl.fused_rows = tl_stage_rows<F>(ctx) != tl_srows<F>(); ctx->tl.max_nslot = 0; ctx->tl.max_nelem = hmax[3]; ctx->tl.max_nrun = 0;
    ctx->tl.tile_elems = te;
    td->stage_bytes = 0;
    td->meta_max = 0;
    td->smem_bytes = hmax[0];                                  // max over tiles of (stage + gather metadata + geometry block)
    td->persist = tl_persist<F>();
    if (td->persist) {      // fixed areas: the next tile's blocks are fetched while the current tile still uses the others
        td->off_meta = hmax[4]; td->off_geo = hmax[4] + hmax[5]; td->off_gs = hmax[4] + hmax[5] + hmax[6];
        td->smem_bytes = hmax[4] + hmax[5] + hmax[6] + hmax[7];     // (hmax[7] = 0 for the one-thread-per-element forms)
    }
    if (hmax[1] > 65533)
        efg_throw(EFG_ERR_LIMIT, "tiled path: a tile stages %d element-matrix entries (> 65535); lower EFG_OPT_TILE_ELEMS", hmax[1]);
    if (hmax[2] > 65535)
        efg_throw(EFG_ERR_LIMIT, "tiled path: a tile gathers %d heavy contributions (> 65535); lower EFG_OPT_TILE_ELEMS", hmax[2]);
    if (td->smem_bytes > 225 * 1024)
        efg_throw(EFG_ERR_LIMIT, "tiled path: a tile needs %d bytes of shared memory; lower EFG_OPT_TILE_ELEMS", td->smem_bytes);
    // Known now: the shared memory the numeric kernel would need with this tile size.  If it does not let two CTAs share an
    // SM the caller retries with a smaller size -- so stop here, before the (expensive) geometry blocks and gather words.
    if (trace.on) fprintf(stderr, "[efg trace] te = %d: %d bytes of shared memory per CTA, budget %d\n", te, td->smem_bytes, ctx->tl_smem_budget);
    if (ctx->opt_tile_elems == 0 && te > 32 && td->smem_bytes > ctx->tl_smem_budget) { td->complete = false; trace.mark("tile descriptors (too large)"); return; }

    tl_excl_scan(ctx, mbytes.p, moff.p, (int64_t)ntiles + 1);
    const int64_t meta_total = tl_read(ctx, moff.p + ntiles);
    LAUNCH(ctx, k_tl_tiles_meta0, grid_for(ntiles, 256), 256, 0, ntiles, moff.p, td->tiles.p);
    if (GEO) {
        tl_excl_scan(ctx, gbytes.p, goffs.p, (int64_t)ntiles + 1);
        td->geo_total = tl_read(ctx, goffs.p + ntiles);
        LAUNCH(ctx, k_tl_tiles_geo0, grid_for(ntiles, 256), 256, 0, ntiles, goffs.p, td->tiles.p);
        td->geo.alloc(pool, (size_t)(td->geo_total > 0 ? td->geo_total : 16));
        CUDA_CHECK(cudaMemsetAsync(td->geo.p, 0, (size_t)(td->geo_total > 0 ? td->geo_total : 16), st));
        LAUNCH(ctx, k_tl_geo_fill<F::GK>, grid_for(ntiles, 1, (int64_t)148 * 16), 128, 0, ntiles, td->tiles.p, tlocal, tnodes, td->tmask.p, gm.xy.p, td->geo.p, form_dim3<F>::value ? gm.z.p : (const double *)nullptr);
    }
    td->meta_bytes = meta_total;

    trace.mark("tile descriptors");
    // T8: gather lists into the metadata blocks
    td->meta.alloc(pool, (size_t)(meta_total > 0 ? meta_total : 16));
    CUDA_CHECK(cudaMemsetAsync(td->meta.p, 0, (size_t)(meta_total > 0 ? meta_total : 16), st));
    LAUNCH(ctx, k_tl_gather_build<F>, grid_for(nowned, 128), 128, 0, tkeys.p, tcols.p, nowned, adjptr_p, adj_p, edof_p, colptr0_p, ctx->rowval.p,
           telem_key.p, telem_ptr.p, emask.p, newpos.p, td->tiles.p, pslot.p, tcol_gidx.p, tcol_heavy.p, td->meta.p, err.p, (int)(ctx->opt_fuse_load && !ctx->have_range && tl_can_fuse<F>()));
    LAUNCH(ctx, k_tl_meta_rows, grid_for(nruns, 128), 128, 0, nruns, runs.p, run_firstk.p, tkeys.p, td->tiles.p, td->meta.p);
    const int e2 = tl_read(ctx, err.p);
    if (e2) efg_throw(EFG_ERR_LIMIT, "tiled path: a tile's gather list exceeds 16-bit offsets (%d); lower EFG_OPT_TILE_ELEMS", e2);

    trace.mark("T8 gather build");
    ctx->tl.ntiles = ntiles;
    ctx->tl.sum_tile_elems = ntelem;
    ctx->tl.numeric_bytes = (GEO ? td->geo_total : ntelem * (F::GK * 4 + 2) + gm.nnodes * 16) + meta_total + nnz * 8 + (int64_t)ntiles * sizeof(TileDescFull);
    if (GEO) { td->tconn.release(); td->tmask.release(); }     // the numeric kernel reads the geometry blocks instead
    td->complete = true;
}

// next smaller tile size to try after `te` needed `smem` bytes against a budget of `budget`
template <class F> static int tl_next_tile_elems(int te, int smem, int budget)
{
    const double scale = (budget > 0 && smem > budget) ? (double)budget / smem : 0.97;
    int next = 32;
    if (F::SPLIT) {
        for (int c : TL_TILE_SIZES) {
            if (c >= te) continue;
            const double rounds = (double)c * tl_items_1b<F>() / tl_block<F>();
            if (rounds / ceil(rounds) < 0.85 && c > 32) continue;
            next = c;
            if (c <= te * scale * 1.03) break;      // (else: this candidate would overflow as well)
        }
    } else {
        next = ((int)(te * (scale < 0.97 ? scale * 0.99 : 0.97))) & ~7;     // the footprint scales with te
        if (next < 32) next = 32;
    }
    return next;
}

// phase B with the tile-size search: the largest tile whose shared-memory footprint lets two CTAs share an SM
template <class F> static void tiled_tiles_search(efg_ctx *ctx)
{
    if (!tiled_sym(ctx)) efg_throw(EFG_ERR_STATE, "tile phase without a pattern phase");
    tiled_order<F>(ctx);
    ctx->tl_smem_budget = tl_smem_budget<F>(ctx);
    if (ctx->opt_tile_elems > 0) { tiled_tiles<F>(ctx, ctx->opt_tile_elems); return; }
    // start from the size the previous symbolic phase of this form settled on (re-assembly after efg_set_mesh, time
    // stepping with a changing mesh): the search below then succeeds at the first attempt
    const int vkind = ctx->mesh[0].kind * 100 + ctx->space[0].fe * 10 + ctx->space[2].fe;      // (mesh kind and the spaces' elements)
    const bool hinted = ctx->te_hint > 0 && ctx->te_hint_form == ctx->form_req && ctx->te_hint_kind == vkind && ctx->te_hint_quad == ctx->quad_req;
    int te = hinted ? ctx->te_hint : tl_default_tile_elems<F>();
    for (;;) {
        int smem = 0;
        try {
            tiled_tiles<F>(ctx, te);
            TiledData *td = tiled_data(ctx);
            smem = td->smem_bytes;
            if (td->complete && (te <= 32 || tl_ctas_per_sm<F>(ctx) >= tl_minb<F>())) {      // co-resident CTAs overlap each other's phases
                ctx->te_hint = te; ctx->te_hint_form = ctx->form_req; ctx->te_hint_kind = vkind; ctx->te_hint_quad = ctx->quad_req;
                return;
            }
        } catch (const EfgError &e) {
            if (e.code != EFG_ERR_LIMIT || te <= 32) throw;
        }
        te = tl_next_tile_elems<F>(te, smem, ctx->tl_smem_budget);
    }
}

template <class F> void tiled_symbolic(efg_ctx *ctx)
{
    if (!(ctx->have_pattern && tiled_sym(ctx))) tiled_pattern<F>(ctx);
    tiled_tiles_search<F>(ctx);
    delete tiled_sym(ctx);          // the pattern-phase tables are not needed once the tiles exist
    tiled_sym(ctx) = nullptr;
    ctx->scratch.release();         // (allocation is free with the arena: nothing is kept "warm" any more)
    // the values go where the sort scratch and the pattern-phase tables were
    const int64_t nnz = ctx->nnz;
    if (ctx->nzval.n < (size_t)(nnz > 0 ? nnz : 1)) ctx->nzval.alloc(ctx->pool, (size_t)(nnz > 0 ? nnz : 1));
    if (getenv("EFG_TRACE")) fprintf(stderr, "[efg trace] arena after the symbolic phase: %.2f GB in use (peak %.2f), %.2f GB reserved in %zu slabs\n",
                                     ctx->pool.bytes / 1e9, ctx->pool.peak / 1e9, ctx->pool.reserved / 1e9, ctx->pool.slabs.size());
}

template <class F, bool S, int BLOCK, int MINB> static void tl_launch_numeric_b(efg_ctx *ctx)
{
    TiledData *td = tiled_data(ctx);
    const MeshDev &gm = ctx->mesh[F::GMESH];
    int per_sm = 0, nsm = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
    if constexpr (tl_persist<F>()) {     // (compile-time: the persistent kernels are only instantiated when selected)
        auto kern = k_tl_numeric_p<F, S, BLOCK, MINB>;
        CUDA_CHECK(tl_raise_smem_limit(ctx, (const void *)kern));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, (size_t)td->smem_bytes));
        if (per_sm < 1) efg_throw(EFG_ERR_LIMIT, "tiled path: the numeric kernel does not fit on an SM (%d bytes of shared memory)", td->smem_bytes);
        const int grid = ctx->tl.ntiles < per_sm * nsm ? ctx->tl.ntiles : per_sm * nsm;      // one CTA per CTA slot
        if (grid > 0)
            LAUNCH(ctx, kern, (unsigned)grid, BLOCK, (size_t)td->smem_bytes, td->tiles.p, ctx->tl.ntiles, td->meta.p, td->geo.p, ctx->nzval.p,
                   td->off_meta, td->off_geo, td->off_gs);
    } else {
        auto kern = k_tl_numeric<F, S, BLOCK, MINB>;
        CUDA_CHECK(tl_raise_smem_limit(ctx, (const void *)kern));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, (size_t)td->smem_bytes));
        if (per_sm < 1) efg_throw(EFG_ERR_LIMIT, "tiled path: the numeric kernel does not fit on an SM (%d bytes of shared memory)", td->smem_bytes);
        const int grid = ctx->tl.ntiles;                    // one CTA per tile
        if (grid > 0)
            LAUNCH(ctx, kern, (unsigned)grid, BLOCK, (size_t)td->smem_bytes, td->tiles.p, ctx->tl.ntiles, td->tconn.p, td->tmask.p, gm.xy.p,
                   td->meta.p, ctx->nzval.p, per_sm * nsm, td->geo.p, (const int32_t *)nullptr, (double *)nullptr);
    }
    ctx->numeric_launches += 1;
}
template <class F, bool S> static void tl_launch_numeric(efg_ctx *ctx)
{
    tl_launch_numeric_b<F, S, tl_block<F>(), tl_minb<F>()>(ctx);
}

template <class F> void tiled_numeric(efg_ctx *ctx)
{
    if (ctx->opt_strict) tl_launch_numeric<F, true>(ctx); else tl_launch_numeric<F, false>(ctx);
}

// K and the load vector in one pass (efg_numeric_with_load): fout is the nrow-long vector, zeroed by the caller
template <class F, bool S> static void tl_launch_numeric_fused(efg_ctx *ctx, double *fout)
{
    if constexpr (tl_can_fuse<F>() && TL_GEO && !tl_persist<F>()) {
        TiledData *td = tiled_data(ctx);
        const MeshDev &gm = ctx->mesh[F::GMESH];
        constexpr int BLOCK = tl_block<F>(), MINB = tl_minb<F>();
        int per_sm = 0, nsm = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
        auto kern = k_tl_numeric<F, S, BLOCK, MINB, true>;
        CUDA_CHECK(tl_raise_smem_limit(ctx, (const void *)kern));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, (size_t)td->smem_bytes));
        if (per_sm < 1) efg_throw(EFG_ERR_LIMIT, "tiled path: the numeric kernel does not fit on an SM (%d bytes of shared memory)", td->smem_bytes);
        if (ctx->tl.ntiles > 0)
            LAUNCH(ctx, kern, (unsigned)ctx->tl.ntiles, BLOCK, (size_t)td->smem_bytes, td->tiles.p, ctx->tl.ntiles, td->tconn.p, td->tmask.p, gm.xy.p,
                   td->meta.p, ctx->nzval.p, per_sm * nsm, td->geo.p, ctx->rowval.p, fout);
        ctx->numeric_launches += 1;
    } else {
        (void)fout;
        efg_throw(EFG_ERR_INVALID, "the fused load vector is available for the heat forms on the tiled path");
    }
}
template <class F> void tiled_numeric_fused(efg_ctx *ctx, double *fout)
{
    if (ctx->opt_strict) tl_launch_numeric_fused<F, true>(ctx, fout); else tl_launch_numeric_fused<F, false>(ctx, fout);
}
