// efg_ctx.cuh -- the assembler context: device-resident mesh, dof maps, symbolic data, result.
#pragma once
#include "efg_common.cuh"
#include "efg_forms.cuh"

// global column -> local column of this ctx (owner-computes sharding), -1 if not owned
struct ColMap {
    int64_t c0, c1;
    int nr;
    const int32_t *first, *last1, *off;
    __device__ __forceinline__ int64_t local(int64_t c) const
    {
        if (nr == 0) return (c >= c0 && c < c1) ? c - c0 : -1;
        int lo = 0, hi = nr - 1;
        if (c < first[0]) return -1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (first[mid] <= c) lo = mid; else hi = mid - 1; }
        return c < last1[lo] ? (int64_t)off[lo] + (c - first[lo]) : -1;
    }
};

struct MeshDev {
    int kind = 0;
    int64_t nel = 0, nnodes = 0;
    DevBuf<int32_t> conn;   // nen x nel, 0-based
    DevBuf<double2> xy;     // nnodes
    DevBuf<double> z;       // nnodes, 3-D meshes only (EFG_T4)
    const double *pending_xy = nullptr;   // EFG_OPT_DEFER_XY: the caller's coordinates, not copied yet
    int nen() const { return kind == EFG_T4 ? 4 : kind; }
};

struct SpaceDev {
    int mesh = -1, ncomp = 0;
    int64_t nnodes = 0;
    DevBuf<int32_t> dof;    // ncomp x nnodes, 0-based (-1 = dof number 0 = unnumbered)
    DevBuf<uint8_t> isdatum; // ncomp x nnodes: prescribed dofs (only for spaces made by efg_gen_space)
    // elements with a dof on the cell (efg_set_space_fe)
    int fe = 0;             // EFG_FE_H1 (the H1 element of the mesh), EFG_FE_T3_BUBBLE, EFG_FE_L2
    int64_t ncells = 0;
    DevBuf<int32_t> cdof;   // ncomp x ncells, 0-based: the dim-2 field's dof numbers (null: no cell field)
};

// --- two-pass path (element matrices to HBM, then segmented gather) ---------------------------
struct TwoPass {
    DevBuf<uint32_t> perm;       // sorted triplet ids (valid ones first)
    DevBuf<int64_t> seg_start;   // nnz+1 offsets into perm
    DevBuf<double> Ke;           // NT x nel (entry k of element e at Ke[k*nel + e])
    int64_t ntrip_valid = 0;
};

// --- tiled fused path (data lives in efg_tiled.cuh's TiledData behind tl_opaque) -------------------
struct TileRun {                // consecutive tile slots that are contiguous in nzval
    int64_t nz0;                // destination in nzval
    int32_t s0, len;            // tile-local first slot, length
};
struct Tiled {
    int ntiles = 0;
    int tile_elems = 0;          // elements per tile
    int64_t sum_tile_elems = 0;  // tile elements incl. halo, summed over tiles
    int max_nq = 0, max_nslot = 0, max_nelem = 0, max_nrun = 0;
    int64_t numeric_bytes = 0;
    bool fused_rows = false;     // the tiling reserves stage rows for the fused load vector (EFG_OPT_FUSE_LOAD)
};

struct efg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evn0 = nullptr, evn1 = nullptr;
    cudaEvent_t ev_tab = nullptr;          // after this ctx's last upload of / launch reading the __constant__ tables (TabGuard)
    bool tab_event_recorded = false;
    std::string err;
    DevPool pool;

    MeshDev mesh[2];
    SpaceDev space[3];

    bool started = false;
    int64_t nrow = 0, ncol = 0;
    // owned matrix columns: one range [c0, c1) (0-based) or a sorted list of disjoint ranges
    int64_t c0 = 0, c1 = 0;
    int64_t ncl = 0;                 // number of owned (local) columns
    bool have_range = false;
    int nranges = 0;                 // 0 = single range
    DevBuf<int32_t> rfirst, rlast1, roff;

    // options
    int opt_path = 0;
    int opt_strict = 0;
    int opt_tile_elems = 0;
    int opt_sfc = 1;
    int opt_fuse_load = 0;
    int opt_defer_xy = 0;
    int opt_host_widen = -1;

    // symbolic state
    bool have_symbolic = false;
    int form = 0, quad = 0, nq = 0, vkind = 0;
    int path = 0;
    int64_t nnz = 0;
    DevBuf<int64_t> colptr;      // ncl+1, 1-based
    DevBuf<int32_t> rowval;      // nnz, 0-based
    DevBuf<double> nzval;        // nnz
    bool have_values = false;
    TwoPass tp;
    Tiled tl;
    void *tl_opaque = nullptr;
    void *tl_sym_opaque = nullptr;       // efg_tiled.cuh: pattern-phase data waiting for the tile phase
    bool have_pattern = false;           // colptr / rowval / nnz are valid (the tiles may still be pending)
    cudaEvent_t ev_pattern = nullptr;    // recorded on `stream` when colptr / rowval are complete
    cudaStream_t copy_stream = nullptr;  // device -> host copies that overlap the rest of the symbolic / numeric phase
    cudaEvent_t ev_copy = nullptr;
    bool copy_pending = false;
    bool widen_failed = false;
    std::vector<cudaEvent_t> widen_events;   // one per chunk of a pattern fetch, created once (their creation costs ~0.4 ms each)
    cudaStream_t in_stream = nullptr;        // deferred coordinate copies (overlap the pattern kernels)
    cudaEvent_t ev_xy = nullptr;
    bool xy_in_flight = false;
    void *mailbox = nullptr, *mailbox_dev = nullptr;    // page-locked mapped words for small read-backs (tl_read)
    void *widen = nullptr;               // HostWiden job of a pattern fetch into a host array (efg_hostcopy.cuh)
    DevBuf<int64_t> cstage[2];           // rowval Int32 -> Int64 staging of the copy stream
    int tl_smem_budget = 0;              // dynamic shared memory per CTA that still lets two CTAs share an SM
    int form_req = 0, quad_req = 0;      // form / rule of the symbolic phase in progress
    int te_hint = 0, te_hint_form = 0, te_hint_kind = 0, te_hint_quad = 0;   // tile size the last symbolic phase settled on
    void *vec_opaque = nullptr;  // efg_vector.cuh: system-vector assembly, K*x, sub-blocks
    DevBuf<char> scratch;        // persistent scratch for the largest symbolic temporaries (kept across calls: the
                                 // multi-GB sort buffers made cudaMallocAsync stall for 0.1-2.5 s when re-allocated every call)

    // stats
    double symbolic_ms = 0, numeric_ms = 0;
    int64_t launches = 0, numeric_launches = 0;
};

#define COLMAP(ctx) ColMap{(ctx)->c0, (ctx)->c1, (ctx)->nranges, (ctx)->rfirst.p, (ctx)->rlast1.p, (ctx)->roff.p}

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                             \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);             \
        (ctx)->launches++;                                                           \
        CUDA_CHECK(cudaGetLastError());                                              \
    } while (0)

// (form, element kind, quadrature rule) -> template instantiation
template <class Fn> inline bool dispatch_form(int form, int vkind, int nq, Fn &&fn)
{
    switch (form) {
    case EFG_FORM_HEAT:
        if (vkind == 3 && nq == 1) { fn(HeatForm<3, 1>{}); return true; }
        if (vkind == 3 && nq == 3) { fn(HeatForm<3, 3>{}); return true; }
        if (vkind == 6 && nq == 1) { fn(HeatForm<6, 1>{}); return true; }
        if (vkind == 6 && nq == 3) { fn(HeatForm<6, 3>{}); return true; }
        if (vkind == 4 && nq == 1) { fn(HeatForm<4, 1>{}); return true; }
        if (vkind == 4 && nq == 4) { fn(HeatForm<4, 4>{}); return true; }
        if (vkind == 4 && nq == 9) { fn(HeatForm<4, 9>{}); return true; }
        // the less common rules: one instantiation per element kind with the number of points read at run time
        if (vkind == 3 && (nq == 4 || nq == 6 || nq == 7 || nq == 9 || nq == 12 || nq == 13)) { fn(HeatForm<3, 0>{}); return true; }
        if (vkind == 6 && (nq == 4 || nq == 6 || nq == 7 || nq == 9 || nq == 12 || nq == 13)) { fn(HeatForm<6, 0>{}); return true; }
        if (vkind == 4 && (nq == 16 || nq == 25)) { fn(HeatForm<4, 0>{}); return true; }
        if (vkind == EFG_T4 && nq == 1) { fn(HeatFormT4<1>{}); return true; }
        if (vkind == EFG_T4 && nq == 4) { fn(HeatFormT4<4>{}); return true; }
        if (vkind == EFG_T4 && nq == 5) { fn(HeatFormT4<5>{}); return true; }
        return false;
    case EFG_FORM_ELASTICITY:
        if (vkind == 3 && nq == 1) { fn(ElasticityForm<3, 1>{}); return true; }
        if (vkind == 3 && nq == 3) { fn(ElasticityForm<3, 3>{}); return true; }
        if (vkind == 6 && nq == 3) { fn(ElasticityForm<6, 3>{}); return true; }
        if (vkind == 4 && nq == 4) { fn(ElasticityForm<4, 4>{}); return true; }
        return false;
    case EFG_FORM_STOKES_GEN:
        if (vkind == 6 && nq == 3) { fn(Stokes2Form<false>{}); return true; }
        return false;
    case EFG_FORM_STOKES_VECLAP_ALT:
        if (vkind == 6 && nq == 3) { fn(Stokes2Form<true>{}); return true; }
        return false;
    // the three-space forms: vkind is the velocity / pressure PAIR (form_kind(): the mesh kind for the H1 pairs)
    case EFG_FORM_STOKES_REDDY:
        if (vkind == EFG_PAIR_T6_T3 && nq == 3) { fn(Stokes3Form<false>{}); return true; }
        if (vkind == EFG_PAIR_T3B_T3 && nq == 3) { fn(Stokes3Form<false, EFG_PAIR_T3B_T3, 3>{}); return true; }
        if (vkind == EFG_PAIR_Q4_L2 && nq == 4) { fn(Stokes3Form<false, EFG_PAIR_Q4_L2, 4>{}); return true; }
        if (vkind == EFG_PAIR_T3_L2 && nq == 3) { fn(Stokes3Form<false, EFG_PAIR_T3_L2, 3>{}); return true; }
        return false;
    case EFG_FORM_STOKES_VECLAP:
        if (vkind == EFG_PAIR_T6_T3 && nq == 3) { fn(Stokes3Form<true>{}); return true; }
        if (vkind == EFG_PAIR_T3B_T3 && nq == 3) { fn(Stokes3Form<true, EFG_PAIR_T3B_T3, 3>{}); return true; }
        if (vkind == EFG_PAIR_Q4_L2 && nq == 4) { fn(Stokes3Form<true, EFG_PAIR_Q4_L2, 4>{}); return true; }
        if (vkind == EFG_PAIR_T3_L2 && nq == 3) { fn(Stokes3Form<true, EFG_PAIR_T3_L2, 3>{}); return true; }
        return false;
    }
    return false;
}

// load the coordinates of the geometry carrier's nodes of element e (global mesh arrays)
template <int GK>
__device__ __forceinline__ void load_xy(const int32_t *__restrict__ conn, const double2 *__restrict__ xy,
                                        int64_t e, double (&X)[GK], double (&Y)[GK])
{
#pragma unroll
    for (int a = 0; a < GK; a++) {
        const double2 p = __ldg(&xy[conn[e * GK + a]]);
        X[a] = p.x; Y[a] = p.y;
    }
}
