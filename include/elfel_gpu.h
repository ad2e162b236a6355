/*
 * elfel_gpu.h -- C ABI of libelfelgpu.so, the B200 (sm_100a) drop-in for Elfel.jl's assembly
 * hot path.  This is exactly what a Julia `SysmatAssemblerGPU <: AbstractSysmatAssembler`
 * binds with `ccall` (see INTEGRATION.md for the .jl shim); the same entry points are driven
 * from Python ctypes in tests/ and bench.py because no Julia runtime exists in this image.
 *
 * Reference interface each entry point replaces (paths relative to the Elfel.jl tree):
 *   efg_create / efg_destroy  SysmatAssemblerSparse(0.0) constructor     src/Assemblers.jl:58-60
 *   efg_set_mesh              FEIterator ctor caching _bir and _geom     src/FEIterators.jl:54-56
 *   efg_set_space             FEIterator's _fld0.dofnums (FEField)       src/FEIterators.jl:70, src/FEFields.jl:15
 *   efg_start                 start!(ass, nrow, ncol)                    src/Assemblers.jl:67-78
 *   efg_assemble              the user's integrate! loop: `for el in elit; init!(ke,...);
 *                             for qp in qpit ... end; assemble!(ass, ke) end`
 *                                                                        examples/heat/poisson/t3.jl:41-64,
 *                                                                        examples/elasticity/stretch/t6.jl:40-63,
 *                                                                        examples/stokes/colliding_flow/ht_p2_p1_gen.jl:46-90,
 *                                                                        ht_p2_p1.jl:55-113, ht_p2_p1_veclap.jl:55-106,
 *                                                                        ht_p2_p1_veclap_alt.jl:60-98
 *                             + assemble!(ass, lma) / transpose(lma)     src/Assemblers.jl:97-114
 *   efg_fetch_csc             finish!(ass) -> sparse(I,J,V,m,n)          src/Assemblers.jl:121-123
 *   efg_vec_assemble          SysvecAssembler start!(av, nrow) + the `init!(fe, eldofs(el)); fe[j] += N[j]*Q*JxW;
 *                             assemble!(av, fe)` half of the same integrate! loop
 *                                                                        src/Assemblers.jl:196-223, src/LocalAssemblers.jl:95-152,
 *                                                                        examples/heat/poisson/t3.jl:44-61, q4.jl:34-51
 *   efg_fetch_vec             finish!(av)                                src/Assemblers.jl:230-232
 *   efg_qp_locations          location(el, qp) for every element / quadrature point   src/FEIterators.jl:227-235
 *   efg_l2_error              evaluate_pressure_error / evaluate_velocity_error      examples/stokes/colliding_flow/ht_p2_p1.jl:120-178,
 *                                                                        ht_p2_p1_gen.jl:124-153, test/test_stokes.jl:438-496
 *   efg_spmv                  `KT = K * T` right after assembly          examples/heat/poisson/t3.jl:78
 *   efg_block_nnz/_fetch_block  `K[1:nu, 1:nu]`, `K[1:nu, nu+1:end]`     examples/heat/poisson/t3.jl:79,
 *                                                                        examples/stokes/colliding_flow/ht_p2_p1_gen.jl (solve!)
 *
 * Conventions
 *   - Plain pointers and sizes only.  All index arrays are Int64 and 1-BASED, in the memory
 *     layout Julia holds them: conn is nen x nel (node ids of element e at conn[e*nen + k]),
 *     xy is 2 x nnodes, dofnums is ncomp x nnodes (Vector{SVector{ncomp,Int64}}).
 *   - Input pointers may be host (pageable or pinned) or device pointers (unified addressing);
 *     they are borrowed for the duration of the call only.  Device inputs are read on the ctx's own
 *     stream: the work that produced them must have completed (synchronize the producer first).
 *   - Output arrays of efg_fetch_csc are allocated by the caller after nnz is known
 *     (two-call pattern) so Julia wraps them in SparseMatrixCSC without a copy.
 *   - Every function returns 0 on success or a negative EFG_ERR_* code; no exception crosses
 *     the boundary.  efg_last_error() gives the message.  A dof number < 1 or > nrow/ncol gives
 *     EFG_ERR_INDEX, mirroring sparse()'s ArgumentError (e.g. a space that was never
 *     data-numbered holds dof number 0, src/FEFields.jl:148).
 *   - One ctx = one device + one CUDA stream; not re-entrant.  Different ctx may be used from
 *     different host threads.  Multi-GPU: either one ctx per GPU driven by the caller (one process or thread per
 *     GPU, each assembling its column ranges: efg_set_column_ranges), or one efg_multi handle (efgm_* below) that
 *     takes the global arrays and shards internally; no communication is needed in either case.
 *   - The quadrature tables and form parameters live in __constant__ memory (one copy per device).  The library
 *     hands them over between the ctx of a device in stream order (per-device lock + events), so several ctx --
 *     on one device or one per device, from one host thread or several -- may be used freely and asynchronously.
 *   - Device memory comes from a private arena per ctx (plain cudaMalloc slabs, sub-allocated on the host); no
 *     process-global allocator state is touched, efg_destroy returns everything to the driver.
 *   - There is no CPU fallback: without a CUDA device efg_create fails with EFG_ERR_CUDA.
 */
#ifndef ELFEL_GPU_H
#define ELFEL_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct efg_ctx efg_ctx;

/* error codes */
#define EFG_OK            0
#define EFG_ERR_INVALID  -1  /* bad argument / unsupported element-form-rule combination */
#define EFG_ERR_CUDA     -2  /* CUDA runtime error (incl. no device) */
#define EFG_ERR_OOM      -3  /* device or host allocation failed */
#define EFG_ERR_INDEX    -4  /* node / dof index out of range (sparse()'s ArgumentError) */
#define EFG_ERR_STATE    -5  /* call order violated (e.g. assemble before start) */
#define EFG_ERR_LIMIT    -6  /* an internal capacity was exceeded (message says which) */

/* element kinds (= nodes per element): FEH1_T3, FEH1_Q4, FEH1_T6 (src/FElements.jl:225-320) */
#define EFG_T3 3
#define EFG_Q4 4
#define EFG_T6 6
/* FEH1_T4, the linear tetrahedron (src/FElements.jl:359-386; examples/heat/poisson/t4.jl): 4 nodes, 3-D coordinates.
 * The only kind code that is not its node count; meshes of this kind are set with efg_set_mesh3. */
#define EFG_T4 40

/* SURVEY 8f row f5 -- finite elements with a dof on the cell itself, for efg_set_space_fe:
 *   EFG_FE_H1         the H1 element of the space's mesh (FEH1_T3 / FEH1_Q4 / FEH1_T6): vertex dofs only
 *   EFG_FE_T3_BUBBLE  FEH1_T3_BUBBLE (src/FElements.jl:324-355): vertex dofs + one dof on the cell (the cubic bubble)
 *   EFG_FE_L2         FEL2_T3 / FEL2_Q4 (src/FElements.jl:394-448): one dof on the cell, no vertex dofs; the geometry
 *                     carrier is the H1 element of the mesh (_geometrycarrier) */
#define EFG_FE_H1         0
#define EFG_FE_L2         1
#define EFG_FE_T3_BUBBLE  7

/* weak forms (the integrate! closures of the reference's examples/tests) */
#define EFG_FORM_HEAT               1 /* space 0: scalar.        params = [kappa]                  */
#define EFG_FORM_ELASTICITY         2 /* space 0: 2 components.  params = D 3x3 column-major (9)   */
#define EFG_FORM_STOKES_GEN         3 /* space 0: u (T6,2), 1: p (T3).  params = D (9)             */
#define EFG_FORM_STOKES_REDDY       4 /* space 0: ux, 1: uy (T6), 2: p (T3).  params = [mu]        */
                                     /* ... and, on ONE mesh (slot 0), the pairs of examples/stokes/colliding_flow/p1b_p1.jl
                                        (ux, uy: EFG_FE_T3_BUBBLE, p: FEH1_T3, npts 3), q1_q0.jl (ux, uy: FEH1_Q4,
                                        p: EFG_FE_L2, Gauss order 2) and FEH1_T3 / EFG_FE_L2 (npts 3); same for _VECLAP */
#define EFG_FORM_STOKES_VECLAP_ALT  5 /* space 0: u (T6,2), 1: p (T3).  params = [mu]              */
#define EFG_FORM_STOKES_VECLAP      6 /* space 0: ux, 1: uy (T6), 2: p (T3).  params = [mu]        */

/* vector (right-hand side) forms */
#define EFG_VFORM_HEAT_LOAD         1 /* space 0: scalar.  fe[j] += N[j]*Q*JxW, params = [Q]            */

/* options for efg_set_option */
#define EFG_OPT_PATH        1 /* 0 = auto (tiled fused kernel; meshes beyond its limits, e.g. a node shared by
                                 > 90 elements, fall back to the two-pass CUDA path), 1 = two-pass (element
                                 matrices to HBM, then segmented gather), 2 = tiled fused kernel only */
#define EFG_OPT_STRICT_FP   2 /* 1 = no FMA contraction: operation order and rounding of the
                                 reference's expressions (bit-identical to the CPU oracle) */
#define EFG_OPT_TILE_ELEMS  3 /* elements per tile of the fused kernel (0 = automatic) */
#define EFG_OPT_SFC_ORDER   4 /* 1 (default) = tiles follow a space-filling-curve order of the
                                 elements, 0 = tiles follow the given element order */
#define EFG_OPT_DEFER_XY    6 /* 1 = efg_set_mesh does not copy a HOST coordinate array: the array is borrowed until the next
                                 efg_pattern / efg_symbolic / efg_assemble call returns (or any later call on the ctx), where it
                                 is copied on a separate stream while the pattern kernels run (they read connectivity and dof
                                 maps only).  For callers that keep the mesh alive across the whole assemble! (the Julia shim) */
#define EFG_OPT_HOST_WIDEN  7 /* row indices fetched into a HOST array: 1 = sent as the device's Int32 and widened in place by library
                                 threads (half the PCIe bytes; best for 1-2 GPUs per host), 0 = widened on the device (best when many
                                 ranks share one host's memory system), -1 (default) = 1 if at most two devices are visible */
#define EFG_OPT_FUSE_LOAD   5 /* 1 = the next symbolic phase of a heat form reserves room for the element load vector next
                                 to the element matrix, so that efg_numeric_with_load can produce K and F in one pass */

/* statistics for efg_get_stat (milliseconds are device times from CUDA events on the ctx stream) */
#define EFG_STAT_SYMBOLIC_MS      1
#define EFG_STAT_NUMERIC_MS       2
#define EFG_STAT_KERNEL_LAUNCHES  3 /* kernels launched by the library so far (all phases) */
#define EFG_STAT_NUMERIC_LAUNCHES 4 /* kernels launched by the last efg_numeric call */
#define EFG_STAT_DEVICE_BYTES     5 /* device memory currently held by the ctx */
#define EFG_STAT_NTILES           6
#define EFG_STAT_TILE_ELEMS       7 /* sum over tiles of elements processed (incl. halo) */
#define EFG_STAT_NUMERIC_BYTES    8 /* bytes the numeric kernel is designed to move per call */
#define EFG_STAT_PATH             9 /* path used by the last symbolic phase (1 or 2) */
#define EFG_STAT_VEC_MS          10 /* device time of the last efg_vec_assemble numeric part (2 kernels) */
#define EFG_STAT_SPMV_MS         11 /* device time of the last efg_spmv kernel */

/* efg_create also brings the library's device module in (tens of ms), so that the first assembly does not pay for it. */
int efg_create(int device, efg_ctx **out);
int efg_destroy(efg_ctx *ctx);
const char *efg_last_error(const efg_ctx *ctx);
int efg_set_option(efg_ctx *ctx, int option, int64_t value);
int efg_get_stat(efg_ctx *ctx, int which, double *out);
/* the CUDA stream (cudaStream_t) all work of this ctx is issued on */
int efg_get_stream(efg_ctx *ctx, void **stream_out);
int efg_synchronize(efg_ctx *ctx);

/* mesh_slot 0: the mesh of space 0 (Stokes: velocity mesh); mesh_slot 1: Stokes pressure mesh. */
int efg_set_mesh(efg_ctx *ctx, int mesh_slot, int elemkind, int64_t nel, int64_t nnodes,
                 const int64_t *conn, const double *xy);
/* The same for a 3-D mesh (row f5: EFG_T4): conn is 4 x nel, xyz is 3 x nnodes (the "geom" attribute of SVector{3}).
 * Available for EFG_FORM_HEAT (triangle-style rule = npts 1, 4 or 5 of src/RefShapes.jl:232-259) and EFG_VFORM_HEAT_LOAD,
 * on both paths (tiled kernel: 3-D Morton tiles, geometry blocks with a third coordinate plane). */
int efg_set_mesh3(efg_ctx *ctx, int mesh_slot, int elemkind, int64_t nel, int64_t nnodes,
                  const int64_t *conn, const double *xyz);
/* space_slot 0..2, living on mesh_slot; dofnums is ncomp x nnodes. */
int efg_set_space(efg_ctx *ctx, int space_slot, int mesh_slot, int ncomp, int64_t nnodes,
                  const int64_t *dofnums);

/* The same for a space whose element carries a dof on the cell (row f5): `fe` = EFG_FE_*; dofnums (ncomp x nnodes) are the
 * dof numbers of the dim-0 field, cell_dofnums (ncomp x nel) those of the dim-2 field -- FESpace._irsfields[0] / [2]
 * (src/FESpaces.jl:75-85), element dof order: vertex dofs first, then the cell's (src/FEIterators.jl:185-194).
 * EFG_FE_L2: nnodes = 0 / dofnums = NULL; EFG_FE_H1: nel = 0 / cell_dofnums = NULL (= efg_set_space). */
int efg_set_space_fe(efg_ctx *ctx, int space_slot, int mesh_slot, int fe, int ncomp, int64_t nnodes, const int64_t *dofnums,
                     int64_t nel, const int64_t *cell_dofnums);

/* start!(ass, nrow, ncol): resets the assembler; mesh/space data are kept. */
int efg_start(efg_ctx *ctx, int64_t nrow, int64_t ncol);
/* Owner-computes sharding: assemble only columns col_first..col_last (1-based, inclusive).
 * Triplets of other columns are dropped; colptr then has (col_last-col_first+2) entries, rebased
 * to start at 1.  Concatenating the blocks of consecutive ranges gives the global matrix. */
int efg_set_column_range(efg_ctx *ctx, int64_t col_first, int64_t col_last);
/* Same, for an owner that holds several disjoint column ranges (ascending, 1-based inclusive), e.g. the
 * vertex-dof rows and the mid-side-dof runs of one horizontal band of a T6 mesh.  The output holds the owned
 * columns in ascending order: colptr has (number of owned columns + 1) entries. */
int efg_set_column_ranges(efg_ctx *ctx, int64_t nranges, const int64_t *col_firsts, const int64_t *col_lasts);

/* Symbolic phase: CSC pattern + scatter maps on the device.  quad_rule = the QPIterator settings: triangles npts
 * (1 | 3 for every form; 4 | 6 | 7 | 9 | 12 | 13 for EFG_FORM_HEAT / EFG_VFORM_HEAT_LOAD / locations / error norms:
 * src/RefShapes.jl:113-230), squares Gauss order (1..3; 4 | 5 for the heat forms: src/RefShapes.jl:85-110, 333-366),
 * tetrahedra npts (1 | 4 | 5: src/RefShapes.jl:232-259).  Cached until mesh/space/start/range/options change. */
int efg_symbolic(efg_ctx *ctx, int form_id, int quad_rule, int64_t *nnz_out);
/* First half of efg_symbolic only: the CSC pattern (colptr / rowval / nnz); the scatter maps are built by the next
 * efg_symbolic / efg_numeric / efg_assemble call.  For callers that overlap: efg_pattern -> allocate the three output
 * arrays (nnz known) -> efg_fetch_pattern_async -> efg_numeric -> efg_fetch_csc(NULL, NULL, nzval).
 * Replaces what finish! gets from sparse() (src/Assemblers.jl:121-123) in two steps: structure first, values later. */
int efg_pattern(efg_ctx *ctx, int form_id, int quad_rule, int64_t *nnz_out);
/* Numeric phase: element quadrature loop fused with the deterministic scatter -> nzval (device). */
int efg_numeric(efg_ctx *ctx, const double *params, int nparams);
/* K and the load vector F of ONE integrate! pass -- the reference's heat loops compute `ke[i,j] += ...` and
 * `fe[j] += N[j]*Q*JxW` in the same quadrature loop and call assemble!(am, ke); assemble!(av, fe) per element
 * (examples/heat/poisson/t3.jl:41-64, q4.jl:31-54).  Needs EFG_FORM_HEAT on the tiled path, an unsharded ctx and
 * EFG_OPT_FUSE_LOAD = 1 at symbolic time; params = [kappa].  F (length nrow; zero for dofs no element touches) is then
 * read with efg_fetch_vec / efg_device_vec like the result of efg_vec_assemble, and is bit-identical to it. */
int efg_numeric_with_load(efg_ctx *ctx, const double *params, int nparams, double Q);
/* symbolic (if not cached) + numeric. */
int efg_assemble(efg_ctx *ctx, int form_id, int quad_rule, const double *params, int nparams,
                 int64_t *nnz_out);

/* finish!: copy out the SparseMatrixCSC fields (Int64 1-based colptr/rowval, Float64 nzval).
 * Any pointer may be NULL to skip that array.  Destination may be host or device memory. */
int efg_fetch_csc(efg_ctx *ctx, int64_t *colptr, int64_t *rowval, double *nzval);
/* Starts copying colptr / rowval (either may be NULL) on the ctx's copy stream and returns at once; the copy overlaps
 * whatever the ctx does next (tile phase, numeric kernel).  The arrays are complete -- and must stay valid until --
 * the next efg_fetch_csc call on this ctx (any arguments; it waits for the copy) or efg_destroy.
 * A HOST rowval array receives the device's Int32 indices unwidened (half the PCIe traffic) in its upper half and a few
 * library threads widen them in place to Int64 1-based while the rest is still travelling (EFG_HOST_THREADS overrides
 * their number: default min(8, cores / visible GPUs), at least 2); page-locked arrays copy at link speed. */
int efg_fetch_pattern_async(efg_ctx *ctx, int64_t *colptr, int64_t *rowval);
/* Device-resident result for a consumer that stays on the GPU (colptr: Int64 1-based,
 * rowval: Int32 0-based, nzval: Float64); valid until the next start/symbolic/destroy. */
int efg_device_csc(efg_ctx *ctx, const int64_t **colptr, const int32_t **rowval, const double **nzval);

/* System VECTOR assembly (SysvecAssembler).  One call = start!(av, nrow) + the element loop of the vector form over
 * mesh 0 / space 0 (set by efg_set_mesh / efg_set_space) + assemble!(av, fe) per element.  Contributions reach
 * val[gi] in the reference's order (ascending element, local index ascending) and every operation is individually
 * rounded, so the result is bit-identical to the CPU loop.  A dof number < 1 or > nrow gives EFG_ERR_INDEX
 * (Julia: BoundsError).  With a column range set on an nrow x nrow system only the owned rows are assembled
 * (owner-computes, same ranges as the matrix columns).  The dof -> contribution map is cached like the matrix
 * pattern. */
int efg_vec_assemble(efg_ctx *ctx, int vform, int quad_rule, const double *params, int nparams, int64_t nrow);
/* finish!(av): nrow doubles (owned rows when sharded), host or device destination. */
int efg_fetch_vec(efg_ctx *ctx, double *out);
int efg_device_vec(efg_ctx *ctx, const double **val, int64_t *n);

/* y = K*x with the assembled matrix, in SparseArrays' accumulation order (y[r] sums its terms by ascending column,
 * product and sum rounded separately): bit-identical to Julia's `K * x`.  x: ncol doubles, y: nrow doubles, host or
 * device.  The row-major view of the pattern is built at the first call and cached until the pattern changes. */
int efg_spmv(efg_ctx *ctx, const double *x, double *y);
/* K[row_first:row_last, col_first:col_last] (1-based inclusive, Julia range semantics; an empty range is
 * first = last+1) as a SparseMatrixCSC: two-call pattern like efg_fetch_csc.  efg_block_nnz fixes the block and
 * returns its stored-entry count; efg_fetch_block fills colptr (ncols_block+1), rowval (rebased to the block) and
 * nzval.  Any pointer may be NULL. */
int efg_block_nnz(efg_ctx *ctx, int64_t row_first, int64_t row_last, int64_t col_first, int64_t col_last, int64_t *nnz_out);
int efg_fetch_block(efg_ctx *ctx, int64_t *colptr, int64_t *rowval, double *nzval);

/* Post-processing integrators of the Stokes examples.  The true-solution closures (truep, trueux, trueuy) stay on the
 * caller's side: efg_qp_locations returns location(el, qp) for every element and quadrature point of a mesh
 * (out: 2 x npts x nel doubles, point q of element e at out[(e*npts + q)*2]; out may be NULL to query npts), the caller
 * evaluates its functions there, and efg_l2_error integrates
 *     sqrt( sum_el sum_qp JxW * sum_c ( sum_j U[eldofs_c[j]] * N_j(qp)  -  truth[(e*npts + q)*ncomp + c] )^2 )
 * for ncomp = 1 or 2 field components given as (space_slot, component) pairs on one mesh -- evaluate_pressure_error:
 * {(p, 0)}; evaluate_velocity_error: {(ux, 0), (uy, 0)} or {(u, 0), (u, 1)}.  U is the system vector (what solve!
 * returns; eldofvals(el) = U[eldofs(el)] after scattersysvec!), nU its length.  Per-element sums follow the reference's
 * operation order; they are added by a fixed-shape tree, so the value is reproducible and equals the CPU loop's running
 * sum to rounding (1e-12 relative in the tests). */
int efg_qp_locations(efg_ctx *ctx, int mesh_slot, int quad_rule, double *out, int64_t *npts_out);
int efg_l2_error(efg_ctx *ctx, int ncomp, const int *space_slots, const int *comps, int quad_rule, const double *U, int64_t nU,
                 const double *truth, double *out);

/* library / build information, e.g. "elfelgpu 0.1 sm_100a" */
const char *efg_version(void);

/* ---- the inputs made on the device (no host arrays, nothing crosses PCIe) ------------------------------
 * efg_gen_mesh: MeshSteward's T3block (orientation :a) / Q4block / T6block(Length, Width, nL, nW) as the examples call them
 *   (examples/heat/poisson/t3.jl:34, q4.jl:24, examples/elasticity/stretch/t6.jl:34): nodes x-fastest, elements i outer /
 *   j inner (test/qmesh-conn.dat), T6: corner nodes first, mid-side nodes in first-encounter order; then every coordinate
 *   is shifted by (xshift, yshift) (the Stokes examples' transform(ir, x -> x - A)).
 * efg_gen_mesh_corners: T6toT3 (examples/stokes/colliding_flow/ht_p2_p1_gen.jl:42): the pressure mesh.
 * efg_gen_space / efg_setebc_box / efg_setebc_nodes / efg_number_dofs: FESpace(mesh, fe, ncomp), setebc! on the nodes of
 *   vselect(geom; box = [x0 x1 y0 y1]) (pass the box already inflated) or on a node list (component 1..ncomp, 0 = all),
 *   numberdofs!(spaces): free dofs first, spaces in the order given, node-major / component-minor
 *   (src/FESpaces.jl:141-173,250-259, src/FEFields.jl:124-177).
 * efg_fetch_mesh / efg_fetch_dofnums: the device-resident inputs copied out in the reference's layout (Int64, 1-based),
 *   for any mesh / space of the ctx (generated or uploaded); NULL array pointers skip the copy (sizes only). */
int efg_gen_mesh(efg_ctx *ctx, int mesh_slot, int elemkind, int64_t nL, int64_t nW, double Length, double Width, double xshift, double yshift);
int efg_gen_mesh_corners(efg_ctx *ctx, int mesh_slot_dst, int mesh_slot_src);
int efg_gen_space(efg_ctx *ctx, int space_slot, int mesh_slot, int ncomp);
int efg_setebc_box(efg_ctx *ctx, int space_slot, int comp, double x0, double x1, double y0, double y1);
int efg_setebc_nodes(efg_ctx *ctx, int space_slot, int comp, int64_t n, const int64_t *node_ids);
int efg_number_dofs(efg_ctx *ctx, int nspaces, const int *space_slots, int64_t *nfree_out, int64_t *ndofs_out);
int efg_fetch_mesh(efg_ctx *ctx, int mesh_slot, int64_t *nel_out, int64_t *nnodes_out, int64_t *conn, double *xy);
int efg_fetch_dofnums(efg_ctx *ctx, int space_slot, int64_t *dofnums);

/* ---- several GPUs behind one handle (SURVEY 8b "efg_create_multi": same calls, the library shards internally; 8e) -------
 * The caller passes the GLOBAL arrays exactly as to the single-GPU functions above (what FEIterator holds:
 * src/FEIterators.jl:54-81); they are borrowed until efgm_assemble has returned (GC.@preserve around the whole
 * assemble!).  Inside, one host thread per device selects -- on the device -- the elements touching the nodes of its band
 * (halo elements replicated), owns the columns of that band's dofs and runs the single-GPU phases on its sub-mesh; there
 * is no communication between the devices.  efgm_fetch_csc interleaves the device blocks by column into the caller's
 * SparseMatrixCSC arrays; the result is bit-identical to the single-GPU result.  `devices` may be NULL (= 0 .. ngpu-1)
 * and may name a device more than once. */
typedef struct efg_multi efg_multi;
int efgm_create(int ngpu, const int *devices, efg_multi **out);
int efgm_destroy(efg_multi *m);
const char *efgm_last_error(const efg_multi *m);
int efgm_device_count(const efg_multi *m);
int efgm_set_option(efg_multi *m, int option, int64_t value);
int efgm_set_mesh(efg_multi *m, int mesh_slot, int elemkind, int64_t nel, int64_t nnodes, const int64_t *conn, const double *xy);
int efgm_set_space(efg_multi *m, int space_slot, int mesh_slot, int ncomp, int64_t nnodes, const int64_t *dofnums);
int efgm_start(efg_multi *m, int64_t nrow, int64_t ncol);
/* shards (first call after efgm_set_mesh / efgm_set_space), then symbolic + numeric on every device; nnz of the global matrix */
int efgm_assemble(efg_multi *m, int form_id, int quad_rule, const double *params, int nparams, int64_t *nnz_out);
/* numeric phase again on the cached shards and patterns (time stepping, Newton) */
int efgm_numeric(efg_multi *m, const double *params, int nparams);
/* global SparseMatrixCSC fields into HOST arrays (colptr: ncol+1, required; rowval / nzval may be NULL) */
int efgm_fetch_csc(efg_multi *m, int64_t *colptr, int64_t *rowval, double *nzval);
/* EFG_STAT_* over the devices: device < 0: milliseconds -> max, counts and bytes -> sum; device >= 0: that device only */
int efgm_get_stat(efg_multi *m, int which, int device, double *out);
/* the ctx of device index `device` (its column block stays on the device: efg_device_csc, efg_spmv is per block) and
 * the column ranges (1-based, inclusive, ascending) that block holds; valid until the next efgm_set_* / efgm_destroy */
int efgm_device_ctx(efg_multi *m, int device, efg_ctx **ctx_out, int64_t *nranges_out, const int64_t **firsts_out, const int64_t **lasts_out);

#ifdef __cplusplus
}
#endif
#endif /* ELFEL_GPU_H */
