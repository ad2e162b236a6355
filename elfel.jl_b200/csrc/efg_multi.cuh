// efg_multi.cuh -- several GPUs behind ONE handle ("efg_create_multi(ngpu, ...): same calls, the library shards
// internally"; 8e: owner-computes column blocks, halo elements replicated, no data-path communication).
//
// The caller hands over the GLOBAL mesh / dof arrays exactly as a single-GPU caller would.  Per device, one host thread
//   1. puts every node of every mesh into one of `ngpu` bands along the longer axis of the bounding box (cuts = quantiles of
//      the node coordinates of mesh 0, so the bands hold equal numbers of nodes),
//   2. selects the elements with at least one node in its band (a flag pass over the connectivity on the device: halo
//      elements end up on both sides of a cut), renumbers their nodes locally and keeps the GLOBAL dof numbers,
//   3. owns the matrix columns of the dofs of its band's nodes (sorted -> a list of column ranges),
//   4. runs the ordinary single-GPU symbolic + numeric phases on its sub-mesh with those column ranges.
// The global CSC is the column-wise interleave of the device blocks: efgm_fetch_csc builds the global colptr from the
// per-column counts and copies every device's column runs straight into their place in the caller's arrays.  Element order
// inside a sub-mesh is the global order, so every nonzero sums its contributions in the same order as on one GPU: the
// result is bit-identical to the single-ctx result (tests/test_gpu_multi.py).
#pragma once
#include <thread>
#include <functional>
#include "efg_ctx.cuh"

struct MultiDev {
    efg_ctx *ctx = nullptr;
    int band = 0;
    // owned column ranges (1-based inclusive, ascending) and the local column offset of each
    std::vector<int64_t> firsts, lasts;
    int64_t nel_local = 0, nnz = 0;
    int rc = EFG_OK;
    std::string err;
};

struct GlobalMesh { int kind = 0; int64_t nel = 0, nnodes = 0; const int64_t *conn = nullptr; const double *xy = nullptr; };
struct GlobalSpace { int mesh = -1, ncomp = 0; int64_t nnodes = 0; const int64_t *dofnums = nullptr; };

struct efg_multi {
    std::vector<MultiDev> dev;
    GlobalMesh mesh[2];
    GlobalSpace space[3];
    int64_t nrow = 0, ncol = 0;
    bool started = false, sharded = false, assembled = false;
    int64_t opt[8] = {0, 0, 0, 0, 1, 0, 0, 0};     // efg_set_option values forwarded to every ctx (index = option id)
    bool opt_set[8] = {false, false, false, false, false, false, false, false};
    int form = 0, quad = 0;
    int64_t nnz = 0;
    std::string err;
};

// ---- device kernels of the sharding pass ---------------------------------------------------------------------------
__global__ void k_mg_band(const double2 *__restrict__ xy, int64_t n, int axis, const double *__restrict__ cuts, int ncuts, uint8_t *__restrict__ band)
{
    GRID_STRIDE(i, n) {
        const double v = axis ? xy[i].y : xy[i].x;
        int b = 0;
        while (b < ncuts && v >= cuts[b]) b++;
        band[i] = (uint8_t)b;
    }
}
__global__ void k_mg_flag_elems(const int32_t *__restrict__ conn, int64_t nel, int nen, const uint8_t *__restrict__ band, int mine, uint8_t *__restrict__ flag)
{
    GRID_STRIDE(e, nel) {
        bool hit = false;
        for (int a = 0; a < nen; a++) hit |= band[conn[e * nen + a]] == mine;
        if (hit) flag[e] = 1;
    }
}
__global__ void k_mg_mark_nodes(const int32_t *__restrict__ conn, int nen, const int32_t *__restrict__ esel, int64_t nsel, uint8_t *__restrict__ used)
{
    GRID_STRIDE(t, nsel * nen) used[conn[(int64_t)esel[t / nen] * nen + t % nen]] = 1;
}
__global__ void k_mg_local_conn(const int32_t *__restrict__ conn, int nen, const int32_t *__restrict__ esel, int64_t nsel,
                                const int32_t *__restrict__ newid, int32_t *__restrict__ lconn)
{
    GRID_STRIDE(t, nsel * nen) lconn[t] = newid[conn[(int64_t)esel[t / nen] * nen + t % nen]];
}
__global__ void k_mg_gather_nodes(const uint8_t *__restrict__ used, const int32_t *__restrict__ newid, int64_t nnodes,
                                  const double2 *__restrict__ gxy, double2 *__restrict__ lxy, int32_t *__restrict__ gnode)
{
    GRID_STRIDE(n, nnodes) if (used[n]) { lxy[newid[n]] = gxy[n]; gnode[newid[n]] = (int32_t)n; }
}
// local dof table (global numbers, 0-based, -1 = unnumbered) + the owned dof numbers of this band (0xFFFFFFFF = not owned)
__global__ void k_mg_local_dofs(const int32_t *__restrict__ gdof, int ncomp, const int32_t *__restrict__ gnode, int64_t nloc,
                                const uint8_t *__restrict__ band, int mine, int32_t *__restrict__ ldof, uint32_t *__restrict__ owned)
{
    GRID_STRIDE(t, nloc * ncomp) {
        const int64_t k = t / ncomp;
        const int32_t g = gnode[k];
        const int32_t d = gdof[(int64_t)g * ncomp + t % ncomp];
        ldof[t] = d;
        owned[t] = (band[g] == mine && d >= 0) ? (uint32_t)d : 0xFFFFFFFFu;
    }
}
__global__ void k_mg_range_heads(const uint32_t *__restrict__ v, int64_t n, int32_t *__restrict__ head)
{
    GRID_STRIDE(i, n) head[i] = (v[i] != 0xFFFFFFFFu && (i == 0 || v[i] > v[i - 1] + 1)) ? 1 : 0;
}
__global__ void k_mg_range_fill(const uint32_t *__restrict__ v, int64_t n, int64_t nvalid, const int32_t *__restrict__ head, const int64_t *__restrict__ idx,
                                int64_t *__restrict__ firsts, int64_t *__restrict__ lasts)
{
    GRID_STRIDE(i, nvalid) {
        const int64_t r = head[i] ? idx[i] : idx[i] - 1;      // idx = exclusive scan of the head flags
        if (head[i]) firsts[r] = (int64_t)v[i] + 1;
        if (i + 1 == nvalid || head[i + 1]) lasts[r] = (int64_t)v[i] + 1;
    }
}
struct IsValidU32 { __device__ __forceinline__ bool operator()(uint32_t v) const { return v != 0xFFFFFFFFu; } };

// Int64 1-based host/device array -> Int32 0-based device array, through a staging buffer (defined in elfel_gpu.cu)
static void ingest_index(efg_ctx *ctx, const int64_t *src, int64_t n, int64_t lo, int64_t hi, int32_t *dst, const char *what);
static void invalidate(efg_ctx *ctx);

// everything one device does between "global arrays given" and "ready for the symbolic phase"
static void multi_shard_device(efg_multi *m, MultiDev &d, int axis, const std::vector<double> &cuts)
{
    efg_ctx *ctx = d.ctx;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevPool &pool = ctx->pool;
    invalidate(ctx);
    const int ncuts = (int)cuts.size();
    DevBuf<double> dcuts;
    dcuts.alloc(pool, (size_t)ncuts + 1);
    if (ncuts) CUDA_CHECK(cudaMemcpyAsync(dcuts.p, cuts.data(), (size_t)ncuts * sizeof(double), cudaMemcpyHostToDevice, st));
    const int nmesh = m->mesh[1].kind ? 2 : 1;
    const int64_t nel = m->mesh[0].nel;
    // 1. bands of the nodes of every mesh, global connectivity as Int32 on the device, element flags
    DevBuf<uint8_t> band[2], eflag;
    DevBuf<int32_t> gconn[2];
    DevBuf<double2> gxy[2];
    eflag.alloc(pool, (size_t)nel + 1);
    CUDA_CHECK(cudaMemsetAsync(eflag.p, 0, (size_t)nel + 1, st));
    for (int s = 0; s < nmesh; s++) {
        const GlobalMesh &gm = m->mesh[s];
        gxy[s].alloc(pool, (size_t)gm.nnodes + 1);
        CUDA_CHECK(cudaMemcpyAsync(gxy[s].p, gm.xy, (size_t)gm.nnodes * sizeof(double2), cudaMemcpyDefault, st));
        band[s].alloc(pool, (size_t)gm.nnodes + 1);
        LAUNCH(ctx, k_mg_band, grid_for(gm.nnodes, 256), 256, 0, gxy[s].p, gm.nnodes, axis, dcuts.p, ncuts, band[s].p);
        gconn[s].alloc(pool, (size_t)(gm.nel * gm.kind) + 1);
        ingest_index(ctx, gm.conn, gm.nel * gm.kind, 1, gm.nnodes, gconn[s].p, "efgm_set_mesh: node id");
        LAUNCH(ctx, k_mg_flag_elems, grid_for(gm.nel, 256), 256, 0, gconn[s].p, gm.nel, gm.kind, band[s].p, d.band, eflag.p);
    }
    // 2. selected elements (ascending global element number: contributions keep the global append order)
    DevBuf<int32_t> esel;
    DevBuf<int64_t> nsel_d;
    esel.alloc(pool, (size_t)nel + 1); nsel_d.alloc(pool, 1);
    {
        cub::CountingInputIterator<int32_t> it(0);
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, it, eflag.p, esel.p, nsel_d.p, nel, st);
        DevBuf<char> tmp;
        tmp.alloc(pool, tb);
        CUDA_CHECK(cub::DeviceSelect::Flagged(tmp.p, tb, it, eflag.p, esel.p, nsel_d.p, nel, st));
        ctx->launches += 2;
    }
    const int64_t nsel = tl_read(ctx, nsel_d.p);
    d.nel_local = nsel;
    eflag.release();
    // 3. per mesh: used nodes -> local numbering, local connectivity, coordinates, local -> global node list
    DevBuf<int32_t> gnode[2];
    int64_t nloc[2] = {0, 0};
    for (int s = 0; s < nmesh; s++) {
        const GlobalMesh &gm = m->mesh[s];
        DevBuf<uint8_t> used;
        DevBuf<int32_t> newid;
        used.alloc(pool, (size_t)gm.nnodes + 1); newid.alloc(pool, (size_t)gm.nnodes + 1);
        CUDA_CHECK(cudaMemsetAsync(used.p, 0, (size_t)gm.nnodes + 1, st));
        LAUNCH(ctx, k_mg_mark_nodes, grid_for(nsel * gm.kind, 256), 256, 0, gconn[s].p, gm.kind, esel.p, nsel, used.p);
        {
            cub::TransformInputIterator<int32_t, cub::CastOp<int32_t>, const uint8_t *> it(used.p, cub::CastOp<int32_t>());
            tl_excl_scan(ctx, it, newid.p, gm.nnodes + 1);
        }
        nloc[s] = (int64_t)tl_read(ctx, newid.p + gm.nnodes);
        MeshDev &md = ctx->mesh[s];
        md.kind = gm.kind; md.nel = nsel; md.nnodes = nloc[s];
        md.conn.alloc(pool, (size_t)(nsel * gm.kind) + 1);
        md.xy.alloc(pool, (size_t)nloc[s] + 1);
        gnode[s].alloc(pool, (size_t)nloc[s] + 1);
        LAUNCH(ctx, k_mg_local_conn, grid_for(nsel * gm.kind, 256), 256, 0, gconn[s].p, gm.kind, esel.p, nsel, newid.p, md.conn.p);
        LAUNCH(ctx, k_mg_gather_nodes, grid_for(gm.nnodes, 256), 256, 0, used.p, newid.p, gm.nnodes, gxy[s].p, md.xy.p, gnode[s].p);
        gconn[s].release(); gxy[s].release();
    }
    if (nmesh == 1) { ctx->mesh[1].kind = 0; ctx->mesh[1].nel = 0; ctx->mesh[1].nnodes = 0; ctx->mesh[1].conn.release(); ctx->mesh[1].xy.release(); }
    // 4. per space: local dof table with GLOBAL dof numbers; the owned dofs of this band -> column ranges
    std::vector<uint32_t *> owned_lists;
    std::vector<int64_t> owned_sizes;
    DevBuf<uint32_t> owned[3];
    int64_t ntot = 0;
    for (int s = 0; s < 3; s++) {
        const GlobalSpace &gs = m->space[s];
        SpaceDev &sd = ctx->space[s];
        if (gs.mesh < 0) { sd.mesh = -1; sd.ncomp = 0; sd.nnodes = 0; sd.dof.release(); continue; }
        const int ms = gs.mesh;
        DevBuf<int32_t> gdof;
        gdof.alloc(pool, (size_t)(gs.nnodes * gs.ncomp) + 1);
        ingest_index(ctx, gs.dofnums, gs.nnodes * gs.ncomp, 0, ((int64_t)1 << 31) - 1, gdof.p, "efgm_set_space: dof number");
        sd.mesh = ms; sd.ncomp = gs.ncomp; sd.nnodes = nloc[ms];
        sd.dof.alloc(pool, (size_t)(nloc[ms] * gs.ncomp) + 1);
        owned[s].alloc(pool, (size_t)(nloc[ms] * gs.ncomp) + 1);
        LAUNCH(ctx, k_mg_local_dofs, grid_for(nloc[ms] * gs.ncomp, 256), 256, 0, gdof.p, gs.ncomp, gnode[ms].p, nloc[ms], band[ms].p, d.band, sd.dof.p, owned[s].p);
        ntot += nloc[ms] * gs.ncomp;
    }
    DevBuf<uint32_t> all, sorted;
    all.alloc(pool, (size_t)ntot + 1); sorted.alloc(pool, (size_t)ntot + 1);
    {
        int64_t o = 0;
        for (int s = 0; s < 3; s++) {
            if (m->space[s].mesh < 0) continue;
            const int64_t n = nloc[m->space[s].mesh] * m->space[s].ncomp;
            CUDA_CHECK(cudaMemcpyAsync(all.p + o, owned[s].p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            o += n;
        }
    }
    tl_sort_keys(ctx, all.p, sorted.p, ntot, 32);            // not-owned entries (0xFFFFFFFF) sort last
    DevBuf<int64_t> nvalid_d;
    nvalid_d.alloc(pool, 1);
    {
        cub::TransformInputIterator<int64_t, IsValidU32, const uint32_t *> it(sorted.p, IsValidU32());
        size_t tb = 0;
        cub::DeviceReduce::Sum(nullptr, tb, it, nvalid_d.p, ntot, st);
        DevBuf<char> tmp;
        tmp.alloc(pool, tb);
        CUDA_CHECK(cub::DeviceReduce::Sum(tmp.p, tb, it, nvalid_d.p, ntot, st));
        ctx->launches += 2;
    }
    const int64_t nvalid = tl_read(ctx, nvalid_d.p);
    DevBuf<int32_t> head;
    DevBuf<int64_t> hidx;
    head.alloc(pool, (size_t)ntot + 2); hidx.alloc(pool, (size_t)ntot + 2);
    CUDA_CHECK(cudaMemsetAsync(head.p, 0, ((size_t)ntot + 2) * sizeof(int32_t), st));
    LAUNCH(ctx, k_mg_range_heads, grid_for(nvalid, 256), 256, 0, sorted.p, nvalid, head.p);
    {
        cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int32_t *> it(head.p, cub::CastOp<int64_t>());
        tl_excl_scan(ctx, it, hidx.p, nvalid + 1);
    }
    const int64_t nranges = tl_read(ctx, hidx.p + nvalid);
    d.firsts.assign((size_t)nranges, 0); d.lasts.assign((size_t)nranges, 0);
    if (nranges > 0) {
        DevBuf<int64_t> f, l;
        f.alloc(pool, (size_t)nranges); l.alloc(pool, (size_t)nranges);
        // a run's last value: the entry before the next head (idx of a non-head entry = idx of its run's head + 1 in the exclusive scan)
        LAUNCH(ctx, k_mg_range_fill, grid_for(nvalid, 256), 256, 0, sorted.p, ntot, nvalid, head.p, hidx.p, f.p, l.p);
        CUDA_CHECK(cudaMemcpyAsync(d.firsts.data(), f.p, (size_t)nranges * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(d.lasts.data(), l.p, (size_t)nranges * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
}

// cuts between the bands: quantiles of the node coordinates of mesh 0 along the longer axis of its bounding box
__global__ void k_mg_axis(const double2 *__restrict__ xy, int64_t n, int axis, double *__restrict__ out)
{
    GRID_STRIDE(i, n) out[i] = axis ? xy[i].y : xy[i].x;
}
static void multi_cuts(efg_multi *m, int &axis, std::vector<double> &cuts)
{
    efg_ctx *ctx = m->dev[0].ctx;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevPool &pool = ctx->pool;
    const GlobalMesh &gm = m->mesh[0];
    const int ng = (int)m->dev.size();
    cuts.clear();
    axis = 1;
    if (ng == 1 || gm.nnodes == 0) return;
    DevBuf<double2> xy;
    xy.alloc(pool, (size_t)gm.nnodes + 1);
    CUDA_CHECK(cudaMemcpyAsync(xy.p, gm.xy, (size_t)gm.nnodes * sizeof(double2), cudaMemcpyDefault, st));
    DevBuf<BBox> bb;
    bb.alloc(pool, 1);
    {
        cub::TransformInputIterator<BBox, XYToBBox, const double2 *> it(xy.p, XYToBBox());
        const BBox init{1e300, 1e300, -1e300, -1e300};
        size_t tb = 0;
        cub::DeviceReduce::Reduce(nullptr, tb, it, bb.p, gm.nnodes, BBoxOp(), init, st);
        DevBuf<char> tmp;
        tmp.alloc(pool, tb);
        CUDA_CHECK(cub::DeviceReduce::Reduce(tmp.p, tb, it, bb.p, gm.nnodes, BBoxOp(), init, st));
        ctx->launches += 2;
    }
    const BBox hb = tl_read(ctx, bb.p);
    axis = (hb.y1 - hb.y0 >= hb.x1 - hb.x0) ? 1 : 0;
    DevBuf<double> k1, k2;
    k1.alloc(pool, (size_t)gm.nnodes + 1); k2.alloc(pool, (size_t)gm.nnodes + 1);
    LAUNCH(ctx, k_mg_axis, grid_for(gm.nnodes, 256), 256, 0, xy.p, gm.nnodes, axis, k1.p);
    xy.release();
    tl_sort_keys(ctx, k1.p, k2.p, gm.nnodes, 64);
    for (int k = 1; k < ng; k++) cuts.push_back(tl_read(ctx, k2.p + (gm.nnodes * k) / ng));
}

// run fn(device index) on one host thread per device; the first failure (lowest device index) is reported
static int multi_parallel(efg_multi *m, const std::function<void(int)> &fn)
{
    const int ng = (int)m->dev.size();
    std::vector<std::thread> th;
    for (int i = 0; i < ng; i++) {
        m->dev[i].rc = EFG_OK;
        th.emplace_back([m, i, &fn]() {
            MultiDev &d = m->dev[i];
            try {
                fn(i);
            } catch (const EfgError &e) {
                d.rc = e.code; d.err = e.msg;
            } catch (const std::bad_alloc &) {
                d.rc = EFG_ERR_OOM; d.err = "host allocation failed";
            } catch (...) {
                d.rc = EFG_ERR_CUDA; d.err = "unknown error";
            }
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < ng; i++)
        if (m->dev[i].rc != EFG_OK) {
            char buf[64];
            snprintf(buf, sizeof buf, "device %d (band %d): ", m->dev[i].ctx->device, i);
            m->err = std::string(buf) + m->dev[i].err;
            return m->dev[i].rc;
        }
    return EFG_OK;
}
// call a single-ctx entry point from a device thread: a failure becomes an exception carrying the ctx's message
static void multi_ck(efg_ctx *ctx, int rc)
{
    if (rc != EFG_OK) throw EfgError{rc, ctx->err};
}
