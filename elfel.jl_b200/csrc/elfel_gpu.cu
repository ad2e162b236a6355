// elfel_gpu.cu -- libelfelgpu.so: C ABI (include/elfel_gpu.h) over the sm_100a kernels.
// Single translation unit (the __constant__ tables are shared by all kernels).
#include "efg_ctx.cuh"
#include "efg_twopass.cuh"
#include "efg_tiled.cuh"
#include "efg_vector.cuh"
#include "efg_multi.cuh"
#include "efg_gen.cuh"
#include "efg_hostcopy.cuh"

#include <cstring>
#include <cmath>

// ------------------------------------------------------------------------------------------------
// quadrature + basis tables of the active rule (what QPIterator's ctor precomputes,
// src/QPIterators.jl:14-50,79-84) -> __constant__ c_tab.  Host arithmetic is plain IEEE double
// (compiled with -ffp-contract=off) in the reference's expression order.
// ------------------------------------------------------------------------------------------------
static int gauss1(int order, double *pc, double *w) // src/RefShapes.jl:91-106 (truncated literals kept)
{
    switch (order) {
    case 1: pc[0] = 0.0; w[0] = 2.0; return 1;
    case 2: pc[0] = -0.577350269189626; pc[1] = 0.577350269189626; w[0] = 1.0; w[1] = 1.0; return 2;
    case 3: pc[0] = -0.774596669241483; pc[1] = 0.0; pc[2] = 0.774596669241483;
            w[0] = 0.5555555555555556; w[1] = 0.8888888888888889; w[2] = 0.5555555555555556; return 3;
    case 4: pc[0] = -0.86113631159405; pc[1] = -0.33998104358486; pc[2] = 0.33998104358486; pc[3] = 0.86113631159405;
            w[0] = 0.34785484513745; w[1] = 0.65214515486255; w[2] = 0.65214515486255; w[3] = 0.34785484513745; return 4;
    case 5: pc[0] = -0.906179845938664; pc[1] = -0.538469310105683; pc[2] = 0.000000000000000;
            pc[3] = 0.538469310105683; pc[4] = 0.906179845938664;
            w[0] = 0.236926885056189; w[1] = 0.478628670499367; w[2] = 0.568888888888889;
            w[3] = 0.478628670499367; w[4] = 0.236926885056189; return 5;
    }
    return -1;   // order > 5 is Golub-Welsch in the reference: out of scope
}

// the higher triangle rules of _triangle, src/RefShapes.jl:120-230: literal tables (incl. the reference's own digits), weights
// divided by 2 where the source does
static int triangle_rule_table(int npts, double (*pc)[2], double *w)
{
    static const double P4[4][2] = {{0.333333333333333, 0.333333333333333}, {0.200000000000000, 0.200000000000000},
                                    {0.600000000000000, 0.200000000000000}, {0.200000000000000, 0.600000000000000}};
    static const double W4[4] = {-0.281250000000000, 0.260416666666667, 0.260416666666667, 0.260416666666667};
    static const double P6[6][2] = {{0.816847572980459, 0.091576213509771}, {0.091576213509771, 0.816847572980459},
                                    {0.091576213509771, 0.091576213509771}, {0.108103018168070, 0.445948490915965},
                                    {0.445948490915965, 0.108103018168070}, {0.445948490915965, 0.445948490915965}};
    static const double W6[6] = {0.109951743655322, 0.109951743655322, 0.109951743655322, 0.223381589678011, 0.223381589678011, 0.223381589678011};
    static const double P7[7][2] = {{0.101286507323456, 0.101286507323456}, {0.797426958353087, 0.101286507323456},
                                    {0.101286507323456, 0.797426958353087}, {0.470142064105115, 0.470142064105115},
                                    {0.059715871789770, 0.470142064105115}, {0.470142064105115, 0.059715871789770},
                                    {0.333333333333333, 0.333333333333333}};
    static const double W7[7] = {0.062969590272414, 0.062969590272414, 0.062969590272414, 0.066197076394253, 0.066197076394253,
                                 0.066197076394253, 0.112500000000000};
    static const double P9[9][2] = {{0.437525248383384, 0.437525248383384}, {0.124949503233232, 0.437525248383384},
                                    {0.437525248383384, 0.124949503233232}, {0.165409927389841, 0.037477420750088},
                                    {0.037477420750088, 0.165409927389841}, {0.797112651860071, 0.165409927389841},
                                    {0.165409927389841, 0.797112651860071}, {0.037477420750088, 0.797112651860071},
                                    {0.797112651860071, 0.037477420750088}};
    static const double W9[9] = {0.205950504760887, 0.205950504760887, 0.205950504760887, 0.063691414286223, 0.063691414286223,
                                 0.063691414286223, 0.063691414286223, 0.063691414286223, 0.063691414286223};
    static const double P12[12][2] = {{0.063089014491502, 0.063089014491502}, {0.873821971016996, 0.063089014491502},
                                      {0.063089014491502, 0.873821971016996}, {0.249286745170910, 0.249286745170910},
                                      {0.501426509658179, 0.249286745170910}, {0.249286745170910, 0.501426509658179},
                                      {0.310352451033785, 0.053145049844816}, {0.053145049844816, 0.310352451033785},
                                      {0.636502499121399, 0.310352451033785}, {0.310352451033785, 0.636502499121399},
                                      {0.053145049844816, 0.636502499121399}, {0.636502499121399, 0.053145049844816}};
    static const double W12[12] = {0.050844906370207, 0.050844906370207, 0.050844906370207, 0.116786275726379, 0.116786275726379,
                                   0.116786275726379, 0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374,
                                   0.082851075618374, 0.082851075618374};
    static const double P13[13][2] = {{0.333333333333333, 0.333333333333333}, {0.479308067841923, 0.260345966079038},
                                      {0.260345966079038, 0.479308067841923}, {0.260345966079038, 0.260345966079038},
                                      {0.869739794195568, 0.065130102902216}, {0.065130102902216, 0.869739794195568},
                                      {0.065130102902216, 0.065130102902216}, {0.638444188569809, 0.312865496004875},
                                      {0.638444188569809, 0.048690315425316}, {0.312865496004875, 0.638444188569809},
                                      {0.312865496004875, 0.048690315425316}, {0.048690315425316, 0.638444188569809},
                                      {0.048690315425316, 0.312865496004875}};
    static const double W13[13] = {-0.149570044467670, 0.175615257433204, 0.175615257433204, 0.175615257433204, 0.053347235608839,
                                   0.053347235608839, 0.053347235608839, 0.077113760890257, 0.077113760890257, 0.077113760890257,
                                   0.077113760890257, 0.077113760890257, 0.077113760890257};
    const double (*P)[2] = nullptr; const double *W = nullptr; bool halve = false;
    switch (npts) {
    case 4: P = P4; W = W4; break;
    case 6: P = P6; W = W6; halve = true; break;
    case 7: P = P7; W = W7; break;
    case 9: P = P9; W = W9; halve = true; break;
    case 12: P = P12; W = W12; halve = true; break;
    case 13: P = P13; W = W13; halve = true; break;
    default: return -1;
    }
    for (int q = 0; q < npts; q++) { pc[q][0] = P[q][0]; pc[q][1] = P[q][1]; w[q] = halve ? W[q] / 2 : W[q]; }
    return npts;
}

// returns npts; pc is npts x 2
static int quadrature_points(int kind, int rule, double (*pc)[2], double *w)
{
    if (kind == EFG_T3 || kind == EFG_T6) {
        if (rule == 1) { // src/RefShapes.jl:114-116
            pc[0][0] = 1.0 / 3.; pc[0][1] = 1.0 / 3.; w[0] = 1.0 / 2.0; return 1;
        }
        if (rule == 3) { // src/RefShapes.jl:117-119
            pc[0][0] = 2.0 / 3; pc[0][1] = 1.0 / 6; pc[1][0] = 1.0 / 6; pc[1][1] = 2.0 / 3;
            pc[2][0] = 1.0 / 6; pc[2][1] = 1.0 / 6;
            w[0] = w[1] = w[2] = (1.0 / 3) / 2; return 3;
        }
        return triangle_rule_table(rule, pc, w);
    }
    if (kind == EFG_Q4) { // src/RefShapes.jl:350-362: i outer, j inner
        double p1[5], w1[5];
        const int np = gauss1(rule, p1, w1);
        if (np < 0) return -1;
        int r = 0;
        for (int i = 0; i < np; i++)
            for (int j = 0; j < np; j++) { pc[r][0] = p1[i]; pc[r][1] = p1[j]; w[r] = w1[i] * w1[j]; r++; }
        return r;
    }
    return -1;
}

// reference shape (element kind of the mesh) behind a form's kind code (the velocity / pressure pairs of row f5 live on one T3 / Q4 mesh)
static int kind_shape(int vkind)
{
    return vkind == EFG_PAIR_T3B_T3 || vkind == EFG_PAIR_T3_L2 ? EFG_T3 : (vkind == EFG_PAIR_Q4_L2 ? EFG_Q4 : vkind);
}

static void basis_tables(int kind, double r, double s, double *N, double (*g)[2])
{
    if (kind == EFG_FE_T3_BUBBLE) { // src/FElements.jl:341-356: ((1 - xi - eta) * xi) * eta and its parametric gradient
        N[0] = (1 - r - s); N[1] = r; N[2] = s; N[3] = (1 - r - s) * r * s;
        g[0][0] = -1.; g[0][1] = -1.; g[1][0] = +1.; g[1][1] = 0.; g[2][0] = 0.; g[2][1] = +1.;
        g[3][0] = (-r * s + (1 - r - s) * s); g[3][1] = (-r * s + (1 - r - s) * r);
    } else if (kind == EFG_FE_L2) { // src/FElements.jl:412-419, 441-448
        N[0] = 1.0; g[0][0] = 0.0; g[0][1] = 0.0;
    } else if (kind == EFG_T3) { // src/FElements.jl:239-246
        N[0] = (1 - r - s); N[1] = r; N[2] = s;
        g[0][0] = -1.; g[0][1] = -1.; g[1][0] = +1.; g[1][1] = 0.; g[2][0] = 0.; g[2][1] = +1.;
    } else if (kind == EFG_T6) { // src/FElements.jl:264-288
        const double t = 1. - r - s;
        N[0] = t * (t + t - 1); N[1] = r * (r + r - 1); N[2] = s * (s + s - 1);
        N[3] = 4 * r * t; N[4] = 4 * r * s; N[5] = 4 * s * t;
        g[0][0] = -3 + 4 * r + 4 * s; g[0][1] = -3 + 4 * r + 4 * s;
        g[1][0] = 4 * r - 1;          g[1][1] = 0.0;
        g[2][0] = 0.0;                g[2][1] = 4 * s - 1;
        g[3][0] = 4 - 8 * r - 4 * s;  g[3][1] = -4 * r;
        g[4][0] = 4 * s;              g[4][1] = 4 * r;
        g[5][0] = -4 * s;             g[5][1] = 4 - 4 * r - 8 * s;
    } else { // Q4: src/FElements.jl:306-320
        N[0] = 0.25 * (1. - r) * (1. - s); N[1] = 0.25 * (1. + r) * (1. - s);
        N[2] = 0.25 * (1. + r) * (1. + s); N[3] = 0.25 * (1. - r) * (1. + s);
        g[0][0] = -(1. - s) * 0.25; g[0][1] = -(1. - r) * 0.25;
        g[1][0] = (1. - s) * 0.25;  g[1][1] = -(1. + r) * 0.25;
        g[2][0] = (1. + s) * 0.25;  g[2][1] = (1. + r) * 0.25;
        g[3][0] = -(1. + s) * 0.25; g[3][1] = (1. - r) * 0.25;
    }
}

// Host copy of the tables of (element kind `vkind`, rule): triangles share their rule between T3 and T6.  Returns npts or -1.
// tetrahedron rules: src/RefShapes.jl:232-259 (weights as literally written); returns npts, pc is npts x 3
static int quadrature_points_t4(int npts, double (*pc)[3], double *w)
{
    if (npts == 1) { pc[0][0] = pc[0][1] = pc[0][2] = 0.25; w[0] = 1.0 / 6.0; return 1; }
    if (npts == 4) {
        const double a = 0.13819660, b = 0.58541020;
        const double P[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
        for (int q = 0; q < 4; q++) { for (int k = 0; k < 3; k++) pc[q][k] = P[q][k]; w[q] = 0.041666666666666666667; }
        return 4;
    }
    if (npts == 5) {
        const double a = 1.0 / 6.0, b = 0.25, c = 0.5, d = -0.8, e = 0.45;
        const double P[5][3] = {{b, b, b}, {c, a, a}, {a, c, a}, {a, a, c}, {a, a, a}};
        const double W[5] = {d, e, e, e, e};
        for (int q = 0; q < 5; q++) { for (int k = 0; k < 3; k++) pc[q][k] = P[q][k]; w[q] = W[q] / 6; }
        return 5;
    }
    return -1;
}

static int build_tables(int vkind, int rule, QTab (&h)[EFG_NTAB])
{
    memset(h, 0, sizeof h);
    if (vkind == EFG_T4) {     // src/FElements.jl:372-378: N = (1 - r - s - t, r, s, t)
        double pc3[EFG_MAXQ][3], w3[EFG_MAXQ];
        const int n3 = quadrature_points_t4(rule, pc3, w3);
        h[EFG_TAB_T4].npts = n3 > 0 ? n3 : 0;
        for (int q = 0; q < n3; q++) {
            h[EFG_TAB_T4].w[q] = w3[q];
            h[EFG_TAB_T4].N[q][0] = (1 - pc3[q][0] - pc3[q][1] - pc3[q][2]); h[EFG_TAB_T4].N[q][1] = pc3[q][0];
            h[EFG_TAB_T4].N[q][2] = pc3[q][1]; h[EFG_TAB_T4].N[q][3] = pc3[q][2];
        }
        return n3;
    }
    double pc[EFG_MAXQ][2], w[EFG_MAXQ];
    vkind = kind_shape(vkind);
    const int npts = quadrature_points(vkind, rule, pc, w);
    if (npts < 0 || npts > EFG_MAXQ) return -1;
    const int kinds[5] = {EFG_T3, EFG_Q4, EFG_T6, EFG_FE_T3_BUBBLE, EFG_FE_L2};     // c_tab slots 0..4 (slot 5: FEH1_T4, above)
    for (int k = 0; k < 5; k++) {
        const bool tri = kinds[k] != EFG_Q4, vtri = vkind != EFG_Q4;
        if (tri != vtri && kinds[k] != EFG_FE_L2) continue;
        h[k].npts = npts;
        for (int q = 0; q < npts; q++) {
            h[k].w[q] = w[q];
            basis_tables(kinds[k], pc[q][0], pc[q][1], h[k].N[q], h[k].gp[q]);
        }
    }
    return npts;
}
static int quad_npts(int vkind, int rule)
{
    double pc[EFG_MAXQ][2], w[EFG_MAXQ];
    if (vkind == EFG_T4) return (rule == 1 || rule == 4 || rule == 5) ? rule : -1;
    const int npts = quadrature_points(kind_shape(vkind), rule, pc, w);
    return (npts < 0 || npts > EFG_MAXQ) ? -1 : npts;
}

// ------------------------------------------------------------------------------------------------
// Ownership of the __constant__ tables.  c_tab / c_prm exist once per DEVICE (every device has its own copy of the
// module's constant bank), shared by all ctx of this process on that device, while every entry point that launches
// a table-reading kernel returns without synchronising.  TabGuard serialises what has to be serialised and nothing
// else:
//   - a per-device mutex is held from "make the tables mine" until this ctx's table-reading launches are enqueued
//     and an event after them is recorded on the ctx's stream;
//   - the device's current content (host mirror) is compared first: identical tables/parameters are not uploaded
//     again (the usual case: repeated efg_numeric calls, or several ctx using the same rule);
//   - before an upload, the ctx's stream waits for the table events of every OTHER live ctx on the device, so no
//     kernel still reading the old tables is overtaken; its own earlier kernels are ordered by the stream itself.
// Ctx on different devices never contend (different mutex, different constant bank): one host thread per GPU works.
// ------------------------------------------------------------------------------------------------
#include <mutex>
#include <algorithm>
#define EFG_MAX_DEVICES 64
struct DevTables {
    std::mutex mu;
    bool valid = false;
    QTab tab[EFG_NTAB];
    double prm[16];
    std::vector<efg_ctx *> users;       // live ctx on this device
    efg_ctx *uploader = nullptr;        // whose stream carried the last upload
};
static DevTables g_devtab[EFG_MAX_DEVICES];

static void tables_register(efg_ctx *ctx)
{
    DevTables &d = g_devtab[ctx->device];
    std::lock_guard<std::mutex> lk(d.mu);
    d.users.push_back(ctx);
}
static void tables_unregister(efg_ctx *ctx)
{
    DevTables &d = g_devtab[ctx->device];
    std::lock_guard<std::mutex> lk(d.mu);
    d.users.erase(std::remove(d.users.begin(), d.users.end(), ctx), d.users.end());
    if (d.uploader == ctx) d.uploader = nullptr;      // (its stream has been synchronised by efg_destroy)
}

struct TabGuard {
    efg_ctx *ctx;
    DevTables &d;
    std::unique_lock<std::mutex> lk;
    int npts = -1;
    // prm == nullptr: the kernels about to be launched do not read c_prm
    TabGuard(efg_ctx *c, int vkind, int rule, const double *prm, int nprm) : ctx(c), d(g_devtab[c->device]), lk(d.mu)
    {
        QTab h[EFG_NTAB];
        npts = build_tables(vkind, rule, h);
        if (npts < 0) return;
        double hp[16] = {0};
        for (int i = 0; i < nprm && i < 16; i++) hp[i] = prm[i];
        const bool same_tab = d.valid && memcmp(h, d.tab, sizeof h) == 0;
        const bool same_prm = !prm || (d.valid && memcmp(hp, d.prm, sizeof hp) == 0);
        if (same_tab && same_prm) { order_after_upload(); return; }
        for (efg_ctx *o : d.users)
            if (o != ctx && o->tab_event_recorded) CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, o->ev_tab, 0));
        d.valid = false;
        // sources are small pageable host buffers: the runtime stages them before the call returns
        if (!same_tab) CUDA_CHECK(cudaMemcpyToSymbolAsync(c_tab, h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
        if (!same_prm) CUDA_CHECK(cudaMemcpyToSymbolAsync(c_prm, hp, sizeof hp, 0, cudaMemcpyHostToDevice, ctx->stream));
        // a ctx that finds "its" content later must still be ordered after this upload: it waits for our event
        memcpy(d.tab, h, sizeof h);
        if (prm) memcpy(d.prm, hp, sizeof hp);
        d.valid = true;
        d.uploader = ctx;
        CUDA_CHECK(cudaEventRecord(ctx->ev_tab, ctx->stream));
        ctx->tab_event_recorded = true;
    }
    // a ctx that did not upload reads tables another ctx's stream wrote: order its kernels after that upload
    void order_after_upload()
    {
        if (d.uploader && d.uploader != ctx && d.uploader->tab_event_recorded)
            CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, d.uploader->ev_tab, 0));
    }
    ~TabGuard()
    {
        if (npts >= 0) {        // everything this ctx enqueued while holding the tables
            if (cudaEventRecord(ctx->ev_tab, ctx->stream) == cudaSuccess) ctx->tab_event_recorded = true;
            else cudaGetLastError();
        }
    }
};

// ------------------------------------------------------------------------------------------------
// ingestion: Int64 1-based host/device arrays -> Int32 0-based device arrays (validated)
// ------------------------------------------------------------------------------------------------
__global__ void k_convert_index(const int64_t *__restrict__ in, int64_t n, int64_t lo, int64_t hi,
                                int32_t *__restrict__ out, int *__restrict__ errflag)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t v = in[i];
        if (v < lo || v > hi) *errflag = 1;
        out[i] = (int32_t)(v - 1);
    }
}

// copy n Int64 from (host or device) src and convert; lo..hi = accepted 1-based range
static void ingest_index(efg_ctx *ctx, const int64_t *src, int64_t n, int64_t lo, int64_t hi, int32_t *dst, const char *what)
{
    const int64_t CH = (int64_t)32 << 20; // 32 Mi entries = 256 MB staging
    DevBuf<int64_t> stage;
    DevBuf<int> errflag;
    stage.alloc(ctx->pool, (size_t)(n < CH ? (n > 0 ? n : 1) : CH));
    errflag.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemsetAsync(errflag.p, 0, sizeof(int), ctx->stream));
    for (int64_t o = 0; o < n; o += CH) {
        const int64_t m = (n - o < CH) ? n - o : CH;
        CUDA_CHECK(cudaMemcpyAsync(stage.p, src + o, (size_t)m * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
        LAUNCH(ctx, k_convert_index, grid_for(m, 256), 256, 0, stage.p, m, lo, hi, dst + o, errflag.p);
    }
    int herr = 0;
    CUDA_CHECK(cudaMemcpyAsync(&herr, errflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (herr) efg_throw(EFG_ERR_INDEX, "%s: index out of range [%lld, %lld]", what, (long long)lo, (long long)hi);
}

static void wait_copies(efg_ctx *ctx)
{
    if (ctx->widen) {               // host threads widening the row indices in the caller's array (efg_hostcopy.cuh)
        HostWiden *w = static_cast<HostWiden *>(ctx->widen);
        w->join();
        if (w->failed != 0) ctx->widen_failed = true;       // reported by the fetch call that waits for the copy
        delete w;
        ctx->widen = nullptr;
    }
    if (ctx->copy_pending && ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        ctx->copy_pending = false;
    }
}
static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
static void invalidate(efg_ctx *ctx)
{
    wait_copies(ctx);              // an asynchronous pattern fetch may still read colptr / rowval
    ctx->have_symbolic = false;
    ctx->have_pattern = false;
    ctx->have_values = false;
    ctx->nnz = 0;
    ctx->colptr.release(); ctx->rowval.release(); ctx->nzval.release();
    ctx->tp.perm.release(); ctx->tp.seg_start.release(); ctx->tp.Ke.release();
    tiled_release(ctx);
    vec_release(ctx);
}

// ------------------------------------------------------------------------------------------------
// output conversion: Int32 0-based -> Int64 1-based
// ------------------------------------------------------------------------------------------------
__global__ void k_rowval_out(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int64_t)in[i] + 1;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// EFG_OPT_DEFER_XY: efg_set_mesh keeps the caller's coordinate pointer instead of copying; the copy is issued on a separate
// input stream when the symbolic phase starts (it then overlaps the pattern kernels, which read connectivity and dof maps
// only), or on the main stream by whichever other entry point comes first.
static void flush_deferred_xy(efg_ctx *ctx, bool overlapped)
{
    for (int k = 0; k < 2; k++) {
        MeshDev &m = ctx->mesh[k];
        if (!m.pending_xy) continue;
        const double *src = m.pending_xy;
        m.pending_xy = nullptr;
        if (m.nnodes <= 0) continue;
        if (overlapped) {
            if (!ctx->in_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->in_stream, cudaStreamNonBlocking));
            if (!ctx->ev_xy) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_xy, cudaEventDisableTiming));
            // ordered after whatever the main stream did to the (re-allocated) buffer so far
            CUDA_CHECK(cudaEventRecord(ctx->ev_xy, ctx->stream));
            CUDA_CHECK(cudaStreamWaitEvent(ctx->in_stream, ctx->ev_xy, 0));
            CUDA_CHECK(cudaMemcpyAsync(m.xy.p, src, (size_t)m.nnodes * sizeof(double2), cudaMemcpyDefault, ctx->in_stream));
            CUDA_CHECK(cudaEventRecord(ctx->ev_xy, ctx->in_stream));
            ctx->xy_in_flight = true;
        } else {
            CUDA_CHECK(cudaMemcpyAsync(m.xy.p, src, (size_t)m.nnodes * sizeof(double2), cudaMemcpyDefault, ctx->stream));
        }
    }
}
// the main stream's next kernels read coordinates: wait for an overlapped copy
static void wait_deferred_xy(efg_ctx *ctx)
{
    if (ctx->xy_in_flight) { CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_xy, 0)); ctx->xy_in_flight = false; }
}

#define API_BEGIN_RAW(ctx)                                                      \
    if (!(ctx)) return EFG_ERR_INVALID;                                         \
    try {                                                                       \
        cudaError_t sd__ = cudaSetDevice((ctx)->device);                        \
        if (sd__ != cudaSuccess) efg_throw(EFG_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(sd__));
/* every entry point that may read coordinates first brings in the ones efg_set_mesh deferred (EFG_OPT_DEFER_XY) */
#define API_BEGIN(ctx)                                                          \
    API_BEGIN_RAW(ctx)                                                          \
        flush_deferred_xy(ctx, false);
#define API_END(ctx)                                                            \
        return EFG_OK;                                                          \
    } catch (const EfgError &e) {                                               \
        (ctx)->err = e.msg;                                                     \
        return e.code;                                                          \
    } catch (const std::bad_alloc &) {                                          \
        (ctx)->err = "host allocation failed";                                  \
        return EFG_ERR_OOM;                                                     \
    } catch (...) {                                                             \
        (ctx)->err = "unknown error";                                           \
        return EFG_ERR_CUDA;                                                    \
    }

extern "C" {

const char *efg_version(void) { return "elfelgpu 0.1 (sm_100a, FP64, CUDA " __DATE__ ")"; }

int efg_create(int device, efg_ctx **out)
{
    if (!out) return EFG_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return EFG_ERR_CUDA; }
    if (device < 0 || device >= ndev || device >= EFG_MAX_DEVICES) return EFG_ERR_INVALID;
    efg_ctx *ctx = new (std::nothrow) efg_ctx();
    if (!ctx) return EFG_ERR_OOM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->evn0) != cudaSuccess || cudaEventCreate(&ctx->evn1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_tab, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_pattern, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return EFG_ERR_CUDA;
    }
    ctx->pool.stream = ctx->stream;
    tables_register(ctx);
    // Bring the library's device module in now, like every handle-creating call of the CUDA libraries does: with CUDA's lazy
    // loading it would otherwise be loaded by the first kernel launch, inside the caller's first assembly.
    {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, (const void *)k_mailbox) != cudaSuccess) cudaGetLastError();
    }
    *out = ctx;
    return EFG_OK;
}

int efg_destroy(efg_ctx *ctx)
{
    if (!ctx) return EFG_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tables_unregister(ctx);
    invalidate(ctx);
    for (auto &m : ctx->mesh) { m.conn.release(); m.xy.release(); m.z.release(); }
    for (auto &s : ctx->space) { s.dof.release(); s.isdatum.release(); s.cdof.release(); }
    ctx->rfirst.release(); ctx->rlast1.release(); ctx->roff.release(); ctx->scratch.release(); ctx->cstage[0].release(); ctx->cstage[1].release();
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    for (cudaEvent_t e : ctx->widen_events) cudaEventDestroy(e);
    if (ctx->in_stream) { cudaStreamSynchronize(ctx->in_stream); cudaStreamDestroy(ctx->in_stream); }
    if (ctx->ev_xy) cudaEventDestroy(ctx->ev_xy);
    ctx->pool.destroy();           // the ctx's private arena goes back to the driver; nothing process-global is touched
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->evn0); cudaEventDestroy(ctx->evn1); cudaEventDestroy(ctx->ev_tab); cudaEventDestroy(ctx->ev_pattern);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return EFG_OK;
}

const char *efg_last_error(const efg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int efg_set_option(efg_ctx *ctx, int option, int64_t value)
{
    API_BEGIN_RAW(ctx)
    switch (option) {
    case EFG_OPT_PATH:
        if (value < 0 || value > 2) efg_throw(EFG_ERR_INVALID, "EFG_OPT_PATH must be 0, 1 or 2");
        if (ctx->opt_path != (int)value) { ctx->opt_path = (int)value; invalidate(ctx); }
        break;
    case EFG_OPT_STRICT_FP: ctx->opt_strict = value ? 1 : 0; ctx->have_values = false; break;
    case EFG_OPT_TILE_ELEMS:
        if (value < 0 || value > 4096) efg_throw(EFG_ERR_INVALID, "EFG_OPT_TILE_ELEMS out of range");
        if (ctx->opt_tile_elems != (int)value) { ctx->opt_tile_elems = (int)value; invalidate(ctx); }
        break;
    case EFG_OPT_SFC_ORDER:
        if (ctx->opt_sfc != (value ? 1 : 0)) { ctx->opt_sfc = value ? 1 : 0; invalidate(ctx); }
        break;
    case EFG_OPT_DEFER_XY: ctx->opt_defer_xy = value ? 1 : 0; break;
    case EFG_OPT_HOST_WIDEN: ctx->opt_host_widen = value < 0 ? -1 : (value ? 1 : 0); break;
    case EFG_OPT_FUSE_LOAD:
        if (ctx->opt_fuse_load != (value ? 1 : 0)) { ctx->opt_fuse_load = value ? 1 : 0; invalidate(ctx); }
        break;
    default: efg_throw(EFG_ERR_INVALID, "unknown option %d", option);
    }
    API_END(ctx)
}

int efg_get_stat(efg_ctx *ctx, int which, double *out)
{
    API_BEGIN_RAW(ctx)
    if (!out) efg_throw(EFG_ERR_INVALID, "null output");
    switch (which) {
    case EFG_STAT_SYMBOLIC_MS: *out = ctx->symbolic_ms; break;
    case EFG_STAT_NUMERIC_MS:
        if (ctx->have_values) {
            float ms = 0;
            CUDA_CHECK(cudaEventSynchronize(ctx->evn1));
            CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->evn0, ctx->evn1));
            ctx->numeric_ms = ms;
        }
        *out = ctx->numeric_ms;
        break;
    case EFG_STAT_KERNEL_LAUNCHES: *out = (double)ctx->launches; break;
    case EFG_STAT_NUMERIC_LAUNCHES: *out = (double)ctx->numeric_launches; break;
    case EFG_STAT_DEVICE_BYTES: *out = (double)ctx->pool.reserved; break;
    case EFG_STAT_NTILES: *out = (double)ctx->tl.ntiles; break;
    case EFG_STAT_TILE_ELEMS: *out = (double)ctx->tl.sum_tile_elems; break;
    case EFG_STAT_NUMERIC_BYTES: *out = (double)ctx->tl.numeric_bytes; break;
    case EFG_STAT_PATH: *out = (double)ctx->path; break;
    case EFG_STAT_VEC_MS: *out = vec_data(ctx) ? vec_data(ctx)->vec_ms : 0.0; break;
    case EFG_STAT_SPMV_MS: *out = vec_data(ctx) ? vec_data(ctx)->spmv_ms : 0.0; break;
    default: efg_throw(EFG_ERR_INVALID, "unknown stat %d", which);
    }
    API_END(ctx)
}

int efg_get_stream(efg_ctx *ctx, void **stream_out)
{
    if (!ctx || !stream_out) return EFG_ERR_INVALID;
    *stream_out = (void *)ctx->stream;
    return EFG_OK;
}

int efg_synchronize(efg_ctx *ctx)
{
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

int efg_set_mesh(efg_ctx *ctx, int slot, int kind, int64_t nel, int64_t nnodes, const int64_t *conn, const double *xy)
{
    API_BEGIN_RAW(ctx)
    if (slot < 0 || slot > 1) efg_throw(EFG_ERR_INVALID, "mesh_slot must be 0 or 1");
    if (kind != EFG_T3 && kind != EFG_Q4 && kind != EFG_T6) efg_throw(EFG_ERR_INVALID, "unsupported element kind %d", kind);
    if (nel < 0 || nnodes < 0 || (nel > 0 && (!conn || !xy))) efg_throw(EFG_ERR_INVALID, "bad mesh arguments");
    if (nnodes >= ((int64_t)1 << 31) || nel >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_LIMIT, "mesh too large for 32-bit device indices");
    invalidate(ctx);
    MeshDev &m = ctx->mesh[slot];
    m.kind = kind; m.nel = nel; m.nnodes = nnodes;
    m.conn.alloc(ctx->pool, (size_t)(nel * kind));
    m.xy.alloc(ctx->pool, (size_t)nnodes);
    m.z.release();
    m.pending_xy = nullptr;
    if (ctx->opt_defer_xy && nnodes > 0 && !is_device_ptr(xy)) m.pending_xy = xy;          // copied when the symbolic phase starts
    else if (nnodes > 0) CUDA_CHECK(cudaMemcpyAsync(m.xy.p, xy, (size_t)nnodes * sizeof(double2), cudaMemcpyDefault, ctx->stream));
    ingest_index(ctx, conn, nel * kind, 1, nnodes, m.conn.p, "efg_set_mesh: node id");
    API_END(ctx)
}

__global__ void k_split_xyz(const double *__restrict__ xyz, int64_t n, double2 *__restrict__ xy, double *__restrict__ z)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        xy[i] = make_double2(xyz[3 * i], xyz[3 * i + 1]);
        z[i] = xyz[3 * i + 2];
    }
}

/* a 3-D mesh (FEH1_T4): xyz is 3 x nnodes */
int efg_set_mesh3(efg_ctx *ctx, int slot, int kind, int64_t nel, int64_t nnodes, const int64_t *conn, const double *xyz)
{
    API_BEGIN(ctx)
    if (slot != 0) efg_throw(EFG_ERR_INVALID, "a 3-D mesh goes into mesh slot 0");
    if (kind != EFG_T4) efg_throw(EFG_ERR_INVALID, "unsupported 3-D element kind %d", kind);
    if (nel < 0 || nnodes < 0 || (nel > 0 && (!conn || !xyz))) efg_throw(EFG_ERR_INVALID, "bad mesh arguments");
    if (nnodes >= ((int64_t)1 << 31) || nel >= ((int64_t)1 << 29)) efg_throw(EFG_ERR_LIMIT, "mesh too large for 32-bit device indices");
    invalidate(ctx);
    MeshDev &m = ctx->mesh[slot];
    m.kind = kind; m.nel = nel; m.nnodes = nnodes;
    m.conn.alloc(ctx->pool, (size_t)(nel * 4));
    m.xy.alloc(ctx->pool, (size_t)nnodes);
    m.z.alloc(ctx->pool, (size_t)nnodes);
    if (nnodes > 0) {
        DevBuf<double> stage;
        const double *src = xyz;
        if (!is_device_ptr(xyz)) {
            stage.alloc(ctx->pool, (size_t)nnodes * 3);
            CUDA_CHECK(cudaMemcpyAsync(stage.p, xyz, (size_t)nnodes * 3 * sizeof(double), cudaMemcpyDefault, ctx->stream));
            src = stage.p;
        }
        LAUNCH(ctx, k_split_xyz, grid_for(nnodes, 256), 256, 0, src, nnodes, m.xy.p, m.z.p);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    ingest_index(ctx, conn, nel * 4, 1, nnodes, m.conn.p, "efg_set_mesh3: node id");
    API_END(ctx)
}

int efg_set_space(efg_ctx *ctx, int slot, int mesh_slot, int ncomp, int64_t nnodes, const int64_t *dofnums)
{
    API_BEGIN_RAW(ctx)
    if (slot < 0 || slot > 2 || mesh_slot < 0 || mesh_slot > 1) efg_throw(EFG_ERR_INVALID, "bad space/mesh slot");
    if (ncomp < 1 || ncomp > 2) efg_throw(EFG_ERR_INVALID, "ncomp must be 1 or 2");
    if (nnodes != ctx->mesh[mesh_slot].nnodes) efg_throw(EFG_ERR_INVALID, "space has %lld terms, its mesh has %lld nodes", (long long)nnodes, (long long)ctx->mesh[mesh_slot].nnodes);
    if (nnodes > 0 && !dofnums) efg_throw(EFG_ERR_INVALID, "null dofnums");
    invalidate(ctx);
    SpaceDev &s = ctx->space[slot];
    s.mesh = mesh_slot; s.ncomp = ncomp; s.nnodes = nnodes;
    s.dof.alloc(ctx->pool, (size_t)(nnodes * ncomp));
    s.isdatum.release();
    s.fe = EFG_FE_H1; s.ncells = 0; s.cdof.release();
    // dof number 0 (= not numbered) is accepted here and rejected by the symbolic phase, like sparse() does
    ingest_index(ctx, dofnums, nnodes * ncomp, 0, ((int64_t)1 << 31) - 1, s.dof.p, "efg_set_space: dof number");
    API_END(ctx)
}

/* FESpace(mesh, FEH1_T3_BUBBLE() | FEL2_T3() | FEL2_Q4(), ncomp): fields on the vertices and / or on the cells */
int efg_set_space_fe(efg_ctx *ctx, int slot, int mesh_slot, int fe, int ncomp, int64_t nnodes, const int64_t *dofnums,
                     int64_t nel, const int64_t *cell_dofnums)
{
    if (fe == EFG_FE_H1) return efg_set_space(ctx, slot, mesh_slot, ncomp, nnodes, dofnums);
    API_BEGIN(ctx)
    if (slot < 0 || slot > 2 || mesh_slot < 0 || mesh_slot > 1) efg_throw(EFG_ERR_INVALID, "bad space/mesh slot");
    if (fe != EFG_FE_T3_BUBBLE && fe != EFG_FE_L2) efg_throw(EFG_ERR_INVALID, "unknown finite element %d", fe);
    if (ncomp != 1) efg_throw(EFG_ERR_INVALID, "spaces with cell dofs have one component here (the examples use one scalar space per velocity component)");
    const MeshDev &m = ctx->mesh[mesh_slot];
    if (fe == EFG_FE_T3_BUBBLE && m.kind != EFG_T3) efg_throw(EFG_ERR_INVALID, "FEH1_T3_BUBBLE needs a T3 mesh");
    if (fe == EFG_FE_L2 && m.kind != EFG_T3 && m.kind != EFG_Q4) efg_throw(EFG_ERR_INVALID, "FEL2 needs a T3 or Q4 mesh");
    const bool nodal = fe == EFG_FE_T3_BUBBLE;
    if (nodal && nnodes != m.nnodes) efg_throw(EFG_ERR_INVALID, "space has %lld vertex terms, its mesh has %lld nodes", (long long)nnodes, (long long)m.nnodes);
    if (nel != m.nel) efg_throw(EFG_ERR_INVALID, "space has %lld cell terms, its mesh has %lld elements", (long long)nel, (long long)m.nel);
    if ((nodal && nnodes > 0 && !dofnums) || (nel > 0 && !cell_dofnums)) efg_throw(EFG_ERR_INVALID, "null dofnums");
    invalidate(ctx);
    SpaceDev &s = ctx->space[slot];
    s.mesh = mesh_slot; s.ncomp = ncomp; s.nnodes = m.nnodes; s.fe = fe; s.ncells = nel;
    s.isdatum.release();
    s.dof.release();
    if (nodal) {
        s.dof.alloc(ctx->pool, (size_t)(nnodes * ncomp));
        ingest_index(ctx, dofnums, nnodes * ncomp, 0, ((int64_t)1 << 31) - 1, s.dof.p, "efg_set_space_fe: dof number");
    }
    s.cdof.alloc(ctx->pool, (size_t)(nel * ncomp));
    ingest_index(ctx, cell_dofnums, nel * ncomp, 0, ((int64_t)1 << 31) - 1, s.cdof.p, "efg_set_space_fe: cell dof number");
    API_END(ctx)
}

int efg_start(efg_ctx *ctx, int64_t nrow, int64_t ncol)
{
    API_BEGIN_RAW(ctx)
    if (nrow < 0 || ncol < 0) efg_throw(EFG_ERR_INVALID, "negative matrix size");
    if (nrow >= ((int64_t)1 << 31) || ncol >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_LIMIT, "matrix dimension exceeds 32-bit device indices");
    const bool same = ctx->started && ctx->nrow == nrow && ctx->ncol == ncol && !ctx->have_range;
    if (!same) invalidate(ctx);
    ctx->nrow = nrow; ctx->ncol = ncol;
    ctx->c0 = 0; ctx->c1 = ncol; ctx->ncl = ncol; ctx->have_range = false; ctx->nranges = 0;
    ctx->started = true;
    ctx->have_values = false;
    API_END(ctx)
}

int efg_set_column_range(efg_ctx *ctx, int64_t first, int64_t last)
{
    API_BEGIN_RAW(ctx)
    if (!ctx->started) efg_throw(EFG_ERR_STATE, "efg_set_column_range before efg_start");
    if (first < 1 || last > ctx->ncol || last < first - 1) efg_throw(EFG_ERR_INVALID, "bad column range");
    invalidate(ctx);
    ctx->c0 = first - 1; ctx->c1 = last; ctx->ncl = last - first + 1; ctx->have_range = true; ctx->nranges = 0;
    API_END(ctx)
}

int efg_set_column_ranges(efg_ctx *ctx, int64_t nranges, const int64_t *firsts, const int64_t *lasts)
{
    API_BEGIN_RAW(ctx)
    if (!ctx->started) efg_throw(EFG_ERR_STATE, "efg_set_column_ranges before efg_start");
    if (nranges < 1 || nranges > (1 << 24) || !firsts || !lasts) efg_throw(EFG_ERR_INVALID, "bad column range list");
    std::vector<int32_t> f((size_t)nranges), l((size_t)nranges), o((size_t)nranges);
    int64_t tot = 0, prev = 0;
    for (int64_t i = 0; i < nranges; i++) {
        if (firsts[i] < 1 || lasts[i] > ctx->ncol || lasts[i] < firsts[i] || firsts[i] <= prev)
            efg_throw(EFG_ERR_INVALID, "column ranges must be non-empty, ascending and disjoint (range %lld)", (long long)i);
        f[(size_t)i] = (int32_t)(firsts[i] - 1); l[(size_t)i] = (int32_t)lasts[i]; o[(size_t)i] = (int32_t)tot;
        tot += lasts[i] - firsts[i] + 1;
        prev = lasts[i];
    }
    invalidate(ctx);
    ctx->rfirst.alloc(ctx->pool, (size_t)nranges); ctx->rlast1.alloc(ctx->pool, (size_t)nranges); ctx->roff.alloc(ctx->pool, (size_t)nranges);
    CUDA_CHECK(cudaMemcpyAsync(ctx->rfirst.p, f.data(), (size_t)nranges * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->rlast1.p, l.data(), (size_t)nranges * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->roff.p, o.data(), (size_t)nranges * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->c0 = firsts[0] - 1; ctx->c1 = lasts[nranges - 1]; ctx->ncl = tot; ctx->have_range = true; ctx->nranges = (int)nranges;
    API_END(ctx)
}

// Kind code of a form's instantiation: the mesh kind, or -- three-space forms -- the velocity / pressure pair
// (EFG_PAIR_*: T6/T3 on two meshes, or a row-f5 pair on one mesh, recognised by the spaces' elements).
static int form_kind(const efg_ctx *ctx, int form)
{
    const int k = ctx->mesh[0].kind;
    if (form != EFG_FORM_STOKES_REDDY && form != EFG_FORM_STOKES_VECLAP) return k;
    const int vfe = ctx->space[0].fe, pfe = ctx->space[2].fe;
    if (vfe == EFG_FE_T3_BUBBLE && pfe == EFG_FE_H1) return EFG_PAIR_T3B_T3;
    if (vfe == EFG_FE_H1 && pfe == EFG_FE_L2) return k == EFG_Q4 ? EFG_PAIR_Q4_L2 : EFG_PAIR_T3_L2;
    return k;
}

static void check_form_inputs(efg_ctx *ctx, int form)
{
    const MeshDev &m0 = ctx->mesh[0];
    if (m0.kind == 0) efg_throw(EFG_ERR_STATE, "mesh 0 not set");
    auto need_space = [&](int s, int mesh, int ncomp, int fe = EFG_FE_H1) {
        const SpaceDev &sp = ctx->space[s];
        if (sp.mesh != mesh || sp.ncomp != ncomp)
            efg_throw(EFG_ERR_INVALID, "form %d needs space %d on mesh %d with %d component(s)", form, s, mesh, ncomp);
        if (sp.fe != fe) efg_throw(EFG_ERR_INVALID, "form %d: space %d has finite element %d, expected %d", form, s, sp.fe, fe);
        if (sp.nnodes != ctx->mesh[mesh].nnodes)     // efg_set_mesh replaced the mesh after efg_set_space
            efg_throw(EFG_ERR_STATE, "space %d numbers %lld nodes, mesh %d now has %lld: call efg_set_space again", s, (long long)sp.nnodes,
                      mesh, (long long)ctx->mesh[mesh].nnodes);
        if (fe != EFG_FE_H1 && sp.ncells != ctx->mesh[mesh].nel)
            efg_throw(EFG_ERR_STATE, "space %d numbers %lld cells, mesh %d now has %lld: call efg_set_space_fe again", s, (long long)sp.ncells,
                      mesh, (long long)ctx->mesh[mesh].nel);
    };
    auto need_pmesh = [&]() {
        const MeshDev &m1 = ctx->mesh[1];
        if (m0.kind != EFG_T6 || m1.kind != EFG_T3 || m1.nel != m0.nel)
            efg_throw(EFG_ERR_INVALID, "Stokes forms need a T6 velocity mesh (slot 0) and a T3 pressure mesh (slot 1) with equal element counts");
    };
    switch (form) {
    case EFG_FORM_HEAT: need_space(0, 0, 1); break;
    case EFG_FORM_ELASTICITY: need_space(0, 0, 2); break;
    case EFG_FORM_STOKES_GEN:
    case EFG_FORM_STOKES_VECLAP_ALT: need_pmesh(); need_space(0, 0, 2); need_space(1, 1, 1); break;
    case EFG_FORM_STOKES_REDDY:
    case EFG_FORM_STOKES_VECLAP:
        switch (form_kind(ctx, form)) {
        case EFG_PAIR_T3B_T3: need_space(0, 0, 1, EFG_FE_T3_BUBBLE); need_space(1, 0, 1, EFG_FE_T3_BUBBLE); need_space(2, 0, 1); break;
        case EFG_PAIR_Q4_L2:
        case EFG_PAIR_T3_L2: need_space(0, 0, 1); need_space(1, 0, 1); need_space(2, 0, 1, EFG_FE_L2); break;
        default: need_pmesh(); need_space(0, 0, 1); need_space(1, 0, 1); need_space(2, 1, 1);
        }
        break;
    default: efg_throw(EFG_ERR_INVALID, "unknown form %d", form);
    }
}

// The symbolic phase in two steps.  want_tiles = false stops after the CSC pattern (colptr / rowval / nnz) on the tiled
// path, so the caller can size its arrays and start moving the pattern to the host while the tiles are built.
static void run_symbolic(efg_ctx *ctx, int form, int quad, bool want_tiles)
{
    if (!ctx->started) efg_throw(EFG_ERR_STATE, "efg_symbolic before efg_start");
    check_form_inputs(ctx, form);
    flush_deferred_xy(ctx, true);          // deferred coordinates travel while the pattern kernels run
    const bool same = ctx->form == form && ctx->quad == quad;
    if (ctx->have_symbolic && same) return;
    if (ctx->have_pattern && same && !want_tiles) return;
    const int vkind = form_kind(ctx, form);
    const int npts = quad_npts(vkind, quad);       // (the symbolic phase does not read the tables)
    if (npts < 0) efg_throw(EFG_ERR_INVALID, "quadrature rule %d not available for element kind %d", quad, vkind);
    const bool resume = ctx->have_pattern && same;      // the pattern exists, the tiles are pending
    if (!resume) { invalidate(ctx); ctx->symbolic_ms = 0; }
    CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    int path = resume ? ctx->path : (ctx->opt_path == 1 ? 1 : 2);
    ctx->form_req = form; ctx->quad_req = quad;
    bool complete = false;
    bool ok = false;
    try {
    ok = dispatch_form(form, vkind, npts, [&](auto F) {
        using Form = decltype(F);
        if (path == 2) {
            try {
                if (!resume) tiled_pattern<Form>(ctx);
                if (want_tiles) { wait_deferred_xy(ctx); tiled_symbolic<Form>(ctx); complete = true; }
            } catch (const EfgError &e) {
                // auto mode: a mesh beyond the tiled path's limits (node valence, ...) takes the general two-pass path
                if (e.code != EFG_ERR_LIMIT || ctx->opt_path == 2) throw;
                invalidate(ctx);
                path = 1;
            }
        }
        if (path == 1) { twopass_symbolic<Form>(ctx); wait_deferred_xy(ctx); ctx->have_pattern = true; complete = true; CUDA_CHECK(cudaEventRecord(ctx->ev_pattern, ctx->stream)); }
    });
    } catch (...) { invalidate(ctx); throw; }       // no half-built state survives an error
    if (!ok) efg_throw(EFG_ERR_INVALID, "form %d is not available for element kind %d with rule %d", form, vkind, quad);
    CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->symbolic_ms += ms;
    ctx->form = form; ctx->quad = quad; ctx->nq = npts; ctx->vkind = vkind; ctx->path = path;
    ctx->have_symbolic = complete;
}

int efg_pattern(efg_ctx *ctx, int form, int quad, int64_t *nnz_out)
{
    API_BEGIN_RAW(ctx)
    run_symbolic(ctx, form, quad, false);
    if (nnz_out) *nnz_out = ctx->nnz;
    API_END(ctx)
}

int efg_symbolic(efg_ctx *ctx, int form, int quad, int64_t *nnz_out)
{
    API_BEGIN_RAW(ctx)
    run_symbolic(ctx, form, quad, true);
    if (nnz_out) *nnz_out = ctx->nnz;
    API_END(ctx)
}

int efg_numeric(efg_ctx *ctx, const double *params, int nparams)
{
    API_BEGIN(ctx)
    wait_deferred_xy(ctx);
    if (!ctx->have_symbolic && ctx->have_pattern) run_symbolic(ctx, ctx->form, ctx->quad, true);     // efg_pattern came first: build the tiles now
    if (!ctx->have_symbolic) efg_throw(EFG_ERR_STATE, "efg_numeric before efg_symbolic");
    const int need = (ctx->form == EFG_FORM_ELASTICITY || ctx->form == EFG_FORM_STOKES_GEN) ? 9 : 1;
    if (!params || nparams != need) efg_throw(EFG_ERR_INVALID, "form %d takes %d parameter(s)", ctx->form, need);
    TabGuard tabs(ctx, ctx->vkind, ctx->quad, params, need);     // tables + parameters current on this device until the launches are enqueued
    ctx->numeric_launches = 0;
    CUDA_CHECK(cudaEventRecord(ctx->evn0, ctx->stream));
    dispatch_form(ctx->form, ctx->vkind, ctx->nq, [&](auto F) {
        using Form = decltype(F);
        if (ctx->path == 1) twopass_numeric<Form>(ctx); else tiled_numeric<Form>(ctx);
    });
    CUDA_CHECK(cudaEventRecord(ctx->evn1, ctx->stream));
    ctx->have_values = true;
    API_END(ctx)
}

/* K and F of one integrate! pass (heat forms, tiled path): the element load vector is staged next to the element matrix and
 * gathered through the diagonal nonzeros' contribution lists */
int efg_numeric_with_load(efg_ctx *ctx, const double *params, int nparams, double Q)
{
    API_BEGIN(ctx)
    if (!ctx->have_symbolic && ctx->have_pattern) run_symbolic(ctx, ctx->form, ctx->quad, true);
    if (!ctx->have_symbolic) efg_throw(EFG_ERR_STATE, "efg_numeric_with_load before efg_symbolic");
    if (ctx->form != EFG_FORM_HEAT || ctx->path != 2) efg_throw(EFG_ERR_INVALID, "the fused load vector needs EFG_FORM_HEAT on the tiled path");
    if (!ctx->tl.fused_rows) efg_throw(EFG_ERR_STATE, "set EFG_OPT_FUSE_LOAD = 1 before the symbolic phase");
    if (ctx->have_range || ctx->nrow != ctx->ncol) efg_throw(EFG_ERR_INVALID, "the fused load vector needs an unsharded square system");
    if (!params || nparams != 1) efg_throw(EFG_ERR_INVALID, "form %d takes 1 parameter (kappa)", ctx->form);
    VecData *vd = vec_get(ctx);
    const int64_t nrow = ctx->nrow;
    if (!(vd->nrl == nrow && vd->nrow == nrow && vd->val.n >= (size_t)(nrow > 0 ? nrow : 1))) {
        vd->have_sym = false;
        vd->val.alloc(ctx->pool, (size_t)(nrow > 0 ? nrow : 1));
        vd->nrl = nrow; vd->nrow = nrow;
    }
    vd->have_val = false;
    const double prm[2] = {params[0], Q};
    TabGuard tabs(ctx, ctx->vkind, ctx->quad, prm, 2);
    ctx->numeric_launches = 0;
    CUDA_CHECK(cudaEventRecord(ctx->evn0, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(vd->val.p, 0, (size_t)(nrow > 0 ? nrow : 1) * sizeof(double), ctx->stream));    // dofs no element touches
    dispatch_form(ctx->form, ctx->vkind, ctx->nq, [&](auto F) {
        using Form = decltype(F);
        if constexpr (form_dim3<Form>::value) efg_throw(EFG_ERR_INVALID, "the fused load vector is not available for 3-D elements");
        else tiled_numeric_fused<Form>(ctx, vd->val.p);
    });
    CUDA_CHECK(cudaEventRecord(ctx->evn1, ctx->stream));
    ctx->have_values = true;
    vd->have_val = true;
    API_END(ctx)
}

int efg_assemble(efg_ctx *ctx, int form, int quad, const double *params, int nparams, int64_t *nnz_out)
{
    int rc = efg_symbolic(ctx, form, quad, nnz_out);
    if (rc != EFG_OK) return rc;
    return efg_numeric(ctx, params, nparams);
}

// ---- result -> caller's arrays.  The pattern travels on the ctx's COPY stream (behind ev_pattern), the values on the main
// stream behind the numeric kernel: a pattern fetch issued right after efg_pattern overlaps the tile phase and the numeric
// kernel, and the PCIe link stays busy from the moment the pattern exists.
static void ensure_copy_stream(efg_ctx *ctx)
{
    if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->ev_copy) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
}
// EFG_OPT_HOST_WIDEN: 1 = row indices cross PCIe as Int32 and host threads widen them (fastest for one or two GPUs per host:
// config 2 e2e 313 -> 262 ms), 0 = widened on the device (8 ranks on one host: the host's memory system is the bottleneck and
// the widening threads add 2x the row-index bytes to it), -1 = by the number of visible devices (<= 2: host).
// The environment variable EFG_HOST_WIDEN overrides the option.
static bool host_widen_wanted(const efg_ctx *ctx)
{
    if (const char *e = getenv("EFG_HOST_WIDEN")) return atoi(e) != 0;
    if (ctx->opt_host_widen >= 0) return ctx->opt_host_widen != 0;
    int ndev = 1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 1; }
    return ndev <= 2;
}
static void enqueue_pattern_copy(efg_ctx *ctx, int64_t *colptr, int64_t *rowval)
{
    ensure_copy_stream(ctx);
    cudaStream_t cs = ctx->copy_stream;
    CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_pattern, 0));
    const int64_t ncl = ctx->ncl;
    if (colptr) CUDA_CHECK(cudaMemcpyAsync(colptr, ctx->colptr.p, (size_t)(ncl + 1) * sizeof(int64_t), cudaMemcpyDefault, cs));
    if (rowval && ctx->nnz > 0) {
        if (is_device_ptr(rowval)) {        // device destination: widen in place, no staging
            k_rowval_out<<<grid_for(ctx->nnz, 256), 256, 0, cs>>>(ctx->rowval.p, ctx->nnz, rowval);
            ctx->launches++;
            CUDA_CHECK(cudaGetLastError());
        } else if (!host_widen_wanted(ctx)) {
            // host destination, widened on the DEVICE: Int32 0-based -> Int64 1-based in two staging buffers (stream order keeps
            // a buffer from being overwritten before its copy has finished).  Twice the PCIe bytes of the path below but no host
            // memory traffic beyond the DMA writes: the better choice when many ranks share one host (EFG_OPT_HOST_WIDEN)
            const int64_t CH = (int64_t)32 << 20;
            const int64_t cap = ctx->nnz < CH ? ctx->nnz : CH;
            for (int b = 0; b < 2; b++)
                if (ctx->cstage[b].n < (size_t)cap) ctx->cstage[b].alloc(ctx->pool, (size_t)cap);
            int b = 0;
            for (int64_t o = 0; o < ctx->nnz; o += CH, b ^= 1) {
                const int64_t m = ctx->nnz - o < CH ? ctx->nnz - o : CH;
                k_rowval_out<<<grid_for(m, 256), 256, 0, cs>>>(ctx->rowval.p + o, m, ctx->cstage[b].p);
                ctx->launches++;
                CUDA_CHECK(cudaGetLastError());
                CUDA_CHECK(cudaMemcpyAsync(rowval + o, ctx->cstage[b].p, (size_t)m * sizeof(int64_t), cudaMemcpyDefault, cs));
            }
        } else {
            // host destination: the Int32 array crosses the link unwidened into the upper half of the caller's array and
            // host threads widen it in place while later chunks / the values are still travelling (efg_hostcopy.cuh)
            wait_copies(ctx);                  // (an earlier fetch into the same array has to be complete)
            HostWiden *w = new HostWiden();
            ctx->widen = w;
            w->dst = rowval; w->nnz = ctx->nnz; w->device = ctx->device;
            int ndev = 1;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 1; }
            w->nthreads = host_widen_threads(ndev);
            w->cuts = widen_cuts(ctx->nnz, (int64_t)16 << 20);      // 64 MB per copy (~1.1 ms on the link)
            const int nch = (int)w->cuts.size() - 1;
            char *up = reinterpret_cast<char *>(rowval) + 4 * ctx->nnz;
            for (int c = 0; c < nch; c++) {
                const int64_t a = w->cuts[(size_t)c], m = w->cuts[(size_t)c + 1] - a;
                CUDA_CHECK(cudaMemcpyAsync(up + 4 * a, ctx->rowval.p + a, (size_t)m * sizeof(int32_t), cudaMemcpyDefault, cs));
                if ((size_t)c >= ctx->widen_events.size()) {
                    cudaEvent_t e;
                    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
                    ctx->widen_events.push_back(e);
                }
                w->events.push_back(ctx->widen_events[(size_t)c]);
                CUDA_CHECK(cudaEventRecord(w->events.back(), cs));
            }
            for (int t = 0; t < w->nthreads; t++) w->threads.emplace_back(widen_worker, w, t);
        }
    }
    ctx->copy_pending = true;
}

int efg_fetch_pattern_async(efg_ctx *ctx, int64_t *colptr, int64_t *rowval)
{
    API_BEGIN_RAW(ctx)
    if (!ctx->have_pattern) efg_throw(EFG_ERR_STATE, "efg_fetch_pattern_async before efg_pattern / efg_symbolic");
    enqueue_pattern_copy(ctx, colptr, rowval);
    API_END(ctx)
}

int efg_fetch_csc(efg_ctx *ctx, int64_t *colptr, int64_t *rowval, double *nzval)
{
    API_BEGIN(ctx)
    if (!ctx->have_pattern) efg_throw(EFG_ERR_STATE, "efg_fetch_csc before efg_symbolic");
    if (nzval && !ctx->have_values) efg_throw(EFG_ERR_STATE, "efg_fetch_csc(nzval) before efg_numeric");
    if (colptr || rowval) enqueue_pattern_copy(ctx, colptr, rowval);
    if (nzval && ctx->nnz > 0) CUDA_CHECK(cudaMemcpyAsync(nzval, ctx->nzval.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    wait_copies(ctx);             // also completes an earlier efg_fetch_pattern_async
    if (ctx->widen_failed) { ctx->widen_failed = false; efg_throw(EFG_ERR_CUDA, "device -> host copy of the row indices failed"); }
    API_END(ctx)
}

int efg_device_csc(efg_ctx *ctx, const int64_t **colptr, const int32_t **rowval, const double **nzval)
{
    API_BEGIN(ctx)
    if (!ctx->have_pattern) efg_throw(EFG_ERR_STATE, "efg_device_csc before efg_symbolic");
    if (colptr) *colptr = ctx->colptr.p;
    if (rowval) *rowval = ctx->rowval.p;
    if (nzval) *nzval = ctx->nzval.p;
    API_END(ctx)
}

// ---- SURVEY 8f rows f1 / f2 (efg_vector.cuh) ----------------------------------------------------------
static float elapsed_sync(efg_ctx *ctx)
{
    float ms = 0;
    CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    return ms;
}

int efg_vec_assemble(efg_ctx *ctx, int vform, int quad, const double *params, int nparams, int64_t nrow)
{
    API_BEGIN(ctx)
    if (vform != EFG_VFORM_HEAT_LOAD) efg_throw(EFG_ERR_INVALID, "unknown vector form %d", vform);
    const MeshDev &m0 = ctx->mesh[0];
    if (m0.kind == 0) efg_throw(EFG_ERR_STATE, "mesh 0 not set");
    if (ctx->space[0].mesh != 0 || ctx->space[0].ncomp != 1) efg_throw(EFG_ERR_INVALID, "the heat load vector needs space 0 on mesh 0 with 1 component");
    if (ctx->space[0].fe != EFG_FE_H1) efg_throw(EFG_ERR_INVALID, "the heat load vector needs an H1 space (vertex dofs only)");
    if (ctx->space[0].nnodes != m0.nnodes) efg_throw(EFG_ERR_STATE, "space 0 numbers %lld nodes, mesh 0 now has %lld: call efg_set_space again", (long long)ctx->space[0].nnodes, (long long)m0.nnodes);
    if (!params || nparams != 1) efg_throw(EFG_ERR_INVALID, "vector form %d takes 1 parameter (Q)", vform);
    if (nrow < 0 || nrow >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_INVALID, "bad vector length");
    const int npts = quad_npts(m0.kind, quad);
    if (npts < 0) efg_throw(EFG_ERR_INVALID, "quadrature rule %d not available for element kind %d", quad, m0.kind);
    VecData *vd = vec_get(ctx);
    if (!(vd->have_sym && vd->nrow == nrow)) {
        vd->have_sym = false; vd->have_val = false;
        switch (m0.kind) {
        case EFG_T3: vec_symbolic<3>(ctx, vd, nrow); break;
        case EFG_Q4: case EFG_T4: vec_symbolic<4>(ctx, vd, nrow); break;
        default: vec_symbolic<6>(ctx, vd, nrow); break;
        }
    }
    const double Q = params[0];
    TabGuard tabs(ctx, m0.kind, quad, nullptr, 0);
    CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int key = m0.kind * 100 + npts;
    switch (key) {
    case 301: vec_numeric_heat<3, 1>(ctx, vd, Q); break;
    case 303: vec_numeric_heat<3, 3>(ctx, vd, Q); break;
    case 601: vec_numeric_heat<6, 1>(ctx, vd, Q); break;
    case 603: vec_numeric_heat<6, 3>(ctx, vd, Q); break;
    case 401: vec_numeric_heat<4, 1>(ctx, vd, Q); break;
    case 404: vec_numeric_heat<4, 4>(ctx, vd, Q); break;
    case 409: vec_numeric_heat<4, 9>(ctx, vd, Q); break;
    case EFG_T4 * 100 + 1: vec_numeric_heat_t4<1>(ctx, vd, Q); break;
    case EFG_T4 * 100 + 4: vec_numeric_heat_t4<4>(ctx, vd, Q); break;
    case EFG_T4 * 100 + 5: vec_numeric_heat_t4<5>(ctx, vd, Q); break;
    default: {
        bool done = false;          // the less common rules: run-time number of points
        vec_dispatch_kq(m0.kind, npts, [&](auto K, auto Qn) {
            if constexpr (decltype(Qn)::value == 0) { vec_numeric_heat<decltype(K)::value, 0>(ctx, vd, Q); done = true; }
        });
        if (!done) efg_throw(EFG_ERR_INVALID, "the heat load vector is not available for element kind %d with rule %d", m0.kind, quad);
    }
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    vd->vec_ms = elapsed_sync(ctx);
    vd->have_val = true;
    API_END(ctx)
}

int efg_fetch_vec(efg_ctx *ctx, double *out)
{
    API_BEGIN(ctx)
    VecData *vd = vec_data(ctx);
    if (!vd || !vd->have_val) efg_throw(EFG_ERR_STATE, "efg_fetch_vec before efg_vec_assemble");
    if (!out) efg_throw(EFG_ERR_INVALID, "null output");
    if (vd->nrl > 0) CUDA_CHECK(cudaMemcpyAsync(out, vd->val.p, (size_t)vd->nrl * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

int efg_device_vec(efg_ctx *ctx, const double **val, int64_t *n)
{
    API_BEGIN(ctx)
    VecData *vd = vec_data(ctx);
    if (!vd || !vd->have_val) efg_throw(EFG_ERR_STATE, "efg_device_vec before efg_vec_assemble");
    if (val) *val = vd->val.p;
    if (n) *n = vd->nrl;
    API_END(ctx)
}

int efg_spmv(efg_ctx *ctx, const double *x, double *y)
{
    API_BEGIN(ctx)
    if (!ctx->have_symbolic || !ctx->have_values) efg_throw(EFG_ERR_STATE, "efg_spmv before efg_numeric");
    if (ctx->have_range) efg_throw(EFG_ERR_STATE, "efg_spmv needs the whole matrix on this ctx (no column range)");
    if (!x || !y) efg_throw(EFG_ERR_INVALID, "null vector");
    VecData *vd = vec_get(ctx);
    // bitwise-symmetric K (heat forms assembled by the tiled kernel, whose element matrix is mirrored from its upper triangle): columns are rows
    const bool sym = ctx->form == EFG_FORM_HEAT && ctx->path == 2 && ctx->nrow == ctx->ncol;
    if (!sym && !vd->have_csr) vec_build_csr(ctx, vd);
    DevBuf<double> xs, ys;
    const double *xd = x;
    double *yd = y;
    if (!is_device_ptr(x)) {
        xs.alloc(ctx->pool, (size_t)ctx->ncol + 1);
        CUDA_CHECK(cudaMemcpyAsync(xs.p, x, (size_t)ctx->ncol * sizeof(double), cudaMemcpyDefault, ctx->stream));
        xd = xs.p;
    }
    if (!is_device_ptr(y)) { ys.alloc(ctx->pool, (size_t)ctx->nrow + 1); yd = ys.p; }
    CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (sym) LAUNCH(ctx, k_spmv_cols_sym, grid_for(ctx->nrow, 256, (int64_t)148 * 32), 256, 0, ctx->colptr.p, ctx->rowval.p, ctx->nzval.p, xd, ctx->nrow, yd);
    else LAUNCH(ctx, k_spmv_rows, grid_for(ctx->nrow, 256, (int64_t)148 * 32), 256, 0, vd->rowptr.p, vd->tperm.p, vd->tcol.p, ctx->nzval.p, xd, ctx->nrow, yd);
    CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    if (yd != y) CUDA_CHECK(cudaMemcpyAsync(y, yd, (size_t)ctx->nrow * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    vd->spmv_ms = elapsed_sync(ctx);
    API_END(ctx)
}

int efg_block_nnz(efg_ctx *ctx, int64_t row_first, int64_t row_last, int64_t col_first, int64_t col_last, int64_t *nnz_out)
{
    API_BEGIN(ctx)
    if (!ctx->have_symbolic) efg_throw(EFG_ERR_STATE, "efg_block_nnz before efg_symbolic");
    if (ctx->have_range) efg_throw(EFG_ERR_STATE, "efg_block_nnz needs the whole matrix on this ctx (no column range)");
    if (row_first < 1 || row_last > ctx->nrow || row_last < row_first - 1 || col_first < 1 || col_last > ctx->ncol || col_last < col_first - 1)
        efg_throw(EFG_ERR_INDEX, "BoundsError: block [%lld:%lld, %lld:%lld] of a %lld x %lld matrix", (long long)row_first, (long long)row_last,
                  (long long)col_first, (long long)col_last, (long long)ctx->nrow, (long long)ctx->ncol);
    VecData *vd = vec_get(ctx);
    const int64_t bncol = col_last - col_first + 1;
    DevBuf<int64_t> cnt;
    cnt.alloc(ctx->pool, (size_t)bncol + 2);
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, ((size_t)bncol + 2) * sizeof(int64_t), ctx->stream));
    vd->bcolptr.alloc(ctx->pool, (size_t)bncol + 2);
    vd->bfirst.alloc(ctx->pool, (size_t)bncol + 2);
    LAUNCH(ctx, k_block_count, grid_for(bncol, 256), 256, 0, ctx->colptr.p, ctx->rowval.p, col_first - 1, bncol, (int32_t)(row_first - 1), (int32_t)row_last, cnt.p, vd->bfirst.p);
    tl_excl_scan(ctx, cnt.p, vd->bcolptr.p, bncol + 1);
    vd->bnnz = tl_read(ctx, vd->bcolptr.p + bncol);
    vd->br0 = row_first - 1; vd->bc0 = col_first - 1; vd->bncol = bncol;
    vd->have_block = true;
    if (nnz_out) *nnz_out = vd->bnnz;
    API_END(ctx)
}

int efg_fetch_block(efg_ctx *ctx, int64_t *colptr, int64_t *rowval, double *nzval)
{
    API_BEGIN(ctx)
    VecData *vd = vec_data(ctx);
    if (!vd || !vd->have_block || !ctx->have_symbolic) efg_throw(EFG_ERR_STATE, "efg_fetch_block before efg_block_nnz");
    if (nzval && !ctx->have_values) efg_throw(EFG_ERR_STATE, "efg_fetch_block(nzval) before efg_numeric");
    const int64_t n = vd->bnnz, bncol = vd->bncol;
    if (colptr) {
        DevBuf<int64_t> cp;
        cp.alloc(ctx->pool, (size_t)bncol + 1);
        LAUNCH(ctx, k_block_colptr_out, grid_for(bncol + 1, 256), 256, 0, vd->bcolptr.p, bncol + 1, cp.p);
        CUDA_CHECK(cudaMemcpyAsync(colptr, cp.p, (size_t)(bncol + 1) * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    if ((rowval || nzval) && n > 0) {
        DevBuf<int64_t> rs;
        DevBuf<double> vs;
        int64_t *rd = nullptr;
        double *vdp = nullptr;
        if (rowval) { if (is_device_ptr(rowval)) rd = rowval; else { rs.alloc(ctx->pool, (size_t)n); rd = rs.p; } }
        if (nzval) { if (is_device_ptr(nzval)) vdp = nzval; else { vs.alloc(ctx->pool, (size_t)n); vdp = vs.p; } }
        LAUNCH(ctx, k_block_copy, grid_for(bncol * 32, 256), 256, 0, vd->bcolptr.p, vd->bfirst.p, bncol, ctx->rowval.p, ctx->nzval.p, vd->br0, rd, vdp);
        if (rowval && rd != rowval) CUDA_CHECK(cudaMemcpyAsync(rowval, rd, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
        if (nzval && vdp != nzval) CUDA_CHECK(cudaMemcpyAsync(nzval, vdp, (size_t)n * sizeof(double), cudaMemcpyDefault, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    API_END(ctx)
}

// ---- location(el, qp) and the evaluate_*_error integrators -------------------------------
int efg_qp_locations(efg_ctx *ctx, int mesh_slot, int quad, double *out, int64_t *npts_out)
{
    API_BEGIN(ctx)
    if (mesh_slot < 0 || mesh_slot > 1 || ctx->mesh[mesh_slot].kind == 0) efg_throw(EFG_ERR_STATE, "mesh %d not set", mesh_slot);
    const MeshDev &m = ctx->mesh[mesh_slot];
    const int npts = quad_npts(m.kind, quad);
    if (npts < 0) efg_throw(EFG_ERR_INVALID, "quadrature rule %d not available for element kind %d", quad, m.kind);
    if (npts_out) *npts_out = npts;
    if (out && m.nel > 0) {
        const size_t n = (size_t)m.nel * npts * 2;
        DevBuf<double> stage;
        double *d = out;
        if (!is_device_ptr(out)) { stage.alloc(ctx->pool, n); d = stage.p; }
        TabGuard tabs(ctx, m.kind, quad, nullptr, 0);
        if (!vec_dispatch_kq(m.kind, npts, [&](auto K, auto Q) { vec_locations<decltype(K)::value, decltype(Q)::value>(ctx, m, d); }))
            efg_throw(EFG_ERR_INVALID, "no location kernel for element kind %d with %d points", m.kind, npts);
        if (d != out) CUDA_CHECK(cudaMemcpyAsync(out, d, n * sizeof(double), cudaMemcpyDefault, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    API_END(ctx)
}

int efg_l2_error(efg_ctx *ctx, int ncomp, const int *space_slots, const int *comps, int quad, const double *U, int64_t nU,
                 const double *truth, double *out)
{
    API_BEGIN(ctx)
    if (ncomp < 1 || ncomp > 2 || !space_slots || !comps || !U || !truth || !out || nU < 0) efg_throw(EFG_ERR_INVALID, "bad efg_l2_error arguments");
    ErrComp ec[2];
    int mslot = -1;
    for (int c = 0; c < ncomp; c++) {
        if (space_slots[c] < 0 || space_slots[c] > 2 || ctx->space[space_slots[c]].mesh < 0) efg_throw(EFG_ERR_STATE, "space %d not set", space_slots[c]);
        const SpaceDev &sp = ctx->space[space_slots[c]];
        if (comps[c] < 0 || comps[c] >= sp.ncomp) efg_throw(EFG_ERR_INVALID, "space %d has no component %d", space_slots[c], comps[c]);
        if (mslot >= 0 && sp.mesh != mslot) efg_throw(EFG_ERR_INVALID, "the components of one error integral must live on the same mesh");
        if (sp.fe == EFG_FE_L2) efg_throw(EFG_ERR_INVALID, "efg_l2_error: space %d has no vertex dofs (FEL2)", space_slots[c]);
        if (sp.nnodes != ctx->mesh[sp.mesh].nnodes) efg_throw(EFG_ERR_STATE, "space %d numbers %lld nodes, its mesh now has %lld: call efg_set_space again", space_slots[c], (long long)sp.nnodes, (long long)ctx->mesh[sp.mesh].nnodes);
        mslot = sp.mesh;
        ec[c] = ErrComp{sp.dof.p, sp.ncomp, comps[c], sp.fe == EFG_FE_T3_BUBBLE ? sp.cdof.p : nullptr};
        if (c > 0 && (ec[c].cdof != nullptr) != (ec[0].cdof != nullptr)) efg_throw(EFG_ERR_INVALID, "the components of one error integral must use the same finite element");
    }
    if (ncomp == 1) ec[1] = ec[0];
    const MeshDev &m = ctx->mesh[mslot];
    if (m.kind == 0) efg_throw(EFG_ERR_STATE, "mesh %d not set", mslot);
    const int npts = quad_npts(m.kind, quad);
    if (npts < 0) efg_throw(EFG_ERR_INVALID, "quadrature rule %d not available for element kind %d", quad, m.kind);
    TabGuard tabs(ctx, m.kind, quad, nullptr, 0);
    const size_t nt = (size_t)m.nel * npts * ncomp;
    DevBuf<double> us, ts, eout, sum;
    DevBuf<int> err;
    const double *Ud = U, *Td = truth;
    if (!is_device_ptr(U)) { us.alloc(ctx->pool, (size_t)nU + 1); CUDA_CHECK(cudaMemcpyAsync(us.p, U, (size_t)nU * sizeof(double), cudaMemcpyDefault, ctx->stream)); Ud = us.p; }
    if (!is_device_ptr(truth)) { ts.alloc(ctx->pool, nt + 1); CUDA_CHECK(cudaMemcpyAsync(ts.p, truth, nt * sizeof(double), cudaMemcpyDefault, ctx->stream)); Td = ts.p; }
    eout.alloc(ctx->pool, (size_t)m.nel + 1); sum.alloc(ctx->pool, 1); err.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    if (!vec_dispatch_kq(m.kind, npts, [&](auto K, auto Q) {
            vec_l2_elem<decltype(K)::value, decltype(Q)::value>(ctx, m, ncomp, ec[0], ec[1], Ud, nU, Td, eout.p, err.p); }))
        efg_throw(EFG_ERR_INVALID, "no error integrator for element kind %d with %d points", m.kind, npts);
    {
        size_t tb = 0;
        cub::DeviceReduce::Sum(nullptr, tb, eout.p, sum.p, m.nel, ctx->stream);
        DevBuf<char> tmp;
        tmp.alloc(ctx->pool, tb);
        CUDA_CHECK(cub::DeviceReduce::Sum(tmp.p, tb, eout.p, sum.p, m.nel, ctx->stream));
        ctx->launches += 2;
    }
    if (tl_read(ctx, err.p)) efg_throw(EFG_ERR_INDEX, "BoundsError: a dof number is < 1 or exceeds the length of U");
    const double E = tl_read(ctx, sum.p);
    *out = sqrt(E);
    API_END(ctx)
}


// ---- several GPUs behind one handle (efg_multi.cuh) ---------------------------------------------------------------------
#define MAPI_BEGIN(m)                                                           \
    if (!(m)) return EFG_ERR_INVALID;                                           \
    try {
#define MAPI_END(m)                                                             \
        return EFG_OK;                                                          \
    } catch (const EfgError &e) {                                               \
        (m)->err = e.msg;                                                       \
        return e.code;                                                          \
    } catch (const std::bad_alloc &) {                                          \
        (m)->err = "host allocation failed";                                    \
        return EFG_ERR_OOM;                                                     \
    } catch (...) {                                                             \
        (m)->err = "unknown error";                                             \
        return EFG_ERR_CUDA;                                                    \
    }

int efgm_create(int ngpu, const int *devices, efg_multi **out)
{
    if (!out) return EFG_ERR_INVALID;
    *out = nullptr;
    if (ngpu < 1 || ngpu > 255) return EFG_ERR_INVALID;
    efg_multi *m = new (std::nothrow) efg_multi();
    if (!m) return EFG_ERR_OOM;
    m->dev.resize((size_t)ngpu);
    for (int i = 0; i < ngpu; i++) {
        m->dev[i].band = i;
        const int rc = efg_create(devices ? devices[i] : i, &m->dev[i].ctx);
        if (rc != EFG_OK) {
            for (int j = 0; j < i; j++) efg_destroy(m->dev[j].ctx);
            delete m;
            return rc;
        }
    }
    *out = m;
    return EFG_OK;
}

int efgm_destroy(efg_multi *m)
{
    if (!m) return EFG_ERR_INVALID;
    for (auto &d : m->dev) efg_destroy(d.ctx);
    delete m;
    return EFG_OK;
}

const char *efgm_last_error(const efg_multi *m) { return m ? m->err.c_str() : "null handle"; }
int efgm_device_count(const efg_multi *m) { return m ? (int)m->dev.size() : 0; }

int efgm_set_option(efg_multi *m, int option, int64_t value)
{
    MAPI_BEGIN(m)
    if (option < 1 || option > 7) efg_throw(EFG_ERR_INVALID, "unknown option %d", option);
    for (auto &d : m->dev) multi_ck(d.ctx, efg_set_option(d.ctx, option, value));
    MAPI_END(m)
}

int efgm_set_mesh(efg_multi *m, int slot, int kind, int64_t nel, int64_t nnodes, const int64_t *conn, const double *xy)
{
    MAPI_BEGIN(m)
    if (slot < 0 || slot > 1) efg_throw(EFG_ERR_INVALID, "mesh_slot must be 0 or 1");
    if (kind != EFG_T3 && kind != EFG_Q4 && kind != EFG_T6) efg_throw(EFG_ERR_INVALID, "unsupported element kind %d", kind);
    if (nel < 0 || nnodes < 0 || (nel > 0 && (!conn || !xy))) efg_throw(EFG_ERR_INVALID, "bad mesh arguments");
    if (nnodes >= ((int64_t)1 << 31) || nel >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_LIMIT, "mesh too large for 32-bit device indices");
    m->mesh[slot] = GlobalMesh{kind, nel, nnodes, conn, xy};
    if (slot == 0) m->mesh[1] = GlobalMesh{};           // like the single-GPU ctx: a new mesh 0 starts a new problem
    m->sharded = false; m->assembled = false;
    MAPI_END(m)
}

int efgm_set_space(efg_multi *m, int slot, int mesh_slot, int ncomp, int64_t nnodes, const int64_t *dofnums)
{
    MAPI_BEGIN(m)
    if (slot < 0 || slot > 2 || mesh_slot < 0 || mesh_slot > 1) efg_throw(EFG_ERR_INVALID, "bad space/mesh slot");
    if (ncomp < 1 || ncomp > 2) efg_throw(EFG_ERR_INVALID, "ncomp must be 1 or 2");
    if (nnodes != m->mesh[mesh_slot].nnodes) efg_throw(EFG_ERR_INVALID, "space has %lld terms, its mesh has %lld nodes", (long long)nnodes, (long long)m->mesh[mesh_slot].nnodes);
    if (nnodes > 0 && !dofnums) efg_throw(EFG_ERR_INVALID, "null dofnums");
    m->space[slot] = GlobalSpace{mesh_slot, ncomp, nnodes, dofnums};
    for (int s = slot + 1; s < 3; s++) if (slot == 0) m->space[s] = GlobalSpace{};
    m->sharded = false; m->assembled = false;
    MAPI_END(m)
}

int efgm_start(efg_multi *m, int64_t nrow, int64_t ncol)
{
    MAPI_BEGIN(m)
    if (nrow < 0 || ncol < 0) efg_throw(EFG_ERR_INVALID, "negative matrix size");
    if (nrow >= ((int64_t)1 << 31) || ncol >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_LIMIT, "matrix dimension exceeds 32-bit device indices");
    if (m->nrow != nrow || m->ncol != ncol) m->assembled = false;
    m->nrow = nrow; m->ncol = ncol; m->started = true;
    MAPI_END(m)
}

int efgm_assemble(efg_multi *m, int form, int quad, const double *params, int nparams, int64_t *nnz_out)
{
    MAPI_BEGIN(m)
    if (!m->started) efg_throw(EFG_ERR_STATE, "efgm_assemble before efgm_start");
    if (m->mesh[0].kind == 0) efg_throw(EFG_ERR_STATE, "mesh 0 not set");
    if (m->mesh[1].kind && m->mesh[1].nel != m->mesh[0].nel) efg_throw(EFG_ERR_INVALID, "the two meshes must have the same elements");
    int rc;
    if (!m->sharded) {
        int axis = 1;
        std::vector<double> cuts;
        multi_cuts(m, axis, cuts);
        rc = multi_parallel(m, [&](int i) { multi_shard_device(m, m->dev[(size_t)i], axis, cuts); });
        if (rc != EFG_OK) throw EfgError{rc, m->err};
        m->sharded = true;
    }
    rc = multi_parallel(m, [&](int i) {
        MultiDev &d = m->dev[(size_t)i];
        d.nnz = 0;
        if (d.firsts.empty()) return;          // an empty band (more devices than distinct coordinates)
        multi_ck(d.ctx, efg_start(d.ctx, m->nrow, m->ncol));
        multi_ck(d.ctx, efg_set_column_ranges(d.ctx, (int64_t)d.firsts.size(), d.firsts.data(), d.lasts.data()));
        multi_ck(d.ctx, efg_assemble(d.ctx, form, quad, params, nparams, &d.nnz));
    });
    if (rc != EFG_OK) throw EfgError{rc, m->err};
    m->nnz = 0;
    for (auto &d : m->dev) m->nnz += d.nnz;
    m->form = form; m->quad = quad; m->assembled = true;
    if (nnz_out) *nnz_out = m->nnz;
    MAPI_END(m)
}

int efgm_numeric(efg_multi *m, const double *params, int nparams)
{
    MAPI_BEGIN(m)
    if (!m->assembled) efg_throw(EFG_ERR_STATE, "efgm_numeric before efgm_assemble");
    const int rc = multi_parallel(m, [&](int i) {
        MultiDev &d = m->dev[(size_t)i];
        if (!d.firsts.empty()) multi_ck(d.ctx, efg_numeric(d.ctx, params, nparams));
    });
    if (rc != EFG_OK) throw EfgError{rc, m->err};
    MAPI_END(m)
}

int efgm_fetch_csc(efg_multi *m, int64_t *colptr, int64_t *rowval, double *nzval)
{
    MAPI_BEGIN(m)
    if (!m->assembled) efg_throw(EFG_ERR_STATE, "efgm_fetch_csc before efgm_assemble");
    if (!colptr) efg_throw(EFG_ERR_INVALID, "efgm_fetch_csc needs colptr (the blocks are placed through it)");
    if (is_device_ptr(colptr) || (rowval && is_device_ptr(rowval)) || (nzval && is_device_ptr(nzval)))
        efg_throw(EFG_ERR_INVALID, "efgm_fetch_csc fills HOST arrays (the device blocks stay available through the per-device ctx)");
    const int64_t ncol = m->ncol;
    std::vector<std::vector<int64_t>> lcp(m->dev.size());
    // per-column counts of every device block into colptr[1..ncol]
    colptr[0] = 1;
    int rc = multi_parallel(m, [&](int i) {
        MultiDev &d = m->dev[(size_t)i];
        if (d.firsts.empty()) return;
        std::vector<int64_t> &cp = lcp[(size_t)i];
        cp.resize((size_t)d.ctx->ncl + 1);
        multi_ck(d.ctx, efg_fetch_csc(d.ctx, cp.data(), nullptr, nullptr));
        int64_t k = 0;
        for (size_t r = 0; r < d.firsts.size(); r++)
            for (int64_t c = d.firsts[r]; c <= d.lasts[r]; c++, k++) colptr[c] = cp[(size_t)k + 1] - cp[(size_t)k];
    });
    if (rc != EFG_OK) throw EfgError{rc, m->err};
    for (int64_t c = 1; c <= ncol; c++) colptr[c] += colptr[c - 1];
    if (colptr[ncol] - 1 != m->nnz) efg_throw(EFG_ERR_STATE, "the device blocks do not cover every column exactly once (nnz %lld vs %lld)", (long long)(colptr[ncol] - 1), (long long)m->nnz);
    if (!rowval && !nzval) return EFG_OK;
    // every column run of every device goes straight to its place in the caller's arrays
    rc = multi_parallel(m, [&](int i) {
        MultiDev &d = m->dev[(size_t)i];
        if (d.firsts.empty()) return;
        efg_ctx *ctx = d.ctx;
        CUDA_CHECK(cudaSetDevice(ctx->device));
        const std::vector<int64_t> &cp = lcp[(size_t)i];
        const int64_t CH = (int64_t)32 << 20;
        DevBuf<int64_t> stage[2];
        if (rowval) { const int64_t cap = ctx->nnz < CH ? (ctx->nnz > 0 ? ctx->nnz : 1) : CH; stage[0].alloc(ctx->pool, (size_t)cap); stage[1].alloc(ctx->pool, (size_t)cap); }
        int b = 0;
        int64_t k = 0;
        for (size_t r = 0; r < d.firsts.size(); r++) {
            const int64_t ncols = d.lasts[r] - d.firsts[r] + 1;
            const int64_t src = cp[(size_t)k] - 1, len = cp[(size_t)(k + ncols)] - cp[(size_t)k];
            const int64_t dst = colptr[d.firsts[r] - 1] - 1;
            k += ncols;
            if (len == 0) continue;
            if (nzval) CUDA_CHECK(cudaMemcpyAsync(nzval + dst, ctx->nzval.p + src, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (rowval)
                for (int64_t o = 0; o < len; o += CH, b ^= 1) {
                    const int64_t n = len - o < CH ? len - o : CH;
                    LAUNCH(ctx, k_rowval_out, grid_for(n, 256), 256, 0, ctx->rowval.p + src + o, n, stage[b].p);
                    CUDA_CHECK(cudaMemcpyAsync(rowval + dst + o, stage[b].p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
                }
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
    if (rc != EFG_OK) throw EfgError{rc, m->err};
    MAPI_END(m)
}

/* which: EFG_STAT_*; device < 0: milliseconds -> max over the devices, counts and bytes -> sum; device >= 0: that device's ctx */
int efgm_get_stat(efg_multi *m, int which, int device, double *out)
{
    MAPI_BEGIN(m)
    if (!out) efg_throw(EFG_ERR_INVALID, "null output");
    if (device >= (int)m->dev.size()) efg_throw(EFG_ERR_INVALID, "no such device index");
    double acc = 0.0;
    for (int i = 0; i < (int)m->dev.size(); i++) {
        if (device >= 0 && i != device) continue;
        double v = 0.0;
        multi_ck(m->dev[(size_t)i].ctx, efg_get_stat(m->dev[(size_t)i].ctx, which, &v));
        const bool is_ms = which == EFG_STAT_SYMBOLIC_MS || which == EFG_STAT_NUMERIC_MS || which == EFG_STAT_VEC_MS || which == EFG_STAT_SPMV_MS;
        acc = is_ms ? (v > acc ? v : acc) : acc + v;
    }
    *out = acc;
    MAPI_END(m)
}

/* the ctx of device index i (its column block: efg_device_csc, efg_fetch_csc, ...) and the owned column ranges of that block */
int efgm_device_ctx(efg_multi *m, int device, efg_ctx **ctx_out, int64_t *nranges_out, const int64_t **firsts_out, const int64_t **lasts_out)
{
    MAPI_BEGIN(m)
    if (device < 0 || device >= (int)m->dev.size()) efg_throw(EFG_ERR_INVALID, "no such device index");
    MultiDev &d = m->dev[(size_t)device];
    if (ctx_out) *ctx_out = d.ctx;
    if (nranges_out) *nranges_out = (int64_t)d.firsts.size();
    if (firsts_out) *firsts_out = d.firsts.data();
    if (lasts_out) *lasts_out = d.lasts.data();
    MAPI_END(m)
}


// ---- meshes, EBCs and numbering made on the device (efg_gen.cuh) ---------------------------------------
int efg_gen_mesh(efg_ctx *ctx, int slot, int kind, int64_t nL, int64_t nW, double Length, double Width, double xshift, double yshift)
{
    API_BEGIN(ctx)
    if (slot < 0 || slot > 1) efg_throw(EFG_ERR_INVALID, "mesh_slot must be 0 or 1");
    if (kind != EFG_T3 && kind != EFG_Q4 && kind != EFG_T6) efg_throw(EFG_ERR_INVALID, "unsupported element kind %d", kind);
    if (nL < 1 || nW < 1 || !(Length > 0.0) || !(Width > 0.0)) efg_throw(EFG_ERR_INVALID, "bad block dimensions");
    const int64_t nv = (nL + 1) * (nW + 1);
    const int64_t nnodes = kind == EFG_T6 ? nv + (4 * nW + 1) + (nL - 1) * (3 * nW + 1) : nv;
    const int64_t nel = kind == EFG_Q4 ? nL * nW : 2 * nL * nW;
    if (nnodes >= ((int64_t)1 << 31) || nel * kind >= ((int64_t)1 << 32)) efg_throw(EFG_ERR_LIMIT, "mesh too large for 32-bit device indices");
    invalidate(ctx);
    MeshDev &m = ctx->mesh[slot];
    m.kind = kind; m.nel = nel; m.nnodes = nnodes;
    m.conn.alloc(ctx->pool, (size_t)(nel * kind));
    m.xy.alloc(ctx->pool, (size_t)nnodes);
    LAUNCH(ctx, k_gen_grid_xy, grid_for(nv, 256), 256, 0, nL, nW, Length, Width, m.xy.p);
    if (kind == EFG_T3) LAUNCH(ctx, k_gen_t3, grid_for(nL * nW, 256), 256, 0, nL, nW, m.conn.p);
    else if (kind == EFG_Q4) LAUNCH(ctx, k_gen_q4, grid_for(nL * nW, 256), 256, 0, nL, nW, m.conn.p);
    else LAUNCH(ctx, k_gen_t6, grid_for(nL * nW, 256), 256, 0, nL, nW, Length, Width, m.conn.p, m.xy.p);
    if (xshift != 0.0 || yshift != 0.0) LAUNCH(ctx, k_gen_shift, grid_for(nnodes, 256), 256, 0, m.xy.p, nnodes, xshift, yshift);
    API_END(ctx)
}

/* T6toT3: the pressure mesh of a Taylor-Hood pair -- the corner nodes of every T6 element, own vertex collection */
int efg_gen_mesh_corners(efg_ctx *ctx, int slot_dst, int slot_src)
{
    API_BEGIN(ctx)
    if (slot_dst < 0 || slot_dst > 1 || slot_src < 0 || slot_src > 1 || slot_dst == slot_src) efg_throw(EFG_ERR_INVALID, "bad mesh slots");
    const MeshDev &src = ctx->mesh[slot_src];
    if (src.kind != EFG_T6) efg_throw(EFG_ERR_STATE, "efg_gen_mesh_corners needs a T6 mesh in the source slot");
    DevBuf<int32_t> mx;
    mx.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemsetAsync(mx.p, 0, sizeof(int32_t), ctx->stream));
    MeshDev &m = ctx->mesh[slot_dst];
    invalidate(ctx);
    m.kind = EFG_T3; m.nel = src.nel;
    m.conn.alloc(ctx->pool, (size_t)(src.nel * 3 + 1));
    LAUNCH(ctx, k_gen_corners, grid_for(src.nel, 256), 256, 0, src.conn.p, src.nel, m.conn.p, mx.p);
    const int64_t nv = (int64_t)tl_read(ctx, mx.p) + 1;
    m.nnodes = nv;
    m.xy.alloc(ctx->pool, (size_t)nv);
    CUDA_CHECK(cudaMemcpyAsync(m.xy.p, src.xy.p, (size_t)nv * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
    API_END(ctx)
}

int efg_gen_space(efg_ctx *ctx, int slot, int mesh_slot, int ncomp)
{
    API_BEGIN(ctx)
    if (slot < 0 || slot > 2 || mesh_slot < 0 || mesh_slot > 1) efg_throw(EFG_ERR_INVALID, "bad space/mesh slot");
    if (ncomp < 1 || ncomp > 2) efg_throw(EFG_ERR_INVALID, "ncomp must be 1 or 2");
    if (ctx->mesh[mesh_slot].kind == 0) efg_throw(EFG_ERR_STATE, "mesh %d not set", mesh_slot);
    invalidate(ctx);
    SpaceDev &s = ctx->space[slot];
    s.mesh = mesh_slot; s.ncomp = ncomp; s.nnodes = ctx->mesh[mesh_slot].nnodes;
    s.fe = EFG_FE_H1; s.ncells = 0; s.cdof.release();
    s.dof.alloc(ctx->pool, (size_t)(s.nnodes * ncomp));
    s.isdatum.alloc(ctx->pool, (size_t)(s.nnodes * ncomp) + 1);      // (+1: the numbering scans read n + 1 flags)
    LAUNCH(ctx, k_tl_fill_i32, grid_for(s.nnodes * ncomp, 256), 256, 0, s.dof.p, s.nnodes * ncomp, -1);     // unnumbered
    CUDA_CHECK(cudaMemsetAsync(s.isdatum.p, 0, (size_t)(s.nnodes * ncomp) + 1, ctx->stream));
    API_END(ctx)
}

static SpaceDev &gen_space_of(efg_ctx *ctx, int slot)
{
    if (slot < 0 || slot > 2 || ctx->space[slot].mesh < 0 || !ctx->space[slot].isdatum.p) efg_throw(EFG_ERR_STATE, "space %d was not made by efg_gen_space", slot);
    return ctx->space[slot];
}

/* comp: 1-based component, 0 = every component */
int efg_setebc_box(efg_ctx *ctx, int space_slot, int comp, double x0, double x1, double y0, double y1)
{
    API_BEGIN(ctx)
    SpaceDev &s = gen_space_of(ctx, space_slot);
    if (comp < 0 || comp > s.ncomp) efg_throw(EFG_ERR_INVALID, "space %d has no component %d", space_slot, comp);
    invalidate(ctx);
    LAUNCH(ctx, k_gen_ebc_box, grid_for(s.nnodes, 256), 256, 0, ctx->mesh[s.mesh].xy.p, s.nnodes, s.ncomp, comp - 1, x0, x1, y0, y1, s.isdatum.p);
    API_END(ctx)
}

int efg_setebc_nodes(efg_ctx *ctx, int space_slot, int comp, int64_t n, const int64_t *node_ids)
{
    API_BEGIN(ctx)
    SpaceDev &s = gen_space_of(ctx, space_slot);
    if (comp < 0 || comp > s.ncomp) efg_throw(EFG_ERR_INVALID, "space %d has no component %d", space_slot, comp);
    if (n < 0 || (n > 0 && !node_ids)) efg_throw(EFG_ERR_INVALID, "bad node list");
    invalidate(ctx);
    if (n == 0) return EFG_OK;
    DevBuf<int64_t> ids;
    DevBuf<int> err;
    ids.alloc(ctx->pool, (size_t)n); err.alloc(ctx->pool, 1);
    CUDA_CHECK(cudaMemcpyAsync(ids.p, node_ids, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_gen_ebc_nodes, grid_for(n, 256), 256, 0, ids.p, n, s.nnodes, s.ncomp, comp - 1, s.isdatum.p, err.p);
    if (tl_read(ctx, err.p)) efg_throw(EFG_ERR_INDEX, "BoundsError: a node id is outside 1..%lld", (long long)s.nnodes);
    API_END(ctx)
}

/* numberdofs!(spaces in the order given): free dofs of space 1, 2, ... first, then the data dofs in the same nesting */
int efg_number_dofs(efg_ctx *ctx, int nspaces, const int *space_slots, int64_t *nfree_out, int64_t *ndofs_out)
{
    API_BEGIN(ctx)
    if (nspaces < 1 || nspaces > 3 || !space_slots) efg_throw(EFG_ERR_INVALID, "bad space list");
    invalidate(ctx);
    std::vector<DevBuf<int64_t>> freepos((size_t)nspaces), datapos((size_t)nspaces);
    std::vector<int64_t> nfree((size_t)nspaces), ndata((size_t)nspaces);
    for (int k = 0; k < nspaces; k++) {
        SpaceDev &s = gen_space_of(ctx, space_slots[k]);
        const int64_t n = s.nnodes * s.ncomp;
        freepos[(size_t)k].alloc(ctx->pool, (size_t)n + 1); datapos[(size_t)k].alloc(ctx->pool, (size_t)n + 1);
        cub::TransformInputIterator<int64_t, NotU8, const uint8_t *> itf(s.isdatum.p, NotU8());
        cub::TransformInputIterator<int64_t, IsU8, const uint8_t *> itd(s.isdatum.p, IsU8());
        tl_excl_scan(ctx, itf, freepos[(size_t)k].p, n + 1);
        tl_excl_scan(ctx, itd, datapos[(size_t)k].p, n + 1);
        nfree[(size_t)k] = tl_read(ctx, freepos[(size_t)k].p + n);
        ndata[(size_t)k] = n - nfree[(size_t)k];
    }
    int64_t free0 = 0, data0 = 0;
    for (int k = 0; k < nspaces; k++) data0 += nfree[(size_t)k];
    const int64_t nf_total = data0;
    for (int k = 0; k < nspaces; k++) {
        SpaceDev &s = ctx->space[space_slots[k]];
        const int64_t n = s.nnodes * s.ncomp;
        if (data0 + ndata[(size_t)k] >= ((int64_t)1 << 31)) efg_throw(EFG_ERR_LIMIT, "more than 2^31 dofs");
        LAUNCH(ctx, k_gen_numbers, grid_for(n, 256), 256, 0, s.isdatum.p, freepos[(size_t)k].p, datapos[(size_t)k].p, n, free0, data0, s.dof.p);
        free0 += nfree[(size_t)k];
        data0 += ndata[(size_t)k];
    }
    if (nfree_out) *nfree_out = nf_total;
    if (ndofs_out) *ndofs_out = data0;
    API_END(ctx)
}

/* copies of the device-resident inputs in the reference's layout (Int64 1-based conn nen x nel, xy 2 x nnodes); sizes via NULL pointers */
int efg_fetch_mesh(efg_ctx *ctx, int slot, int64_t *nel_out, int64_t *nnodes_out, int64_t *conn, double *xy)
{
    API_BEGIN(ctx)
    if (slot < 0 || slot > 1 || ctx->mesh[slot].kind == 0) efg_throw(EFG_ERR_STATE, "mesh %d not set", slot);
    const MeshDev &m = ctx->mesh[slot];
    if (nel_out) *nel_out = m.nel;
    if (nnodes_out) *nnodes_out = m.nnodes;
    if (conn) {
        if (m.kind == EFG_T4) efg_throw(EFG_ERR_INVALID, "efg_fetch_mesh: 2-D meshes only");
        const int64_t n = m.nel * m.kind;
        DevBuf<int64_t> tmp;
        int64_t *d = conn;
        if (!is_device_ptr(conn)) { tmp.alloc(ctx->pool, (size_t)n + 1); d = tmp.p; }
        LAUNCH(ctx, k_gen_index_out, grid_for(n, 256), 256, 0, m.conn.p, n, d);
        if (d != conn) CUDA_CHECK(cudaMemcpyAsync(conn, d, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    if (xy) {
        CUDA_CHECK(cudaMemcpyAsync(xy, m.xy.p, (size_t)m.nnodes * sizeof(double2), cudaMemcpyDefault, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    API_END(ctx)
}

int efg_fetch_dofnums(efg_ctx *ctx, int space_slot, int64_t *dofnums)
{
    API_BEGIN(ctx)
    if (space_slot < 0 || space_slot > 2 || ctx->space[space_slot].mesh < 0) efg_throw(EFG_ERR_STATE, "space %d not set", space_slot);
    if (!dofnums) efg_throw(EFG_ERR_INVALID, "null output");
    const SpaceDev &s = ctx->space[space_slot];
    if (!s.dof.p) efg_throw(EFG_ERR_INVALID, "space %d has no vertex dofs (FEL2)", space_slot);
    const int64_t n = s.nnodes * s.ncomp;
    DevBuf<int64_t> tmp;
    int64_t *d = dofnums;
    if (!is_device_ptr(dofnums)) { tmp.alloc(ctx->pool, (size_t)n + 1); d = tmp.p; }
    LAUNCH(ctx, k_gen_index_out, grid_for(n, 256), 256, 0, s.dof.p, n, d);
    if (d != dofnums) CUDA_CHECK(cudaMemcpyAsync(dofnums, d, (size_t)n * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

} // extern "C"
