/* c_abi_smoke.c -- the drop-in boundary used from plain C99, without ctypes: proves include/elfel_gpu.h is valid C and
 * that libelfelgpu.so is usable by any FFI (this is what Julia's ccall does).
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include tests/c_abi_smoke.c -o smoke -L elfel.jl_b200 -lelfelgpu -lm
 *
 * Without a CUDA device: efg_create must fail with EFG_ERR_CUDA (no CPU fallback) -> prints "NO DEVICE", exit 0.
 * With one: assembles BASELINE config 1 (heat FEH1_T3 on the 100 x 100 unit-square T3block, all boundary nodes
 * prescribed; examples/heat/poisson/t3.jl) through efg_set_mesh / efg_set_space / efg_start / efg_pattern /
 * efg_fetch_pattern_async / efg_numeric / efg_fetch_csc and checks nnz = 7N^2+6N+1, the pattern invariants of sparse()
 * (colptr monotone from 1, rows ascending inside a column, Int64 1-based), row sums = 0 and symmetry of K. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "elfel_gpu.h"

#define N 100

static int fail(efg_ctx *ctx, const char *what, int rc)
{
    fprintf(stderr, "FAILED %s: rc=%d %s\n", what, rc, ctx ? efg_last_error(ctx) : "");
    return 1;
}

int main(void)
{
    efg_ctx *ctx = NULL;
    int rc = efg_create(0, &ctx);
    if (rc == EFG_ERR_CUDA) { printf("NO DEVICE (efg_create -> EFG_ERR_CUDA, there is no CPU fallback): %s\n", efg_version()); return 0; }
    if (rc != EFG_OK) return fail(ctx, "efg_create", rc);

    const int64_t nn = (int64_t)(N + 1) * (N + 1), nel = 2 * (int64_t)N * N;
    int64_t *conn = (int64_t *)malloc((size_t)nel * 3 * sizeof(int64_t));
    double *xy = (double *)malloc((size_t)nn * 2 * sizeof(double));
    int64_t *dof = (int64_t *)malloc((size_t)nn * sizeof(int64_t));
    if (!conn || !xy || !dof) return 2;
    for (int j = 0; j <= N; j++)
        for (int i = 0; i <= N; i++) { xy[2 * (j * (N + 1) + i)] = (double)i / N; xy[2 * (j * (N + 1) + i) + 1] = (double)j / N; }
    int64_t e = 0;                                   /* T3block orientation :a, elements i outer / j inner (test/qmesh-conn.dat) */
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            const int64_t f = (int64_t)j * (N + 1) + i + 1;
            conn[3 * e] = f; conn[3 * e + 1] = f + 1; conn[3 * e + 2] = f + N + 2; e++;
            conn[3 * e] = f; conn[3 * e + 1] = f + N + 2; conn[3 * e + 2] = f + N + 1; e++;
        }
    int64_t nfree = 0, ndata = 0;                    /* numberfreedofs! then numberdatadofs! (src/FEFields.jl:137-177) */
    for (int64_t k = 0; k < nn; k++) { const int i = (int)(k % (N + 1)), j = (int)(k / (N + 1)); if (i > 0 && i < N && j > 0 && j < N) dof[k] = ++nfree; else dof[k] = 0; }
    for (int64_t k = 0; k < nn; k++) if (dof[k] == 0) dof[k] = nfree + (++ndata);

    if ((rc = efg_set_mesh(ctx, 0, EFG_T3, nel, nn, conn, xy)) != EFG_OK) return fail(ctx, "efg_set_mesh", rc);
    if ((rc = efg_set_space(ctx, 0, 0, 1, nn, dof)) != EFG_OK) return fail(ctx, "efg_set_space", rc);
    if ((rc = efg_start(ctx, nn, nn)) != EFG_OK) return fail(ctx, "efg_start", rc);
    int64_t nnz = 0;
    if ((rc = efg_pattern(ctx, EFG_FORM_HEAT, 1, &nnz)) != EFG_OK) return fail(ctx, "efg_pattern", rc);
    if (nnz != 7 * (int64_t)N * N + 6 * N + 1) { fprintf(stderr, "FAILED nnz = %lld\n", (long long)nnz); return 1; }
    int64_t *colptr = (int64_t *)malloc((size_t)(nn + 1) * sizeof(int64_t)), *rowval = (int64_t *)malloc((size_t)nnz * sizeof(int64_t));
    double *nzval = (double *)malloc((size_t)nnz * sizeof(double));
    if (!colptr || !rowval || !nzval) return 2;
    if ((rc = efg_fetch_pattern_async(ctx, colptr, rowval)) != EFG_OK) return fail(ctx, "efg_fetch_pattern_async", rc);
    const double kappa = 1.0;
    if ((rc = efg_numeric(ctx, &kappa, 1)) != EFG_OK) return fail(ctx, "efg_numeric", rc);
    if ((rc = efg_fetch_csc(ctx, NULL, NULL, nzval)) != EFG_OK) return fail(ctx, "efg_fetch_csc", rc);

    if (colptr[0] != 1 || colptr[nn] != nnz + 1) { fprintf(stderr, "FAILED colptr ends\n"); return 1; }
    double *rowsum = (double *)calloc((size_t)nn, sizeof(double));
    double maxabs = 0.0, asym = 0.0;
    for (int64_t c = 0; c < nn; c++) {
        if (colptr[c + 1] < colptr[c]) { fprintf(stderr, "FAILED colptr not monotone\n"); return 1; }
        for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; p++) {
            const int64_t r = rowval[p];
            if (r < 1 || r > nn || (p > colptr[c] - 1 && rowval[p - 1] >= r)) { fprintf(stderr, "FAILED rows of column %lld\n", (long long)c); return 1; }
            rowsum[r - 1] += nzval[p];
            if (fabs(nzval[p]) > maxabs) maxabs = fabs(nzval[p]);
            /* K[r,c] == K[c,r]: find row c+1 in column r-1 */
            int64_t lo = colptr[r - 1] - 1, hi = colptr[r] - 1;
            while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (rowval[mid] < c + 1) lo = mid + 1; else hi = mid; }
            if (lo >= colptr[r] - 1 || rowval[lo] != c + 1) { fprintf(stderr, "FAILED pattern not symmetric\n"); return 1; }
            if (fabs(nzval[lo] - nzval[p]) > asym) asym = fabs(nzval[lo] - nzval[p]);
        }
    }
    double maxrow = 0.0;
    for (int64_t r = 0; r < nn; r++) if (fabs(rowsum[r]) > maxrow) maxrow = fabs(rowsum[r]);
    if (!(maxabs > 1.0 && maxabs < 8.0) || maxrow > 1e-12 || asym > 0.0) { fprintf(stderr, "FAILED values: max %g rowsum %g asym %g\n", maxabs, maxrow, asym); return 1; }
    double ms = 0.0;
    efg_get_stat(ctx, EFG_STAT_NUMERIC_MS, &ms);
    printf("OK %s: T3 N=%d nnz=%lld max|K|=%.3f max|rowsum|=%.1e numeric %.3f ms\n", efg_version(), N, (long long)nnz, maxabs, maxrow, ms);
    efg_destroy(ctx);
    free(conn); free(xy); free(dof); free(colptr); free(rowval); free(nzval); free(rowsum);
    return 0;
}
