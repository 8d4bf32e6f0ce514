/* Diagnostic entry points (host arrays in/out) that expose the dense building blocks one at a time, so
 * the GPU parity tests can localise a failure to a kernel: DMMA GEMM, blocked Cholesky, blocked TRSMs.
 * Not part of the drop-in boundary (not declared under include/). */
#include "htn_internal.h"

#include <string.h>

static int up(htn_ctx_t* ctx, void** d, const void* h, size_t bytes)
{
    int rc = 0;
    HTN_TRY(htn_dalloc(ctx, d, bytes));
    if (h) HTN_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
done:
    return rc;
}

/* C = alpha A B^T + beta C + bias, row-major host arrays with the given (even) leading dimensions */
int htn_dbg_gemm_tn(int M, int N, int K, double alpha, const double* A, size_t lda, const double* B, size_t ldb,
                    double beta, double* Cm, size_t ldc, const double* bias, int mode, int ktri, int kext0,
                    int bn)
{
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, -1, NULL, 1);
    if (rc) return rc;
    void *dA = NULL, *dB = NULL, *dC = NULL, *dbias = NULL;
    HTN_TRY(up(&ctx, &dA, A, (size_t)M * lda * 8));
    HTN_TRY(up(&ctx, &dB, B, (size_t)N * ldb * 8));
    HTN_TRY(up(&ctx, &dC, Cm, (size_t)M * ldc * 8));
    if (bias) HTN_TRY(up(&ctx, &dbias, bias, (size_t)N * 8));
    HTN_TRY(htnk_gemm_tn(ctx.stream, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, dbias, mode, ktri, kext0, bn));
    HTN_CUDA(cudaMemcpyAsync(Cm, dC, (size_t)M * ldc * 8, cudaMemcpyDeviceToHost, ctx.stream));
done:
    htn_dfree(&ctx, dbias);
    htn_dfree(&ctx, dC);
    htn_dfree(&ctx, dB);
    htn_dfree(&ctx, dA);
    {
        int rc2 = htn_ctx_end(&ctx, 1);
        if (!rc) rc = rc2;
    }
    return rc;
}

/* a: n x n symmetric (upper triangle read).  l: n x n, receives L (zeros above).  If xrows > 0, x is
 * xrows x n and is overwritten: op 1: x L^-T, op 2: x L^-1, op 3: x (L L^T)^-1.  Returns info (>0) or <0. */
int htn_dbg_potrf_solve(int n, const double* a, double* l, int op, int xrows, double* x)
{
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, -1, NULL, 1);
    if (rc) return rc;
    const size_t ld = htn_round_up((size_t)n, 16);
    void *dA = NULL, *dF = NULL, *dX = NULL, *dinfo = NULL;
    htn_factor_t fa;
    memset(&fa, 0, sizeof(fa));
    int h_info = 0;
    HTN_TRY(up(&ctx, &dA, a, (size_t)n * n * 8));
    HTN_TRY(htn_dalloc(&ctx, &dF, (size_t)n * ld * 8));
    HTN_CUDA(cudaMemsetAsync(dF, 0, (size_t)n * ld * 8, ctx.stream));
    HTN_TRY(htn_dalloc(&ctx, &dinfo, sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx.stream));
    HTN_TRY(htnk_sym_expand(ctx.stream, dA, (size_t)n, n, 0, NULL, 0, dF, ld));
    HTN_TRY(htn_factor_init(&ctx, &fa, dF, ld, n, 1));
    HTN_TRY(htn_potrf(&ctx, &fa, dinfo));
    HTN_TRY(htn_factor_make_u(&ctx, &fa));
    HTN_CUDA(cudaMemcpy2DAsync(l, (size_t)n * 8, dF, ld * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost, ctx.stream));
    HTN_CUDA(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    if (xrows > 0 && x) {
        HTN_TRY(htn_dalloc(&ctx, &dX, (size_t)xrows * ld * 8));
        HTN_CUDA(cudaMemsetAsync(dX, 0, (size_t)xrows * ld * 8, ctx.stream));
        HTN_CUDA(cudaMemcpy2DAsync(dX, ld * 8, x, (size_t)n * 8, (size_t)n * 8, xrows, cudaMemcpyHostToDevice,
                                   ctx.stream));
        if (op == 1) HTN_TRY(htn_trsm_right_lt(&ctx, &fa, dX, ld, xrows));
        if (op == 2) HTN_TRY(htn_trsm_right_l(&ctx, &fa, dX, ld, xrows));
        if (op == 3) HTN_TRY(htn_potrs_rows(&ctx, &fa, dX, ld, xrows));
        HTN_CUDA(cudaMemcpy2DAsync(x, (size_t)n * 8, dX, ld * 8, (size_t)n * 8, xrows, cudaMemcpyDeviceToHost,
                                   ctx.stream));
    }
    HTN_CUDA(cudaStreamSynchronize(ctx.stream));
done:
    htn_factor_free(&ctx, &fa);
    htn_dfree(&ctx, dX);
    htn_dfree(&ctx, dinfo);
    htn_dfree(&ctx, dF);
    htn_dfree(&ctx, dA);
    {
        int rc2 = htn_ctx_end(&ctx, 1);
        if (!rc) rc = rc2;
    }
    return rc ? rc : h_info;
}

int htnk_latency_probe(cudaStream_t st, int nthreads, double* host_out);
/* cycles per dependent op, see dev/peak.cu */
int htn_dbg_latency(int nthreads, double* out8)
{
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, -1, NULL, 1);
    if (rc) return rc;
    rc = htnk_latency_probe(ctx.stream, nthreads, out8);
    htn_ctx_end(&ctx, 1);
    return rc;
}

double htnk_dmma_sweep(cudaStream_t st, int warps_per_sm, int ilp);
double htn_dbg_dmma_sweep(int warps_per_sm, int ilp)
{
    htn_ctx_t ctx;
    if (htn_ctx_begin(&ctx, -1, NULL, 1)) return -1.0;
    double v = htnk_dmma_sweep(ctx.stream, warps_per_sm, ilp);
    htn_ctx_end(&ctx, 1);
    return v;
}

double htnk_dmma_tile_probe(cudaStream_t st, int ctas_per_sm);
double htn_dbg_dmma_tile_probe(int ctas_per_sm)
{
    htn_ctx_t ctx;
    if (htn_ctx_begin(&ctx, -1, NULL, 1)) return -1.0;
    double v = htnk_dmma_tile_probe(ctx.stream, ctas_per_sm);
    htn_ctx_end(&ctx, 1);
    return v;
}

int htnk_fill_test_spd(cudaStream_t st, double* f, size_t ld, int n);
/* milliseconds of ONE blocked Cholesky (htn_potrf) of an n x n SPD test matrix generated on the device,
 * best of `reps`, CUDA events; also returns the LAPACK info in *info_out */
double htn_dbg_potrf_time(int n, int reps, int* info_out)
{
    htn_ctx_t ctx;
    if (htn_ctx_begin(&ctx, -1, NULL, 1)) return -1.0;
    int rc = 0;
    const size_t ld = htn_round_up((size_t)n, 16);
    double* F = NULL;
    int* d_info = NULL;
    double best = -1.0;
    htn_factor_t fa;
    memset(&fa, 0, sizeof(fa));
    cudaEvent_t e0 = NULL, e1 = NULL;
    HTN_TRY(htn_dalloc(&ctx, (void**)&F, (size_t)n * ld * 8));
    HTN_TRY(htn_dalloc(&ctx, (void**)&d_info, sizeof(int)));
    HTN_TRY(htn_factor_init(&ctx, &fa, F, ld, n, 0));
    HTN_CUDA(cudaEventCreate(&e0));
    HTN_CUDA(cudaEventCreate(&e1));
    for (int r = 0; r < reps; r++) {
        HTN_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), ctx.stream));
        HTN_CUDA(cudaMemsetAsync(F, 0, (size_t)n * ld * 8, ctx.stream));
        HTN_TRY(htnk_fill_test_spd(ctx.stream, F, ld, n));
        HTN_CUDA(cudaEventRecord(e0, ctx.stream));
        HTN_TRY(htn_potrf(&ctx, &fa, d_info));
        HTN_CUDA(cudaEventRecord(e1, ctx.stream));
        HTN_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        HTN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (best < 0 || ms < best) best = ms;
    }
    if (info_out) HTN_CUDA(cudaMemcpy(info_out, d_info, sizeof(int), cudaMemcpyDeviceToHost));
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    htn_factor_free(&ctx, &fa);
    htn_dfree(&ctx, d_info);
    htn_dfree(&ctx, F);
    htn_ctx_end(&ctx, 1);
    return rc ? -1.0 : best;
}
