/* Generator-only entry points and diagnostics of include/htnorm_batch.h. */
#include "htn_internal.h"

#include <string.h>

static int streams_setup(htn_ctx_t* ctx, const htn_batch_t* b, htnk_streams_t* sv, void** d_seeds)
{
    int rc = 0;
    *d_seeds = NULL;
    sv->gen = (int)b->gen;
    sv->seed0 = b->seed0;
    sv->states = NULL;
    sv->zpre = NULL;
    sv->seeds = b->seeds;
    if (b->seeds && b->mem == HTN_MEM_HOST) {
        HTN_TRY(htn_dalloc(ctx, d_seeds, b->nsamples * 8));
        HTN_CUDA(cudaMemcpyAsync(*d_seeds, b->seeds, b->nsamples * 8, cudaMemcpyHostToDevice, ctx->stream));
        sv->seeds = (const uint64_t*)*d_seeds;
    }
done:
    return rc;
}

int htn_rng_raw_fill(const htn_batch_t* b, size_t nwords, uint64_t* out)
{
    if (!b || !out) return HTN_E_ARG;
    if ((int)b->gen != HTN_GEN_PCG64 && (int)b->gen != HTN_GEN_XRS128P) return HTN_E_ARG;
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, b->device, b->stream, b->mem == HTN_MEM_HOST);
    if (rc) return rc;
    void *d_seeds = NULL, *d_out = NULL;
    htnk_streams_t sv = {0};
    const size_t total = b->nsamples * nwords;
    if (total == 0) goto done;
    HTN_TRY(streams_setup(&ctx, b, &sv, &d_seeds));
    if (b->mem == HTN_MEM_HOST) {
        HTN_TRY(htn_dalloc(&ctx, &d_out, total * 8));
        HTN_TRY(htnk_raw_fill(ctx.stream, &sv, b->nsamples, nwords, (uint64_t*)d_out));
        HTN_CUDA(cudaMemcpyAsync(out, d_out, total * 8, cudaMemcpyDeviceToHost, ctx.stream));
    } else {
        HTN_TRY(htnk_raw_fill(ctx.stream, &sv, b->nsamples, nwords, out));
    }
done:
    htn_dfree(&ctx, d_out);
    htn_dfree(&ctx, d_seeds);
    {
        int rc2 = htn_ctx_end(&ctx, b->mem == HTN_MEM_HOST);
        if (!rc) rc = rc2;
    }
    return rc;
}

int htn_std_normal_fill(const htn_batch_t* b, size_t n, double* out, uint32_t* ndraws)
{
    if (!b || !out) return HTN_E_ARG;
    if ((int)b->gen != HTN_GEN_PCG64 && (int)b->gen != HTN_GEN_XRS128P) return HTN_E_ARG;
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, b->device, b->stream, b->mem == HTN_MEM_HOST);
    if (rc) return rc;
    void *d_seeds = NULL, *d_out = NULL, *d_nd = NULL;
    htnk_streams_t sv = {0};
    const size_t total = b->nsamples * n;
    if (total == 0) goto done;
    HTN_TRY(streams_setup(&ctx, b, &sv, &d_seeds));
    if (b->mem == HTN_MEM_HOST) {
        HTN_TRY(htn_dalloc(&ctx, &d_out, total * 8));
        if (ndraws) HTN_TRY(htn_dalloc(&ctx, &d_nd, b->nsamples * 4));
        htnk_segment_t seg = {n, (double*)d_out, n, NULL, NULL, NULL};
        HTN_TRY(htnk_normal_fill(ctx.stream, &sv, b->nsamples, 1, &seg, (uint32_t*)d_nd));
        HTN_CUDA(cudaMemcpyAsync(out, d_out, total * 8, cudaMemcpyDeviceToHost, ctx.stream));
        if (ndraws)
            HTN_CUDA(cudaMemcpyAsync(ndraws, d_nd, b->nsamples * 4, cudaMemcpyDeviceToHost, ctx.stream));
    } else {
        htnk_segment_t seg = {n, out, n, NULL, NULL, NULL};
        HTN_TRY(htnk_normal_fill(ctx.stream, &sv, b->nsamples, 1, &seg, ndraws));
    }
done:
    htn_dfree(&ctx, d_nd);
    htn_dfree(&ctx, d_out);
    htn_dfree(&ctx, d_seeds);
    {
        int rc2 = htn_ctx_end(&ctx, b->mem == HTN_MEM_HOST);
        if (!rc) rc = rc2;
    }
    return rc;
}

static double peak(int device, double ms, int tensor)
{
    htn_ctx_t ctx;
    if (htn_ctx_begin(&ctx, device, NULL, 1)) return -1.0;
    double v = tensor == 2 ? htnk_int_issue_peak_gops(ctx.stream, ms)
                           : (tensor ? htnk_dmma_peak_tflops(ctx.stream, ms) : htnk_dfma_peak_tflops(ctx.stream, ms));
    htn_ctx_end(&ctx, 1);
    return v;
}
double htn_int_issue_peak_gops(int device, double ms) { return peak(device, ms > 0 ? ms : 50.0, 2); }

double htn_fp64_tensor_peak_tflops(int device, double ms) { return peak(device, ms > 0 ? ms : 50.0, 1); }
double htn_fp64_fma_peak_tflops(int device, double ms) { return peak(device, ms > 0 ? ms : 50.0, 0); }
