/* Hyperplane-truncated MVN sampler: x ~ N(mean, cov) restricted to {x : G x = r}.
 *
 * Reference path restated for a batch (src/htnorm.c:85-187, src/htnorm_distributions.c:58-88):
 *     y = mean + L z,   x = y + cov G^T (G cov G^T)^-1 (r - G y)           (Cong et al. 2017, Alg. 2)
 * With L = chol(cov), W = G L and c0 = r - G mean this is, for sample b,
 *     alpha_b = (G cov G^T)^-1 (c0 - W z_b),      x_b = mean + [ L | cov G^T ] [ z_b ; alpha_b ]
 * i.e. ONE triangular-trapezoidal DMMA GEMM over the whole batch, after a skinny GEMM for W z_b.
 * Everything that depends only on (cov, G, r, mean) - the O(k^3) factorisation the reference repeats on
 * every call - is done once per batch call.
 *
 * Device buffers (row-major, leading dimensions multiples of 16 doubles):
 *     F  : k x ldz   [ L (lower, zeros above) | 0 pad | cov G^T (k x k2) | 0 pad ],  ldz = kp + k2p
 *     Z~ : B x ldz   [ z_b (reverse fill)     | 0 pad | alpha_b          | 0 pad ]
 * so out = Z~ F^T + mean with the K range of each N tile clipped to L's non-zero part.
 */
#include "htn_internal.h"

#include <stdlib.h>
#include <string.h>

#define SMALL_K_MAX 128
#define SMALL_K2_MAX 16
#define ALPHA_K2_MAX 16

/* Per-call options live in the context (htn_internal.h):
 *   host_sink      - host-memory batches: the wrapper points it at the caller's output array; the shared dense path
 *                    then copies each finished sub-chunk device -> host on its own stream WHILE the next one is in
 *                    the sampling GEMM (the PCIe copy of the samples is the longest leg of a host-to-host call);
 *   async_info_out - device-memory batches run fully asynchronously: the LAPACK codes stay on the device, the
 *                    sampling GEMM is guarded by them, and this per-sample info array is filled by a kernel.
 *                    Host-memory callers (which must leave `out` untouched on failure and return the code) leave it
 *                    NULL and get the synchronous behaviour. */
#define SINK_CHUNK 32768 /* upper bound on the rows of a host-sink chunk */

static void streams_view(const htn_streams_t* s, size_t offset, htnk_streams_t* v)
{
    v->gen = s->gen;
    v->seeds = s->seeds ? s->seeds + offset : NULL;
    v->seed0 = s->seed0 + (uint64_t)offset;
    v->states = s->states ? s->states + 4 * offset : NULL;
    v->zpre = s->zpre; /* single-draw use only (offset 0) */
}

/* k2 x k2 system (G cov G^T) for k2 > 16: D := (c0 - D) (L_s L_s^T)^-1 through the blocked TRSMs */
static int alpha_large(htn_ctx_t* ctx, htn_factor_t* fs, double* d, size_t ldd, int nrows, int k2,
                       const double* c0)
{
    int rc = 0;
    HTN_TRY(htnk_rsub_rows(ctx->stream, d, ldd, nrows, k2, c0));
    HTN_TRY(htn_potrs_rows(ctx, fs, d, ldd, nrows));
done:
    return rc;
}

/* zero the pad / coefficient columns of Z~, then z_b into columns [0, k) back to front
 * (std_normal_rand_fill, src/htnorm_distributions.c:12-13, 79) */
static int ht_draw_chunk(cudaStream_t st, const htn_streams_t* s, size_t b0, size_t nb, int k, size_t ldz,
                         double* Z)
{
    int rc = 0;
    htnk_streams_t sv;
    streams_view(s, b0, &sv);
    HTN_CUDA(cudaMemset2DAsync(Z + k, ldz * 8, 0, (ldz - (size_t)k) * 8, nb, st));
    htnk_segment_t seg = {(size_t)k, Z, ldz, NULL, NULL, NULL};
    htnk_normal_fill_shares_device_next();
    HTN_TRY(htnk_normal_fill(st, &sv, nb, 1, &seg, NULL));
done:
    return rc;
}

static int ht_shared_dense(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, unsigned flags, int k2, int k,
                           const double* mean, const double* cov, const double* g, const double* r, double* out,
                           int* first_info)
{
    int rc = 0;
    const int lower_src = (flags & HTN_FLAG_COLMAJOR) ? 1 : 0;
    const size_t kp = htn_round_up((size_t)k, 16), k2p = htn_round_up((size_t)k2, 16), ldz = kp + k2p;
    double *F = NULL, *Sfull = NULL, *Gp = NULL, *CG = NULL, *Sm = NULL, *WT = NULL, *W = NULL, *c0 = NULL;
    double *Z[2] = {NULL, NULL}, *D = NULL;
    int* d_info = NULL;
    /* the generator runs on a side stream: it needs nothing from the factorisation, so chunk 0's normals
     * are drawn WHILE cov is being factorised, and chunk c+1's while chunk c is in the sampling GEMM */
    cudaStream_t side = NULL, side2 = NULL; /* side2: the G cov G^T branch, independent of chol(cov) */
    cudaEvent_t ev_sym = NULL, ev_sbranch = NULL;
    cudaEvent_t ev_fork = NULL, ev_rng[2] = {NULL, NULL}, ev_gemm[2] = {NULL, NULL}, ev_sub = NULL;
    int rng_pending = -1; /* buffer whose generator launch the main stream has not yet waited for */
    int side2_used = 0, copy_used = 0;
    /* A caller's own generator object (single draw, `states`) must be left untouched when chol(cov) fails, as the
     * reference leaves it (src/htnorm_distributions.c:77-79 returns before any draw): its normals are drawn only
     * after the factorisation's code is known.  Seeded batch streams have no caller-visible state and are drawn
     * up front, beside the factorisation. */
    const int defer_draw = s->states != NULL;
    htn_factor_t fa, fs;
    memset(&fa, 0, sizeof(fa));
    memset(&fs, 0, sizeof(fs));
    cudaStream_t st = ctx->stream;

    HTN_TRY(htn_dalloc(ctx, (void**)&d_info, 2 * sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(d_info, 0, 2 * sizeof(int), st));
    HTN_TRY(htn_dalloc(ctx, (void**)&F, (size_t)k * ldz * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Sfull, (size_t)k * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Gp, (size_t)k2 * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&CG, (size_t)k2 * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&W, (size_t)k2 * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&WT, (size_t)k * k2p * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Sm, (size_t)k2 * k2p * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&c0, (size_t)k2p * 8));
    /* sym_expand writes every element of the k x k squares of F and Sfull: only the columns beyond k need zeros (the
     * padding, and F's cov G^T block until its GEMM has run) - at k = 16384 that is 4 GB of memsets less */
    HTN_CUDA(cudaMemset2DAsync(F + k, ldz * 8, 0, (ldz - (size_t)k) * 8, (size_t)k, st));
    if (kp > (size_t)k) HTN_CUDA(cudaMemset2DAsync(Sfull + k, kp * 8, 0, (kp - (size_t)k) * 8, (size_t)k, st));
    HTN_CUDA(cudaMemsetAsync(Gp, 0, (size_t)k2 * kp * 8, st));
    HTN_CUDA(cudaMemsetAsync(CG, 0, (size_t)k2 * kp * 8, st));
    HTN_CUDA(cudaMemsetAsync(W, 0, (size_t)k2 * kp * 8, st));
    HTN_CUDA(cudaMemsetAsync(Sm, 0, (size_t)k2 * k2p * 8, st));

    size_t chunk = nsamples;
    const size_t budget = htn_chunk_budget(); /* bytes of Z~ per chunk (two chunks are in flight) */
    if (chunk * ldz * 8 > budget) chunk = htn_round_up(budget / (ldz * 8), 128);
    double* const sink = ctx->host_sink;
    cudaStream_t copy_st = NULL;
    /* Host-memory callers: the normals of a whole Z~ chunk are still drawn in ONE launch (every stream in parallel,
     * beside the factorisation), but the coefficient pass, the sampling GEMM and the device->host copy then go
     * through it in SUB-chunks, so that the first copy starts early and the copy engine never waits: the first
     * sub-chunk is small (about 8 MB of samples), the following ones double up to about 64 MB. */
    size_t sub0 = 0, submax = 0;
    if (sink) {
        sub0 = htn_round_up(((size_t)8 << 20) / ((size_t)k * 8) + 1, 128);
        submax = htn_round_up(((size_t)64 << 20) / ((size_t)k * 8) + 1, 128);
        const char* e = getenv("HTN_SINK_CHUNK"); /* experiments: fixed sub-chunk rows */
        if (e && atol(e) >= 128) sub0 = submax = (size_t)atol(e);
    }
    const size_t nchunks = (nsamples + chunk - 1) / chunk;
    HTN_TRY(htn_dalloc(ctx, (void**)&Z[0], chunk * ldz * 8));
    if (nchunks > 1) HTN_TRY(htn_dalloc(ctx, (void**)&Z[1], chunk * ldz * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&D, chunk * k2p * 8));
    {
        /* the side streams and events are created once per host thread and device, and kept: creating and destroying
         * three streams and eight events costs more host time than a small batch takes on the GPU */
        static _Thread_local struct {
            int device;
            cudaStream_t side, side2, copy;
            cudaEvent_t ev[8];
        } sy = {-1, NULL, NULL, NULL, {NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL}};
        if (sy.device != ctx->device) {
            if (sy.device >= 0) { /* another device was current before: let go of its objects */
                for (int e = 0; e < 8; e++) cudaEventDestroy(sy.ev[e]);
                cudaStreamDestroy(sy.side);
                cudaStreamDestroy(sy.side2);
                cudaStreamDestroy(sy.copy);
                sy.device = -1;
            }
            HTN_CUDA(cudaStreamCreateWithFlags(&sy.side, cudaStreamNonBlocking));
            HTN_CUDA(cudaStreamCreateWithFlags(&sy.side2, cudaStreamNonBlocking));
            HTN_CUDA(cudaStreamCreateWithFlags(&sy.copy, cudaStreamNonBlocking));
            for (int e = 0; e < 8; e++) HTN_CUDA(cudaEventCreateWithFlags(&sy.ev[e], cudaEventDisableTiming));
            sy.device = ctx->device;
        }
        side = sy.side;
        side2 = sy.side2;
        if (sink) copy_st = sy.copy;
        ev_fork = sy.ev[0];
        ev_sym = sy.ev[1];
        ev_sbranch = sy.ev[2];
        ev_sub = sy.ev[3];
        ev_rng[0] = sy.ev[4];
        ev_rng[1] = sy.ev[5];
        ev_gemm[0] = sy.ev[6];
        ev_gemm[1] = sy.ev[7];
    }
    htn_prof_mark(ctx, 0);
    HTN_CUDA(cudaEventRecord(ev_fork, st));
    HTN_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    if (!defer_draw) {
        HTN_TRY(ht_draw_chunk(side, s, 0, nsamples < chunk ? nsamples : chunk, k, ldz, Z[0]));
        HTN_CUDA(cudaEventRecord(ev_rng[0], side));
        rng_pending = 0;
    }

    /* ---- once per batch: everything that depends on (cov, G, r, mean) only ---- */
    /* symmetric copy of cov (as the reference's uplo='L' calls see it) and its lower triangle in F */
    HTN_TRY(htnk_sym_expand(st, cov, (size_t)k, k, lower_src, Sfull, kp, F, ldz));
    if (flags & HTN_FLAG_COLMAJOR) /* g is k2 x k column-major = k x k2 row-major */
        HTN_TRY(htnk_transpose(st, g, (size_t)k2, k, k2, Gp, kp));
    else
        HTN_TRY(htnk_copy2d(st, g, (size_t)k, k2, k, Gp, kp, NULL));
    HTN_CUDA(cudaEventRecord(ev_sym, st));
    /* ---- branch on side2 (needs cov and G only, runs under the factorisation): cov G^T, G cov G^T and
     * its Cholesky, r - G mean ---- */
    HTN_CUDA(cudaStreamWaitEvent(side2, ev_sym, 0));
    side2_used = 1;
    /* cov G^T -> F[:, kp : kp+k2]   (dsymm, src/htnorm.c:154; dsymv :73).  Disjoint from L's columns. */
    HTN_TRY(htnk_gemm_tn(side2, k, k2, k, 1.0, Sfull, kp, Gp, kp, 0.0, F + kp, ldz, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    HTN_TRY(htnk_transpose(side2, F + kp, ldz, k, k2, CG, kp));
    /* G cov G^T   (dgemm TN, src/htnorm.c:164; ddot :77) */
    HTN_TRY(htnk_gemm_tn(side2, k2, k2, k, 1.0, Gp, kp, CG, kp, 0.0, Sm, k2p, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    HTN_TRY(htnk_r_minus_gmean(side2, Gp, kp, k2, k, mean, r, c0));
    {
        htn_ctx_t c2 = *ctx; /* same device, the branch's stream */
        c2.stream = side2;
        if (k2 > 1) { /* dposv's factorisation, src/htnorm.c:169; k2 == 1 divides by the scalar (:77) */
            HTN_TRY(htn_factor_init(&c2, &fs, Sm, k2p, k2, k2 > ALPHA_K2_MAX));
            HTN_TRY(htn_potrf(&c2, &fs, d_info + 1));
            if (k2 > ALPHA_K2_MAX) HTN_TRY(htn_factor_make_u(&c2, &fs));
        }
    }
    HTN_CUDA(cudaEventRecord(ev_sbranch, side2));
    /* ---- main stream: chol(cov)  (dpotrf, src/htnorm_distributions.c:77) ---- */
    HTN_TRY(htn_factor_init(ctx, &fa, F, ldz, k, 0));
    HTN_TRY(htn_potrf(ctx, &fa, d_info));
    /* join the G cov G^T branch (it reads Sfull, which is recycled next) */
    HTN_CUDA(cudaStreamWaitEvent(st, ev_sbranch, 0));
    if (k2 <= 16) {
        /* W = G L: a few rows against the triangle, column-parallel with a fixed-order reduction; the partial sums
         * park in the Sfull buffer (free from here on) when they fit: chunks * k2 * kp <= k * kp doubles */
        const size_t need = (size_t)htnk_wl_chunks(k) * (size_t)k2 * kp;
        if (need <= (size_t)k * kp) {
            HTN_TRY(htnk_wl(st, Gp, kp, k2, F, ldz, k, Sfull, kp, W, kp));
        } else {
            double* part = NULL;
            HTN_TRY(htn_dalloc(ctx, (void**)&part, need * 8));
            rc = htnk_wl(st, Gp, kp, k2, F, ldz, k, part, kp, W, kp);
            htn_dfree(ctx, part);
            if (rc) goto done;
        }
    } else {
        /* W = G L  via  W^T = L^T G^T  (L^T parked in the Sfull buffer, free from here on) */
        HTN_TRY(htnk_transpose_lower(st, F, ldz, k, Sfull, kp));
        HTN_TRY(htnk_gemm_tn(st, k, k2, k, 1.0, Sfull, kp, Gp, kp, 0.0, WT, k2p, NULL, HTNK_GEMM_FULL, 0, 0, 0));
        HTN_TRY(htnk_transpose(st, WT, k2p, k, k2, W, kp));
    }

    int h_info[2] = {0, 0};
    const int async = !defer_draw && k2 <= ALPHA_K2_MAX && ctx->async_info_out != NULL;
    if (!async) {
        HTN_CUDA(cudaMemcpyAsync(h_info, d_info, sizeof(h_info), cudaMemcpyDeviceToHost, st));
        HTN_CUDA(cudaStreamSynchronize(st));
    }
    if (h_info[0]) { /* cov not positive definite: no draw, out untouched (src/htnorm_distributions.c:78) */
        *first_info = h_info[0];
        ctx->cov_failed = 1;
        goto done;
    }
    *first_info = h_info[1]; /* G cov G^T not PD: the unprojected y is returned (src/htnorm.c:169-177) */
    if (defer_draw) { /* chol(cov) succeeded (the code was read above: a `states` call is never asynchronous) */
        HTN_TRY(ht_draw_chunk(st, s, 0, nsamples < chunk ? nsamples : chunk, k, ldz, Z[0]));
        HTN_CUDA(cudaEventRecord(ev_rng[0], st));
    }
    htn_prof_mark(ctx, 1);

    /* ---- per chunk of samples ---- */
    for (size_t c = 0; c < nchunks; c++) {
        const size_t b0 = c * chunk;
        const size_t nb = nsamples - b0 < chunk ? nsamples - b0 : chunk;
        const int buf = (int)(c & 1);
        double* Zc = Z[buf];
        if (c + 1 < nchunks) { /* draw the next chunk on the side stream, into the other buffer */
            const size_t b1 = b0 + chunk;
            const int nbuf = buf ^ 1;
            if (c >= 1) HTN_CUDA(cudaStreamWaitEvent(side, ev_gemm[nbuf], 0)); /* its previous user is done */
            HTN_TRY(ht_draw_chunk(side, s, b1, nsamples - b1 < chunk ? nsamples - b1 : chunk, k, ldz, Z[nbuf]));
            HTN_CUDA(cudaEventRecord(ev_rng[nbuf], side));
        }
        HTN_CUDA(cudaStreamWaitEvent(st, ev_rng[buf], 0));
        rng_pending = (c + 1 < nchunks) ? (buf ^ 1) : -1;
        htn_prof_mark(ctx, 2);
        size_t sub = sink ? sub0 : nb;
        for (size_t s0 = 0; s0 < nb; s0 += sub, sub = sub * 2 < submax ? sub * 2 : (submax ? submax : sub)) {
            const size_t sn = nb - s0 < sub ? nb - s0 : sub;
            double* Zs = Zc + s0 * ldz;
            double* outs = out + (b0 + s0) * (size_t)k;
            if (!h_info[1]) {
                if (k2 <= 4) {
                    /* one pass over Z~: D = Z W^T (the G y of src/htnorm.c:132 without the mean part) and the
                     * k2 x k2 solve (dposv :169), fused */
                    HTN_TRY(htnk_ht_coeff(st, sn, k, k2, Zs, ldz, kp, W, kp, c0, Sm, k2p, async ? d_info + 1 : NULL));
                } else {
                    /* D = Z W^T   (the G y of src/htnorm.c:132 without the mean part) */
                    HTN_TRY(htnk_gemm_tn(st, (int)sn, k2, k, 1.0, Zs, ldz, W, kp, 0.0, D, k2p, NULL, HTNK_GEMM_FULL, 0, 0,
                                         0));
                    if (k2 <= ALPHA_K2_MAX) {
                        HTN_TRY(htnk_ht_alpha(st, sn, k2, D, k2p, c0, Sm, k2p, async ? d_info + 1 : NULL, Zs, ldz, kp));
                    } else {
                        HTN_TRY(alpha_large(ctx, &fs, D, k2p, (int)sn, k2, c0));
                        HTN_TRY(htnk_copy2d(st, D, k2p, (int)sn, k2, Zs + kp, ldz, NULL));
                    }
                }
            }
            htn_prof_mark(ctx, 3);
            /* out = Z~ F^T + mean   (dtrmv + daxpy, src/htnorm_distributions.c:81-83; dgemv, src/htnorm.c:176) */
            if (async) htnk_gemm_guard_next(d_info); /* cov not PD: the launch is a no-op, out stays untouched */
            HTN_TRY(htnk_gemm_tn(st, (int)sn, k, (int)(kp + (size_t)k2), 1.0, Zs, ldz, F, ldz, 0.0, outs, (size_t)k, mean,
                                 HTNK_GEMM_B_LOWER, k, (int)kp, 0));
            htn_prof_mark(ctx, 4);
            if (sink) { /* ship this sub-chunk home while the next one is being computed */
                HTN_CUDA(cudaEventRecord(ev_sub, st));
                HTN_CUDA(cudaStreamWaitEvent(copy_st, ev_sub, 0));
                copy_used = 1;
                HTN_CUDA(cudaMemcpyAsync(sink + (b0 + s0) * (size_t)k, outs, sn * (size_t)k * 8, cudaMemcpyDeviceToHost,
                                         copy_st));
            }
        }
        HTN_CUDA(cudaEventRecord(ev_gemm[buf], st));
    }
    if (sink) {
        HTN_CUDA(cudaStreamSynchronize(copy_st));
        copy_used = 0;
        ctx->host_sink_used = 1;
    }
    if (async) HTN_TRY(htnk_fill_info(st, ctx->async_info_out, nsamples, d_info));
done:
    /* Never free a buffer another stream may still be using.  The frees below are ordered on the main stream, so
     * whatever is queued on the side streams - in whatever state an error exit left them - is joined first: a FRESH
     * event per stream (a stale one from an earlier call would be a no-op), a full synchronisation if that fails. */
    if (rng_pending >= 0 && ev_rng[rng_pending] &&
        (cudaEventRecord(ev_rng[rng_pending], side) != cudaSuccess ||
         cudaStreamWaitEvent(st, ev_rng[rng_pending], 0) != cudaSuccess))
        cudaStreamSynchronize(side);
    if (side2_used && (cudaEventRecord(ev_sbranch, side2) != cudaSuccess ||
                       cudaStreamWaitEvent(st, ev_sbranch, 0) != cudaSuccess))
        cudaStreamSynchronize(side2);
    if (copy_used) cudaStreamSynchronize(copy_st); /* error exit with device -> host copies of d_out in flight */
    htn_factor_free(ctx, &fa);
    htn_factor_free(ctx, &fs);
    htn_dfree(ctx, Z[0]);
    htn_dfree(ctx, Z[1]);
    htn_dfree(ctx, D);
    /* (streams and events belong to the per-thread cache) */
    htn_dfree(ctx, c0);
    htn_dfree(ctx, Sm);
    htn_dfree(ctx, WT);
    htn_dfree(ctx, W);
    htn_dfree(ctx, CG);
    htn_dfree(ctx, Gp);
    htn_dfree(ctx, Sfull);
    htn_dfree(ctx, F);
    htn_dfree(ctx, d_info);
    return rc;
}

/* diagonal covariance (conf->diag): no factorisation, HBM-bound elementwise kernels
 * (src/htnorm_distributions.c:62-66, src/htnorm.c:68-71, 138-148) */
static int ht_shared_diag(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, unsigned flags, int k2, int k,
                          const double* mean, const double* cov, const double* g, const double* r, double* out,
                          int* first_info)
{
    int rc = 0;
    const size_t kp = htn_round_up((size_t)k, 16), k2p = htn_round_up((size_t)k2, 16);
    double *var = NULL, *sd = NULL, *Gp = NULL, *CG = NULL, *CGT = NULL, *Sm = NULL, *D = NULL;
    int* d_info = NULL;
    htn_factor_t fs;
    memset(&fs, 0, sizeof(fs));
    cudaStream_t st = ctx->stream;
    HTN_TRY(htn_dalloc(ctx, (void**)&d_info, sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
    HTN_TRY(htn_dalloc(ctx, (void**)&var, kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&sd, kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Gp, (size_t)k2 * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&CG, (size_t)k2 * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&CGT, (size_t)k * k2p * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Sm, (size_t)k2 * k2p * 8));
    HTN_CUDA(cudaMemsetAsync(Gp, 0, (size_t)k2 * kp * 8, st));
    HTN_CUDA(cudaMemsetAsync(CG, 0, (size_t)k2 * kp * 8, st));
    HTN_CUDA(cudaMemsetAsync(Sm, 0, (size_t)k2 * k2p * 8, st));
    HTN_TRY(htnk_diag_extract(st, cov, (size_t)k, k, 0, var));
    HTN_TRY(htnk_diag_extract(st, cov, (size_t)k, k, 1, sd));
    if (flags & HTN_FLAG_COLMAJOR)
        HTN_TRY(htnk_transpose(st, g, (size_t)k2, k, k2, Gp, kp));
    else
        HTN_TRY(htnk_copy2d(st, g, (size_t)k, k2, k, Gp, kp, NULL));
    HTN_TRY(htnk_colmul(st, Gp, kp, k2, k, var, CG, kp));
    HTN_TRY(htnk_transpose(st, CG, kp, k2, k, CGT, k2p));
    HTN_TRY(htnk_gemm_tn(st, k2, k2, k, 1.0, Gp, kp, CG, kp, 0.0, Sm, k2p, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    if (k2 > 1) {
        HTN_TRY(htn_factor_init(ctx, &fs, Sm, k2p, k2, k2 > ALPHA_K2_MAX));
        HTN_TRY(htn_potrf(ctx, &fs, d_info));
        if (k2 > ALPHA_K2_MAX) HTN_TRY(htn_factor_make_u(ctx, &fs));
    }
    int h_info = 0;
    HTN_CUDA(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    HTN_CUDA(cudaStreamSynchronize(st));
    *first_info = h_info;

    HTN_TRY(htn_dalloc(ctx, (void**)&D, nsamples * k2p * 8));
    htnk_streams_t sv;
    streams_view(s, 0, &sv);
    htnk_segment_t seg = {(size_t)k, out, (size_t)k, NULL, sd, mean};
    HTN_TRY(htnk_normal_fill(st, &sv, nsamples, 1, &seg, NULL)); /* y = mean + sqrt(cov_ii) z, straight into out */
    if (!h_info) {
        HTN_TRY(htnk_rows_dot(st, nsamples, k, k2, Gp, kp, out, (size_t)k, D, k2p));
        if (k2 <= ALPHA_K2_MAX)
            HTN_TRY(htnk_ht_alpha(st, nsamples, k2, D, k2p, r, Sm, k2p, NULL, D, k2p, 0));
        else
            HTN_TRY(alpha_large(ctx, &fs, D, k2p, (int)nsamples, k2, r));
        HTN_TRY(htnk_ht_diag_project(st, nsamples, k, k2, CGT, k2p, D, k2p, out, (size_t)k));
    }
done:
    htn_factor_free(ctx, &fs);
    htn_dfree(ctx, D);
    htn_dfree(ctx, Sm);
    htn_dfree(ctx, CGT);
    htn_dfree(ctx, CG);
    htn_dfree(ctx, Gp);
    htn_dfree(ctx, sd);
    htn_dfree(ctx, var);
    htn_dfree(ctx, d_info);
    return rc;
}

/* Independent problems of mid size (128 < k <= 1024, k2 <= 4, dense cov), ONE sample each: all problems go through one
 * launch sequence - per 64-column panel step a warp-per-problem factorisation of the diagonal blocks, a batched panel solve and
 * a batched rank-64 update - then one projection kernel (dev/mid_batched.cu).  Replaces the host loop of one shared-path
 * call per problem (>= 0.5 ms each).  Reference per problem: src/htnorm.c:85-187, src/htnorm_distributions.c:58-88. */
#define MID_K_MAX 1024
#define MID_K2_MAX 4
/* the launch sequence of one range of problems on one stream (workspace pointers already offset to the range) */
static int ht_mid_range(cudaStream_t st, size_t np, unsigned flags, int k2, int k, const double* mean, const double* cov,
                        const double* g, const double* r, const double* zs, double* F, size_t ld, size_t sF, double* rd,
                        size_t sr, unsigned* ctr, double* out, int* info)
{
    int rc = 0;
    const int nsteps = (k + 63) / 64, lo = (flags & HTN_FLAG_COLMAJOR) ? 1 : 0;
    /* left-looking: every panel is formed from the caller's covariance and the panels before it, written once */
    for (int q = 0; q < nsteps; q++) {
        const int j0 = 64 * q, nb = k - j0 < 64 ? k - j0 : 64;
        /* the first diagonal block is the caller's own: the factorisation reads it in place */
        if (q > 0) HTN_TRY(htnk_chol_rows_batched(st, cov, (size_t)k * k, lo, F, ld, sF, k, j0, rd, sr, info, 1, np));
        HTN_TRY(htnk_potf64_batched(st, F, ld, sF, j0, nb, rd, sr, info, np, ctr + q, q == 0 ? cov : NULL, (size_t)k * k,
                                    (size_t)k, lo));
        HTN_TRY(htnk_chol_rows_batched(st, cov, (size_t)k * k, lo, F, ld, sF, k, j0, rd, sr, info, 0, np));
    }
    HTN_TRY(htnk_ht_mid_project(st, np, k, k2, F, ld, sF, mean, g, r, zs, lo, out, info));
done:
    return rc;
}

#define MID_PARTS_MAX 8
static int ht_independent_mid(htn_ctx_t* ctx, const htn_streams_t* s, size_t P, unsigned flags, int k2, int k,
                              const double* mean, const double* cov, const double* g, const double* r, double* out,
                              int* info_dev)
{
    int rc = 0;
    cudaStream_t st = ctx->stream;
    const size_t ld = htn_round_up((size_t)k, 16), sF = (size_t)k * ld, sr = htn_round_up((size_t)k, 64);
    const int nsteps = (k + 63) / 64;
    double *F = NULL, *rd = NULL, *zs = NULL;
    unsigned* ctr = NULL;
    int side_used = 0;
    /* Two halves of a large batch go down two streams: the diagonal-block kernel is one latency-bound warp per problem
     * (eight warps per SM), so the other half's panel kernels fill the machine beside it. */
    static _Thread_local struct { int device; cudaStream_t side; cudaEvent_t ev_fork, ev_join; } ms = {-1, NULL, NULL, NULL};
    size_t split_min = 1024;
    int mid_parts = 2;
    {
        const char* e = getenv("HTN_MID_SPLIT"); /* experiments: smallest batch that is split (0 = never) */
        if (e) split_min = (size_t)atol(e);
        e = getenv("HTN_MID_PARTS");
        if (e && atoi(e) >= 1 && atoi(e) <= MID_PARTS_MAX) mid_parts = atoi(e);
    }
    /* the workspace holds a full k x ld factor per problem: very large batches go through in slices of <= 2 GiB of factors */
    size_t slice_bytes = (size_t)2 << 30;
    {
        const char* e = getenv("HTN_MID_SLICE_BYTES"); /* the tests shrink it to force the multi-slice path */
        if (e && atoll(e) > 0) slice_bytes = (size_t)atoll(e);
    }
    size_t slice = slice_bytes / (sF * 8);
    if (slice < 1) slice = 1;
    if (slice > P) slice = P;
    if (split_min && P >= split_min && ms.device != ctx->device) {
        if (ms.device >= 0) {
            cudaEventDestroy(ms.ev_fork);
            cudaEventDestroy(ms.ev_join);
            cudaStreamDestroy(ms.side);
            ms.device = -1;
        }
        HTN_CUDA(cudaStreamCreateWithFlags(&ms.side, cudaStreamNonBlocking));
        HTN_CUDA(cudaEventCreateWithFlags(&ms.ev_fork, cudaEventDisableTiming));
        HTN_CUDA(cudaEventCreateWithFlags(&ms.ev_join, cudaEventDisableTiming));
        ms.device = ctx->device;
    }
    HTN_TRY(htn_dalloc(ctx, (void**)&F, slice * sF * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&rd, slice * sr * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&zs, slice * (size_t)k * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&ctr, (size_t)MID_PARTS_MAX * nsteps * sizeof(unsigned)));
    for (size_t p0 = 0; p0 < P && !rc; p0 += slice) {
        const size_t np = P - p0 < slice ? P - p0 : slice;
        htnk_streams_t sv;
        streams_view(s, p0, &sv);
        htnk_segment_t seg = {(size_t)k, zs, (size_t)k, NULL, NULL, NULL};
        HTN_TRY(htnk_normal_fill(st, &sv, np, 1, &seg, NULL));
        HTN_CUDA(cudaMemsetAsync(info_dev + p0, 0, np * sizeof(int), st));
        HTN_CUDA(cudaMemsetAsync(ctr, 0, (size_t)MID_PARTS_MAX * nsteps * sizeof(unsigned), st));
        const int nparts = (split_min && np >= split_min) ? mid_parts : 1;
        if (nparts > 1) {
            HTN_CUDA(cudaEventRecord(ms.ev_fork, st));
            HTN_CUDA(cudaStreamWaitEvent(ms.side, ms.ev_fork, 0));
            side_used = 1;
        }
        for (int h = 0; h < nparts && !rc; h++) { /* even parts on the caller's stream, odd parts on the side stream */
            const size_t q0 = np * (size_t)h / nparts, nq = np * (size_t)(h + 1) / nparts - q0, a0 = p0 + q0;
            rc = ht_mid_range((h & 1) ? ms.side : st, nq, flags, k2, k, mean + a0 * (size_t)k, cov + a0 * (size_t)k * k,
                              g + a0 * (size_t)k2 * k, r + a0 * (size_t)k2, zs + q0 * (size_t)k, F + q0 * sF, ld, sF,
                              rd + q0 * sr, sr, ctr + h * nsteps, out + a0 * (size_t)k, info_dev + a0);
        }
        if (nparts > 1) { /* the slice's workspace is reused by the next slice, and freed in stream order on st */
            if (cudaEventRecord(ms.ev_join, ms.side) != cudaSuccess || cudaStreamWaitEvent(st, ms.ev_join, 0) != cudaSuccess) {
                cudaStreamSynchronize(ms.side);
                if (!rc) rc = HTN_E_CUDA;
            }
            side_used = 0;
        }
    }
done:
    if (side_used) cudaStreamSynchronize(ms.side); /* an error between fork and join */
    htn_dfree(ctx, ctr);
    htn_dfree(ctx, zs);
    htn_dfree(ctx, rd);
    htn_dfree(ctx, F);
    return rc;
}

int htn_ht_device(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, int independent, unsigned flags,
                  size_t k2, size_t k, const double* mean, const double* cov, const double* g, const double* r,
                  int diag, double* out, int* d_info_out, int* first_info)
{
    int rc = 0;
    *first_info = 0;
    ctx->cov_failed = 0;
    if (nsamples == 0) return 0;
    if (k == 0 || k2 == 0 || !mean || !cov || !g || !r || !out || k > 0x7fffffffu / 2 || k2 > k ||
        nsamples > 0x7fffffffu) {
        htn_set_error("hyperplane: bad argument (k=%zu k2=%zu nsamples=%zu)", k, k2, nsamples);
        return HTN_E_ARG;
    }
    if (independent) {
        /* (64, 128]: the CTA-per-problem kernel of round 1, unless the batched mid-size path is asked to take over from a
         * smaller order (HTN_MID_MIN_K, experiments; it needs k2 <= 4, a dense cov and at least two problems) */
        static int mid_min_k = -1;
        if (mid_min_k < 0) {
            const char* e = getenv("HTN_MID_MIN_K");
            mid_min_k = e && atoi(e) > 64 ? atoi(e) : SMALL_K_MAX + 1;
        }
        const int mid_takes = !diag && (int)k >= mid_min_k && k2 <= MID_K2_MAX && nsamples >= 2 && !s->states && !s->zpre;
        if (k <= SMALL_K_MAX && k2 <= SMALL_K2_MAX && !mid_takes) {
            int* info_dev = d_info_out;
            int* tmp = NULL;
            double* zs = NULL;
            if (!info_dev) {
                HTN_TRY(htn_dalloc(ctx, (void**)&tmp, nsamples * sizeof(int)));
                info_dev = tmp;
            }
            /* every problem's normals first, one stream per thread (fully parallel), then one CTA per
             * problem for the algebra */
            rc = htn_dalloc(ctx, (void**)&zs, nsamples * k * sizeof(double));
            if (!rc) {
                htnk_streams_t sv;
                streams_view(s, 0, &sv);
                htnk_segment_t seg = {k, zs, k, NULL, NULL, NULL};
                rc = htnk_normal_fill(ctx->stream, &sv, nsamples, 1, &seg, NULL);
            }
            if (!rc) {
                /* the warp-per-problem kernel for the shapes it is specialised for, else the generic CTA-per-problem one */
                unsigned* ctr = NULL;
                rc = 1;
                if (!diag && !(flags & HTN_FLAG_COLMAJOR) && k <= 64 && (k & 7) == 0 && k2 <= 4) {
                    rc = htn_dalloc(ctx, (void**)&ctr, sizeof(unsigned));
                    if (!rc && cudaMemsetAsync(ctr, 0, sizeof(unsigned), ctx->stream) != cudaSuccess) rc = HTN_E_CUDA;
                    if (!rc)
                        rc = htnk_ht_small_warp(ctx->stream, nsamples, (int)k, (int)k2, mean, cov, g, r, zs, out, info_dev,
                                                ctr);
                    htn_dfree(ctx, ctr);
                }
                if (rc == 1)
                    rc = htnk_ht_small_batched(ctx->stream, nsamples, (int)k, (int)k2, mean, cov, g, r, zs, diag,
                                               (flags & HTN_FLAG_COLMAJOR) ? 1 : 0, out, info_dev);
            }
            htn_dfree(ctx, zs);
            htn_dfree(ctx, tmp);
            *first_info = 0; /* per-problem codes live in the device array */
            return rc;
        }
        if (!diag && k <= MID_K_MAX && k2 <= MID_K2_MAX && nsamples >= 2 && !s->states && !s->zpre &&
            !(getenv("HTN_MID_BATCHED") && atoi(getenv("HTN_MID_BATCHED")) == 0)) {
            /* mid-size problems, batched: one launch sequence over all of them */
            int* info_dev = d_info_out;
            int* tmp = NULL;
            if (!info_dev) {
                HTN_TRY(htn_dalloc(ctx, (void**)&tmp, nsamples * sizeof(int)));
                info_dev = tmp;
            }
            rc = ht_independent_mid(ctx, s, nsamples, flags, (int)k2, (int)k, mean, cov, g, r, out, info_dev);
            htn_dfree(ctx, tmp);
            *first_info = 0; /* per-problem codes live in the device array */
            return rc;
        }
        /* large independent problems: one factorisation each, through the shared-input path */
        for (size_t i = 0; i < nsamples && !rc; i++) {
            htn_streams_t si = *s;
            si.seeds = s->seeds ? s->seeds + i : NULL;
            si.seed0 = s->seed0 + i;
            si.states = s->states ? s->states + 4 * i : NULL;
            si.zpre = NULL;
            int fi = 0;
            rc = htn_ht_device(ctx, &si, 1, 0, flags, k2, k, mean + i * k, cov + i * k * k, g + i * k2 * k,
                               r + i * k2, diag, out + i * k, NULL, &fi);
            if (!rc && d_info_out) rc = htnk_fill_i32(ctx->stream, d_info_out + i, 1, fi);
            if (fi && !*first_info) *first_info = fi;
        }
        return rc;
    }
    if (diag)
        rc = ht_shared_diag(ctx, s, nsamples, flags, (int)k2, (int)k, mean, cov, g, r, out, first_info);
    else
        rc = ht_shared_dense(ctx, s, nsamples, flags, (int)k2, (int)k, mean, cov, g, r, out, first_info);
    if (!rc && d_info_out && !(ctx->async_info_out && !diag)) rc = htnk_fill_i32(ctx->stream, d_info_out, nsamples, *first_info);
done:
    return rc;
}

/* LAPACK status of chol(cov) alone (used to decide whether a foreign generator may be advanced) */
int htn_ht_cov_info(htn_ctx_t* ctx, int k, const double* cov, unsigned flags, int* info)
{
    int rc = 0;
    const size_t kp = htn_round_up((size_t)k, 16);
    double* F = NULL;
    int* d_info = NULL;
    htn_factor_t fa;
    memset(&fa, 0, sizeof(fa));
    HTN_TRY(htn_dalloc(ctx, (void**)&F, (size_t)k * kp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&d_info, sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(F, 0, (size_t)k * kp * 8, ctx->stream));
    HTN_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
    HTN_TRY(htnk_sym_expand(ctx->stream, cov, (size_t)k, k, (flags & HTN_FLAG_COLMAJOR) ? 1 : 0, NULL, 0, F, kp));
    HTN_TRY(htn_factor_init(ctx, &fa, F, kp, k, 0));
    HTN_TRY(htn_potrf(ctx, &fa, d_info));
    HTN_CUDA(cudaMemcpyAsync(info, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HTN_CUDA(cudaStreamSynchronize(ctx->stream));
done:
    htn_factor_free(ctx, &fa);
    htn_dfree(ctx, d_info);
    htn_dfree(ctx, F);
    return rc;
}

/* ---- public entry points ---- */

static int upload(htn_ctx_t* ctx, void** dst, const void* src, size_t bytes)
{
    int rc = 0;
    HTN_TRY(htn_dalloc(ctx, dst, bytes));
    HTN_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
done:
    return rc;
}

int htn_hyperplane_truncated_mvn_batch(const htn_batch_t* b, const ht_config_t* conf, double* out, int* info)
{
    if (!b || !conf || !out) {
        htn_set_error("hyperplane batch: NULL argument");
        return HTN_E_ARG;
    }
    const size_t B = b->nsamples, k = conf->gncol, k2 = conf->gnrow;
    const size_t np = b->independent ? B : 1;
    const unsigned flags = b->flags | HTN_DEFAULT_FLAGS;
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, b->device, b->stream, b->mem == HTN_MEM_HOST);
    if (rc) return rc;
    int first = 0;
    htn_streams_t s = {(int)b->gen, b->seeds, b->seed0, NULL, NULL};
    if (b->mem == HTN_MEM_DEVICE) {
        /* asynchronous: nothing in this branch waits for the device (shared dense inputs, k2 <= 16, info given) */
        ctx.async_info_out = (!b->independent && !conf->diag && k2 <= ALPHA_K2_MAX) ? info : NULL;
        rc = htn_ht_device(&ctx, &s, B, b->independent, flags, k2, k, conf->mean, conf->cov, conf->g, conf->r,
                           conf->diag, out, info, &first);
        int rc2 = htn_ctx_end(&ctx, 0);
        return rc ? rc : rc2;
    }
    /* host memory: stage in, run, stage out (synchronous) */
    void *d_mean = NULL, *d_cov = NULL, *d_g = NULL, *d_r = NULL, *d_seeds = NULL, *d_out = NULL, *d_info = NULL;
    if (B == 0) goto done;
    if (k == 0 || k2 == 0 || !conf->mean || !conf->cov || !conf->g || !conf->r) {
        htn_set_error("hyperplane batch: empty or NULL input");
        rc = HTN_E_ARG;
        goto done;
    }
    HTN_TRY(upload(&ctx, &d_mean, conf->mean, np * k * 8));
    HTN_TRY(upload(&ctx, &d_cov, conf->cov, np * k * k * 8));
    HTN_TRY(upload(&ctx, &d_g, conf->g, np * k2 * k * 8));
    HTN_TRY(upload(&ctx, &d_r, conf->r, np * k2 * 8));
    if (b->seeds) {
        HTN_TRY(upload(&ctx, &d_seeds, b->seeds, B * 8));
        s.seeds = d_seeds;
    }
    HTN_TRY(htn_dalloc(&ctx, &d_out, B * k * 8));
    HTN_TRY(htn_dalloc(&ctx, &d_info, B * sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(d_info, 0, B * sizeof(int), ctx.stream));
    if (b->independent) /* a problem whose cov is not PD leaves its row of `out` as the caller had it */
        HTN_CUDA(cudaMemcpyAsync(d_out, out, B * k * 8, cudaMemcpyHostToDevice, ctx.stream));
    ctx.host_sink = (!b->independent && !conf->diag) ? out : NULL;
    ctx.host_sink_used = 0;
    rc = htn_ht_device(&ctx, &s, B, b->independent, flags, k2, k, d_mean, d_cov, d_g, d_r, conf->diag, d_out,
                       d_info, &first);
    ctx.host_sink = NULL;
    if (rc) goto done;
    {
        int* h_info = info ? info : malloc(B * sizeof(int));
        if (!h_info) { rc = HTNORM_ALLOC_ERROR; goto done; }
        HTN_CUDA(cudaMemcpyAsync(h_info, d_info, B * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
        HTN_CUDA(cudaStreamSynchronize(ctx.stream));
        for (size_t i = 0; i < B && !first; i++)
            if (h_info[i]) first = h_info[i];
        if (!info) free(h_info);
    }
    /* shared inputs with cov not PD: nothing was drawn, `out` stays untouched (reference behaviour) */
    if (!ctx.host_sink_used && !(first > 0 && !b->independent && !conf->diag && ctx.cov_failed))
        HTN_CUDA(cudaMemcpyAsync(out, d_out, B * k * 8, cudaMemcpyDeviceToHost, ctx.stream));
done:
    htn_dfree(&ctx, d_info);
    htn_dfree(&ctx, d_out);
    htn_dfree(&ctx, d_seeds);
    htn_dfree(&ctx, d_r);
    htn_dfree(&ctx, d_g);
    htn_dfree(&ctx, d_cov);
    htn_dfree(&ctx, d_mean);
    {
        int rc2 = htn_ctx_end(&ctx, 1);
        if (!rc) rc = rc2;
    }
    return rc ? rc : first;
}

/* Single draw through the reference's own signature (include/htnorm.h).  The generator must be one of
 * this library's (function-pointer identity): its state is uploaded, advanced on the GPU by exactly the
 * draws the reference would make, and written back. */
int htn_hyperplane_truncated_mvn(rng_t* rng, const ht_config_t* conf, double* HTN_RESTRICT out)
{
    if (!rng || !conf || !out) {
        htn_set_error("hyperplane: NULL argument");
        return HTN_E_ARG;
    }
    const int kind = htn_rng_kind(rng);
    const size_t k = conf->gncol, k2 = conf->gnrow;
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, -1, NULL, 1);
    if (rc) return rc;
    int first = 0;
    void *d_mean = NULL, *d_cov = NULL, *d_g = NULL, *d_r = NULL, *d_out = NULL, *d_state = NULL, *d_z = NULL;
    double* h_z = NULL;
    uint64_t h_state[4] = {0, 0, 0, 0};
    if (kind >= 0) memcpy(h_state, rng->base, kind == HTN_GEN_PCG64 ? 32 : 16);
    if (k == 0 || k2 == 0) { rc = HTN_E_ARG; goto done; }
    HTN_TRY(upload(&ctx, &d_mean, conf->mean, k * 8));
    HTN_TRY(upload(&ctx, &d_cov, conf->cov, k * k * 8));
    HTN_TRY(upload(&ctx, &d_g, conf->g, k2 * k * 8));
    HTN_TRY(upload(&ctx, &d_r, conf->r, k2 * 8));
    HTN_TRY(upload(&ctx, &d_out, out, k * 8)); /* a failed factorisation must leave out as it was */
    if (kind >= 0) {
        HTN_TRY(upload(&ctx, &d_state, h_state, sizeof(h_state)));
        htn_streams_t s = {kind, NULL, 0, (uint64_t*)d_state, NULL};
        HTN_TRY(htn_ht_device(&ctx, &s, 1, 0, HTN_DEFAULT_FLAGS, k2, k, d_mean, d_cov, d_g, d_r, conf->diag, d_out, NULL,
                              &first));
    } else {
        /* foreign generator: its callbacks run on the host.  The reference draws only after a
         * successful factorisation (src/htnorm_distributions.c:78), so factorise first with no draw ... */
        if (!conf->diag) {
            HTN_TRY(htn_ht_cov_info(&ctx, (int)k, d_cov, HTN_DEFAULT_FLAGS, &first));
        }
        if (first == 0) { /* ... then draw k normals through the callbacks and run the device path on them */
            h_z = malloc(k * sizeof(double));
            if (!h_z) { rc = HTNORM_ALLOC_ERROR; goto done; }
            htn_foreign_draw(rng, k, h_z);
            HTN_TRY(upload(&ctx, &d_z, h_z, k * 8));
            htn_streams_t s = {0, NULL, 0, NULL, (const double*)d_z};
            HTN_TRY(htn_ht_device(&ctx, &s, 1, 0, HTN_DEFAULT_FLAGS, k2, k, d_mean, d_cov, d_g, d_r, conf->diag, d_out,
                                  NULL, &first));
        }
    }
    HTN_CUDA(cudaMemcpyAsync(out, d_out, k * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (kind >= 0)
        HTN_CUDA(cudaMemcpyAsync(h_state, d_state, sizeof(h_state), cudaMemcpyDeviceToHost, ctx.stream));
    HTN_CUDA(cudaStreamSynchronize(ctx.stream));
    if (kind >= 0) memcpy(rng->base, h_state, kind == HTN_GEN_PCG64 ? 32 : 16);
done:
    free(h_z);
    htn_dfree(&ctx, d_z);
    htn_dfree(&ctx, d_state);
    htn_dfree(&ctx, d_out);
    htn_dfree(&ctx, d_r);
    htn_dfree(&ctx, d_g);
    htn_dfree(&ctx, d_cov);
    htn_dfree(&ctx, d_mean);
    {
        int rc2 = htn_ctx_end(&ctx, 1);
        if (!rc) rc = rc2;
    }
    return rc ? rc : first;
}
