/* Structured-precision MVN sampler: x ~ N(mean, (A + Phi^T Omega Phi)^-1), optionally with the
 * structured mean of Example 4 (Cong et al. 2017, Alg. 4).
 *
 * Reference path restated for a batch (src/htnorm.c:217-356, src/htnorm_distributions.c:91-125):
 *     y1 ~ N(0, A^-1),  y2 ~ N(0, Omega^-1)      (two runs of normals from the SAME stream, y1 first)
 *     M = Omega^-1 + Phi A^-1 Phi^T,   b = Phi y1 + y2   (struct mean: b = t - b)
 *     x = mean + y1 + X M^-1 b         (struct mean: x = y1 - X M^-1 b),   X = A^-1 Phi^T
 * The signs are the REFERENCE's (+1 / -1, src/htnorm.c:229,320,336,343) unless
 * HTN_FLAG_PAPER_SIGN is set.  An IDENTITY Omega reproduces the reference's memset quirk.
 *
 * Once per batch: chol(A), chol(Omega), X^T = Phi A^-1 (two blocked TRSMs instead of the reference's
 * explicit dpotri + dsymm), M (one DMMA GEMM), chol(M).  Per sample everything is a row of a GEMM / TRSM
 * over the batch: Y1 = Z1 L_A^-1, B = Z2' + Y1 Phi^T, Alpha = B M^-1, out = Y1 + mean + coef Alpha X^T.
 */
#include "htn_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <time.h>
/* HTN_SP_TRACE=1 (development only): synchronise the device at the phase boundaries of htn_sp_device and print how long
 * each phase took on its own (no overlap between phases is left in such a run) */
static int sp_trace_on(void)
{
    static int v = -1;
    if (v < 0) v = getenv("HTN_SP_TRACE") != NULL;
    return v;
}
static void sp_trace(const char* label)
{
    static _Thread_local double t_prev = 0.0;
    if (!sp_trace_on()) return;
    cudaDeviceSynchronize();
    struct timespec ts;
    timespec_get(&ts, TIME_UTC);
    const double t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (label) fprintf(stderr, "[sp] %-28s %8.3f ms\n", label, t - t_prev);
    t_prev = t;
}

#include <stdlib.h>
#include <string.h>

/* value of every double of the reference's IDENTITY "factor": memset(factor, 1, n * sizeof(double))
 * (src/htnorm_distributions.c:99) -> bytes 0x01 -> 7.7486e-304; Omega^-1 gets 1 / that (src/htnorm.c:203) */
static double identity_factor_garbage(void)
{
    double d;
    memset(&d, 1, sizeof(d));
    return d;
}

/* stage a dense SPD input: lower triangle into a fresh n x ld buffer, factorise, and build the explicit
 * inverse factor that turns every later many-row solve into one triangular GEMM */
/* rows: how many rows will be solved against this factor in all.  Few rows against a large factor: only the 2048-wide
 * diagonal blocks of the inverse are built (cfg4: p = 8192 against 1024 + 256 rows), and the solves substitute block by
 * block.  HTN_INV_BLOCK (development / tests): 0 = always the whole inverse, W = that block width regardless. */
static int inverse_block_width(int n, size_t rows)
{
    const char* e = getenv("HTN_INV_BLOCK");
    if (e) return atoi(e);
    return (n >= 4096 && rows * 3 <= (size_t)n) ? 2048 : 0;
}

static int factor_dense(htn_ctx_t* ctx, const double* src, int n, int lower_src, double** buf, htn_factor_t* fa,
                        int* d_info, int wblock)
{
    int rc = 0;
    const size_t ld = htn_round_up((size_t)n, 16);
    HTN_TRY(htn_dalloc(ctx, (void**)buf, (size_t)n * ld * 8));
    if (ld != (size_t)n) /* sym_expand writes the whole n x n square (zeros above the diagonal): only padding needs the memset */
        HTN_CUDA(cudaMemsetAsync(*buf, 0, (size_t)n * ld * 8, ctx->stream));
    HTN_TRY(htnk_sym_expand(ctx->stream, src, (size_t)n, n, lower_src, NULL, 0, *buf, ld));
    HTN_TRY(htn_factor_init(ctx, fa, *buf, ld, n, 1));
    sp_trace("  (before potrf)");
    HTN_TRY(htn_potrf(ctx, fa, d_info));
    sp_trace("  potrf");
    HTN_TRY(htn_factor_make_u(ctx, fa));
    sp_trace("  U = L^T");
    HTN_TRY(htn_factor_make_inverse_blocks(ctx, fa, wblock)); /* wblock = 0: the whole inverse */
    sp_trace("  inverse factor");
done:
    return rc;
}

/* side streams and events of the structured path: created once per host thread and device and kept (creating and
 * destroying them costs more host time than a small batch takes) */
static _Thread_local struct {
    int device;
    cudaStream_t side, side2;
    cudaEvent_t ev_fork, ev_rng, ev_a, ev_y1, ev_o;
} ss = {-1, NULL, NULL, NULL, NULL, NULL, NULL, NULL};

static int sp_streams_ready(int device)
{
    if (ss.device == device) return 0;
    if (ss.device >= 0) { /* another device was current before: let go of its objects */
        cudaStreamDestroy(ss.side);
        cudaStreamDestroy(ss.side2);
        cudaEventDestroy(ss.ev_fork);
        cudaEventDestroy(ss.ev_rng);
        cudaEventDestroy(ss.ev_a);
        cudaEventDestroy(ss.ev_y1);
        cudaEventDestroy(ss.ev_o);
        ss.device = -1;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&ss.side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ss.side2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ss.ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ss.ev_rng, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ss.ev_a, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ss.ev_y1, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ss.ev_o, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        htn_set_error("structured precision: side streams: %s", cudaGetErrorString(e));
        return HTN_E_CUDA;
    }
    ss.device = device;
    return 0;
}

int htn_sp_device(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, unsigned flags, size_t n_, size_t p_,
                  const double* mean, const double* a, const double* phi, const double* omega, int struct_mean,
                  int a_id, int o_id, double* out, int* d_info_out, int* first_info)
{
    int rc = 0;
    *first_info = 0;
    if (nsamples == 0) return 0;
    if (n_ == 0 || p_ == 0 || !mean || !phi || !out || (a_id != IDENTITY && !a) || (o_id != IDENTITY && !omega) ||
        n_ > 0x3fffffffu || p_ > 0x3fffffffu || nsamples > 0x7fffffffu || a_id < 0 || a_id > 2 || o_id < 0 ||
        o_id > 2) {
        htn_set_error("structured precision: bad argument (n=%zu p=%zu nsamples=%zu)", n_, p_, nsamples);
        return HTN_E_ARG;
    }
    const int n = (int)n_, p = (int)p_;
    const int lower_src = (flags & HTN_FLAG_COLMAJOR) ? 1 : 0;
    const size_t pp = htn_round_up(p_, 16), np = htn_round_up(n_, 16);
    cudaStream_t st = ctx->stream;
    double *Phip = NULL, *FA = NULL, *FO = NULL, *XT = NULL, *X = NULL, *M = NULL, *T1 = NULL;
    double *da = NULL, *sda = NULL, *sdo = NULL, *invo = NULL;
    double *Z1 = NULL, *Z2 = NULL, *Y1 = NULL, *Y2 = NULL;
    int* d_info = NULL;
    /* side: the generator - chunk 0 is drawn while A, Omega and M are being factorised; side2: Y1 of chunk 0, one big
     * triangular GEMM that needs only chol(A), beside the chain of small launches that builds X and M */
    cudaStream_t side = NULL, side2 = NULL;
    cudaEvent_t ev_fork = NULL, ev_rng = NULL, ev_a = NULL, ev_y1 = NULL, ev_o = NULL;
    int rng_in_flight = 0, y1_in_flight = 0, side2_used = 0, omega_forked = 0;
    htn_factor_t fa, fo, fm;
    memset(&fa, 0, sizeof(fa));
    memset(&fo, 0, sizeof(fo));
    memset(&fm, 0, sizeof(fm));

    size_t chunk = nsamples;
    const size_t budget = htn_chunk_budget();
    if (chunk * (pp + np) * 8 > budget) chunk = htn_round_up(budget / ((pp + np) * 8), 128);

    HTN_TRY(htn_dalloc(ctx, (void**)&d_info, 3 * sizeof(int)));
    HTN_CUDA(cudaMemsetAsync(d_info, 0, 3 * sizeof(int), st));
    HTN_TRY(htn_dalloc(ctx, (void**)&Phip, (size_t)n * pp * 8));
    HTN_CUDA(cudaMemsetAsync(Phip, 0, (size_t)n * pp * 8, st));
    if (flags & HTN_FLAG_COLMAJOR) /* phi is n x p column-major = p x n row-major */
        HTN_TRY(htnk_transpose(st, phi, (size_t)n, p, n, Phip, pp));
    else
        HTN_TRY(htnk_copy2d(st, phi, (size_t)p, n, p, Phip, pp, NULL));
    /* diagonal types: the divide by sqrt(d_ii) is folded into the generator's write-out
     * (mvn_rand_prec, src/htnorm_distributions.c:101-109) */
    if (a_id == DIAGONAL) {
        HTN_TRY(htn_dalloc(ctx, (void**)&da, pp * 8));
        HTN_TRY(htn_dalloc(ctx, (void**)&sda, pp * 8));
        HTN_TRY(htnk_diag_extract(st, a, (size_t)p, p, 0, da));
        HTN_TRY(htnk_diag_extract(st, a, (size_t)p, p, 1, sda));
    }
    if (o_id == DIAGONAL) {
        HTN_TRY(htn_dalloc(ctx, (void**)&sdo, np * 8));
        HTN_TRY(htn_dalloc(ctx, (void**)&invo, np * 8));
        HTN_TRY(htnk_diag_extract(st, omega, (size_t)n, n, 1, sdo));
        HTN_TRY(htnk_diag_extract(st, omega, (size_t)n, n, 2, invo));
    }
    HTN_TRY(htn_dalloc(ctx, (void**)&Z1, chunk * pp * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Z2, chunk * np * 8));
    HTN_TRY(htn_dalloc(ctx, (void**)&Y2, chunk * np * 8));
    if (a_id == NORMAL) HTN_TRY(htn_dalloc(ctx, (void**)&Y1, chunk * pp * 8));
    HTN_TRY(sp_streams_ready(ctx->device));
    side = ss.side, side2 = ss.side2;
    ev_fork = ss.ev_fork, ev_rng = ss.ev_rng, ev_a = ss.ev_a, ev_y1 = ss.ev_y1, ev_o = ss.ev_o;
    /* Both runs of normals of chunk 0 are drawn up front (y1's then y2's, from the same stream).  If chol(A)
     * or chol(omega) fails the reference would have stopped earlier: only the single-draw entry cares (it
     * writes the generator back), and it re-checks the codes below before touching the caller's state. */
    {
        const size_t nb0 = nsamples < chunk ? nsamples : chunk;
        htnk_streams_t sv = {s->gen, s->seeds, s->seed0, s->states, s->zpre};
        htnk_segment_t seg[2] = {{(size_t)p, Z1, pp, a_id == DIAGONAL ? sda : NULL, NULL, NULL},
                                 {(size_t)n, Z2, np, o_id == DIAGONAL ? sdo : NULL, NULL, NULL}};
        HTN_CUDA(cudaEventRecord(ev_fork, st));
        if (!s->states) { /* batched (seeded) streams: overlap with the factorisations */
            HTN_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
            htnk_normal_fill_shares_device_next();
            HTN_TRY(htnk_normal_fill(side, &sv, nb0, 2, seg, NULL));
            HTN_CUDA(cudaEventRecord(ev_rng, side));
            rng_in_flight = 1;
        }
    }

    sp_trace(NULL);
    /* ---- A, Omega: factor (dense) (mvn_rand_prec, src/htnorm_distributions.c:110-123) ---- */
    if (a_id == NORMAL && o_id == NORMAL) { /* two dense factorisations: Omega's (the smaller one) runs beside A's */
        htn_ctx_t c2 = *ctx;
        c2.stream = side2;
        HTN_CUDA(cudaEventRecord(ev_a, st));
        HTN_CUDA(cudaStreamWaitEvent(side2, ev_a, 0));
        side2_used = 1;
        HTN_TRY(factor_dense(&c2, omega, n, lower_src, &FO, &fo, d_info + 1, 0));
        HTN_CUDA(cudaEventRecord(ev_o, side2));
        omega_forked = 1;
    }
    int early0 = 0; /* chunk 0: Y1, X and b = Phi y1 + y2 were formed on side2, beside the chain that builds and factors M */
    if (a_id == NORMAL) {
        HTN_TRY(factor_dense(ctx, a, p, lower_src, &FA, &fa, d_info + 0, inverse_block_width(p, (size_t)2 * n_ + nsamples)));
        /* T1 = Phi L_A^-T (n x p): M needs only T1 (Phi A^-1 Phi^T = T1 T1^T); X^T = T1 L_A^-1 is not needed before the
         * last GEMM of a chunk, so it leaves the critical path */
        HTN_TRY(htn_dalloc(ctx, (void**)&T1, (size_t)n * pp * 8));
        HTN_TRY(htn_dalloc(ctx, (void**)&XT, (size_t)n * pp * 8));
        HTN_TRY(htn_dalloc(ctx, (void**)&X, (size_t)p * np * 8));
        HTN_CUDA(cudaMemsetAsync(XT, 0, (size_t)n * pp * 8, st));
        HTN_CUDA(cudaMemsetAsync(X, 0, (size_t)p * np * 8, st));
        HTN_TRY(htn_solve_lt(ctx, &fa, Phip, pp, T1, pp, n));
        if (rng_in_flight) {
            /* side2: L^T y1 = z for chunk 0 (see the loop below), X^T = T1 L_A^-1, X, and chunk 0's b = Phi y1 + y2 -
             * four device-filling GEMMs that need nothing of M, beside the latency-bound chain M -> chol(M) -> M^-1 */
            const size_t nb0 = nsamples < chunk ? nsamples : chunk;
            htn_ctx_t c2 = *ctx;
            c2.stream = side2;
            HTN_CUDA(cudaEventRecord(ev_a, st));
            HTN_CUDA(cudaStreamWaitEvent(side2, ev_a, 0));
            HTN_CUDA(cudaStreamWaitEvent(side2, ev_rng, 0));
            side2_used = 1;
            y1_in_flight = 1; /* joined at the exits from here on */
            HTN_TRY(htn_solve_l(&c2, &fa, Z1, pp, Y1, pp, (int)nb0));
            HTN_TRY(htn_solve_l(&c2, &fa, T1, pp, XT, pp, n));
            HTN_TRY(htnk_transpose(side2, XT, pp, n, p, X, np));
            {
                double* bm0 = Z2;
                if (o_id == NORMAL) { /* chol(Omega) ran on side2 as well: its factor is there in stream order */
                    HTN_TRY(htn_solve_l(&c2, &fo, Z2, np, Y2, np, (int)nb0));
                    bm0 = Y2;
                }
                HTN_TRY(htnk_gemm_tn(side2, (int)nb0, n, p, 1.0, Y1, pp, Phip, pp, 1.0, bm0, np, NULL, HTNK_GEMM_FULL, 0, 0, 0));
                if (struct_mean) HTN_TRY(htnk_rsub_rows(side2, bm0, np, (int)nb0, n, mean));
            }
            HTN_CUDA(cudaEventRecord(ev_y1, side2));
            early0 = 1;
        } else {
            HTN_TRY(htn_solve_l(ctx, &fa, T1, pp, XT, pp, n));
            HTN_TRY(htnk_transpose(st, XT, pp, n, p, X, np));
        }
    }
    sp_trace("factor A (+inverse), Y1, Omega");
    if (omega_forked) /* Omega^-1 itself is formed from the whole inverse factor below */
        HTN_CUDA(cudaStreamWaitEvent(st, ev_o, 0));
    else if (o_id == NORMAL)
        HTN_TRY(factor_dense(ctx, omega, n, lower_src, &FO, &fo, d_info + 1, 0));
    /* ---- X^T = Phi A^-1 (n x p)   (src/htnorm.c:261-302; dpotri + dsymm replaced by two triangular GEMMs: above) */
    if (a_id == IDENTITY) {
        XT = Phip;
    } else if (a_id == DIAGONAL) {
        HTN_TRY(htn_dalloc(ctx, (void**)&XT, (size_t)n * pp * 8));
        HTN_CUDA(cudaMemsetAsync(XT, 0, (size_t)n * pp * 8, st));
        HTN_TRY(htnk_coldiv(st, Phip, pp, n, p, da, XT, pp));
    }
    sp_trace("X^T = Phi A^-1");
    /* ---- M = Omega^-1 + X^T Phi^T  (compute_precision_inverse, src/htnorm.c:190-214; dgemm/dsyrk :267-300) */
    HTN_TRY(htn_dalloc(ctx, (void**)&M, (size_t)n * np * 8));
    HTN_CUDA(cudaMemsetAsync(M, 0, (size_t)n * np * 8, st));
    if (o_id == NORMAL) /* Omega^-1 = L^-T L^-1 */
        HTN_TRY(htnk_gemm_tn(st, n, n, n, 1.0, fo.linvt, fo.ld, fo.linvt, fo.ld, 0.0, M, np, NULL, HTNK_GEMM_FULL, 0,
                             0, 0));
    else if (o_id == DIAGONAL)
        HTN_TRY(htnk_diag_set(st, M, np, n, invo, 0.0));
    else
        HTN_TRY(htnk_diag_set(st, M, np, n, NULL, 1 / identity_factor_garbage()));
    if (a_id == NORMAL) /* Phi A^-1 Phi^T = T1 T1^T */
        HTN_TRY(htnk_gemm_tn(st, n, n, p, 1.0, T1, pp, T1, pp, 1.0, M, np, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    else
        HTN_TRY(htnk_gemm_tn(st, n, n, p, 1.0, XT, pp, Phip, pp, 1.0, M, np, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    sp_trace("M");
    HTN_TRY(htn_factor_init(ctx, &fm, M, np, n, 1));
    HTN_TRY(htn_potrf(ctx, &fm, d_info + 2)); /* dposv, src/htnorm.c:330 */
    HTN_TRY(htn_factor_make_u(ctx, &fm));
    HTN_TRY(htn_factor_make_inverse(ctx, &fm));
    sp_trace("factor M + inverse");
    /* X = (X^T)^T, p x n: the B operand of the final GEMM (A dense: formed above, beside this chain) */
    if (a_id != NORMAL) {
        HTN_TRY(htn_dalloc(ctx, (void**)&X, (size_t)p * np * 8));
        HTN_CUDA(cudaMemsetAsync(X, 0, (size_t)p * np * 8, st));
        HTN_TRY(htnk_transpose(st, XT, pp, n, p, X, np));
    }

    int h_info[3] = {0, 0, 0};
    HTN_CUDA(cudaMemcpyAsync(h_info, d_info, sizeof(h_info), cudaMemcpyDeviceToHost, st));
    HTN_CUDA(cudaStreamSynchronize(st));
    if (h_info[0]) { /* A not PD: returned before any draw (src/htnorm.c:245) */
        *first_info = h_info[0];
        goto finish;
    }

    sp_trace("transpose X, sync");
    double coef = struct_mean ? -1.0 : 1.0;
    if (flags & HTN_FLAG_PAPER_SIGN) coef = -coef;
    for (size_t b0 = 0; b0 < nsamples; b0 += chunk) {
        const int nb = (int)(nsamples - b0 < chunk ? nsamples - b0 : chunk);
        htnk_streams_t sv = {s->gen, s->seeds ? s->seeds + b0 : NULL, s->seed0 + (uint64_t)b0,
                             s->states ? s->states + 4 * b0 : NULL, s->zpre};
        htnk_segment_t seg[2] = {{(size_t)p, Z1, pp, a_id == DIAGONAL ? sda : NULL, NULL, NULL},
                                 {(size_t)n, Z2, np, o_id == DIAGONAL ? sdo : NULL, NULL, NULL}};
        if (b0 == 0 && rng_in_flight) {
            HTN_CUDA(cudaStreamWaitEvent(st, ev_rng, 0));
            rng_in_flight = 0;
            if (h_info[1]) continue; /* Omega not PD: the call fails after y1's draws (:246) */
        } else if (h_info[1]) { /* single draw continuing a generator: advance it by y1's draws only */
            HTN_TRY(htnk_normal_fill(st, &sv, (size_t)nb, 1, seg, NULL));
            continue;
        } else {
            HTN_TRY(htnk_normal_fill(st, &sv, (size_t)nb, 2, seg, NULL));
        }
        /* L^T y1 = z (dtrsv, src/htnorm_distributions.c:120) for the whole chunk: Y1 = Z1 L_A^-1 */
        const double* y1 = Z1;
        if (a_id == NORMAL) {
            if (b0 == 0 && y1_in_flight) {
                HTN_CUDA(cudaStreamWaitEvent(st, ev_y1, 0));
                y1_in_flight = 0;
            } else {
                HTN_TRY(htn_solve_l(ctx, &fa, Z1, pp, Y1, pp, nb));
            }
            y1 = Y1;
        }
    sp_trace("chunk: normals / Y1");
        const int early = (b0 == 0 && early0); /* Y2, b and (structured mean) t - b of this chunk are already there */
        double* bm = Z2;
        if (o_id == NORMAL) {
            if (!early) HTN_TRY(htn_solve_l(ctx, &fo, Z2, np, Y2, np, nb));
            bm = Y2;
        }
        double* scratch = (bm == Z2) ? Y2 : Z2;
        /* b = Phi y1 + y2 (dgemm, src/htnorm.c:310) */
        if (!early) HTN_TRY(htnk_gemm_tn(st, nb, n, p, 1.0, y1, pp, Phip, pp, 1.0, bm, np, NULL, HTNK_GEMM_FULL, 0, 0, 0));
    sp_trace("chunk: Y2, b = Phi y1 + y2");
        double* o = out + b0 * p_;
        if (struct_mean) { /* src/htnorm.c:314-321 */
            if (!early) HTN_TRY(htnk_rsub_rows(st, bm, np, nb, n, mean));
            HTN_TRY(htnk_copy2d(st, y1, pp, nb, p, o, p_, NULL));
        } else { /* src/htnorm.c:322-326 */
            HTN_TRY(htnk_copy2d(st, y1, pp, nb, p, o, p_, mean));
        }
        if (h_info[2]) continue; /* M not PD: out keeps y1 (+ mean), as after a failed dposv (:330-346) */
    sp_trace("chunk: copy y1 to out");
        /* alpha = b M^-1: two triangular GEMMs, result back in bm */
        HTN_TRY(htn_solve_lt(ctx, &fm, bm, np, scratch, np, nb));
        HTN_TRY(htn_solve_l(ctx, &fm, scratch, np, bm, np, nb));
    sp_trace("chunk: alpha solves");
        /* out += coef * alpha X^T  (dgemv / dgemm, src/htnorm.c:332-346) */
        HTN_TRY(htnk_gemm_tn(st, nb, p, n, coef, bm, np, X, np, 1.0, o, p_, NULL, HTNK_GEMM_FULL, 0, 0, 0));
        sp_trace("chunk: out += alpha X^T");
    }
    *first_info = h_info[1] ? h_info[1] : h_info[2];
finish:
    if (d_info_out) rc = htnk_fill_i32(st, d_info_out, nsamples, *first_info);
done:
    if (rng_in_flight && ev_rng) cudaStreamWaitEvent(st, ev_rng, 0); /* never free under the generator */
    if (side2_used && (y1_in_flight || omega_forked || rc)) { /* ... nor under the early Y1 product, whatever state it was left in */
        if (cudaEventRecord(ev_y1, side2) != cudaSuccess || cudaStreamWaitEvent(st, ev_y1, 0) != cudaSuccess)
            cudaStreamSynchronize(side2);
    }
    htn_factor_free(ctx, &fa);
    htn_factor_free(ctx, &fo);
    htn_factor_free(ctx, &fm);
    htn_dfree(ctx, Z1);
    htn_dfree(ctx, Z2);
    htn_dfree(ctx, Y1);
    htn_dfree(ctx, Y2);
    htn_dfree(ctx, T1);
    htn_dfree(ctx, X);
    htn_dfree(ctx, M);
    if (XT != Phip) htn_dfree(ctx, XT);
    htn_dfree(ctx, invo);
    htn_dfree(ctx, sdo);
    htn_dfree(ctx, sda);
    htn_dfree(ctx, da);
    htn_dfree(ctx, FO);
    htn_dfree(ctx, FA);
    htn_dfree(ctx, Phip);
    htn_dfree(ctx, d_info);
    return rc;
}

static int upload(htn_ctx_t* ctx, void** dst, const void* src, size_t bytes)
{
    int rc = 0;
    *dst = NULL;
    if (!src) return 0;
    HTN_TRY(htn_dalloc(ctx, dst, bytes));
    HTN_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
done:
    return rc;
}

/* host arrays -> device, run, -> host.  states_host (4 words) non-NULL: single draw continuing a generator */
static int sp_host(const sp_config_t* conf, int gen, const uint64_t* seeds, uint64_t seed0, uint64_t* states_host,
                   size_t B, unsigned flags, int device, double* out, int* info, rng_t* foreign)
{
    const size_t n = conf->pnrow, p = conf->pncol;
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, device, NULL, 1);
    if (rc) return rc;
    int first = 0;
    void *d_mean = NULL, *d_a = NULL, *d_phi = NULL, *d_om = NULL, *d_seeds = NULL, *d_out = NULL, *d_info = NULL,
         *d_state = NULL, *d_z = NULL;
    double* h_z = NULL;
    if (B == 0) goto done;
    if (n == 0 || p == 0 || !conf->mean || !conf->phi || !out) {
        htn_set_error("structured precision: empty or NULL input");
        rc = HTN_E_ARG;
        goto done;
    }
    HTN_TRY(upload(&ctx, &d_mean, conf->mean, (conf->struct_mean ? n : p) * 8));
    /* DIAGONAL / IDENTITY matrices are still full square arrays in the reference's API; only the
     * diagonal is read (src/htnorm_distributions.c:106-107) */
    if (conf->a_id != IDENTITY) HTN_TRY(upload(&ctx, &d_a, conf->a, p * p * 8));
    if (conf->o_id != IDENTITY) HTN_TRY(upload(&ctx, &d_om, conf->omega, n * n * 8));
    HTN_TRY(upload(&ctx, &d_phi, conf->phi, n * p * 8));
    if (seeds) HTN_TRY(upload(&ctx, &d_seeds, seeds, B * 8));
    if (states_host) HTN_TRY(upload(&ctx, &d_state, states_host, 32));
    HTN_TRY(upload(&ctx, &d_out, out, B * p * 8)); /* failed factorisations leave `out` as it was */
    HTN_TRY(htn_dalloc(&ctx, &d_info, B * sizeof(int)));
    if (foreign) {
        /* foreign generator: the reference draws y1's normals only after chol(A) succeeded and
         * y2's only after chol(omega) did (src/htnorm.c:245-246): check both, then draw on the host */
        int ia = 0, io = 0;
        if (conf->a_id == NORMAL) HTN_TRY(htn_ht_cov_info(&ctx, (int)p, d_a, flags, &ia));
        if (ia) { first = ia; goto done; }
        h_z = malloc((p + n) * sizeof(double));
        if (!h_z) { rc = HTNORM_ALLOC_ERROR; goto done; }
        htn_foreign_draw(foreign, p, h_z);
        if (conf->o_id == NORMAL) HTN_TRY(htn_ht_cov_info(&ctx, (int)n, d_om, flags, &io));
        if (io) { first = io; goto done; }
        htn_foreign_draw(foreign, n, h_z + p);
        HTN_TRY(upload(&ctx, &d_z, h_z, (p + n) * 8));
        htn_streams_t s = {0, NULL, 0, NULL, (const double*)d_z};
        HTN_TRY(htn_sp_device(&ctx, &s, 1, flags, n, p, d_mean, d_a, d_phi, d_om, conf->struct_mean, conf->a_id,
                              conf->o_id, d_out, d_info, &first));
    } else {
        htn_streams_t s = {gen, (const uint64_t*)d_seeds, seed0, (uint64_t*)d_state, NULL};
        HTN_TRY(htn_sp_device(&ctx, &s, B, flags, n, p, d_mean, d_a, d_phi, d_om, conf->struct_mean, conf->a_id,
                              conf->o_id, d_out, d_info, &first));
    }
    if (info) HTN_CUDA(cudaMemcpyAsync(info, d_info, B * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    HTN_CUDA(cudaMemcpyAsync(out, d_out, B * p * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (states_host) HTN_CUDA(cudaMemcpyAsync(states_host, d_state, 32, cudaMemcpyDeviceToHost, ctx.stream));
done:
    free(h_z);
    htn_dfree(&ctx, d_z);
    htn_dfree(&ctx, d_state);
    htn_dfree(&ctx, d_info);
    htn_dfree(&ctx, d_out);
    htn_dfree(&ctx, d_seeds);
    htn_dfree(&ctx, d_om);
    htn_dfree(&ctx, d_phi);
    htn_dfree(&ctx, d_a);
    htn_dfree(&ctx, d_mean);
    {
        int rc2 = htn_ctx_end(&ctx, 1);
        if (!rc) rc = rc2;
    }
    return rc ? rc : first;
}

int htn_structured_precision_mvn_batch(const htn_batch_t* b, const sp_config_t* conf, double* out, int* info)
{
    if (!b || !conf || !out) {
        htn_set_error("structured precision batch: NULL argument");
        return HTN_E_ARG;
    }
    if (b->independent) {
        htn_set_error("structured precision: independent problems are not supported");
        return HTN_E_ARG;
    }
    if (b->mem == HTN_MEM_HOST)
        return sp_host(conf, (int)b->gen, b->seeds, b->seed0, NULL, b->nsamples, b->flags | HTN_DEFAULT_FLAGS, b->device, out,
                       info, NULL);
    htn_ctx_t ctx;
    int rc = htn_ctx_begin(&ctx, b->device, b->stream, 0);
    if (rc) return rc;
    int first = 0;
    htn_streams_t s = {(int)b->gen, b->seeds, b->seed0, NULL, NULL};
    rc = htn_sp_device(&ctx, &s, b->nsamples, b->flags | HTN_DEFAULT_FLAGS, conf->pnrow, conf->pncol, conf->mean, conf->a, conf->phi,
                       conf->omega, conf->struct_mean, conf->a_id, conf->o_id, out, info, &first);
    int rc2 = htn_ctx_end(&ctx, 0);
    return rc ? rc : rc2;
}

int htn_structured_precision_mvn(rng_t* rng, const sp_config_t* conf, double* HTN_RESTRICT out)
{
    if (!rng || !conf || !out) {
        htn_set_error("structured precision: NULL argument");
        return HTN_E_ARG;
    }
    const int kind = htn_rng_kind(rng);
    if (kind < 0) /* foreign generator: host draws through its callbacks, device algebra */
        return sp_host(conf, 0, NULL, 0, NULL, 1, HTN_DEFAULT_FLAGS, -1, out, NULL, rng);
    uint64_t h_state[4] = {0, 0, 0, 0};
    memcpy(h_state, rng->base, kind == HTN_GEN_PCG64 ? 32 : 16);
    int rc = sp_host(conf, kind, NULL, 0, h_state, 1, HTN_DEFAULT_FLAGS, -1, out, NULL, NULL);
    if (rc >= 0 || rc == HTNORM_ALLOC_ERROR) memcpy(rng->base, h_state, kind == HTN_GEN_PCG64 ? 32 : 16);
    return rc;
}
