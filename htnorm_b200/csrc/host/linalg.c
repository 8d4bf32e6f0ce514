/* Blocked dense drivers of the host layer: every O(n^3) step is a sequence of DMMA GEMM launches
 * (htnk_gemm_tn) around the shared-memory diagonal-block kernel (htnk_potf2_inv).
 *
 *   htn_potrf          <-> dpotrf('L')   reference src/htnorm_distributions.c:77,115
 *   htn_trsm_right_*   <-> dtrsv / the solve phases of dposv, dpotri   src/htnorm_distributions.c:120,
 *                                                                     src/htnorm.c:169,210,330
 * All matrices row-major.  "right" solves act on ROWS of X (one row = one sample / one right-hand
 * side), which keeps every operand K-contiguous for the TN GEMM.
 */
#include "htn_internal.h"

#include <stdlib.h>
#include <string.h>

int htn_factor_init(htn_ctx_t* ctx, htn_factor_t* fa, double* f, size_t ld, int n, int want_u)
{
    int rc = 0;
    memset(fa, 0, sizeof(*fa));
    fa->n = n;
    fa->ld = ld;
    fa->f = f;
    fa->nblk = (n + HTN_NB - 1) / HTN_NB;
    fa->solves = want_u;
    size_t blk = (size_t)HTN_NB * HTN_NB * sizeof(double);
    HTN_TRY(htn_dalloc(ctx, (void**)&fa->rdiag, (size_t)fa->nblk * HTN_NB * sizeof(double)));
    HTN_TRY(htn_dalloc(ctx, (void**)&fa->pscr, (16 * 64 + 32) * sizeof(double)));
    HTN_CUDA(cudaMemsetAsync(fa->pscr, 0, (16 * 64 + 32) * sizeof(double), ctx->stream));
    if (fa->solves) {
        HTN_TRY(htn_dalloc(ctx, (void**)&fa->dinv, blk * fa->nblk));
        HTN_TRY(htn_dalloc(ctx, (void**)&fa->dinvt, blk * fa->nblk));
    }
    if (want_u) {
        HTN_TRY(htn_dalloc(ctx, (void**)&fa->u, (size_t)n * ld * sizeof(double)));
    }
done:
    return rc;
}

void htn_factor_free(htn_ctx_t* ctx, htn_factor_t* fa)
{
    htn_dfree(ctx, fa->dinv);
    htn_dfree(ctx, fa->dinvt);
    htn_dfree(ctx, fa->u);
    htn_dfree(ctx, fa->rdiag);
    htn_dfree(ctx, fa->pscr);
    htn_dfree(ctx, fa->linv);
    htn_dfree(ctx, fa->linvt);
    fa->dinv = fa->dinvt = fa->u = fa->rdiag = fa->linv = fa->linvt = fa->pscr = NULL;
}

int htn_factor_make_u(htn_ctx_t* ctx, htn_factor_t* fa)
{
    int rc = 0;
    if (!fa->u) HTN_TRY(htn_dalloc(ctx, (void**)&fa->u, (size_t)fa->n * fa->ld * sizeof(double)));
    /* the transpose writes every element of the n x n square (zeros below the diagonal without reading L's upper half):
     * only the padding columns, if any, need the memset */
    if ((size_t)fa->n != fa->ld)
        HTN_CUDA(cudaMemsetAsync(fa->u, 0, (size_t)fa->n * fa->ld * sizeof(double), ctx->stream));
    HTN_TRY(htnk_transpose_lower(ctx->stream, fa->f, fa->ld, fa->n, fa->u, fa->ld));
done:
    return rc;
}

/* side stream of the super-panel look-ahead (one per host thread and device, kept) */
static _Thread_local struct { int device; cudaStream_t side; cudaEvent_t ev_chain, ev_b; } sl = {-1, NULL, NULL, NULL};

/* Block column [c0, c1) of the factorisation (all rows from c0 down), RECURSIVELY: left half, one update of the right
 * half's columns by the left half (a GEMM with K = half the width - long and fat), right half.  The flat left-looking
 * loop over 128-panels updated every panel by all earlier panels of the super-panel: only skinny (N = 128) GEMMs, which
 * ran at 14 TFLOP/s (n = 8192: 2.4 ms of the 9.6); the recursion does the same flops with N = width / 2. */
static int potrf_block_column(htn_ctx_t* ctx, htn_factor_t* fa, int c0, int c1, int* d_info)
{
    int rc = 0;
    const int n = fa->n;
    const size_t ld = fa->ld;
    if (c1 - c0 <= HTN_NB) {
        const int nb = c1 - c0, rows = n - c1;
        double* ajj = fa->f + (size_t)c0 * ld + c0;
        int rows_done = 0;
        HTN_TRY(htnk_potf2_panel(ctx->stream, ajj, ld, nb, c0, fa->rdiag + c0, d_info, rows, fa->pscr,
                                 (unsigned*)(fa->pscr + 16 * 64), &rows_done));
        if (rows > 0 && !rows_done)
            HTN_TRY(htnk_trsm_rows(ctx->stream, fa->f + (size_t)c1 * ld + c0, ld, rows, ajj, ld, nb, fa->rdiag + c0));
        return 0;
    }
    const int cm = c0 + (int)htn_round_up((size_t)((c1 - c0) / 2), HTN_NB);
    HTN_TRY(potrf_block_column(ctx, fa, c0, cm, d_info));
    /* A[cm:, cm:c1] -= L[cm:, c0:cm] L[cm:c1, c0:cm]^T (the tiles above the diagonal of the square part are skipped) */
    HTN_TRY(htnk_gemm_tn(ctx->stream, n - cm, c1 - cm, cm - c0, -1.0, fa->f + (size_t)cm * ld + c0, ld,
                         fa->f + (size_t)cm * ld + c0, ld, 1.0, fa->f + (size_t)cm * ld + cm, ld, NULL, HTNK_GEMM_C_LOWER, 0, 0, 0));
    HTN_TRY(potrf_block_column(ctx, fa, cm, c1, d_info));
done:
    return rc;
}

int htn_potrf(htn_ctx_t* ctx, htn_factor_t* fa, int* d_info)
{
    int rc = 0;
    int sla_used = 0; /* work was queued on the look-ahead side stream: join it before returning, whatever happened */
    const int n = fa->n;
    const size_t ld = fa->ld;
    /* Two-level blocking.  Super-panels of W columns are factorised left-looking INSIDE the panel (block
     * column j gets the updates of the panel's earlier columns in one GEMM, then potf2 + row solve), and the
     * matrix to the right of the super-panel gets ONE right-looking rank-W update.  W = 128 is the plain
     * right-looking algorithm (most parallel tiles per launch: best for small n); large n uses wide panels so
     * that the trailing update has a long K loop (W / 16 chunks instead of 8) and the trailing matrix is
     * re-read and re-written n / W instead of n / 128 times. */
    static int la_max = -1; /* HTN_POTRF_LA (experiments): largest n factorised by the look-ahead variant */
    if (la_max < 0) {
        const char* e = getenv("HTN_POTRF_LA");
        la_max = e ? atoi(e) : 4096;
    }
    int W = n >= 8192 ? 1024 : (n >= 2048 ? 512 : HTN_NB);
    {
        const char* e = getenv("HTN_POTRF_W"); /* experiments: super-panel width */
        if (e && atoi(e) >= HTN_NB) W = (int)htn_round_up((size_t)atoi(e), HTN_NB);
    }
    static int la_min = -1;
    if (la_min < 0) {
        const char* e = getenv("HTN_POTRF_LA_MIN");
        la_min = e ? atoi(e) : 2 * HTN_NB + 1;
    }
    if (n >= la_min && n > 2 * HTN_NB && n <= la_max) {
        /* Up to n = 4096 (from n = 8192 the wide super-panels below win): plain right-looking with LOOK-AHEAD.  The
         * rank-128 update of panel j is split into (a) the next block column - all that potf2(j+1) and the row solve
         * (j+1) need - on the main stream, and (b) the rest of the trailing matrix on a side stream, beside potf2(j+1) +
         * solve(j+1), which are single-CTA / latency bound.  (a)_{j+1} writes what (b)_j writes, so it waits for it;
         * nothing else is shared.  (a) has its own kernel (htnk_panel_update: 7 us instead of the 15 us of the general
         * GEMM on a rows x 128 x 128 problem); with it the look-ahead pays from the smallest multi-panel orders on
         * (n = 500 / 1000 / 2000 / 4096: 0.159 / 0.314 / 0.69 / 2.04 ms against 0.170 / 0.353 / 0.76 / 2.11 ms for the
         * plain order or the general GEMM). */
        /* the side stream and the two events are created once per host thread and device, and kept */
        static _Thread_local struct { int device; cudaStream_t side; cudaEvent_t ev_t, ev_b; } la = {-1, NULL, NULL, NULL};
        if (la.device != ctx->device) {
            if (la.device >= 0) { /* another device was current before: let go of its objects */
                cudaEventDestroy(la.ev_t);
                cudaEventDestroy(la.ev_b);
                cudaStreamDestroy(la.side);
                la.device = -1;
            }
            if (cudaStreamCreateWithFlags(&la.side, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.ev_t, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.ev_b, cudaEventDisableTiming) == cudaSuccess)
                la.device = ctx->device;
            else
                rc = HTN_E_CUDA;
        }
        cudaStream_t side = la.side;
        cudaEvent_t ev_t = la.ev_t, ev_b = la.ev_b;
        static int pu_on = -1; /* HTN_POTRF_PU=0: the general GEMM for the next block column */
        if (pu_on < 0) {
            const char* e = getenv("HTN_POTRF_PU");
            pu_on = (e && *e == '0') ? 0 : 1;
        }
        int pending_b = 0;
        for (int j0 = 0; j0 < n && !rc; j0 += HTN_NB) {
            const int nb = n - j0 < HTN_NB ? n - j0 : HTN_NB;
            const int J1 = j0 + nb, rows = n - J1;
            double* ajj = fa->f + (size_t)j0 * ld + j0;
            int rows_done = 0;
            rc = htnk_potf2_panel(ctx->stream, ajj, ld, nb, j0, fa->rdiag + j0, d_info, rows, fa->pscr,
                                  (unsigned*)(fa->pscr + 16 * 64), &rows_done);
            if (rc || rows <= 0) break;
            if (!rows_done)
                rc = htnk_trsm_rows(ctx->stream, fa->f + (size_t)J1 * ld + j0, ld, rows, ajj, ld, nb, fa->rdiag + j0);
            if (rc) break;
            const int nb2 = rows < HTN_NB ? rows : HTN_NB, J2 = J1 + nb2, R2 = n - J2;
            const double* pan = fa->f + (size_t)J1 * ld + j0; /* L[J1:, j0:J1] */
            if (R2 > 0 && cudaEventRecord(ev_t, ctx->stream) != cudaSuccess) { rc = HTN_E_CUDA; break; }
            if (pending_b && cudaStreamWaitEvent(ctx->stream, ev_b, 0) != cudaSuccess) { rc = HTN_E_CUDA; break; }
            pending_b = 0;
            /* (a) next block column: A[J1:, J1:J2] -= L[J1:, j] * L[J1:J2, j]^T */
            rc = nb == HTN_NB && pu_on ? htnk_panel_update(ctx->stream, fa->f + (size_t)J1 * ld + J1, ld, pan, ld, rows, nb2) : 1;
            if (rc == 1)
                rc = htnk_gemm_tn(ctx->stream, rows, nb2, nb, -1.0, pan, ld, pan, ld, 1.0, fa->f + (size_t)J1 * ld + J1, ld,
                                  NULL, HTNK_GEMM_FULL, 0, 0, 0);
            if (rc) break;
            if (R2 > 0) { /* (b) the rest, lower tiles: A[J2:, J2:] -= L[J2:, j] * L[J2:, j]^T */
                const double* pan2 = fa->f + (size_t)J2 * ld + j0;
                if (cudaStreamWaitEvent(side, ev_t, 0) != cudaSuccess) { rc = HTN_E_CUDA; break; }
                rc = htnk_gemm_tn(side, R2, R2, nb, -1.0, pan2, ld, pan2, ld, 1.0, fa->f + (size_t)J2 * ld + J2, ld, NULL,
                                  HTNK_GEMM_C_LOWER, 0, 0, 0);
                if (rc) break;
                if (cudaEventRecord(ev_b, side) != cudaSuccess) { rc = HTN_E_CUDA; break; }
                pending_b = 1;
            }
        }
        if (pending_b) cudaStreamWaitEvent(ctx->stream, ev_b, 0); /* the factor is complete on the main stream */
        if (rc == HTN_E_CUDA) htn_set_error("potrf look-ahead: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc) goto done;
        if (fa->solves) HTN_TRY(htnk_trtri_batched(ctx->stream, fa->f, ld, n, fa->dinv, fa->dinvt));
        goto done;
    }
    /* Super-panel LOOK-AHEAD (n >= HTN_POTRF_SLA, default 4097): the rank-W update of super-panel J is split into (a) the
     * columns of super-panel J + 1 - all that its factorisation chain reads - on the caller's stream, and (b) the rest
     * of the trailing matrix on a LOW-PRIORITY side stream.  The chain of super-panel J + 1 (64 fused panel launches of
     * ~50 us at n = 8192: 3.5 ms of latency-bound work on a few SMs) then runs BESIDE (b), the one kernel that can fill the
     * device.  (a) of J + 1 writes what (b) of J writes, so it waits for it; nothing else is shared. */
    static int sla_min = -1;
    if (sla_min < 0) {
        const char* e = getenv("HTN_POTRF_SLA");
        sla_min = e ? atoi(e) : 4097;
    }
    int use_sla = n >= sla_min && n > 2 * W;
    if (use_sla && sl.device != ctx->device) {
        if (sl.device >= 0) {
            cudaEventDestroy(sl.ev_chain);
            cudaEventDestroy(sl.ev_b);
            cudaStreamDestroy(sl.side);
            sl.device = -1;
        }
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio); /* numerically largest = lowest priority */
        if (cudaStreamCreateWithPriority(&sl.side, cudaStreamNonBlocking, lo_prio) == cudaSuccess &&
            cudaEventCreateWithFlags(&sl.ev_chain, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&sl.ev_b, cudaEventDisableTiming) == cudaSuccess)
            sl.device = ctx->device;
        else {
            cudaGetLastError();
            use_sla = 0; /* no side stream: the plain order below */
        }
    }
    static int rec_on = -1; /* HTN_POTRF_REC=0: the flat left-looking loop inside a super-panel */
    if (rec_on < 0) {
        const char* e = getenv("HTN_POTRF_REC");
        rec_on = (e && *e == '0') ? 0 : 1;
    }
    int pending_b = 0;
    for (int J0 = 0; J0 < n; J0 += W) {
        const int J1 = J0 + W < n ? J0 + W : n;
        if (rec_on && J1 - J0 > HTN_NB) {
            HTN_TRY(potrf_block_column(ctx, fa, J0, J1, d_info));
        } else
        for (int j0 = J0; j0 < J1; j0 += HTN_NB) {
            const int nb = n - j0 < HTN_NB ? n - j0 : HTN_NB;
            const int rows = n - j0 - nb;
            double* ajj = fa->f + (size_t)j0 * ld + j0;
            if (j0 > J0) /* A[j0:, j0:j0+nb] -= L[j0:, J0:j0] * L[j0:j0+nb, J0:j0]^T */
                HTN_TRY(htnk_gemm_tn(ctx->stream, n - j0, nb, j0 - J0, -1.0, fa->f + (size_t)j0 * ld + J0, ld,
                                     fa->f + (size_t)j0 * ld + J0, ld, 1.0, ajj, ld, NULL, HTNK_GEMM_FULL, 0, 0, 0));
            /* diagonal block and, in the same launch when the block is full, the panel under it
             * (L21 = A21 * L11^-T, rows independent, in place) */
            int rows_done = 0;
            HTN_TRY(htnk_potf2_panel(ctx->stream, ajj, ld, nb, j0, fa->rdiag + j0, d_info, rows, fa->pscr,
                                     (unsigned*)(fa->pscr + 16 * 64), &rows_done));
            if (rows > 0 && !rows_done)
                HTN_TRY(htnk_trsm_rows(ctx->stream, fa->f + (size_t)(j0 + nb) * ld + j0, ld, rows, ajj, ld, nb,
                                       fa->rdiag + j0));
        }
        const int R = n - J1;
        if (R <= 0) break;
        const double* pan = fa->f + (size_t)J1 * ld + J0; /* L[J1:, J0:J1] */
        if (!use_sla) { /* trailing update: A[J1:, J1:] -= L[J1:, J0:J1] * L[J1:, J0:J1]^T on the lower tiles (DMMA) */
            HTN_TRY(htnk_gemm_tn(ctx->stream, R, R, J1 - J0, -1.0, pan, ld, pan, ld, 1.0, fa->f + (size_t)J1 * ld + J1, ld, NULL,
                                 HTNK_GEMM_C_LOWER, 0, 0, 0));
            continue;
        }
        const int J2 = J1 + W < n ? J1 + W : n, R2 = n - J2;
        if (R2 > 0) HTN_CUDA(cudaEventRecord(sl.ev_chain, ctx->stream));
        if (pending_b) { /* (b) of the previous super-panel wrote the block (a) is about to update */
            HTN_CUDA(cudaStreamWaitEvent(ctx->stream, sl.ev_b, 0));
            pending_b = 0;
        }
        /* (a) A[J1:, J1:J2] -= L[J1:, J] L[J1:J2, J]^T: the rows from J1 on, the next super-panel's columns */
        HTN_TRY(htnk_gemm_tn(ctx->stream, R, J2 - J1, J1 - J0, -1.0, pan, ld, pan, ld, 1.0, fa->f + (size_t)J1 * ld + J1, ld, NULL,
                             HTNK_GEMM_C_LOWER, 0, 0, 0));
        if (R2 > 0) { /* (b) A[J2:, J2:] -= L[J2:, J] L[J2:, J]^T, beside the next chain */
            const double* pan2 = fa->f + (size_t)J2 * ld + J0;
            HTN_CUDA(cudaStreamWaitEvent(sl.side, sl.ev_chain, 0));
            sla_used = 1;
            rc = htnk_gemm_tn(sl.side, R2, R2, J1 - J0, -1.0, pan2, ld, pan2, ld, 1.0, fa->f + (size_t)J2 * ld + J2, ld, NULL,
                              HTNK_GEMM_C_LOWER, 0, 0, 0);
            if (!rc && cudaEventRecord(sl.ev_b, sl.side) != cudaSuccess) rc = HTN_E_CUDA;
            pending_b = 1; /* joined below even on failure: the side stream may have work queued */
            if (rc) break;
        }
    }
    if (rc) goto done;
    /* inverses of the diagonal blocks, all at once, only when blocked solves will follow */
    if (sla_used) { /* the factor is complete on the caller's stream only once (b) of the last super-panels has landed */
        if (cudaEventRecord(sl.ev_b, sl.side) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, sl.ev_b, 0) != cudaSuccess)
            cudaStreamSynchronize(sl.side);
        sla_used = 0;
    }
    if (fa->solves) HTN_TRY(htnk_trtri_batched(ctx->stream, fa->f, ld, n, fa->dinv, fa->dinvt));
done:
    if (sla_used && (cudaEventRecord(sl.ev_b, sl.side) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, sl.ev_b, 0) != cudaSuccess))
        cudaStreamSynchronize(sl.side); /* error exit: never leave the side stream running under buffers about to be freed */
    return rc;
}

/* X_new * L^T = X : forward over column blocks */
int htn_trsm_right_lt(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m)
{
    int rc = 0;
    for (int jb = 0; jb < fa->nblk; jb++) {
        const int j0 = jb * HTN_NB;
        const int nb = fa->n - j0 < HTN_NB ? fa->n - j0 : HTN_NB;
        double* xj = x + j0;
        if (j0 > 0) /* X_j -= X_{<j} * L[j, <j]^T */
            HTN_TRY(htnk_gemm_tn(ctx->stream, m, nb, j0, -1.0, x, ldx, fa->f + (size_t)j0 * fa->ld, fa->ld, 1.0,
                                 xj, ldx, NULL, HTNK_GEMM_FULL, 0, 0, 0));
        /* X_j = X_j * inv(L_jj)^T : B[n][kk] = dinv[n][kk].  In place (C == A): one 128-wide N tile per
         * row slab, whose K loop has consumed every input before its epilogue writes */
        HTN_TRY(htnk_gemm_tn(ctx->stream, m, nb, nb, 1.0, xj, ldx, fa->dinv + (size_t)jb * HTN_NB * HTN_NB, HTN_NB,
                             0.0, xj, ldx, NULL, HTNK_GEMM_FULL, 0, 0, 128));
    }
done:
    return rc;
}

/* X_new * L = X : backward over column blocks; B[n][kk] = L[j1+kk][j0+n] = U[j0+n][j1+kk] */
int htn_trsm_right_l(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m)
{
    int rc = 0;
    if (!fa->u) {
        htn_set_error("internal: factor has no transposed copy");
        return HTN_E_ARG;
    }
    for (int jb = fa->nblk - 1; jb >= 0; jb--) {
        const int j0 = jb * HTN_NB;
        const int nb = fa->n - j0 < HTN_NB ? fa->n - j0 : HTN_NB;
        const int j1 = j0 + nb;
        double* xj = x + j0;
        if (j1 < fa->n)
            HTN_TRY(htnk_gemm_tn(ctx->stream, m, nb, fa->n - j1, -1.0, x + j1, ldx,
                                 fa->u + (size_t)j0 * fa->ld + j1, fa->ld, 1.0, xj, ldx, NULL, HTNK_GEMM_FULL, 0,
                                 0, 0));
        /* X_j = X_j * inv(L_jj) : B[n][kk] = dinv[kk][n] = dinvt[n][kk] */
        HTN_TRY(htnk_gemm_tn(ctx->stream, m, nb, nb, 1.0, xj, ldx, fa->dinvt + (size_t)jb * HTN_NB * HTN_NB,
                             HTN_NB, 0.0, xj, ldx, NULL, HTNK_GEMM_FULL, 0, 0, 128));
    }
done:
    return rc;
}

/* Explicit inverse of the factor, so that solves with MANY right-hand-side rows become one triangular DMMA
 * GEMM each (full-width grids) instead of n/128 dependent panel steps: linv = L^-1 (lower) by the blocked
 * solve X L = I, linvt = its transpose. */
/* inv(L[lo:hi, lo:hi]) into linv (and its transpose into linvt), by halving:
 *   L = [L11 0; L21 L22]  ->  L^-1 = [L11^-1 0; -L22^-1 L21 L11^-1, L22^-1]
 * The leaves are the 128 x 128 diagonal-block inverses of trtri_batched; every inner node is two large GEMMs
 * (instead of a 16-step column sweep whose GEMMs have 64 CTAs each): n = 2000: 1.5 -> 0.5 ms. */
/* The two halves of a node are independent until the node's own products: the left half of the top two levels runs
 * on a side stream (four sub-trees in flight; the lower levels are chains of small launches, n = 2000: 92 of them).
 * Streams and events are created once per host thread and device and kept. */
#define INV_FORKS 3 /* root, and its two children */
static _Thread_local struct {
    int device;
    cudaStream_t side[INV_FORKS];
    cudaEvent_t ev_fork[INV_FORKS], ev_join[INV_FORKS];
} iv = {-1, {NULL, NULL, NULL}, {NULL, NULL, NULL}, {NULL, NULL, NULL}};

static int inverse_forks_ready(int device)
{
    if (iv.device == device) return 1;
    if (iv.device >= 0) { /* another device was current before: let go of its objects */
        for (int i = 0; i < INV_FORKS; i++) {
            cudaStreamDestroy(iv.side[i]);
            cudaEventDestroy(iv.ev_fork[i]);
            cudaEventDestroy(iv.ev_join[i]);
        }
        iv.device = -1;
    }
    int made = 0;
    for (; made < INV_FORKS; made++) {
        if (cudaStreamCreateWithFlags(&iv.side[made], cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaEventCreateWithFlags(&iv.ev_fork[made], cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamDestroy(iv.side[made]);
            break;
        }
        if (cudaEventCreateWithFlags(&iv.ev_join[made], cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamDestroy(iv.side[made]);
            cudaEventDestroy(iv.ev_fork[made]);
            break;
        }
    }
    if (made < INV_FORKS) { /* no forks then: the recursion simply stays on one stream */
        for (int i = 0; i < made; i++) {
            cudaStreamDestroy(iv.side[i]);
            cudaEventDestroy(iv.ev_fork[i]);
            cudaEventDestroy(iv.ev_join[i]);
        }
        cudaGetLastError();
        return 0;
    }
    iv.device = device;
    return 1;
}

/* slot: fork resources this node may use (0 root, 1 / 2 its children), -1 = stay on the stream.  The leaves (diagonal-
 * block inverses) are already in place (htnk_scatter_diag_blocks). */
static int inverse_rec(htn_ctx_t* ctx, htn_factor_t* fa, int lo, int hi, double* tt, size_t ldt, int slot)
{
    int rc = 0;
    const size_t ld = fa->ld;
    const int n = hi - lo;
    if (n <= HTN_NB) return 0;
    const int mid = lo + (int)htn_round_up((size_t)(n / 2), HTN_NB);
    if (slot >= 0 && n >= 4 * HTN_NB) {
        const size_t half = htn_round_up((size_t)((mid - lo) / 2), HTN_NB) + HTN_NB;
        const size_t ldt2 = htn_round_up(half, 16);
        double* tt2 = NULL;
        HTN_TRY(htn_dalloc(ctx, (void**)&tt2, 2 * ldt2 * ldt2 * sizeof(double)));
        htn_ctx_t left = *ctx;
        left.stream = iv.side[slot];
        int forked = cudaEventRecord(iv.ev_fork[slot], ctx->stream) == cudaSuccess &&
                     cudaStreamWaitEvent(left.stream, iv.ev_fork[slot], 0) == cudaSuccess;
        if (!forked) left.stream = ctx->stream;
        const int child = slot == 0 ? 1 : -1; /* the root's children fork once more */
        int rc_l = inverse_rec(&left, fa, lo, mid, tt2, ldt2, forked ? child : -1);
        int rc_r = inverse_rec(ctx, fa, mid, hi, tt, ldt, forked && slot == 0 ? 2 : -1);
        if (forked && (cudaEventRecord(iv.ev_join[slot], left.stream) != cudaSuccess ||
                       cudaStreamWaitEvent(ctx->stream, iv.ev_join[slot], 0) != cudaSuccess)) {
            cudaStreamSynchronize(left.stream); /* never free a buffer the side stream may still be using */
            if (!rc_l) rc_l = HTN_E_CUDA;
        }
        htn_dfree(ctx, tt2);
        if (rc_l || rc_r) return rc_l ? rc_l : rc_r;
    } else {
        HTN_TRY(inverse_rec(ctx, fa, lo, mid, tt, ldt, -1));
        HTN_TRY(inverse_rec(ctx, fa, mid, hi, tt, ldt, -1));
    }
    {
        /* both products have a triangular operand, placed where the GEMM can skip its zero blocks (the B side) */
        const int M = hi - mid, N = mid - lo;
        double* t = tt;                      /* T   (M x N) */
        double* ttr = tt + (size_t)ldt * ldt; /* T^T (N x M) */
        /* T = L21 * inv(L11):  B[n][k] = inv(L11)[k][n] = linvt11[n][k], zero for k < n */
        HTN_TRY(htnk_gemm_tn(ctx->stream, M, N, N, 1.0, fa->f + (size_t)mid * ld + lo, ld,
                             fa->linvt + (size_t)lo * ld + lo, ld, 0.0, t, ldt, NULL, HTNK_GEMM_B_UPPER, 0, 0, 0));
        HTN_TRY(htnk_transpose(ctx->stream, t, ldt, M, N, ttr, ldt));
        /* X21^T (N x M) = -T^T * inv(L22)^T:  B[m][k] = inv(L22)[m][k], zero for k > m; straight into linvt */
        double* x21t = fa->linvt + (size_t)lo * ld + mid;
        HTN_TRY(htnk_gemm_tn(ctx->stream, N, M, M, -1.0, ttr, ldt, fa->linv + (size_t)mid * ld + mid, ld, 0.0, x21t, ld,
                             NULL, HTNK_GEMM_B_LOWER, M, M, 0));
        HTN_TRY(htnk_transpose(ctx->stream, x21t, ld, N, M, fa->linv + (size_t)mid * ld + lo, ld));
    }
done:
    return rc;
}

int htn_factor_make_inverse(htn_ctx_t* ctx, htn_factor_t* fa)
{
    int rc = 0;
    const size_t bytes = (size_t)fa->n * fa->ld * sizeof(double);
    double* tt = NULL;
    if (!fa->linv) HTN_TRY(htn_dalloc(ctx, (void**)&fa->linv, bytes));
    if (!fa->linvt) HTN_TRY(htn_dalloc(ctx, (void**)&fa->linvt, bytes));
    HTN_CUDA(cudaMemsetAsync(fa->linv, 0, bytes, ctx->stream));
    HTN_CUDA(cudaMemsetAsync(fa->linvt, 0, bytes, ctx->stream));
    fa->inv_block = 0;
    if (fa->dinv && fa->dinvt) {
        const size_t half = htn_round_up((size_t)(fa->n / 2), HTN_NB) + HTN_NB;
        const size_t ldt = htn_round_up(half, 16);
        HTN_TRY(htn_dalloc(ctx, (void**)&tt, 2 * ldt * ldt * sizeof(double))); /* T and T^T */
        rc = htnk_scatter_diag_blocks(ctx->stream, fa->dinv, fa->dinvt, HTN_NB, fa->n, fa->linv, fa->linvt, fa->ld);
        if (!rc) rc = inverse_rec(ctx, fa, 0, fa->n, tt, ldt, inverse_forks_ready(ctx->device) ? 0 : -1);
        htn_dfree(ctx, tt);
        return rc;
    }
    /* no block inverses at hand: column sweep on the identity */
    HTN_TRY(htnk_diag_set(ctx->stream, fa->linv, fa->ld, fa->n, NULL, 1.0));
    HTN_TRY(htn_trsm_right_l(ctx, fa, fa->linv, fa->ld, fa->n));
    HTN_TRY(htnk_transpose(ctx->stream, fa->linv, fa->ld, fa->n, fa->n, fa->linvt, fa->ld));
done:
    return rc;
}

int htn_factor_make_inverse_blocks(htn_ctx_t* ctx, htn_factor_t* fa, int w)
{
    int rc = 0;
    const size_t bytes = (size_t)fa->n * fa->ld * sizeof(double);
    double* tt = NULL;
    if (w <= 0 || w % HTN_NB || !fa->dinv || !fa->dinvt || !fa->u || w >= fa->n) return htn_factor_make_inverse(ctx, fa);
    if (!fa->linv) HTN_TRY(htn_dalloc(ctx, (void**)&fa->linv, bytes));
    if (!fa->linvt) HTN_TRY(htn_dalloc(ctx, (void**)&fa->linvt, bytes));
    /* only the diagonal w x w blocks are ever read (solve_*_blocks): zero those, not the whole matrices (p = 8192: 268 MB
     * instead of 1 GB) */
    for (int lo = 0; lo < fa->n; lo += w) {
        const size_t wj = (size_t)(lo + w < fa->n ? w : fa->n - lo);
        HTN_CUDA(cudaMemset2DAsync(fa->linv + (size_t)lo * fa->ld + lo, fa->ld * sizeof(double), 0, wj * sizeof(double), wj, ctx->stream));
        HTN_CUDA(cudaMemset2DAsync(fa->linvt + (size_t)lo * fa->ld + lo, fa->ld * sizeof(double), 0, wj * sizeof(double), wj, ctx->stream));
    }
    {
        const size_t half = htn_round_up((size_t)(w / 2), HTN_NB) + HTN_NB;
        const size_t ldt = htn_round_up(half, 16);
        const int forks = inverse_forks_ready(ctx->device);
        const int nblk = (fa->n + w - 1) / w;
        const size_t per = 2 * ldt * ldt; /* T and T^T of one block's recursion */
        static int par_on = -1; /* HTN_INV_PAR=0: the diagonal blocks one after the other (each forking inside) */
        if (par_on < 0) {
            const char* e = getenv("HTN_INV_PAR");
            par_on = (e && *e == '0') ? 0 : 1;
        }
        if (forks && par_on && nblk >= 2) {
            /* The diagonal blocks are independent: block b runs on stream b mod 4 (the caller's and the three fork streams),
             * each with its own scratch and without forking inside - the recursion is a chain of small launches, so four
             * chains side by side overlap their latencies (cfg4, four 2048-blocks: the call 18.83 -> 18.59 ms; the products themselves,
             * 11.5 GFLOP at mid-size GEMM efficiency, and the 1 GB of memsets remain). */
            HTN_TRY(htn_dalloc(ctx, (void**)&tt, (size_t)nblk * per * sizeof(double)));
            rc = htnk_scatter_diag_blocks(ctx->stream, fa->dinv, fa->dinvt, HTN_NB, fa->n, fa->linv, fa->linvt, fa->ld);
            int forked = 0;
            if (!rc) {
                forked = cudaEventRecord(iv.ev_fork[0], ctx->stream) == cudaSuccess;
                for (int i = 0; i < INV_FORKS && forked; i++)
                    if (cudaStreamWaitEvent(iv.side[i], iv.ev_fork[0], 0) != cudaSuccess) rc = HTN_E_CUDA;
            }
            for (int b = 0; b < nblk && !rc; b++) {
                const int lo = b * w, hi = lo + w < fa->n ? lo + w : fa->n, si = forked ? b % (INV_FORKS + 1) : 0;
                htn_ctx_t c2 = *ctx;
                if (si > 0) c2.stream = iv.side[si - 1];
                rc = inverse_rec(&c2, fa, lo, hi, tt + (size_t)b * per, ldt, -1);
            }
            for (int i = 0; i < INV_FORKS && forked; i++) /* join, whatever happened: the scratch is freed on the caller's stream */
                if (cudaEventRecord(iv.ev_join[i], iv.side[i]) != cudaSuccess ||
                    cudaStreamWaitEvent(ctx->stream, iv.ev_join[i], 0) != cudaSuccess) {
                    cudaStreamSynchronize(iv.side[i]);
                    if (!rc) rc = HTN_E_CUDA;
                }
        } else {
            HTN_TRY(htn_dalloc(ctx, (void**)&tt, per * sizeof(double)));
            rc = htnk_scatter_diag_blocks(ctx->stream, fa->dinv, fa->dinvt, HTN_NB, fa->n, fa->linv, fa->linvt, fa->ld);
            for (int lo = 0; lo < fa->n && !rc; lo += w)
                rc = inverse_rec(ctx, fa, lo, lo + w < fa->n ? lo + w : fa->n, tt, ldt, forks ? 0 : -1);
        }
        htn_dfree(ctx, tt);
        if (!rc) fa->inv_block = w;
    }
done:
    return rc;
}

/* block substitution with the diagonal-block inverses (fa->inv_block = W): per block column one update GEMM with the
 * columns already solved and one triangular GEMM with the block's inverse */
static int solve_l_blocks(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m)
{
    int rc = 0;
    const int n = fa->n, w = fa->inv_block;
    const size_t ld = fa->ld, ldt = htn_round_up((size_t)w, 16);
    double* t = NULL;
    HTN_TRY(htn_dalloc(ctx, (void**)&t, (size_t)m * ldt * sizeof(double)));
    for (int c0 = (n - 1) / w * w; c0 >= 0; c0 -= w) { /* Y L = X: last block column first */
        const int wj = n - c0 < w ? n - c0 : w, r1 = c0 + wj, k = n - r1;
        const double* src = x + c0;
        size_t lds = ldx;
        if (k > 0) { /* T = X_j - Y[:, r1:] L[r1:, c0:c0+wj];  B[nn][kk] = L[r1+kk][c0+nn] = U[c0+nn][r1+kk] */
            HTN_TRY(htnk_copy2d(ctx->stream, x + c0, ldx, m, wj, t, ldt, NULL));
            HTN_TRY(htnk_gemm_tn(ctx->stream, m, wj, k, -1.0, y + r1, ldy, fa->u + (size_t)c0 * ld + r1, ld, 1.0, t, ldt,
                                 NULL, HTNK_GEMM_FULL, 0, 0, 0));
            src = t;
            lds = ldt;
        }
        HTN_TRY(htnk_gemm_tn(ctx->stream, m, wj, wj, 1.0, src, lds, fa->linvt + (size_t)c0 * ld + c0, ld, 0.0, y + c0, ldy,
                             NULL, HTNK_GEMM_B_UPPER, 0, 0, 0));
    }
done:
    htn_dfree(ctx, t);
    return rc;
}

static int solve_lt_blocks(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m)
{
    int rc = 0;
    const int n = fa->n, w = fa->inv_block;
    const size_t ld = fa->ld, ldt = htn_round_up((size_t)w, 16);
    double* t = NULL;
    HTN_TRY(htn_dalloc(ctx, (void**)&t, (size_t)m * ldt * sizeof(double)));
    for (int c0 = 0; c0 < n; c0 += w) { /* Y L^T = X: first block column first */
        const int wj = n - c0 < w ? n - c0 : w;
        const double* src = x + c0;
        size_t lds = ldx;
        if (c0 > 0) { /* T = X_j - Y[:, :c0] L[c0:c0+wj, :c0]^T;  B[nn][kk] = L[c0+nn][kk] */
            HTN_TRY(htnk_copy2d(ctx->stream, x + c0, ldx, m, wj, t, ldt, NULL));
            HTN_TRY(htnk_gemm_tn(ctx->stream, m, wj, c0, -1.0, y, ldy, fa->f + (size_t)c0 * ld, ld, 1.0, t, ldt, NULL,
                                 HTNK_GEMM_FULL, 0, 0, 0));
            src = t;
            lds = ldt;
        }
        HTN_TRY(htnk_gemm_tn(ctx->stream, m, wj, wj, 1.0, src, lds, fa->linv + (size_t)c0 * ld + c0, ld, 0.0, y + c0, ldy,
                             NULL, HTNK_GEMM_B_LOWER, wj, wj, 0));
    }
done:
    htn_dfree(ctx, t);
    return rc;
}

/* Y (m x n) = X L^-1      (solves Y L = X; the row form of L^T y = x, dtrsv 'L','T') */
int htn_solve_l(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m)
{
    if (fa->inv_block) return solve_l_blocks(ctx, fa, x, ldx, y, ldy, m);
    return htnk_gemm_tn(ctx->stream, m, fa->n, fa->n, 1.0, x, ldx, fa->linvt, fa->ld, 0.0, y, ldy, NULL,
                        HTNK_GEMM_B_UPPER, 0, 0, 0);
}

/* Y = X L^-T     (solves Y L^T = X) */
int htn_solve_lt(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m)
{
    if (fa->inv_block) return solve_lt_blocks(ctx, fa, x, ldx, y, ldy, m);
    return htnk_gemm_tn(ctx->stream, m, fa->n, fa->n, 1.0, x, ldx, fa->linv, fa->ld, 0.0, y, ldy, NULL,
                        HTNK_GEMM_B_LOWER, fa->n, fa->n, 0);
}

/* Y = X (L L^T)^-1 through a scratch buffer t (m x n, leading dimension ldx) */
int htn_solve_spd(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* t, double* y,
                  size_t ldy, int m)
{
    int rc = htn_solve_lt(ctx, fa, x, ldx, t, ldx, m);
    if (!rc) rc = htn_solve_l(ctx, fa, t, ldx, y, ldy, m);
    return rc;
}

int htn_potrs_rows(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m)
{
    int rc = htn_trsm_right_lt(ctx, fa, x, ldx, m);
    if (!rc) rc = htn_trsm_right_l(ctx, fa, x, ldx, m);
    return rc;
}
