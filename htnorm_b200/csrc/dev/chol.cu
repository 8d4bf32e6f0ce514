// Shared-memory / register kernels of the blocked Cholesky and of the blocked triangular solves:
//   potf2_kernel         factor one nb x nb (nb <= 128) diagonal block (dpotrf's unblocked base case;
//                        reference call sites src/htnorm_distributions.c:77,115 via src/htnorm_blas.h:111)
//   trsm_rows_kernel     panel solve L21 = A21 L11^-T, rows independent (no inverse needed)
//   trtri_batched_kernel inverse of every diagonal block at once, so that the many-right-hand-side
//                        solves of the samplers (dtrsv / dposv / dpotri's solve phases,
//                        src/htnorm_distributions.c:120, src/htnorm.c:169,210,330) become DMMA GEMMs
// Pivot rule: fail on `ajj <= 0` only - NaN passes and propagates with info = 0, as the reference +
// OpenBLAS do (reference tests/test_pyhtnorm.py:46-48, 91-95).
//
// One CTA of 512 threads, the block in REGISTERS (32 per thread, one owner per element): a column costs
// one barrier + one fast rsqrt + the owner-local rank-1 update.
#include <cuda_runtime.h>

#include "launchers.h"

namespace {

constexpr int NB = 128;
constexpr int LDB = NB + 4;  // 132
constexpr int NT = 512;

// 1/sqrt(x): MUFU seed (rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps in FP64 (~1 ulp).  About a third
// of the latency of the library rsqrt()/sqrt()+div pair, and it sits on the critical path of EVERY column.
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double h = 0.5 * y;
        const double e = fma(-x * y, h, 0.5);  // 0.5 - 0.5 x y^2
        y = fma(y, e, y);
    }
    return y;
}

// Thread (i = tid / 4, cq = tid % 4) OWNS the elements (i, c), c == cq (mod 4), and keeps them in 32
// REGISTERS used as a ROTATING window: during column group jj (columns 4jj .. 4jj+3) R[u] holds the
// element of column cq + 4 (jj + u); after the group R[0] is final and the window shifts by one, so the
// loop body is identical for every group (a fully unrolled 128-column version measured I-cache bound).
//
// Step j is software-pipelined around ONE barrier:
//   urgent : pivot rsqrt (redundantly in every thread), then the owners of column j+1 apply step j to
//            that column only and publish it (unscaled) to shared memory        -> barrier
//   lazy   : the rest of the rank-1 update of step j runs after the barrier, i.e. in the shadow of the
//            other warps' next urgent part.
// The published columns rotate through three buffers (a warp still in lazy(j) reads column j while
// column j+2 is being published).  No per-element predicates: updates that fall above the diagonal or
// beyond the block land in registers / zero padding that are never read.
__global__ void __launch_bounds__(NT, 1)
potf2_kernel(double* __restrict__ a, size_t ld, int nb, int col0, double* __restrict__ rdiag_out,
             int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]  staging for the coalesced load / store
    double* colbuf = sm + NB * LDB;  // [3][2 * NB] published columns, second half stays zero
    double* rdiag = colbuf + 6 * NB; // [NB]       1 / L[j][j]
    const int tid = threadIdx.x;
    const int i = tid >> 2, cq = tid & 3;

    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : 0.0;
    }
    for (int idx = tid; idx < 6 * NB; idx += NT) colbuf[idx] = 0.0;
    __syncthreads();
    double R[32];
#pragma unroll
    for (int jj = 0; jj < 32; jj++) R[jj] = S[i * LDB + cq + 4 * jj];
    if (cq == 0) colbuf[i] = R[0];  // column 0
    __syncthreads();

    int nfact = nb;
    int bj = 0;  // buffer that holds column j
    const int ngroups = (nb + 3) >> 2;
    for (int jj = 0; jj < ngroups; jj++) {
        const int c0 = cq + 4 * jj;  // the column this thread retires in this group
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = 4 * jj + q;
            if (j < nb && nfact == nb) {  // uniform
                const double* col = colbuf + bj * 2 * NB;
                const int bn = bj == 2 ? 0 : bj + 1;
                double* coln = colbuf + bn * 2 * NB;
                const double ajj = col[j];
                double m = 0.0, aij = 0.0, rinv = 0.0;
                const bool ok = !(ajj <= 0.0);  // uniform; NaN passes
                if (ok) {
                    rinv = fast_rsqrt(ajj);
                    if (i > j) {
                        aij = col[i];
                        m = aij * (rinv * rinv);
                        // urgent: column j+1 (register R[0] of the thread cq == q+1, or R[1] of cq == 0)
                        if (q < 3) {
                            if (cq == q + 1) { R[0] = fma(-m, col[j + 1], R[0]); coln[i] = R[0]; }
                        } else {
                            if (cq == 0) { R[1] = fma(-m, col[j + 1], R[1]); coln[i] = R[1]; }
                        }
                    }
                } else {
                    if (tid == 0) atomicCAS(info, 0, col0 + j + 1);
                    nfact = j;
                }
                __syncthreads();
                if (ok) {
                    if (i > j) {
                        // lazy: every other owned column > j+1
                        if (cq > q + 1) R[0] = fma(-m, col[c0], R[0]);
                        if (!(q == 3 && cq == 0)) R[1] = fma(-m, col[c0 + 4], R[1]);
#pragma unroll
                        for (int u = 2; u < 8; u++) R[u] = fma(-m, col[c0 + 4 * u], R[u]);
#pragma unroll
                        for (int ub = 1; ub < 4; ub++) {
                            if (jj + 8 * ub < 32) {  // uniform: later columns do not exist
#pragma unroll
                                for (int u = 8 * ub; u < 8 * ub + 8; u++) R[u] = fma(-m, col[c0 + 4 * u], R[u]);
                            }
                        }
                        if (cq == q) R[0] = aij * rinv;  // L[i][j]
                    } else if (i == j && cq == q) {
                        double d = ajj * rinv;  // sqrt(ajj), one Heron correction
                        d = fma(fma(-d, d, ajj), 0.5 * rinv, d);
                        R[0] = d;
                    }
                    if (tid == 0) rdiag[j] = rinv;
                }
                bj = bn;
            }
        }
        // retire column c0: L[i][c0] (zeros above the diagonal / beyond a failed column)
        S[i * LDB + c0] = (c0 <= i && c0 < nfact) ? R[0] : 0.0;
#pragma unroll
        for (int u = 0; u < 31; u++) R[u] = R[u + 1];
        R[31] = 0.0;
    }
    const bool failed = nfact < nb;
    for (int c = cq + 4 * ngroups; c < NB; c += 4) S[i * LDB + c] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < nb * nb; idx += NT) {
        int r = idx / nb, c = idx - r * nb;
        a[(size_t)r * ld + c] = S[r * LDB + c];
    }
    if (rdiag_out != nullptr && !failed)
        for (int c = tid; c < nb; c += NT) rdiag_out[c] = rdiag[c];
}

// ---- inverse of diagonal blocks, batched: one CTA per block, all blocks in parallel ----
// inv(L) by forward elimination on the identity held in registers (thread (i, cq) owns columns
// c == cq mod 4 of row i); row t of the inverse is final after step t and is published by 128 threads
// (coalesced) while the others update.  Only launched for factors that feed the blocked TRSMs.
__global__ void __launch_bounds__(NT, 1)
trtri_batched_kernel(const double* __restrict__ f, size_t ld, int n, double* __restrict__ dinv_all,
                     double* __restrict__ dinvt_all)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]
    double* xrow = sm + NB * LDB;    // [2][NB]
    double* rdiag = xrow + 2 * NB;   // [NB]
    const int tid = threadIdx.x;
    const int i = tid >> 2, cq = tid & 3;
    const int j0 = blockIdx.x * NB;
    const int nb = min(NB, n - j0);
    const double* a = f + (size_t)j0 * ld + j0;
    double* dinv = dinv_all + (size_t)blockIdx.x * NB * NB;
    double* dinvt = dinvt_all + (size_t)blockIdx.x * NB * NB;

    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    if (tid < nb) rdiag[tid] = 1.0 / S[tid * LDB + tid];
    double R[32];
#pragma unroll
    for (int jj = 0; jj < 32; jj++) R[jj] = (cq + 4 * jj == i) ? 1.0 : 0.0;
    __syncthreads();
    for (int t = 0; t < nb; t++) {
        double* xr = xrow + (t & 1) * NB;
        if (i == t) {  // raw row t of the right-hand side; scaling by 1/L[t][t] is folded into its readers
#pragma unroll
            for (int jj = 0; jj < 32; jj++) xr[cq + 4 * jj] = R[jj];
        }
        __syncthreads();
        const double rinv = rdiag[t];
        if (tid < nb) {
            const double x = xr[tid] * rinv;
            dinv[(size_t)t * NB + tid] = x;
            dinvt[(size_t)tid * NB + t] = x;
        }
        if (i > t && i < nb) {
            const double lit = S[i * LDB + t] * rinv;
            if (t < 64) {  // uniform: columns beyond t are still zero
#pragma unroll
                for (int jj = 0; jj < 16; jj++) R[jj] = fma(-lit, xr[cq + 4 * jj], R[jj]);
            } else {
#pragma unroll
                for (int jj = 0; jj < 32; jj++) R[jj] = fma(-lit, xr[cq + 4 * jj], R[jj]);
            }
        }
        // xrow is double-buffered: step t+1 writes the other half, and the write of step t+2 is
        // ordered after this step's reads by the barrier of step t+1
    }
}

// ---- panel solve of the blocked Cholesky: X := X * L^-T for rows of a panel, rows independent ----
// SIXTEEN threads per row: thread (row i, cq) keeps the residuals of columns c == cq (mod 16) in a rotating
// window of 8 registers (R[0] = the 16-column group being solved); the solved entry travels to the other
// 15 threads of the row by shuffle; L (the freshly factored diagonal block) sits in shared memory with an
// odd row stride, so the 16 different rows a half-warp reads per step hit 16 different banks.
// No block-level synchronisation after the load: rows never interact.  32 rows per CTA.
constexpr int TR_ROWS = NT / 16;
constexpr int LDT = NB + 1;
__global__ void __launch_bounds__(NT, 1)
trsm_rows_kernel(double* __restrict__ x, size_t ldx, int nrows, const double* __restrict__ l, size_t ldl,
                 int nb, const double* __restrict__ rdiag_g)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDT] : L
    double* rdiag = sm + NB * LDT;   // [NB]
    const int tid = threadIdx.x;
    const int i = tid >> 4, cq = tid & 15;
    const int lane = tid & 31;
    const size_t row = (size_t)blockIdx.x * TR_ROWS + i;
    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDT + c] = (r < nb && c <= r) ? l[(size_t)r * ldl + c] : 0.0;
    }
    if (tid < NB) rdiag[tid] = tid < nb ? rdiag_g[tid] : 0.0;
    __syncthreads();
    const bool live = row < (size_t)nrows;
    double* xr = x + (live ? row : 0) * ldx;
    double R[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int c = cq + 16 * u;
        R[u] = (live && c < nb) ? xr[c] : 0.0;
    }
    const int ngroups = (nb + 15) >> 4;
    for (int jj = 0; jj < ngroups; jj++) {
        const int c0 = cq + 16 * jj;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int t = 16 * jj + q;  // column being solved (uniform)
            // x_t = residual_t / L[t][t], owned by the row's thread cq == q.  Columns t >= nb have
            // rdiag = 0 and zero rows of L: they solve to 0 and change nothing.
            double xt = R[0] * rdiag[t];
            xt = __shfl_sync(0xffffffffu, xt, (lane & 16) | q);
            if (cq == q) R[0] = xt;
            else if (cq > q) R[0] = fma(-xt, S[c0 * LDT + t], R[0]);
#pragma unroll
            for (int u = 1; u < 8; u++) {
                if (jj + u < 8) R[u] = fma(-xt, S[(c0 + 16 * u) * LDT + t], R[u]);  // uniform guard
            }
        }
        if (live && c0 < nb) xr[c0] = R[0];
#pragma unroll
        for (int u = 0; u < 7; u++) R[u] = R[u + 1];
        R[7] = 0.0;
    }
}

}  // namespace

extern "C" int htnk_potf2(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info)
{
    if (nb <= 0) return 0;
    if (nb > NB) return -202;
    const size_t smem = (size_t)(NB * LDB + 7 * NB) * sizeof(double);
    if (cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    potf2_kernel<<<1, NT, smem, st>>>(a, ld, nb, col0, rdiag, info);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_trtri_batched(cudaStream_t st, const double* f, size_t ld, int n, double* dinv, double* dinvt)
{
    if (n <= 0) return 0;
    const size_t smem = (size_t)(NB * LDB + 3 * NB) * sizeof(double);
    if (cudaFuncSetAttribute(trtri_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return -201;
    trtri_batched_kernel<<<(n + NB - 1) / NB, NT, smem, st>>>(f, ld, n, dinv, dinvt);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_trsm_rows(cudaStream_t st, double* x, size_t ldx, int nrows, const double* l, size_t ldl, int nb,
                              const double* rdiag)
{
    if (nrows <= 0 || nb <= 0) return 0;
    if (nb > NB) return -202;
    const size_t smem = (size_t)(NB * LDT + NB) * sizeof(double);
    if (cudaFuncSetAttribute(trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    trsm_rows_kernel<<<(nrows + TR_ROWS - 1) / TR_ROWS, NT, smem, st>>>(x, ldx, nrows, l, ldl, nb, rdiag);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
