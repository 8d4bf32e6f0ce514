// Diagonal-block kernel of the blocked Cholesky: factor one nb x nb (nb <= 128) block held in shared
// memory and produce inv(L) as well, so that every triangular solve around it (the panel solve of the
// factorisation, the blocked TRSMs of the samplers) becomes a DMMA GEMM with inv(L_jj).
//
// Replaces dpotrf's unblocked base case (reference call sites src/htnorm_distributions.c:77,115 via
// src/htnorm_blas.h:111-112) and dposv / dpotri's triangular kernels (src/htnorm.c:169,210,330).
// Pivot rule: fail on `ajj <= 0` only - NaN passes and propagates with info = 0, as the reference +
// OpenBLAS do (reference tests/test_pyhtnorm.py:46-48, 91-95).
//
// One CTA of 512 threads.  Phase 1 (right-looking, block in smem, rows padded to 129 doubles so both
// row and column walks are bank-conflict free): per column one rsqrt-free pivot, a column scale and a
// rank-1 update of the trailing triangle.  Phase 2: inv(L) by forward elimination on the identity with
// the 128 x 128 right-hand side held in REGISTERS (32 doubles per thread); row t of the inverse is
// final after step t and is written straight to global memory (and transposed to dinvt).
#include <cuda_runtime.h>

#include "launchers.h"

namespace {

constexpr int NB = 128;
constexpr int LDB = NB + 1;
constexpr int NT = 512;

__global__ void __launch_bounds__(NT, 1)
potf2_inv_kernel(double* __restrict__ a, size_t ld, int nb, int col0, double* __restrict__ dinv,
                 double* __restrict__ dinvt, int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]
    double* xrow = sm + NB * LDB;    // [2][NB]
    double* dg = xrow + 2 * NB;      // [NB] diagonal of L (S keeps the pre-sqrt pivots until write-back)
    const int tid = threadIdx.x;

    // load the lower triangle (upper part of the block is never read)
    for (int idx = tid; idx < nb * nb; idx += NT) {
        int i = idx / nb, c = idx - i * nb;
        S[i * LDB + c] = (c <= i) ? a[(size_t)i * ld + c] : 0.0;
    }
    __syncthreads();

    // ---- phase 1: Cholesky ----
    const int ty = tid >> 5, tx = tid & 31;
    bool failed = false;
    int nfact = nb;  // columns factorised before a failure
    for (int j = 0; j < nb; j++) {
        const double ajj = S[j * LDB + j];
        if (ajj <= 0.0) {
            if (tid == 0) atomicCAS(info, 0, col0 + j + 1);
            failed = true;
            nfact = j;
            break;
        }
        const double d = sqrt(ajj);
        const double rinv = 1.0 / d;
        for (int i = j + 1 + tid; i < nb; i += NT) S[i * LDB + j] *= rinv;
        if (tid == 0) dg[j] = d;
        __syncthreads();
        for (int i = j + 1 + ty; i < nb; i += NT / 32) {
            const double lij = S[i * LDB + j];
            for (int c = j + 1 + tx; c <= i; c += 32) S[i * LDB + c] -= lij * S[c * LDB + j];
        }
        __syncthreads();  // the next pivot read S[j+1][j+1] and column scale need these updates
    }
    __syncthreads();

    // write L back: lower triangle, zeros above
    for (int idx = tid; idx < nb * nb; idx += NT) {
        int i = idx / nb, c = idx - i * nb;
        a[(size_t)i * ld + c] = (c < i) ? S[i * LDB + c] : (c == i && !(failed && i >= nfact) ? dg[i] : 0.0);
    }
    if (dinv == nullptr || failed) return;

    // ---- phase 2: inv(L) ----
    // thread owns row i = tid / 4 and columns c = cq + 4*j (j = 0..31), cq = tid % 4
    const int i = tid >> 2, cq = tid & 3;
    double R[32];
#pragma unroll
    for (int j = 0; j < 32; j++) R[j] = (cq + 4 * j == i) ? 1.0 : 0.0;
    for (int t = 0; t < nb; t++) {
        double* xr = xrow + (t & 1) * NB;
        if (i == t) {
            const double rinv = 1.0 / dg[t];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int c = cq + 4 * j;
                const double x = (c <= t) ? R[j] * rinv : 0.0;
                xr[c] = x;
                if (c < nb) {
                    dinv[(size_t)t * NB + c] = x;
                    if (dinvt) dinvt[(size_t)c * NB + t] = x;
                }
            }
        }
        __syncthreads();
        if (i > t && i < nb) {
            const double lit = S[i * LDB + t];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int c = cq + 4 * j;
                if (c <= t) R[j] -= lit * xr[c];
            }
        }
        // xrow is double-buffered: the write of step t+1 goes to the other half, and the write of
        // step t+2 is ordered after this read by the barrier of step t+1
    }
}

}  // namespace

extern "C" int htnk_potf2_inv(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* dinv,
                              double* dinvt, int* info)
{
    if (nb <= 0) return 0;
    if (nb > NB) return -202;
    const size_t smem = (size_t)(NB * LDB + 3 * NB) * sizeof(double);
    if (cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return -201;
    potf2_inv_kernel<<<1, NT, smem, st>>>(a, ld, nb, col0, dinv, dinvt, info);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
