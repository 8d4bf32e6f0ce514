// Diagonal-block kernel of the blocked Cholesky: factor one nb x nb (nb <= 128) block held in shared
// memory and produce inv(L) as well, so that every triangular solve around it (the panel solve of the
// factorisation, the blocked TRSMs of the samplers) becomes a DMMA GEMM with inv(L_jj).
//
// Replaces dpotrf's unblocked base case (reference call sites src/htnorm_distributions.c:77,115 via
// src/htnorm_blas.h:111-112) and dposv / dpotri's triangular kernels (src/htnorm.c:169,210,330).
// Pivot rule: fail on `ajj <= 0` only - NaN passes and propagates with info = 0, as the reference +
// OpenBLAS do (reference tests/test_pyhtnorm.py:46-48, 91-95).
//
// One CTA of 512 threads, the block in REGISTERS (32 per thread, one owner per element): a column costs
// one barrier + one fast rsqrt + the owner-local rank-1 update.  Phase 2: inv(L) by forward elimination
// on the identity, also register-resident; row t of the inverse is final after step t and is published
// by 128 threads (coalesced) while the others update.
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
// REGISTERS for the whole factorisation (the column loop is fully unrolled so every register index is a
// compile-time constant).  Step j: the owners of column j publish it (unscaled) to a double-buffered
// shared-memory column, ONE barrier, then every thread computes the pivot's rsqrt redundantly and applies
// the rank-1 update to its own registers: no shared-memory traffic for the matrix itself.
__global__ void __launch_bounds__(NT, 1)
potf2_inv_kernel(double* __restrict__ a, size_t ld, int nb, int col0, double* __restrict__ dinv,
                 double* __restrict__ dinvt, int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]  staging for the coalesced load / store, L for phase 2
    double* colbuf = sm + NB * LDB;  // [2][NB]    published column (phase 1) / published row (phase 2)
    double* rdiag = colbuf + 2 * NB; // [NB]       1 / L[j][j]
    const int tid = threadIdx.x;
    const int i = tid >> 2, cq = tid & 3;

    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    double R[32];
#pragma unroll
    for (int jj = 0; jj < 32; jj++) R[jj] = S[i * LDB + cq + 4 * jj];

    // ---- phase 1 ----
    int nfact = nb;
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = 4 * jj + q;
            if (j < nb && nfact == nb) {  // uniform
                double* col = colbuf + (j & 1) * NB;
                if (cq == q && i >= j) col[i] = R[jj];
                __syncthreads();
                const double ajj = col[j];
                if (ajj <= 0.0) {  // uniform; NaN passes
                    if (tid == 0) atomicCAS(info, 0, col0 + j + 1);
                    nfact = j;
                } else {
                    const double rinv = fast_rsqrt(ajj);
                    if (i > j) {
                        const double aij = col[i];
                        const double m = aij * (rinv * rinv);
#pragma unroll
                        for (int jj2 = jj; jj2 < 32; jj2++) {
                            const int c = cq + 4 * jj2;
                            if ((jj2 > jj || cq > q) && c <= i) R[jj2] = fma(-m, col[c], R[jj2]);
                        }
                        if (cq == q) R[jj] = aij * rinv;  // L[i][j]
                    } else if (i == j && cq == q) {
                        double d = ajj * rinv;                       // sqrt(ajj), one Heron correction
                        d = fma(fma(-d, d, ajj), 0.5 * rinv, d);
                        R[jj] = d;
                    }
                    if (tid == 0) rdiag[j] = rinv;
                }
            }
        }
    }
    const bool failed = nfact < nb;
    __syncthreads();
    // L -> smem (lower triangle, zeros above; nothing beyond a failed column is meaningful)
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
        const int c = cq + 4 * jj;
        S[i * LDB + c] = (c <= i && c < nfact) ? R[jj] : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < nb * nb; idx += NT) {
        int r = idx / nb, c = idx - r * nb;
        a[(size_t)r * ld + c] = S[r * LDB + c];
    }
    if (dinv == nullptr || failed) return;

    // ---- phase 2: inv(L) by forward elimination on the identity held in registers ----
#pragma unroll
    for (int jj = 0; jj < 32; jj++) R[jj] = (cq + 4 * jj == i) ? 1.0 : 0.0;
    for (int t = 0; t < nb; t++) {
        double* xr = colbuf + (t & 1) * NB;
        if (i == t) {  // raw row t of the right-hand side; scaling by 1/L[t][t] is folded into its readers
#pragma unroll
            for (int jj = 0; jj < 32; jj++) xr[cq + 4 * jj] = R[jj];
        }
        __syncthreads();
        const double rinv = rdiag[t];
        if (tid < nb) {  // row t of the inverse is final: publish (coalesced for dinv)
            const double x = xr[tid] * rinv;
            dinv[(size_t)t * NB + tid] = x;
            if (dinvt) dinvt[(size_t)tid * NB + t] = x;
        }
        if (i > t && i < nb) {
            const double lit = S[i * LDB + t] * rinv;
#pragma unroll
            for (int g4 = 0; g4 < 8; g4++) {
                if (16 * g4 <= t) {  // uniform: columns beyond t are still zero
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int jj = 4 * g4 + u;
                        R[jj] = fma(-lit, xr[cq + 4 * jj], R[jj]);
                    }
                }
            }
        }
        // colbuf is double-buffered: step t+1 writes the other half, and the write of step t+2 is
        // ordered after this step's reads by the barrier of step t+1
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
