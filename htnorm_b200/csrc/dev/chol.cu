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

// potf2: the 128 x 128 block is factorised in 16 column panels of width 8.
//   panel   : threads 0..127 (thread = row) each read the 8 x 8 diagonal sub-block and factor it LOCALLY in
//             registers (redundantly - no communication on the 8-pivot chain), then solve their own row of
//             the panel against it (28 FMAs) and publish the row (final L) to shared memory;
//   update  : the trailing lower triangle gets its rank-8 update on the tensor cores: every 8 x 8 tile is
//             two DMMAs (m8n8k4) on fragments read straight from the published panel; warp w owns tile row
//             b + 1 + w.
// Two barriers per panel, i.e. 32 per block instead of one per column; the pivot chain (128 dependent
// rsqrt + a few FMAs each) is all that remains serial.
__device__ __forceinline__ void dmma_sub(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int LDP = 12;  // panel row stride (doubles): 8 + 4 keeps the DMMA fragment reads conflict-free

__global__ void __launch_bounds__(NT, 1)
potf2_kernel(double* __restrict__ a, size_t ld, int nb, int col0, double* __restrict__ rdiag_out,
             int* __restrict__ info)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]  the block (lower triangle meaningful)
    double* P = sm + NB * LDB;       // [NB][LDP]  current panel: L[:, 8b .. 8b+7]
    double* rdiag = P + NB * LDP;    // [NB]       1 / L[j][j]
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : 0.0;
    }
    for (int idx = tid; idx < NB * LDP; idx += NT) P[idx] = 0.0;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    const int npanels = (nb + 7) >> 3;
    int nfact = nb;
    for (int b = 0; b < npanels; b++) {
        const int j0 = 8 * b;
        if (tid < NB) {
            // ---- local Cholesky of the 8 x 8 diagonal sub-block (identical in every thread) ----
            double l[8][8];
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++) l[r][c] = (c <= r) ? S[(j0 + r) * LDB + j0 + c] : 0.0;
            double ri[8];
            int bad = -1;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double ajj = l[j][j];
                // columns beyond nb (partial last panel) are padding: treat them as the identity
                if (j0 + j >= nb) ajj = 1.0;
                if (ajj <= 0.0 && bad < 0) bad = j;  // NaN passes
                const double rinv = fast_rsqrt(ajj);
                ri[j] = rinv;
                double d = ajj * rinv;
                d = fma(fma(-d, d, ajj), 0.5 * rinv, d);
                l[j][j] = d;
#pragma unroll
                for (int r = 0; r < 8; r++)
                    if (r > j) l[r][j] *= rinv;
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        if (r > j && c > j && c <= r) l[r][c] = fma(-l[r][j], l[c][j], l[r][c]);
            }
            if (bad >= 0) {  // uniform over the 128 threads: every one factored the same block
                if (tid == 0) {
                    atomicCAS(info, 0, col0 + j0 + bad + 1);
                    s_fail = j0 + bad + 1;
                }
            } else {
                // ---- my row of the panel: x = a L11^-T (rows below), or L11's own row (diagonal rows) ----
                const int r = tid;
                if (r >= j0 && r < nb) {
                    double x[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) x[c] = S[r * LDB + j0 + c];
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        double v = x[c];
#pragma unroll
                        for (int t = 0; t < 8; t++)
                            if (t < c) v = fma(-x[t], l[c][t], v);
                        const double vraw = v, ric = ri[c];
                        v *= ric;
                        // a diagonal row's own entry is the pivot (sqrt with one Heron step); nothing above it
                        const double piv = fma(fma(-v, v, vraw), 0.5 * ric, v);
                        const int dr = r - j0;
                        v = (dr == c) ? piv : ((dr < c) ? 0.0 : v);
                        if (j0 + c >= nb) v = 0.0;
                        x[c] = v;
                    }
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        S[r * LDB + j0 + c] = x[c];
                        P[r * LDP + c] = x[c];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (tid == j && j0 + j < nb) rdiag[j0 + j] = ri[j];
            }
        }
        __syncthreads();
        if (s_fail) { nfact = s_fail - 1; break; }
        // ---- rank-8 update of the trailing lower triangle on the tensor cores ----
        {
            const int tr = b + 1 + warp;  // this warp's tile row
            if (tr < npanels) {
                const int g = lane >> 2, t4 = lane & 3;
                const double a0 = -P[(tr * 8 + g) * LDP + t4];
                const double a1 = -P[(tr * 8 + g) * LDP + 4 + t4];
                for (int tc = b + 1; tc <= tr; tc++) {
                    const double b0 = P[(tc * 8 + g) * LDP + t4];
                    const double b1 = P[(tc * 8 + g) * LDP + 4 + t4];
                    double2* cp = reinterpret_cast<double2*>(&S[(tr * 8 + g) * LDB + tc * 8 + 2 * t4]);
                    double2 c = *cp;
                    dmma_sub(c.x, c.y, a0, b0);
                    dmma_sub(c.x, c.y, a1, b1);
                    *cp = c;
                }
            }
        }
        __syncthreads();
    }
    const bool failed = nfact < nb;
    // write L back: lower triangle, zeros above; nothing beyond a failed column is meaningful
    for (int idx = tid; idx < nb * nb; idx += NT) {
        int r = idx / nb, c = idx - r * nb;
        a[(size_t)r * ld + c] = (c <= r && c < ((nfact + 7) & ~7)) ? S[r * LDB + c] : 0.0;
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
    const size_t smem = (size_t)(NB * LDB + NB * LDP + NB) * sizeof(double);
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
