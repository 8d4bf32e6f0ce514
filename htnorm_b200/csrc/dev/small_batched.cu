// Many small independent problems (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems, each with
// its own mean / cov / G / r and its own generator stream).
//
// One CTA of 128 threads per problem runs the WHOLE of the reference's htn_hyperplane_truncated_mvn
// (src/htnorm.c:85-187 + mvn_rand_cov, src/htnorm_distributions.c:58-88) out of shared memory, so every
// input byte is read from HBM exactly once (35 872 B per problem at k = 64, k2 = 4) and only the k-vector
// result goes back: the roofline of this path is HBM, not the 127 kflop of arithmetic per problem.
// Several CTAs are resident per SM (about 39 KB of shared memory each at k = 64), so one problem's
// 64-step pivot chain overlaps the others' loads and updates.
//
// The normals z of every problem are drawn beforehand by normal_fill_kernel (one stream per thread, fully
// parallel - a per-problem stream drawn inside this kernel would leave 127 threads waiting for one).
//
//   load cov (coalesced) and symmetrise it through the triangle the reference reads
//   cov * G^T                               dsymm,  src/htnorm.c:154   (before the factor overwrites cov)
//   Cholesky in smem, one owner thread per element, one barrier per column   dpotrf, distributions.c:77
//   y = mean + L z                          dtrmv + daxpy, distributions.c:81-83
//   r - G y, G cov G^T, k2 x k2 solve       dgemm :132, :164, dposv :169
//   x = y + cov G^T alpha                   dgemv :176
//
// Error codes per problem follow the reference: cov not PD -> its pivot index and out untouched;
// G cov G^T not PD -> its pivot index with the unprojected y in out.
#include <cuda_runtime.h>
#include <stdint.h>

#include "launchers.h"

namespace {

constexpr int NT = 128;
constexpr int MAXK2 = 16;

__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double h = 0.5 * y;
        const double e = fma(-x * y, h, 0.5);
        y = fma(y, e, y);
    }
    return y;
}

__global__ void __launch_bounds__(NT)
ht_small_kernel(size_t nproblems, int k, int k2, const double* __restrict__ mean, const double* __restrict__ cov,
                const double* __restrict__ gm, const double* __restrict__ rv, const double* __restrict__ zs,
                int diag, int lower_src, double* __restrict__ out, int* __restrict__ info)
{
    extern __shared__ double sm[];
    const int ld = k + 2;                 // even pad: row-owner accesses (i, cq + 2u) are conflict-free
    double* A = sm;                       // [k][ld]   cov, then its factor in the lower triangle
    double* G = A + (size_t)k * ld;       // [k2][k]
    double* CG = G + (size_t)k2 * k;      // [k2][k]   row c = cov * g[c]^T
    double* z = CG + (size_t)k2 * k;      // [k]
    double* y = z + k;                    // [k]
    double* mu = y + k;                   // [k]
    double* S2 = mu + k;                  // [MAXK2][MAXK2]
    double* gy = S2 + MAXK2 * MAXK2;      // [MAXK2]
    __shared__ int s_fail, s_info2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpr = NT / k >= 2 ? 2 : 1;  // threads per row in the factorisation (2 for k <= 64)
    const int row = tid / tpr, cq = tid % tpr;

    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)k * k;
        const double* gp = gm + pb * (size_t)k2 * k;
        // ---- stage inputs (coalesced) ----
        for (int idx = tid; idx < k * k; idx += NT) {
            int i = idx / k, j = idx - i * k;
            A[i * ld + j] = covp[idx];
        }
        for (int idx = tid; idx < k2 * k; idx += NT) G[idx] = gp[idx];
        for (int i = tid; i < k; i += NT) {
            mu[i] = mean[pb * (size_t)k + i];
            z[i] = zs[pb * (size_t)k + i];
        }
        if (tid == 0) { s_fail = 0; s_info2 = 0; }
        __syncthreads();
        if (!diag) {  // symmetrise through the triangle the reference reads
            for (int idx = tid; idx < k * k; idx += NT) {
                int i = idx / k, j = idx - i * k;
                bool keep = lower_src ? (j <= i) : (j >= i);
                if (!keep) A[i * ld + j] = A[j * ld + i];
            }
            __syncthreads();
        }
        // ---- cov * G^T (before the factorisation overwrites the lower triangle) ----
        for (int idx = tid; idx < k2 * k; idx += NT) {
            int c = idx / k, i = idx - c * k;
            double acc = 0.0;
            if (diag) {
                acc = A[i * ld + i] * G[c * k + i];
            } else {
                const double* ai = A + i * ld;
                const double* gc = G + c * k;
#pragma unroll 8
                for (int j = 0; j < k; j++) acc = fma(ai[j], gc[j], acc);
            }
            CG[idx] = acc;
        }
        __syncthreads();
        // ---- Cholesky: thread (row, cq) owns A[row][c], c == cq (mod tpr); one barrier per column ----
        if (!diag) {
            for (int j = 0; j < k; j++) {
                const double ajj = A[j * ld + j];
                if (ajj <= 0.0) {  // uniform; NaN passes
                    if (tid == 0) s_fail = j + 1;
                    break;
                }
                const double rinv = fast_rsqrt(ajj);
                if (row > j && row < k) {
                    const double aij = A[row * ld + j];
                    const double m = aij * (rinv * rinv);
                    int c = cq + tpr * ((j + 1 - cq + tpr - 1) / tpr);  // first owned column > j
                    double* ar = A + row * ld;
#pragma unroll 4
                    for (; c <= row; c += tpr) ar[c] = fma(-m, A[c * ld + j], ar[c]);
                    if (cq == (j % tpr)) A[j * ld + row] = aij * rinv;  // L[row][j] parked above the diagonal
                } else if (row == j && cq == (j % tpr)) {
                    double d = ajj * rinv;
                    d = fma(fma(-d, d, ajj), 0.5 * rinv, d);
                    y[j] = d;  // diagonal of L (y is free until the triangular product)
                }
                __syncthreads();
            }
            __syncthreads();
            if (s_fail) {  // reference returns before drawing (src/htnorm_distributions.c:78); out untouched
                if (tid == 0 && info) info[pb] = s_fail;
                __syncthreads();
                continue;
            }
            // unpark: L[i][j] = A[j][i] for j < i, diagonal from y
            for (int idx = tid; idx < k * k; idx += NT) {
                int i = idx / k, j = idx - i * k;
                if (j < i) A[i * ld + j] = A[j * ld + i];
            }
            __syncthreads();
            if (tid < k) A[tid * ld + tid] = y[tid];
            __syncthreads();
        }
        // ---- y = mean + L z   /   mean + sqrt(cov_ii) z ----
        for (int i = tid; i < k; i += NT) {
            double acc;
            if (diag) {
                acc = __dadd_rn(mu[i], __dmul_rn(sqrt(A[i * ld + i]), z[i]));
            } else {
                acc = 0.0;
                const double* ai = A + i * ld;
                for (int t = 0; t <= i; t++) acc = fma(ai[t], z[t], acc);
                acc += mu[i];
            }
            y[i] = acc;
        }
        __syncthreads();
        // ---- gy = r - G y  and  S2 = (cov G^T)^T G^T ----
        for (int c = warp; c < k2; c += NT / 32) {
            double acc = 0.0;
            for (int i = lane; i < k; i += 32) acc = fma(G[c * k + i], y[i], acc);
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) gy[c] = rv[pb * (size_t)k2 + c] - acc;
        }
        for (int idx = tid; idx < k2 * k2; idx += NT) {
            int a = idx / k2, c = idx - a * k2;
            double acc = 0.0;
            for (int i = 0; i < k; i++) acc = fma(CG[a * k + i], G[c * k + i], acc);
            S2[a * MAXK2 + c] = acc;
        }
        __syncthreads();
        // ---- k2 x k2 solve by one thread (tiny) ----
        if (tid == 0) {
            int f2 = 0;
            if (k2 == 1) {
                gy[0] = gy[0] / S2[0];
            } else {
                for (int j = 0; j < k2 && !f2; j++) {
                    double ajj = S2[j * MAXK2 + j];
                    for (int t = 0; t < j; t++) ajj -= S2[j * MAXK2 + t] * S2[j * MAXK2 + t];
                    if (ajj <= 0.0) { f2 = j + 1; break; }
                    ajj = sqrt(ajj);
                    S2[j * MAXK2 + j] = ajj;
                    for (int i = j + 1; i < k2; i++) {
                        double s = S2[i * MAXK2 + j];
                        for (int t = 0; t < j; t++) s -= S2[i * MAXK2 + t] * S2[j * MAXK2 + t];
                        S2[i * MAXK2 + j] = s / ajj;
                    }
                }
                if (!f2) {
                    for (int i = 0; i < k2; i++) {
                        double s = gy[i];
                        for (int t = 0; t < i; t++) s -= S2[i * MAXK2 + t] * gy[t];
                        gy[i] = s / S2[i * MAXK2 + i];
                    }
                    for (int i = k2 - 1; i >= 0; i--) {
                        double s = gy[i];
                        for (int t = i + 1; t < k2; t++) s -= S2[t * MAXK2 + i] * gy[t];
                        gy[i] = s / S2[i * MAXK2 + i];
                    }
                }
            }
            s_info2 = f2;
            if (info) info[pb] = f2;
        }
        __syncthreads();
        const int f2 = s_info2;
        // ---- x = y + cov G^T alpha ----
        for (int i = tid; i < k; i += NT) {
            double acc = y[i];
            if (!f2) {
                double p = 0.0;
                for (int c = 0; c < k2; c++) p = fma(CG[c * k + i], gy[c], p);
                acc += p;
            }
            out[pb * (size_t)k + i] = acc;
        }
        __syncthreads();
    }
}

}  // namespace

// zs: nproblems x k normals already drawn (reverse-fill order), one stream per problem
extern "C" int htnk_ht_small_batched(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                                     const double* cov, const double* g, const double* r, const double* zs,
                                     int diag, int lower_src, double* out, int* info)
{
    if (nproblems == 0) return 0;
    if (k < 1 || k > 128 || k2 < 1 || k2 > MAXK2) return -202;
    const size_t smem = ((size_t)k * (k + 2) + 2 * (size_t)k2 * k + 3 * (size_t)k + MAXK2 * MAXK2 + MAXK2) * 8;
    if (cudaFuncSetAttribute(ht_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ht_small_kernel, NT, smem);
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)sms * per_sm;  // persistent: every resident slot loops over problems
    if (grid > nproblems) grid = nproblems;
    ht_small_kernel<<<(unsigned)grid, NT, smem, st>>>(nproblems, k, k2, mean, cov, g, r, zs, diag, lower_src, out,
                                                     info);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
