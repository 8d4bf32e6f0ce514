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
//   cov * G^T on the tensor cores (DMMA)    dsymm,  src/htnorm.c:154   (before the factor overwrites cov)
//   Cholesky in 8-wide panels: local 8 x 8 factor + row solve per thread, DMMA rank-8 trailing update
//                                           dpotrf, distributions.c:77
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

constexpr int LDP = 12;

__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}

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
    const int kpad = (k + 7) & ~7, ntiles = kpad >> 3;
    const int ld = kpad + 4;              // 4 mod 16: DMMA A-fragment reads of the rows are conflict-free
    const int ldg = kpad + 4;
    const int k2p = (k2 + 7) & ~7;        // G rows / CG row stride, padded to the 8-wide DMMA tile
    double* A = sm;                       // [kpad][ld]   cov, then its factor in the lower triangle
    double* G = A + (size_t)kpad * ld;    // [k2p][ldg] zero-padded rows / columns
    double* CG = G + (size_t)k2p * ldg;   // [kpad][k2p] row i = (cov * G^T)[i][:]
    double* P = CG + (size_t)kpad * k2p;  // [kpad][LDP]  current panel of L
    double* z = P + (size_t)kpad * LDP;   // [kpad]
    double* y = z + kpad;                 // [kpad]
    double* mu = y + kpad;                // [kpad]
    double* S2 = mu + kpad;               // [MAXK2][MAXK2]
    double* gy = S2 + MAXK2 * MAXK2;      // [MAXK2]
    __shared__ int s_fail, s_info2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // zero everything once: padding rows / columns must stay zero for the tensor-core tiles
    for (int idx = tid; idx < kpad * ld + k2p * ldg + kpad * k2p + kpad * LDP + 3 * kpad; idx += NT) sm[idx] = 0.0;
    __syncthreads();

    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)k * k;
        const double* gp = gm + pb * (size_t)k2 * k;
        // ---- stage inputs (coalesced) ----
        // asynchronous 8-byte copies global -> shared (LDGSTS): every element of the problem is in flight at
        // once instead of 32 dependent load -> store round trips per thread
        for (int i = warp; i < k; i += NT / 32) {  // one row per warp: 32 lanes x consecutive doubles
            const double* src = covp + (size_t)i * k;
            for (int j = lane; j < k; j += 32) cp_async8(&A[i * ld + j], src + j);
        }
        for (int c = warp; c < k2; c += NT / 32)
            for (int j = lane; j < k; j += 32) cp_async8(&G[c * ldg + j], gp + (size_t)c * k + j);
        for (int i = tid; i < k; i += NT) {
            cp_async8(&mu[i], mean + pb * (size_t)k + i);
            cp_async8(&z[i], zs + pb * (size_t)k + i);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        if (tid == 0) { s_fail = 0; s_info2 = 0; }
        __syncthreads();
        if (!diag) {  // symmetrise through the triangle the reference reads
            for (int i = warp; i < k; i += NT / 32)
                for (int j = lane; j < k; j += 32) {
                    const bool keep = lower_src ? (j <= i) : (j >= i);
                    if (!keep) A[i * ld + j] = A[j * ld + i];
                }
            __syncthreads();
        }
        // ---- cov * G^T (before the factorisation overwrites the lower triangle) ----
        if (diag) {
            for (int idx = tid; idx < k2 * k; idx += NT) {
                int c = idx / k, i = idx - c * k;
                CG[i * k2p + c] = A[i * ld + i] * G[c * ldg + i];
            }
        } else {
            // tensor cores: CG (k x k2) = cov (k x k) * G^T; warp w owns 8-row tiles w, w+4, ...
            const int g = lane >> 2, t4 = lane & 3;
            for (int tr = warp; tr < ntiles; tr += NT / 32) {
                for (int tc = 0; tc < (k2 + 7) / 8; tc++) {
                    double c0 = 0.0, c1 = 0.0;
                    const double* ar = A + (tr * 8 + g) * ld + t4;
                    const double* gr = G + (tc * 8 + g) * ldg + t4;
                    for (int kk = 0; kk < kpad; kk += 4) dmma_acc(c0, c1, ar[kk], gr[kk]);
                    CG[(tr * 8 + g) * k2p + tc * 8 + 2 * t4] = c0;
                    CG[(tr * 8 + g) * k2p + tc * 8 + 2 * t4 + 1] = c1;
                }
            }
        }
        __syncthreads();
        // ---- Cholesky in 8-wide panels: thread = row factors the 8 x 8 diagonal block locally and solves its
        // own panel row; the trailing lower triangle gets its rank-8 update by DMMA (see chol.cu:potf2) ----
        if (!diag) {
            const int g = lane >> 2, t4 = lane & 3;
            for (int b = 0; b < ntiles; b++) {
                const int j0 = 8 * b;
                if (tid >= j0 && tid < kpad) {
                    double l[8][8];
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++) l[r][c] = (c <= r) ? A[(j0 + r) * ld + j0 + c] : 0.0;
                    double ri[8];
                    int bad = -1;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        double ajj = l[j][j];
                        if (j0 + j >= k) ajj = 1.0;  // padding columns of the last panel
                        if (ajj <= 0.0 && bad < 0) bad = j;  // NaN passes
                        const double rinv = fast_rsqrt(ajj);
                        ri[j] = rinv;
#pragma unroll
                        for (int r = 0; r < 8; r++)
                            if (r > j) l[r][j] *= rinv;
#pragma unroll
                        for (int r = 0; r < 8; r++)
#pragma unroll
                            for (int c = 0; c < 8; c++)
                                if (r > j && c > j && c <= r) l[r][c] = fma(-l[r][j], l[c][j], l[r][c]);
                    }
                    if (bad >= 0) {
                        if (tid == j0) s_fail = j0 + bad + 1;  // every participating thread sees the same block
                    } else {
                        const int r = tid;
                        double x[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) x[c] = A[r * ld + j0 + c];
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            double v = x[c];
#pragma unroll
                            for (int t = 0; t < 8; t++)
                                if (t < c) v = fma(-x[t], l[c][t], v);
                            const double vraw = v, ric = ri[c];
                            v *= ric;
                            const double piv = fma(fma(-v, v, vraw), 0.5 * ric, v);
                            const int dr = r - j0;
                            v = (dr == c) ? piv : ((dr < c) ? 0.0 : v);
                            if (j0 + c >= k || r >= k) v = 0.0;
                            x[c] = v;
                        }
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            A[r * ld + j0 + c] = x[c];
                            P[r * LDP + c] = x[c];
                        }
                    }
                }
                __syncthreads();
                if (s_fail) break;
                for (int tr = b + 1 + warp; tr < ntiles; tr += NT / 32) {
                    const double a0 = -P[(tr * 8 + g) * LDP + t4];
                    const double a1 = -P[(tr * 8 + g) * LDP + 4 + t4];
#pragma unroll 2
                    for (int tc = b + 1; tc <= tr; tc++) {
                        const double b0 = P[(tc * 8 + g) * LDP + t4];
                        const double b1 = P[(tc * 8 + g) * LDP + 4 + t4];
                        double2* cp = reinterpret_cast<double2*>(&A[(tr * 8 + g) * ld + tc * 8 + 2 * t4]);
                        double2 c = *cp;
                        dmma_acc(c.x, c.y, a0, b0);
                        dmma_acc(c.x, c.y, a1, b1);
                        *cp = c;
                    }
                }
                __syncthreads();
            }
            if (s_fail) {  // reference returns before drawing (src/htnorm_distributions.c:78); out untouched
                if (tid == 0 && info) info[pb] = s_fail;
                __syncthreads();
                continue;
            }
        }
        // ---- y = mean + L z   /   mean + sqrt(cov_ii) z ----
        if (diag) {
            for (int i = tid; i < k; i += NT) y[i] = __dadd_rn(mu[i], __dmul_rn(sqrt(A[i * ld + i]), z[i]));
        } else {
            // triangular product on the tensor cores: z is column 0 of an (k x 8) right-hand side; the
            // diagonal 8 x 8 blocks of A are clean lower triangles, blocks above the diagonal are not read
            const int g = lane >> 2, t4 = lane & 3;
            for (int tr = warp; tr < ntiles; tr += NT / 32) {
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                const double* ar = A + (tr * 8 + g) * ld + t4;
                for (int kk = 0; kk < 8 * (tr + 1); kk += 8) {
                    dmma_acc(c0, c1, ar[kk], g == 0 ? z[kk + t4] : 0.0);
                    dmma_acc(d0, d1, ar[kk + 4], g == 0 ? z[kk + 4 + t4] : 0.0);
                }
                if (t4 == 0 && tr * 8 + g < k) y[tr * 8 + g] = c0 + d0 + mu[tr * 8 + g];
            }
        }
        __syncthreads();
        // ---- gy = r - G y  and  S2 = (cov G^T)^T G^T ----
        for (int c = warp; c < k2; c += NT / 32) {
            double acc = 0.0;
            for (int i = lane; i < k; i += 32) acc = fma(G[c * ldg + i], y[i], acc);
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) gy[c] = rv[pb * (size_t)k2 + c] - acc;
        }
        if (warp == NT / 32 - 1) {  // S2 = G (cov G^T): 8 x 8 tiles on the tensor cores, the last warp's job
            const int g = lane >> 2, t4 = lane & 3;
            for (int ta = 0; ta < k2p / 8; ta++)
                for (int tc = 0; tc < k2p / 8; tc++) {
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // two chains: even / odd k4 steps
                    const double* ga = G + (ta * 8 + g) * ldg + t4;
                    const double* cb = CG + t4 * k2p + tc * 8 + g;
                    for (int kk = 0; kk < kpad; kk += 8) {
                        dmma_acc(c0, c1, ga[kk], cb[kk * k2p]);
                        dmma_acc(d0, d1, ga[kk + 4], cb[(kk + 4) * k2p]);
                    }
                    S2[(ta * 8 + g) * MAXK2 + tc * 8 + 2 * t4] = c0 + d0;
                    S2[(ta * 8 + g) * MAXK2 + tc * 8 + 2 * t4 + 1] = c1 + d1;
                }
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
                for (int c = 0; c < k2; c++) p = fma(CG[i * k2p + c], gy[c], p);
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
    const size_t kpad = (size_t)((k + 7) & ~7);
    const size_t k2p = (size_t)((k2 + 7) & ~7);
    const size_t smem = (kpad * (kpad + 4) + k2p * (kpad + 4) + kpad * k2p + kpad * LDP + 3 * kpad + MAXK2 * MAXK2 + MAXK2) * 8;
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
