// Many small independent problems (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems, each with
// its own mean / cov / G / r and its own generator stream).
//
// One CTA of 128 threads per problem runs the WHOLE of the reference's htn_hyperplane_truncated_mvn
// (src/htnorm.c:85-187 + mvn_rand_cov, src/htnorm_distributions.c:58-88) out of shared memory, so every
// input byte is read from HBM at most once (35 872 B per problem at k = 64, k2 = 4) and only the k-vector
// result goes back: the roofline of this path is HBM, not the 127 kflop of arithmetic per problem.
// Several CTAs are resident per SM, so one problem's 64-step pivot chain overlaps the others' loads and
// updates.
//
// The normals z of every problem are drawn beforehand by normal_fill_kernel (one stream per thread, fully
// parallel - a per-problem stream drawn inside this kernel would leave 127 threads waiting for one).
//
// "Upper world": the reference reads the UPPER triangle of a row-major cov, so the rows are staged
// as they lie in HBM (16-byte cp.async of the part right of the diagonal, nothing is symmetrised or transposed)
// and the factorisation is carried out as cov = U^T U on that triangle; U = L^T holds exactly the numbers of
// the reference's lower factor.  A lower-triangle source (HTN_FLAG_COLMAJOR) is transposed while staging.
//
//   cov * G^T on the tensor cores (DMMA), symmetric reads through the one triangle   dsymm,  src/htnorm.c:154
//   Cholesky in 8-wide panels: local 8 x 8 factor + column solve per thread, DMMA rank-8 trailing update
//                                           dpotrf, distributions.c:77
//   y = mean + U^T z                        dtrmv + daxpy, distributions.c:81-83
//   r - G y, G cov G^T, k2 x k2 solve       dgemm :132, :164, dposv :169
//   x = y + cov G^T alpha                   dgemv :176
//
// Error codes per problem follow the reference: cov not PD -> its pivot index and out untouched;
// G cov G^T not PD -> its pivot index with the unprojected y in out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "launchers.h"

namespace {

constexpr int NT = 128;
constexpr int MAXK2 = 16;

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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
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

// shared-memory geometry, the same on host and device
struct SmallGeom {
    int kpad, ntiles, ld, ldp, k2p;
    size_t doubles;
};
__host__ __device__ inline SmallGeom small_geom(int k, int k2)
{
    SmallGeom g;
    g.kpad = (k + 7) & ~7;
    g.ntiles = g.kpad >> 3;
    // ld = 8 mod 16: the 16-byte accumulator accesses of the trailing update are conflict-free;
    // ldp = 4 mod 16: so are the DMMA operand reads of the 8 x kpad panel
    g.ld = g.kpad + ((8 - g.kpad) & 15);
    if (g.ld == g.kpad) g.ld += 16;
    g.ldp = g.kpad + ((4 - g.kpad) & 15);
    g.k2p = (k2 + 7) & ~7;
    g.doubles = (size_t)g.kpad * g.ld + (size_t)g.k2p * g.ld + (size_t)g.kpad * g.k2p + 8 * (size_t)g.ldp +
                3 * (size_t)g.kpad + MAXK2 * MAXK2 + MAXK2;
    return g;
}

// PROF (HTN_SMALL_PROF=1, development only): thread 0 of every CTA adds the clock64() span of each phase to prof[]
#define PHASE(i)                                               \
    if (PROF && tid == 0) {                                    \
        const long long t_now = clock64();                     \
        atomicAdd((unsigned long long*)&prof[i], (unsigned long long)(t_now - t_prev)); \
        t_prev = t_now;                                        \
    }

template <bool PROF>
__global__ void __launch_bounds__(NT, 4)
ht_small_kernel(size_t nproblems, int k, int k2, const double* __restrict__ mean, const double* __restrict__ cov,
                const double* __restrict__ gm, const double* __restrict__ rv, const double* __restrict__ zs,
                int diag, int lower_src, int vec_ok, double* __restrict__ out, int* __restrict__ info, long long* prof)
{
    extern __shared__ __align__(16) double sm[];
    long long t_prev = PROF ? clock64() : 0;
    const SmallGeom ge = small_geom(k, k2);
    const int kpad = ge.kpad, ntiles = ge.ntiles, ld = ge.ld, ldp = ge.ldp, k2p = ge.k2p;
    double* A = sm;                       // [kpad][ld]   cov (upper triangle valid), then U in place
    double* G = A + (size_t)kpad * ld;    // [k2p][ld]    zero-padded rows / columns
    double* CG = G + (size_t)k2p * ld;    // [kpad][k2p]  row i = (cov * G^T)[i][:]
    double* P = CG + (size_t)kpad * k2p;  // [8][ldp]     current row panel of U
    double* z = P + 8 * (size_t)ldp;      // [kpad]
    double* y = z + kpad;                 // [kpad]
    double* mu = y + kpad;                // [kpad]
    double* S2 = mu + kpad;               // [MAXK2][MAXK2]
    double* gy = S2 + MAXK2 * MAXK2;      // [MAXK2]
    __shared__ int s_fail, s_info2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    // zero everything once: padding rows / columns must stay zero for the tensor-core tiles
    for (size_t idx = tid; idx < ge.doubles; idx += NT) sm[idx] = 0.0;
    __syncthreads();

    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)k * k;
        const double* gp = gm + pb * (size_t)k2 * k;
        // ---- stage inputs: asynchronous copies global -> shared (LDGSTS), everything in flight at once ----
        if (diag) {
            for (int i = tid; i < k; i += NT) cp_async8(&A[i * ld + i], covp + (size_t)i * k + i);
        } else if (lower_src) {  // element (i, j <= i) is the valid one: park it at its mirror position
            for (int i = warp; i < k; i += NT / 32)
                for (int j = lane; j <= i; j += 32) cp_async8(&A[j * ld + i], covp + (size_t)i * k + j);
        } else if (vec_ok) {  // row i from the even column at or left of the diagonal, 16 bytes at a time
            for (int i = warp; i < k; i += NT / 32) {
                const double* src = covp + (size_t)i * k;
                for (int j = (i & ~1) + 2 * lane; j < k; j += 64) cp_async16(&A[i * ld + j], src + j);
            }
        } else {
            for (int i = warp; i < k; i += NT / 32) {
                const double* src = covp + (size_t)i * k;
                for (int j = i + lane; j < k; j += 32) cp_async8(&A[i * ld + j], src + j);
            }
        }
        if (vec_ok) {
            for (int idx = tid; idx < k2 * (k >> 1); idx += NT) {
                const int c = idx / (k >> 1), j = 2 * (idx - c * (k >> 1));
                cp_async16(&G[c * ld + j], gp + (size_t)c * k + j);
            }
            for (int j = 2 * tid; j < k; j += 2 * NT) {
                cp_async16(&mu[j], mean + pb * (size_t)k + j);
                cp_async16(&z[j], zs + pb * (size_t)k + j);
            }
        } else {
            for (int c = warp; c < k2; c += NT / 32)
                for (int j = lane; j < k; j += 32) cp_async8(&G[c * ld + j], gp + (size_t)c * k + j);
            for (int i = tid; i < k; i += NT) {
                cp_async8(&mu[i], mean + pb * (size_t)k + i);
                cp_async8(&z[i], zs + pb * (size_t)k + i);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        if (tid == 0) { s_fail = 0; s_info2 = 0; }
        __syncthreads();
        PHASE(0)
        // ---- cov * G^T (before the factorisation overwrites the triangle) ----
        if (diag) {
            for (int idx = tid; idx < k2 * k; idx += NT) {
                int c = idx / k, i = idx - c * k;
                CG[i * k2p + c] = A[i * ld + i] * G[c * ld + i];
            }
        } else {
            // tensor cores: CG (k x k2) = cov (k x k) * G^T; warp w owns 8-row tiles w, w+4, ...;
            // cov(row, col) lives at A[min][max]: direct right of the diagonal block, transposed left of it
            for (int tr = warp; tr < ntiles; tr += NT / 32) {
                const int row = tr * 8 + g;
                for (int tc = 0; tc < k2p / 8; tc++) {
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // two accumulator chains
                    const double* gr = G + (tc * 8 + g) * ld + t4;
                    for (int kb = 0; kb < tr; kb++) {
                        const double* at = A + (kb * 8 + t4) * ld + row;
                        dmma_acc(c0, c1, at[0], gr[kb * 8]);
                        dmma_acc(d0, d1, at[4 * ld], gr[kb * 8 + 4]);
                    }
                    {
                        const int ca = tr * 8 + t4, cb = ca + 4;
                        const double a0 = ca >= row ? A[row * ld + ca] : A[ca * ld + row];
                        const double a1 = cb >= row ? A[row * ld + cb] : A[cb * ld + row];
                        dmma_acc(c0, c1, a0, gr[tr * 8]);
                        dmma_acc(d0, d1, a1, gr[tr * 8 + 4]);
                    }
                    const double* ar = A + row * ld + t4;
                    for (int kb = tr + 1; kb < ntiles; kb++) {
                        dmma_acc(c0, c1, ar[kb * 8], gr[kb * 8]);
                        dmma_acc(d0, d1, ar[kb * 8 + 4], gr[kb * 8 + 4]);
                    }
                    CG[row * k2p + tc * 8 + 2 * t4] = c0 + d0;
                    CG[row * k2p + tc * 8 + 2 * t4 + 1] = c1 + d1;
                }
            }
        }
        __syncthreads();
        PHASE(2)
        // ---- Cholesky cov = U^T U in 8-row panels: every thread factors the 8 x 8 diagonal block locally
        // (redundantly) and solves its own COLUMN of the panel; the trailing upper triangle gets its rank-8
        // update by DMMA (see chol.cu:potf2 for the lower-triangle twin) ----
        if (!diag) {
            for (int b = 0; b < ntiles; b++) {
                const int j0 = 8 * b;
                if (tid >= j0 && tid < kpad) {
                    double l[8][8];  // l[r][c] = U[j0 + c][j0 + r], c <= r
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++) l[r][c] = (c <= r) ? A[(j0 + c) * ld + j0 + r] : 0.0;
                    double x[8];  // this thread's column of the panel
#pragma unroll
                    for (int c = 0; c < 8; c++) x[c] = A[(j0 + c) * ld + tid];
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
                        const int dr = tid - j0;
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            double v = x[c];
#pragma unroll
                            for (int t = 0; t < 8; t++)
                                if (t < c) v = fma(-x[t], l[c][t], v);
                            const double vraw = v, ric = ri[c];
                            v *= ric;
                            const double piv = fma(fma(-v, v, vraw), 0.5 * ric, v);
                            v = (dr == c) ? piv : ((dr < c) ? 0.0 : v);
                            if (j0 + c >= k || tid >= k) v = 0.0;
                            x[c] = v;
                        }
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            A[(j0 + c) * ld + tid] = x[c];
                            P[c * ldp + tid] = x[c];
                        }
                    }
                }
                __syncthreads();
                PHASE(3)
                if (s_fail) break;
                // trailing update: tile (ti <= tj), both beyond the panel; warp w takes tile rows b+1+w, b+5+w, ...
                // The next tile's accumulator is loaded before the current one is stored (the compiler cannot
                // hoist the load itself: it must assume the store aliases it).
                for (int ti = b + 1 + warp; ti < ntiles; ti += NT / 32) {
                    const double a0 = -P[t4 * ldp + ti * 8 + g];
                    const double a1 = -P[(4 + t4) * ldp + ti * 8 + g];
                    double2* cp = reinterpret_cast<double2*>(&A[(ti * 8 + g) * ld + ti * 8 + 2 * t4]);
                    double2 c = *cp;
                    for (int tj = ti; tj < ntiles; tj++, cp += 4) {
                        const double b0 = P[t4 * ldp + tj * 8 + g];
                        const double b1 = P[(4 + t4) * ldp + tj * 8 + g];
                        double2 cn = c;
                        if (tj + 1 < ntiles) cn = cp[4];
                        double e0 = 0.0, e1 = 0.0;  // second, independent chain
                        dmma_acc(c.x, c.y, a0, b0);
                        dmma_acc(e0, e1, a1, b1);
                        *cp = make_double2(c.x + e0, c.y + e1);
                        c = cn;
                    }
                }
                __syncthreads();
                PHASE(4)
            }
            if (s_fail) {  // reference returns before drawing (src/htnorm_distributions.c:78); out untouched
                if (tid == 0 && info) info[pb] = s_fail;
                __syncthreads();
                continue;
            }
        }
        // ---- y = mean + U^T z   /   mean + sqrt(cov_ii) z ----
        if (diag) {
            for (int i = tid; i < k; i += NT) y[i] = __dadd_rn(mu[i], __dmul_rn(sqrt(A[i * ld + i]), z[i]));
        } else {
            // triangular product on the tensor cores: z is column 0 of a (k x 8) right-hand side; the operand
            // is U transposed, so row i of the tile reads column i of U; below-diagonal entries of the diagonal
            // block are masked
            for (int tr = warp; tr < ntiles; tr += NT / 32) {
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                const int row = tr * 8 + g;
                for (int kb = 0; kb < tr; kb++) {
                    const double* at = A + (kb * 8 + t4) * ld + row;
                    dmma_acc(c0, c1, at[0], g == 0 ? z[kb * 8 + t4] : 0.0);
                    dmma_acc(d0, d1, at[4 * ld], g == 0 ? z[kb * 8 + 4 + t4] : 0.0);
                }
                {
                    const double* at = A + (tr * 8 + t4) * ld + row;
                    const double a0 = t4 <= g ? at[0] : 0.0;
                    const double a1 = 4 + t4 <= g ? at[4 * ld] : 0.0;
                    dmma_acc(c0, c1, a0, g == 0 ? z[tr * 8 + t4] : 0.0);
                    dmma_acc(d0, d1, a1, g == 0 ? z[tr * 8 + 4 + t4] : 0.0);
                }
                if (t4 == 0 && row < k) y[row] = c0 + d0 + mu[row];
            }
        }
        __syncthreads();
        PHASE(5)
        // ---- gy = r - G y  and  S2 = (cov G^T)^T G^T ----
        for (int c = warp; c < k2; c += NT / 32) {
            double acc = 0.0;
            for (int i = lane; i < k; i += 32) acc = fma(G[c * ld + i], y[i], acc);
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) gy[c] = rv[pb * (size_t)k2 + c] - acc;
        }
        if (warp == NT / 32 - 1) {  // S2 = G (cov G^T): 8 x 8 tiles on the tensor cores, the last warp's job
            for (int ta = 0; ta < k2p / 8; ta++)
                for (int tc = 0; tc < k2p / 8; tc++) {
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // two chains: even / odd k4 steps
                    const double* ga = G + (ta * 8 + g) * ld + t4;
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
        PHASE(6)
        // ---- k2 x k2 solve (dposv, src/htnorm.c:169) ----
        int f2 = 0;
        double al[4] = {0.0, 0.0, 0.0, 0.0};
        if (k2 <= 4) {
            // every thread solves the tiny system in registers (broadcast reads): no hand-off, no barrier.
            // Unused rows / columns are an identity block.
            double s[4][4], bv[4], ri[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                bv[i] = i < k2 ? gy[i] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j <= i) s[i][j] = (i < k2) ? S2[i * MAXK2 + j] : (i == j ? 1.0 : 0.0);
            }
            if (k2 == 1) {
                al[0] = bv[0] / s[0][0];  // the reference divides by the scalar (src/htnorm.c:77)
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double ajj = s[j][j];
                    if (ajj <= 0.0 && f2 == 0) f2 = j + 1;
                    const double rinv = fast_rsqrt(ajj);
                    ri[j] = rinv;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (i > j) s[i][j] *= rinv;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            if (i > j && c > j && c <= i) s[i][c] = fma(-s[i][j], s[c][j], s[i][c]);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {  // forward: L w = b
                    double v = bv[i];
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < i) v = fma(-s[i][t], bv[t], v);
                    bv[i] = v * ri[i];
                }
#pragma unroll
                for (int i = 3; i >= 0; i--) {  // backward: L^T alpha = w
                    double v = bv[i];
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t > i) v = fma(-s[t][i], bv[t], v);
                    bv[i] = v * ri[i];
                }
#pragma unroll
                for (int i = 0; i < 4; i++) al[i] = bv[i];
            }
            if (tid == 0 && info) info[pb] = f2;
        } else {
            if (tid == 0) {
                for (int j = 0; j < k2 && !f2; j++) {
                    double ajj = S2[j * MAXK2 + j];
                    for (int t = 0; t < j; t++) ajj -= S2[j * MAXK2 + t] * S2[j * MAXK2 + t];
                    if (ajj <= 0.0) { f2 = j + 1; break; }
                    const double rinv = fast_rsqrt(ajj);
                    S2[j * MAXK2 + j] = rinv;  // the reciprocal of the pivot
                    for (int i = j + 1; i < k2; i++) {
                        double s = S2[i * MAXK2 + j];
                        for (int t = 0; t < j; t++) s -= S2[i * MAXK2 + t] * S2[j * MAXK2 + t];
                        S2[i * MAXK2 + j] = s * rinv;
                    }
                }
                if (!f2) {
                    for (int i = 0; i < k2; i++) {
                        double s = gy[i];
                        for (int t = 0; t < i; t++) s -= S2[i * MAXK2 + t] * gy[t];
                        gy[i] = s * S2[i * MAXK2 + i];
                    }
                    for (int i = k2 - 1; i >= 0; i--) {
                        double s = gy[i];
                        for (int t = i + 1; t < k2; t++) s -= S2[t * MAXK2 + i] * gy[t];
                        gy[i] = s * S2[i * MAXK2 + i];
                    }
                }
                s_info2 = f2;
                if (info) info[pb] = f2;
            }
            __syncthreads();
            f2 = s_info2;
        }
        PHASE(7)
        // ---- x = y + cov G^T alpha ----
        for (int i = tid; i < k; i += NT) {
            double acc = y[i];
            if (!f2) {
                double p = 0.0;
                if (k2 <= 4) {
#pragma unroll
                    for (int c = 0; c < 4; c++) p = fma(CG[i * k2p + c], al[c], p);  // columns >= k2 are zero
                } else {
                    for (int c = 0; c < k2; c++) p = fma(CG[i * k2p + c], gy[c], p);
                }
                acc += p;
            }
            out[pb * (size_t)k + i] = acc;
        }
        __syncthreads();
        PHASE(8)
    }
}
#undef PHASE

}  // namespace

// zs: nproblems x k normals already drawn (reverse-fill order), one stream per problem
extern "C" int htnk_ht_small_batched(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                                     const double* cov, const double* g, const double* r, const double* zs,
                                     int diag, int lower_src, double* out, int* info)
{
    if (nproblems == 0) return 0;
    if (k < 1 || k > 128 || k2 < 1 || k2 > MAXK2) return -202;
    const SmallGeom ge = small_geom(k, k2);
    const size_t smem = ge.doubles * 8;
    // 16-byte copies need an even row length and 16-byte aligned arrays (torch / cudaMalloc give 256)
    const int vec_ok = (k % 2 == 0) && ((((uintptr_t)cov | (uintptr_t)g | (uintptr_t)mean | (uintptr_t)zs) & 15) == 0);
    static const bool want_prof = getenv("HTN_SMALL_PROF") != nullptr;
    auto kern = want_prof ? ht_small_kernel<true> : ht_small_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("HTN_SMALL_CTAS")) per_sm = atoi(e) < per_sm ? (atoi(e) > 0 ? atoi(e) : 1) : per_sm;  // experiments
    size_t grid = (size_t)sms * per_sm;  // persistent: every resident slot loops over problems
    if (grid > nproblems) grid = nproblems;
    long long* prof = nullptr;
    if (want_prof) {
        if (cudaMalloc(&prof, 16 * sizeof(long long)) != cudaSuccess) return -201;
        cudaMemsetAsync(prof, 0, 16 * sizeof(long long), st);
    }
    kern<<<(unsigned)grid, NT, smem, st>>>(nproblems, k, k2, mean, cov, g, r, zs, diag, lower_src, vec_ok, out, info,
                                           prof);
    htnk_count_launch();
    if (want_prof) {  // development aid: synchronous on purpose
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        static const char* names[9] = {"stage", "-", "covG", "panel", "update", "trmv", "gy+S2", "solve", "out"};
        fprintf(stderr, "[ht_small] %d CTAs/SM (%zu B smem), clk per problem:", per_sm, smem);
        for (int i = 0; i < 9; i++) fprintf(stderr, " %s %.0f", names[i], (double)h[i] / (double)nproblems);
        fprintf(stderr, "\n");
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
