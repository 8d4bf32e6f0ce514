// Many small independent problems (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems, each with
// its own mean / cov / G / r and its own generator stream).
//
// One CTA per problem runs the WHOLE of the reference's htn_hyperplane_truncated_mvn
// (src/htnorm.c:85-187 + mvn_rand_cov, src/htnorm_distributions.c:58-88) out of shared memory, so every
// input byte is read from HBM exactly once (35 872 B per problem at k = 64, k2 = 4) and only the k-vector
// result goes back: the kernel is bound by HBM, not by the 127 kflop of arithmetic per problem.
//
//   warp 0 lane 0 : the problem's ziggurat stream (k sequential draws - latency hidden behind ...)
//   all other warps: ... the coalesced load of cov (symmetrised through the upper triangle),
//                   G, r, mean and cov * G^T  (dsymm, src/htnorm.c:154)
//   then together : Cholesky in smem (dpotrf, distributions.c:77), y = mean + L z (dtrmv/daxpy, :81-83),
//                   r - G y (:132), G cov G^T (:164), k2 x k2 solve (dposv, :169), x = y + cov G^T alpha (:176)
//
// Error codes per problem follow the reference: cov not PD -> its pivot index, out untouched and the
// stream not advanced; G cov G^T not PD -> its pivot index with the unprojected y in out.
#include <cuda_runtime.h>
#include <stdint.h>

#include "launchers.h"
#include "rng.cuh"

namespace {

using namespace htn;

constexpr int NT = 128;
constexpr int MAXK2 = 16;

template <int GEN>
__global__ void __launch_bounds__(NT)
ht_small_kernel(const uint64_t* __restrict__ seeds, uint64_t seed0, size_t nproblems, int k, int k2,
                const double* __restrict__ mean, const double* __restrict__ cov, const double* __restrict__ gm,
                const double* __restrict__ rv, int diag, int lower_src, double* __restrict__ out,
                int* __restrict__ info)
{
    extern __shared__ double sm[];
    const int ld = k + 1;
    double* A = sm;                       // [k][ld]   cov, then its factor in the lower triangle
    double* G = A + (size_t)k * ld;       // [k2][k]
    double* CG = G + (size_t)k2 * k;      // [k2][k]   row c = cov * g[c]^T
    double* z = CG + (size_t)k2 * k;      // [k]
    double* y = z + k;                    // [k]
    double* mu = y + k;                   // [k]
    double* S2 = mu + k;                  // [MAXK2][MAXK2]
    double* gy = S2 + MAXK2 * MAXK2;      // [MAXK2]
    __shared__ uint64_t s_ki[256];
    __shared__ double s_wi[256];
    __shared__ int s_info;

    const int tid = threadIdx.x;
    zig_stage_tables(s_ki, s_wi);
    if (tid == 0) s_info = 0;
    __syncthreads();

    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)k * k;
        const double* gp = gm + pb * (size_t)k2 * k;
        // ---- stage inputs ----
        for (int idx = tid; idx < k * k; idx += NT) {
            int i = idx / k, j = idx - i * k;
            A[i * ld + j] = covp[idx];
        }
        for (int idx = tid; idx < k2 * k; idx += NT) G[idx] = gp[idx];
        for (int i = tid; i < k; i += NT) mu[i] = mean[pb * (size_t)k + i];
        __syncthreads();
        // symmetrise through the triangle the reference reads
        for (int idx = tid; idx < k * k; idx += NT) {
            int i = idx / k, j = idx - i * k;
            bool keep = lower_src ? (j <= i) : (j >= i);
            if (!keep && !diag) A[i * ld + j] = A[j * ld + i];
        }
        __syncthreads();
        // ---- cov * G^T (before the factorisation overwrites the lower triangle) ----
        for (int idx = tid; idx < k2 * k; idx += NT) {
            int c = idx / k, i = idx - c * k;
            double acc = 0.0;
            if (diag) {
                acc = A[i * ld + i] * G[c * k + i];
            } else {
                for (int j = 0; j < k; j++) acc += A[i * ld + j] * G[c * k + j];
            }
            CG[idx] = acc;
        }
        // ---- Cholesky (dense) ----
        int fail = 0;
        if (!diag) {
            __syncthreads();
            for (int j = 0; j < k; j++) {
                const double ajj = A[j * ld + j];
                if (ajj <= 0.0) { fail = j + 1; break; }
                const double d = sqrt(ajj);
                const double rinv = 1.0 / d;
                __syncthreads();
                for (int i = j + 1 + tid; i < k; i += NT) A[i * ld + j] *= rinv;
                if (tid == 0) A[j * ld + j] = d;
                __syncthreads();
                // trailing update of the lower triangle: thread owns rows i = j+1+tid, ...
                for (int i = j + 1 + tid; i < k; i += NT) {
                    const double lij = A[i * ld + j];
                    for (int c = j + 1; c <= i; c++) A[i * ld + c] -= lij * A[c * ld + j];
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (fail) {  // reference returns before drawing (src/htnorm_distributions.c:78)
            if (tid == 0 && info) info[pb] = fail;
            continue;
        }
        // ---- the stream: z in reverse fill order ----
        if (tid == 0) {
            GenState g;
            gen_seed<GEN>(g, seeds ? seeds[pb] : seed0 + (uint64_t)pb);
            uint32_t ncalls = 0;
            for (int j = 0; j < k; j++) {
                double v;
                for (;;) {
                    ZigDraw d;
                    if (zig_fast(gen_next<GEN>(g), s_ki, s_wi, d)) { v = d.x; break; }
                    if (zig_slow<GEN>(g, d, v, ncalls)) break;
                }
                z[k - 1 - j] = v;
            }
        }
        __syncthreads();
        // ---- y = mean + L z   /   mean + sqrt(cov_ii) z ----
        for (int i = tid; i < k; i += NT) {
            double acc;
            if (diag) {
                acc = __dadd_rn(mu[i], __dmul_rn(sqrt(A[i * ld + i]), z[i]));
            } else {
                acc = 0.0;
                for (int t = 0; t <= i; t++) acc += A[i * ld + t] * z[t];
                acc += mu[i];
            }
            y[i] = acc;
        }
        __syncthreads();
        // ---- gy = r - G y  and  S2 = (cov G^T)^T G^T ----
        {
            const int lane = tid & 31, warp = tid >> 5;
            for (int c = warp; c < k2; c += NT / 32) {
                double acc = 0.0;
                for (int i = lane; i < k; i += 32) acc += G[c * k + i] * y[i];
                for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) gy[c] = rv[pb * (size_t)k2 + c] - acc;
            }
            for (int idx = tid; idx < k2 * k2; idx += NT) {
                int a = idx / k2, c = idx - a * k2;
                double acc = 0.0;
                for (int i = 0; i < k; i++) acc += CG[a * k + i] * G[c * k + i];
                S2[a * MAXK2 + c] = acc;
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
            s_info = f2;
            if (info) info[pb] = f2;
        }
        __syncthreads();
        const int f2 = s_info;
        // ---- x = y + cov G^T alpha ----
        for (int i = tid; i < k; i += NT) {
            double acc = y[i];
            if (!f2) {
                double p = 0.0;
                for (int c = 0; c < k2; c++) p += CG[c * k + i] * gy[c];
                acc += p;
            }
            out[pb * (size_t)k + i] = acc;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int htnk_ht_small_batched(cudaStream_t st, const htnk_streams_t* s, size_t nproblems, int k, int k2,
                                     const double* mean, const double* cov, const double* g, const double* r,
                                     int diag, int lower_src, double* out, int* info)
{
    if (nproblems == 0) return 0;
    if (k < 1 || k > 128 || k2 < 1 || k2 > MAXK2) return -202;
    const size_t smem = ((size_t)k * (k + 1) + 2 * (size_t)k2 * k + 3 * (size_t)k + MAXK2 * MAXK2 + MAXK2) * 8;
    auto kern = s->gen == 0 ? ht_small_kernel<0> : ht_small_kernel<1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)sms * per_sm * 4;  // persistent-ish: a few CTAs per slot, grid-stride over problems
    if (grid > nproblems) grid = nproblems;
    kern<<<(unsigned)grid, NT, smem, st>>>(s->seeds, s->seed0, nproblems, k, k2, mean, cov, g, r, diag, lower_src,
                                          out, info);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
