// Layout and elementwise helpers around the GEMM / Cholesky kernels (all HBM-bound, coalesced along
// the contiguous dimension; transposing ones go through a padded 32 x 33 shared-memory tile).
// Each cites the reference loop or BLAS-1/2 call it stands in for.
#include <cuda_runtime.h>
#include <stdint.h>

#include "launchers.h"

namespace {

inline int status()
{
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

// ---- symmetric expansion: the reference's uplo='L' Fortran calls read a row-major array through its
// UPPER triangle (-DHTNORM_COLMAJOR: the lower one).  full = symmetric copy, low = lower copy.
__global__ void __launch_bounds__(256)
sym_expand_kernel(const double* __restrict__ src, size_t lds, int n, int lower_src, double* __restrict__ full,
                  size_t ldf, double* __restrict__ low, size_t ldl, size_t ssrc, size_t slow)
{
    src += (size_t)blockIdx.z * ssrc;  // problem blockIdx.z of a batched launch (strides in doubles, 0 for one problem)
    if (low) low += (size_t)blockIdx.z * slow;
    __shared__ double d_t[32][33];  // direct tile   src[bi-rows][bj-cols]
    __shared__ double t_t[32][33];  // mirrored tile src[bj-rows][bi-cols]
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    // an off-diagonal tile lies wholly on one side of the diagonal: it needs the direct tile OR the mirrored one, and a
    // tile above the diagonal of a lower-only copy needs neither (it is all zeros): half the reads, or none
    const bool use_direct = bi == bj || (lower_src ? bj < bi : bj > bi);
    const bool use_mirror = bi == bj || !use_direct;
    const bool zeros_only = full == nullptr && bj > bi;
    for (int r = ty; r < 32; r += 8) {
        int i = bi * 32 + r, j = bj * 32 + tx;
        d_t[r][tx] = (use_direct && !zeros_only && i < n && j < n) ? src[(size_t)i * lds + j] : 0.0;
        int i2 = bj * 32 + r, j2 = bi * 32 + tx;
        t_t[r][tx] = (use_mirror && !zeros_only && i2 < n && j2 < n) ? src[(size_t)i2 * lds + j2] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int i = bi * 32 + r, j = bj * 32 + tx;
        if (i >= n || j >= n) continue;
        bool direct = lower_src ? (j <= i) : (j >= i);
        double v = direct ? d_t[r][tx] : t_t[tx][r];
        if (full) full[(size_t)i * ldf + j] = v;
        if (low) low[(size_t)i * ldl + j] = (j <= i) ? v : 0.0;
    }
}

// src_lower != 0: src is lower triangular (zeros above its diagonal, e.g. a Cholesky factor): the tiles above the diagonal
// are not read, their transposes are written as zeros
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ src, size_t lds, int rows, int cols, double* __restrict__ dst,
                 size_t ldd, int src_lower)
{
    __shared__ double t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const bool skip = src_lower && c0 > r0 + 31;
    for (int r = ty; r < 32; r += 8) {
        int i = r0 + r, j = c0 + tx;
        t[r][tx] = (!skip && i < rows && j < cols) ? src[(size_t)i * lds + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int j = c0 + r, i = r0 + tx;  // dst[j][i] = src[i][j]
        if (j < cols && i < rows) dst[(size_t)j * ldd + i] = t[tx][r];
    }
}

// op: 0 copy (+bias[c]), 1 divide by d[c], 2 multiply by d[c], 3 dst = d[c] - dst
__global__ void __launch_bounds__(256)
rowwise_kernel(const double* __restrict__ src, size_t lds, int rows, int cols, const double* __restrict__ d,
               double* __restrict__ dst, size_t ldd, int op)
{
    const size_t total = (size_t)rows * cols;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / cols;
        int c = (int)(idx - r * cols);
        double v;
        switch (op) {
            case 0: v = src[r * lds + c]; if (d) v += d[c]; break;
            case 1: v = src[r * lds + c] / d[c]; break;
            case 2: v = src[r * lds + c] * d[c]; break;
            default: v = d[c] - dst[r * ldd + c]; break;
        }
        dst[r * ldd + c] = v;
    }
}

__global__ void diag_extract_kernel(const double* __restrict__ a, size_t lda, int n, int op, double* __restrict__ d)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = a[(size_t)i * lda + i];
    if (op == 1) v = sqrt(v);
    else if (op == 2) v = 1.0 / v;
    d[i] = v;
}

__global__ void diag_set_kernel(double* __restrict__ m, size_t ld, int n, const double* __restrict__ v, double cval)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    m[(size_t)i * ld + i] = v ? v[i] : cval;
}

__global__ void fill_i32_kernel(int* p, size_t n, int v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// y[c] = r[c] - g[c] . mean : one block per row c (ddot, reference src/htnorm.c:65 folded with the mean)
__global__ void __launch_bounds__(256)
r_minus_gmean_kernel(const double* __restrict__ g, size_t ldg, int k, const double* __restrict__ mean,
                     const double* __restrict__ r, double* __restrict__ y)
{
    __shared__ double red[8];
    const int c = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) acc += g[(size_t)c * ldg + i] * mean[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        y[c] = r[c] - s;
    }
}

// per-sample k2 x k2 solve, k2 <= 16 (dposv's solve phase, reference src/htnorm.c:169; k2 == 1: :77)
constexpr int MAXK2 = 16;
__global__ void __launch_bounds__(128)
ht_alpha_kernel(size_t nsamples, int k2, const double* d, size_t ldd, const double* __restrict__ c0,
                const double* __restrict__ ls, size_t ldls, const int* __restrict__ info_s,
                double* z, size_t ldz, size_t koff)
{
    __shared__ double s_l[MAXK2 * MAXK2];
    __shared__ double s_c0[MAXK2];
    for (int i = threadIdx.x; i < k2 * k2; i += blockDim.x) s_l[i] = ls[(size_t)(i / k2) * ldls + (i % k2)];
    for (int i = threadIdx.x; i < k2; i += blockDim.x) s_c0[i] = c0[i];
    __syncthreads();
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nsamples) return;
    const bool skip = info_s && *info_s != 0;
    double v[MAXK2];
#pragma unroll
    for (int c = 0; c < MAXK2; c++) v[c] = (c < k2 && !skip) ? s_c0[c] - d[b * ldd + c] : 0.0;
    if (!skip) {
        if (k2 == 1) {
            v[0] = v[0] / s_l[0];
        } else {
#pragma unroll
            for (int i = 0; i < MAXK2; i++) {
                if (i < k2) {
                    double s = v[i];
#pragma unroll
                    for (int t = 0; t < MAXK2; t++)
                        if (t < i) s -= s_l[i * k2 + t] * v[t];
                    v[i] = s / s_l[i * k2 + i];
                }
            }
#pragma unroll
            for (int i = MAXK2 - 1; i >= 0; i--) {
                if (i < k2) {
                    double s = v[i];
#pragma unroll
                    for (int t = 0; t < MAXK2; t++)
                        if (t > i && t < k2) s -= s_l[t * k2 + i] * v[t];
                    v[i] = s / s_l[i * k2 + i];
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < MAXK2; c++)
        if (c < k2) z[b * ldz + koff + c] = v[c];
}

// Fused coefficient pass for k2 <= 4 (replaces the skinny GEMM D = Z W^T + ht_alpha_kernel): one warp per sample
// streams its z row ONCE from HBM (16-byte loads, W comes from L1/L2), reduces the k2 dot products by shuffles,
// and lane 0 solves the k2 x k2 system (factor ls of G cov G^T) and stores alpha into the row's coefficient
// columns.  HBM-bound: 8 (k + k2) B per sample.  (G y of src/htnorm.c:132 without the mean part; dposv :169.)
template <int K2T>
__global__ void __launch_bounds__(256)
ht_coeff_kernel(size_t nsamples, int k, int k2, double* __restrict__ z, size_t ldz, size_t koff,
                const double* __restrict__ w, size_t ldw, const double* __restrict__ c0,
                const double* __restrict__ ls, size_t ldls, const int* __restrict__ info_s)
{
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nsamples) return;
    const double* zr = z + b * ldz;
    double acc[K2T];
#pragma unroll
    for (int c = 0; c < K2T; c++) acc[c] = 0.0;
    const int kev = k & ~1;
#pragma unroll 4
    for (int j = 2 * lane; j < kev; j += 64) {
        const double2 zv = *reinterpret_cast<const double2*>(zr + j);
#pragma unroll
        for (int c = 0; c < K2T; c++) {
            if (c < k2) {
                const double2 wv = __ldg(reinterpret_cast<const double2*>(w + (size_t)c * ldw + j));
                acc[c] = fma(zv.x, wv.x, acc[c]);
                acc[c] = fma(zv.y, wv.y, acc[c]);
            }
        }
    }
    if ((k & 1) && lane == 0) {
        const double zv = zr[k - 1];
#pragma unroll
        for (int c = 0; c < K2T; c++)
            if (c < k2) acc[c] = fma(zv, w[(size_t)c * ldw + k - 1], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < K2T; c++)
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane == 0) {
        const bool skip = info_s && *info_s != 0;
        double v[K2T];
#pragma unroll
        for (int c = 0; c < K2T; c++) v[c] = (c < k2 && !skip) ? c0[c] - acc[c] : 0.0;
        if (!skip) {
            if (k2 == 1) {
                v[0] = v[0] / ls[0];
            } else {
#pragma unroll
                for (int i = 0; i < K2T; i++) {
                    if (i < k2) {
                        double sv = v[i];
#pragma unroll
                        for (int t = 0; t < K2T; t++)
                            if (t < i) sv -= ls[(size_t)i * ldls + t] * v[t];
                        v[i] = sv / ls[(size_t)i * ldls + i];
                    }
                }
#pragma unroll
                for (int i = K2T - 1; i >= 0; i--) {
                    if (i < k2) {
                        double sv = v[i];
#pragma unroll
                        for (int t = 0; t < K2T; t++)
                            if (t > i && t < k2) sv -= ls[(size_t)t * ldls + i] * v[t];
                        v[i] = sv / ls[(size_t)i * ldls + i];
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < K2T; c++)
            if (c < k2) z[b * ldz + koff + c] = v[c];
    }
}

// out[b][i] += sum_c cgt[i][c] * alpha[b][c]   (dgemv, reference src/htnorm.c:176, diagonal-cov path)
__global__ void __launch_bounds__(256)
ht_diag_project_kernel(size_t nsamples, int k, int k2, const double* __restrict__ cgt, size_t ldc,
                       const double* __restrict__ alpha, size_t lda, double* __restrict__ out, size_t ldo)
{
    const size_t total = nsamples * (size_t)k;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t b = idx / k;
        int i = (int)(idx - b * k);
        double acc = 0.0;
        for (int c = 0; c < k2; c++) acc += cgt[(size_t)i * ldc + c] * alpha[b * lda + c];
        out[b * ldo + i] += acc;
    }
}

// dcol[b][c] = g[c] . y[b]: one warp per sample, up to 16 rows of g per pass
__global__ void __launch_bounds__(256)
rows_dot_kernel(size_t nsamples, int k, int k2, const double* __restrict__ g, size_t ldg,
                const double* __restrict__ y, size_t ldy, double* __restrict__ dcol, size_t ldd)
{
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nsamples) return;
    for (int c0 = 0; c0 < k2; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] = 0.0;
        for (int i = lane; i < k; i += 32) {
            const double yv = y[b * ldy + i];
#pragma unroll
            for (int c = 0; c < 16; c++)
                if (c0 + c < k2) acc[c] += g[(size_t)(c0 + c) * ldg + i] * yv;
        }
#pragma unroll
        for (int c = 0; c < 16; c++) {
            double a = acc[c];
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0 && c0 + c < k2) dcol[b * ldd + c0 + c] = a;
        }
    }
}

unsigned grid_for(size_t total, int block)
{
    size_t g = (total + block - 1) / block;
    const size_t cap = 148u * 32u;  // grid-stride: a few waves of the 148 SMs
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}


// W = G L for a few constraint rows (k2 <= 16) and a lower-triangular L: W[c][j] = sum_{i >= j} G[c][i] L[i][j].
// Thread = column j (rows of L are read coalesced), the row range is split over blockIdx.y; partial sums go to
// part[chunk][c][j] and are added up in a FIXED order by wl_reduce_kernel (bit-reproducible, no atomics).
template <int K2T>
__global__ void __launch_bounds__(128)
wl_partial_kernel(const double* __restrict__ g, size_t ldg, int k2, const double* __restrict__ l, size_t ldl, int k,
                  int rows_per_chunk, double* __restrict__ part, size_t ldp)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int i1 = min(k, i0 + rows_per_chunk);
    __shared__ double gs[K2T][32];  // G[:, s0 .. s0 + 31] of the current 32-row slice
    double acc[K2T];
#pragma unroll
    for (int c = 0; c < K2T; c++) acc[c] = 0.0;
    // rows above this block's first column contribute nothing (L is zero there): skipped slice-wise; inside a
    // slice all 32 loads are issued before the FMAs (the stored zeros above the diagonal make that safe)
    for (int s0 = max(i0, (int)(blockIdx.x * 128) & ~31); s0 < i1; s0 += 32) {
        __syncthreads();
        for (int e = threadIdx.x; e < K2T * 32; e += 128) {
            const int c = e >> 5, i = s0 + (e & 31);
            gs[c][e & 31] = (c < k2 && i < i1) ? g[(size_t)c * ldg + i] : 0.0;
        }
        __syncthreads();
        if (j < k) {
            double lv[32];
#pragma unroll
            for (int u = 0; u < 32; u++) lv[u] = (s0 + u < i1) ? l[(size_t)(s0 + u) * ldl + j] : 0.0;
#pragma unroll
            for (int u = 0; u < 32; u++)
#pragma unroll
                for (int c = 0; c < K2T; c++) acc[c] = fma(gs[c][u], lv[u], acc[c]);
        }
    }
    if (j < k) {
#pragma unroll
        for (int c = 0; c < K2T; c++)
            if (c < k2) part[((size_t)blockIdx.y * k2 + c) * ldp + j] = acc[c];
    }
}

__global__ void wl_reduce_kernel(const double* __restrict__ part, size_t ldp, int nchunks, int k2, int k,
                                 double* __restrict__ w, size_t ldw)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= k) return;
    double v = 0.0;
    for (int q = 0; q < nchunks; q++) v += part[((size_t)q * k2 + c) * ldp + j];
    w[(size_t)c * ldw + j] = v;
}
}  // namespace

extern "C" int htnk_sym_expand(cudaStream_t st, const double* src, size_t lds, int n, int lower_src,
                               double* full, size_t ldf, double* low, size_t ldl)
{
    if (n <= 0) return 0;
    dim3 grid((n + 31) / 32, (n + 31) / 32);
    sym_expand_kernel<<<grid, 256, 0, st>>>(src, lds, n, lower_src, full, ldf, low, ldl, 0, 0);
    return status();
}

extern "C" int htnk_transpose(cudaStream_t st, const double* src, size_t lds, int rows, int cols, double* dst,
                              size_t ldd)
{
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, st>>>(src, lds, rows, cols, dst, ldd, 0);
    return status();
}

extern "C" int htnk_transpose_lower(cudaStream_t st, const double* src, size_t lds, int n, double* dst, size_t ldd)
{
    if (n <= 0) return 0;
    dim3 grid((n + 31) / 32, (n + 31) / 32);
    transpose_kernel<<<grid, 256, 0, st>>>(src, lds, n, n, dst, ldd, 1);
    return status();
}

namespace {
// every 128 x 128 diagonal-block inverse (and its transpose), stored block after block, goes to its place on the diagonal
// of the two n x n inverse-factor matrices: one launch for all the leaves of the recursive inversion
__global__ void scatter_diag_blocks_kernel(const double* __restrict__ dinv, const double* __restrict__ dinvt, int nb, int n,
                                           double* __restrict__ linv, double* __restrict__ linvt, size_t ld)
{
    const int jb = blockIdx.y;
    const int rows = min(nb, n - jb * nb);
    const double* s0 = dinv + (size_t)jb * nb * nb;
    const double* s1 = dinvt + (size_t)jb * nb * nb;
    double* d0 = linv + ((size_t)jb * nb) * ld + (size_t)jb * nb;
    double* d1 = linvt + ((size_t)jb * nb) * ld + (size_t)jb * nb;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < rows * nb; idx += gridDim.x * blockDim.x) {
        const int r = idx / nb, c = idx - r * nb;
        if (c < rows) {
            d0[(size_t)r * ld + c] = s0[(size_t)r * nb + c];
            d1[(size_t)r * ld + c] = s1[(size_t)r * nb + c];
        }
    }
}
}  // namespace

extern "C" int htnk_scatter_diag_blocks(cudaStream_t st, const double* dinv, const double* dinvt, int nb, int n,
                                        double* linv, double* linvt, size_t ld)
{
    if (n <= 0) return 0;
    const int nblk = (n + nb - 1) / nb;
    scatter_diag_blocks_kernel<<<dim3(8, nblk), 256, 0, st>>>(dinv, dinvt, nb, n, linv, linvt, ld);
    return status();
}

extern "C" int htnk_copy2d(cudaStream_t st, const double* src, size_t lds, int rows, int cols, double* dst,
                           size_t ldd, const double* bias)
{
    if (rows <= 0 || cols <= 0) return 0;
    rowwise_kernel<<<grid_for((size_t)rows * cols, 256), 256, 0, st>>>(src, lds, rows, cols, bias, dst, ldd, 0);
    return status();
}

extern "C" int htnk_coldiv(cudaStream_t st, const double* src, size_t lds, int rows, int cols, const double* d,
                           double* dst, size_t ldd)
{
    if (rows <= 0 || cols <= 0) return 0;
    rowwise_kernel<<<grid_for((size_t)rows * cols, 256), 256, 0, st>>>(src, lds, rows, cols, d, dst, ldd, 1);
    return status();
}

extern "C" int htnk_colmul(cudaStream_t st, const double* src, size_t lds, int rows, int cols, const double* d,
                           double* dst, size_t ldd)
{
    if (rows <= 0 || cols <= 0) return 0;
    rowwise_kernel<<<grid_for((size_t)rows * cols, 256), 256, 0, st>>>(src, lds, rows, cols, d, dst, ldd, 2);
    return status();
}

extern "C" int htnk_rsub_rows(cudaStream_t st, double* dst, size_t ldd, int rows, int cols, const double* t)
{
    if (rows <= 0 || cols <= 0) return 0;
    rowwise_kernel<<<grid_for((size_t)rows * cols, 256), 256, 0, st>>>(dst, ldd, rows, cols, t, dst, ldd, 3);
    return status();
}

extern "C" int htnk_diag_extract(cudaStream_t st, const double* a, size_t lda, int n, int op, double* d)
{
    if (n <= 0) return 0;
    diag_extract_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, lda, n, op, d);
    return status();
}

extern "C" int htnk_diag_set(cudaStream_t st, double* m, size_t ld, int n, const double* v, double cval)
{
    if (n <= 0) return 0;
    diag_set_kernel<<<(n + 255) / 256, 256, 0, st>>>(m, ld, n, v, cval);
    return status();
}

namespace {
__global__ void fill_info_kernel(int* p, size_t n, const int* codes)
{
    const int v = codes[0] ? codes[0] : codes[1];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}
}  // namespace
extern "C" int htnk_fill_info(cudaStream_t st, int* p, size_t n, const int* codes)
{
    if (n == 0) return 0;
    fill_info_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, n, codes);
    return status();
}

extern "C" int htnk_fill_i32(cudaStream_t st, int* p, size_t n, int v)
{
    if (n == 0) return 0;
    fill_i32_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, n, v);
    return status();
}

extern "C" int htnk_r_minus_gmean(cudaStream_t st, const double* g, size_t ldg, int k2, int k, const double* mean,
                                  const double* r, double* y)
{
    if (k2 <= 0) return 0;
    r_minus_gmean_kernel<<<k2, 256, 0, st>>>(g, ldg, k, mean, r, y);
    return status();
}

extern "C" int htnk_ht_alpha(cudaStream_t st, size_t nsamples, int k2, const double* d, size_t ldd,
                             const double* c0, const double* ls, size_t ldls, const int* info_s, double* z,
                             size_t ldz, size_t koff)
{
    if (nsamples == 0) return 0;
    if (k2 < 1 || k2 > MAXK2) return -202;
    ht_alpha_kernel<<<(unsigned)((nsamples + 127) / 128), 128, 0, st>>>(nsamples, k2, d, ldd, c0, ls, ldls,
                                                                      info_s, z, ldz, koff);
    return status();
}

extern "C" int htnk_ht_diag_project(cudaStream_t st, size_t nsamples, int k, int k2, const double* cgt,
                                    size_t ldc, const double* alpha, size_t lda, double* out, size_t ldo)
{
    if (nsamples == 0) return 0;
    ht_diag_project_kernel<<<grid_for(nsamples * (size_t)k, 256), 256, 0, st>>>(nsamples, k, k2, cgt, ldc, alpha,
                                                                               lda, out, ldo);
    return status();
}

extern "C" int htnk_rows_dot(cudaStream_t st, size_t nsamples, int k, int k2, const double* g, size_t ldg,
                             const double* y, size_t ldy, double* dcol, size_t ldd)
{
    if (nsamples == 0) return 0;
    rows_dot_kernel<<<(unsigned)((nsamples + 7) / 8), 256, 0, st>>>(nsamples, k, k2, g, ldg, y, ldy, dcol, ldd);
    return status();
}

// ---- test matrix for the factorisation benchmark: lower triangle of the SPD Toeplitz matrix
// a[i][j] = 1 / (1 + |i - j|) + [i == j]   (zeros above the diagonal) ----
namespace {
__global__ void fill_test_spd_kernel(double* f, size_t ld, int n)
{
    const size_t total = (size_t)n * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t i = idx / n, j = idx - i * n;
        f[i * ld + j] = j <= i ? 1.0 / (1.0 + (double)(i - j)) + (i == j ? 1.0 : 0.0) : 0.0;
    }
}
}  // namespace
extern "C" int htnk_fill_test_spd(cudaStream_t st, double* f, size_t ld, int n)
{
    fill_test_spd_kernel<<<148 * 8, 256, 0, st>>>(f, ld, n);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

/* W (k2 x ldw) = G (k2 x ldg) * L (k x ldl, lower triangular), k2 <= 16.  part: workspace of
 * htnk_wl_workspace(k) * k2 * ldp doubles with ldp >= k. */
extern "C" int htnk_wl_chunks(int k)
{
    int n = (k + 31) / 32;  // one 32-row slice per chunk up to 64 chunks, then longer chunks
    return n < 1 ? 1 : (n > 64 ? 64 : n);
}

extern "C" int htnk_wl(cudaStream_t st, const double* g, size_t ldg, int k2, const double* l, size_t ldl, int k,
                       double* part, size_t ldp, double* w, size_t ldw)
{
    if (k2 <= 0 || k <= 0) return 0;
    if (k2 > 16) return -202;
    const int nch = htnk_wl_chunks(k);
    int rows = (k + nch - 1) / nch;
    rows = (rows + 31) / 32 * 32;
    dim3 grid((unsigned)((k + 127) / 128), (unsigned)nch);
    if (k2 == 1) wl_partial_kernel<1><<<grid, 128, 0, st>>>(g, ldg, k2, l, ldl, k, rows, part, ldp);
    else if (k2 <= 4) wl_partial_kernel<4><<<grid, 128, 0, st>>>(g, ldg, k2, l, ldl, k, rows, part, ldp);
    else wl_partial_kernel<16><<<grid, 128, 0, st>>>(g, ldg, k2, l, ldl, k, rows, part, ldp);
    if (status()) return -201;
    wl_reduce_kernel<<<dim3((unsigned)((k + 255) / 256), (unsigned)k2), 256, 0, st>>>(part, ldp, nch, k2, k, w, ldw);
    return status();
}

/* alpha_b = (G cov G^T)^-1 (c0 - W z_b) for every sample row of Z~ in one HBM pass (k2 <= 4); alpha goes to
 * z[b * ldz + koff + c].  ls: Cholesky factor of G cov G^T (k2 == 1: the scalar itself); *info_s != 0 -> alpha = 0. */
extern "C" int htnk_ht_coeff(cudaStream_t st, size_t nsamples, int k, int k2, double* z, size_t ldz, size_t koff,
                             const double* w, size_t ldw, const double* c0, const double* ls, size_t ldls,
                             const int* info_s)
{
    if (nsamples == 0) return 0;
    if (k2 < 1 || k2 > 4 || (ldz & 1) || (ldw & 1) || (reinterpret_cast<uintptr_t>(z) & 15) ||
        (reinterpret_cast<uintptr_t>(w) & 15))
        return -202;
    const unsigned grid = (unsigned)((nsamples + 7) / 8);
    if (k2 == 1)
        ht_coeff_kernel<1><<<grid, 256, 0, st>>>(nsamples, k, k2, z, ldz, koff, w, ldw, c0, ls, ldls, info_s);
    else
        ht_coeff_kernel<4><<<grid, 256, 0, st>>>(nsamples, k, k2, z, ldz, koff, w, ldw, c0, ls, ldls, info_s);
    return status();
}
