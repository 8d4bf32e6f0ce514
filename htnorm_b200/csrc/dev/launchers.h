/* Internal C ABI between the C host layer (csrc/host) and the CUDA translation units
 * (csrc/dev).  One function = one kernel launch (or a fixed short sequence) on `stream`;
 * plain pointers and sizes only; every pointer is a DEVICE pointer unless it says host.
 * Return value: 0 or HTN_E_CUDA (-201) if the launch failed (cudaGetLastError).
 */
#ifndef HTNB_LAUNCHERS_H
#define HTNB_LAUNCHERS_H

#include <stddef.h>
#include <stdint.h>

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bookkeeping ---- */
void htnk_count_launch(void);
unsigned long long htnk_launch_count(void);

/* ---- generators + ziggurat (rng_zig.cu) ---- */

/* where stream b starts: state words (4 per stream, see htn_gen_state) or seeds / seed0 + b */
typedef struct {
    int gen;                /* 0 PCG64 (2 x PCG32), 1 Xoroshiro128+ */
    const uint64_t* seeds;  /* device, nstreams, or NULL */
    uint64_t seed0;         /* seed_b = seed0 + b when seeds == NULL */
    uint64_t* states;       /* device, 4 words per stream, or NULL.  If non-NULL the streams START
                               from these states (seeds ignored) and the advanced states are written back */
    const double* zpre;     /* device or NULL: normals already drawn (draw order, all segments of stream b
                               back to back); the generator is not run, they are only scattered */
} htnk_streams_t;

/* one run of normals per stream: element j (draw order) goes to out[b*ld + (n-1-j)] after
 * v = z; if (div) v = z / div[i]; if (mul) v = z * mul[i]; if (add) v = add[i] + v   (i = n-1-j) */
typedef struct {
    size_t n;
    double* out;
    size_t ld;
    const double* div;
    const double* mul;
    const double* add;
} htnk_segment_t;

int htnk_raw_fill(cudaStream_t st, const htnk_streams_t* s, size_t nstreams, size_t nwords, uint64_t* out);
/* nseg = 1 or 2 segments drawn back to back from each stream; ndraws: u32 per stream or NULL */
/* one-shot hint: the NEXT htnk_normal_fill of this thread overlaps other kernels and leaves some SMs to them */
void htnk_normal_fill_shares_device_next(void);
int htnk_normal_fill(cudaStream_t st, const htnk_streams_t* s, size_t nstreams, int nseg,
                     const htnk_segment_t* seg, uint32_t* ndraws);

/* ---- FP64 tensor-core GEMM (gemm_f64.cu) ----
 * C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * C + bias[n]      (all row-major)
 * A, B: 16-byte aligned, lda/ldb even.  C: any ldc.  bias may be NULL.
 * mode: HTNK_GEMM_FULL; HTNK_GEMM_B_LOWER - B is lower-trapezoidal: B[n][kk] == 0 for
 *       n < kk < ktri, and columns [kext0, K) are dense (K range per tile is clipped accordingly);
 *       HTNK_GEMM_C_LOWER - only tiles that touch the lower triangle (m >= n) are computed;
 *       HTNK_GEMM_B_UPPER - B[n][kk] == 0 for kk < n: the K loop of a tile starts at its first column.
 * bn: N-tile width, 128 / 32 / 8 (0 = pick from N). */
enum { HTNK_GEMM_FULL = 0, HTNK_GEMM_B_LOWER = 1, HTNK_GEMM_C_LOWER = 2, HTNK_GEMM_B_UPPER = 3 };
int htnk_gemm_tn(cudaStream_t st, int M, int N, int K, double alpha, const double* A, size_t lda,
                 const double* B, size_t ldb, double beta, double* C, size_t ldc, const double* bias,
                 int mode, int ktri, int kext0, int bn);

/* TMA-fed variant for wide N (gemm_tma.cu); returns 1 when the operands do not admit a tensor map */
int htnk_gemm_tn_tma(cudaStream_t st, int M, int N, int K, double alpha, const double* A, size_t lda,
                     const double* B, size_t ldb, double beta, double* C, size_t ldc, const double* bias,
                     int mode, int ktri, int kext0, const int* skip);
/* one-shot guard: the NEXT htnk_gemm_tn launch of this thread becomes a no-op on the device if *flag != 0 */
void htnk_gemm_guard_next(const int* flag);

/* ---- factorisation building blocks (chol.cu) ---- */
/* Factor the nb x nb (nb <= 128) diagonal block at a (row-major, ld) in place: lower triangle := L,
 * strictly-upper := 0.  rdiag (device, nb doubles, may be NULL) receives 1 / L[j][j].  *info (device int)
 * receives col0 + j + 1 at the first pivot <= 0 (only if it still holds 0). */
int htnk_potf2(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info);
/* diagonal block + the nrows_below rows under it (L21 = A21 L11^-T) in ONE launch when the block is full (nb = 128)
 * and 16-byte aligned: *rows_done = 1; otherwise only the block is factorised and *rows_done = 0 (run htnk_trsm_rows).
 * tscr: 16 * 64 doubles, flags: 17 unsigned - scratch private to the factorisation, no initialisation needed */
int htnk_potf2_panel(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info, int nrows_below,
                     double* tscr, unsigned* flags, int* rows_done);
/* next block column of the look-ahead factorisation: C (rows x nb2) -= X (rows x 128) X[0:nb2, :]^T (chol.cu); returns 1
 * when the operands are not 16-byte aligned (take the general GEMM) */
int htnk_panel_update(cudaStream_t st, double* c, size_t ldc, const double* x, size_t ldx, int rows, int nb2);
/* x (nrows x nb, leading dimension ldx) := x * L^-T with the nb x nb lower block l; rdiag = 1 / diag(l) */
int htnk_trsm_rows(cudaStream_t st, double* x, size_t ldx, int nrows, const double* l, size_t ldl, int nb,
                   const double* rdiag);
/* inverse of every 128 x 128 diagonal block of the n x n lower factor f: block b -> dinv + b * 128 * 128
 * (row-major, leading dimension 128) and its transpose -> dinvt + b * 128 * 128 */
int htnk_trtri_batched(cudaStream_t st, const double* f, size_t ld, int n, double* dinv, double* dinvt);

/* ---- layout / elementwise helpers (elementwise.cu) ---- */
/* src: n x n row-major (ld lds), symmetric, read through its UPPER triangle (lower_src = 0) or LOWER
 * triangle (lower_src = 1).  full (may be NULL): n x n symmetric copy (ld ldf).  low (may be NULL):
 * lower triangle copy with zeros above the diagonal (ld ldl). */
int htnk_sym_expand(cudaStream_t st, const double* src, size_t lds, int n, int lower_src, double* full,
                    size_t ldf, double* low, size_t ldl);
/* dst[c][r] = src[r][c], src rows x cols */
int htnk_transpose(cudaStream_t st, const double* src, size_t lds, int rows, int cols, double* dst,
                   size_t ldd);
/* the same for a lower-triangular n x n source (zeros above the diagonal are not read; every element of dst is written) */
int htnk_transpose_lower(cudaStream_t st, const double* src, size_t lds, int n, double* dst, size_t ldd);
/* dst[r][c] = src[r][c] (rows x cols), optional bias[c] added; strided 2-D copy */
/* all nb x nb diagonal-block inverses (stored block after block) and their transposes -> the diagonals of linv / linvt */
int htnk_scatter_diag_blocks(cudaStream_t st, const double* dinv, const double* dinvt, int nb, int n, double* linv,
                             double* linvt, size_t ld);
int htnk_copy2d(cudaStream_t st, const double* src, size_t lds, int rows, int cols, double* dst,
                size_t ldd, const double* bias);
/* dst[r][c] = src[r][c] / d[c]  (structured precision, diagonal A: src/htnorm.c:272-280) */
int htnk_coldiv(cudaStream_t st, const double* src, size_t lds, int rows, int cols, const double* d,
                double* dst, size_t ldd);
/* dst[r][c] = src[r][c] * d[c] */
int htnk_colmul(cudaStream_t st, const double* src, size_t lds, int rows, int cols, const double* d,
                double* dst, size_t ldd);
/* dst[r][c] = t[c] - dst[r][c] */
int htnk_rsub_rows(cudaStream_t st, double* dst, size_t ldd, int rows, int cols, const double* t);
/* d[i] = a[i*lda + i] (op 0), sqrt of it (op 1), 1 / it (op 2) */
int htnk_diag_extract(cudaStream_t st, const double* a, size_t lda, int n, int op, double* d);
/* m[i*ld + i] += (or =) v[i] / constant */
int htnk_diag_set(cudaStream_t st, double* m, size_t ld, int n, const double* v, double cval);
/* fused coefficient pass (k2 <= 4): alpha_b = (G cov G^T)^-1 (c0 - W z_b) into z[b * ldz + koff + c], one HBM pass */
int htnk_ht_coeff(cudaStream_t st, size_t nsamples, int k, int k2, double* z, size_t ldz, size_t koff, const double* w,
                  size_t ldw, const double* c0, const double* ls, size_t ldls, const int* info_s);
/* W (k2 x ldw) = G (k2 x ldg) * L (k x ldl, lower triangular), k2 <= 16; part: htnk_wl_chunks(k) * k2 * ldp doubles */
int htnk_wl_chunks(int k);
int htnk_wl(cudaStream_t st, const double* g, size_t ldg, int k2, const double* l, size_t ldl, int k, double* part,
            size_t ldp, double* w, size_t ldw);
int htnk_fill_i32(cudaStream_t st, int* p, size_t n, int v);
/* p[i] = codes[0] ? codes[0] : codes[1]  (codes: two device ints, e.g. chol(cov) and chol(G cov G^T)) */
int htnk_fill_info(cudaStream_t st, int* p, size_t n, const int* codes);
/* y[c] = r[c] - sum_i g[c][i] * mean[i]   (k2 rows) */
int htnk_r_minus_gmean(cudaStream_t st, const double* g, size_t ldg, int k2, int k, const double* mean,
                       const double* r, double* y);
/* hyperplane coefficients, k2 <= 16: per sample b solve (L L^T) alpha = c0 - d[b][:] with the k2 x k2
 * lower factor ls (leading dimension ldls); k2 == 1 divides by the scalar s11 instead (src/htnorm.c:77).  alpha is
 * written to z[b*ldz + koff + c].  If *info_s != 0 alpha = 0 (projection skipped, src/htnorm.c:170). */
int htnk_ht_alpha(cudaStream_t st, size_t nsamples, int k2, const double* d, size_t ldd, const double* c0,
                  const double* ls, size_t ldls, const int* info_s, double* z, size_t ldz, size_t koff);
/* diagonal covariance: out[b][i] = y[b][i] + sum_c cgt[i*ldc + c] * alpha[b][c]
 * where y[b][i] is already in out (mean + sqrt(cov_ii) z) */
int htnk_ht_diag_project(cudaStream_t st, size_t nsamples, int k, int k2, const double* cgt, size_t ldc,
                         const double* alpha, size_t lda, double* out, size_t ldo);
/* dcol[b][c] = sum_i g[c][i] * y[b][i]  (k2 <= 16; diagonal-covariance path) */
int htnk_rows_dot(cudaStream_t st, size_t nsamples, int k, int k2, const double* g, size_t ldg,
                  const double* y, size_t ldy, double* dcol, size_t ldd);

/* ---- small independent problems (small_batched.cu) ---- */
/* One CTA per problem, k <= 128, k2 <= 16: the whole of src/htnorm.c:85-187 for problem b with its own
 * mean/cov/g/r (back to back); zs holds the problem's k normals (already drawn, reverse-fill order);
 * info[b] as the reference's return code. */
int htnk_ht_small_batched(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                          const double* cov, const double* g, const double* r, const double* zs, int diag,
                          int lower_src, double* out, int* info);

/* One WARP per problem, k = 64, k2 <= 4, dense row-major cov (small_warp.cu): same contract as above.  Returns 1 when the
 * shape is not one it is built for (the caller then takes htnk_ht_small_batched).  counter: one zero-initialised device
 * unsigned, the kernel's dynamic problem queue. */
int htnk_ht_small_warp(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean, const double* cov,
                       const double* g, const double* r, const double* zs, double* out, int* info, unsigned* counter);

/* ---- batched building blocks of the mid-size independent problems (one launch over all problems; strides in doubles) ---- */
/* left-looking panel step of the batched mid-size factorisation (mid_batched.cu): diagonal block / rows below it */
int htnk_chol_rows_batched(cudaStream_t st, const double* cov, size_t scov, int lower_src, double* F, size_t ld, size_t sF, int k,
                           int j0, const double* rdiag, size_t sr, const int* info, int diag, size_t nbatch);
/* diagonal 64-block of every problem's lower factorisation, one WARP per problem (mid_batched.cu): factorises in place the
 * nb x nb (nb <= 64) block at row / column j0 of F + b * sF (lower storage, leading dimension ld); rdiag + b * sr + j0 receives
 * 1 / L[j][j]; info[b] receives j0 + j + 1 at the first pivot <= 0 (if it still holds 0; problems whose info is set are skipped).
 * counter: one zero-initialised device unsigned.  src != NULL (j0 = 0 only): the block is read from the caller's matrices
 * (src + b * ssrc, leading dimension src_ld, lower or upper triangle valid) instead of from F. */
int htnk_potf64_batched(cudaStream_t st, double* F, size_t ld, size_t sF, int j0, int nb, double* rdiag, size_t sr, int* info,
                        size_t nbatch, unsigned* counter, const double* src, size_t ssrc, size_t src_ld, int src_lower);
/* projection of one sample per problem given the factor (mid_batched.cu): x = y + cov G^T (G cov G^T)^-1 (r - G y), y = mean + L z,
 * formed around the factor (V = L^T G^T); k2 <= 4.  info[b]: set by the factorisation (row untouched) or by the k2 x k2 solve. */
int htnk_ht_mid_project(cudaStream_t st, size_t nbatch, int k, int k2, const double* F, size_t ld, size_t sF,
                        const double* mean, const double* g, const double* r, const double* zs, int g_colmajor, double* out,
                        int* info);

/* ---- peak microbenchmarks (peak.cu) ---- */
double htnk_dmma_peak_tflops(cudaStream_t st, double ms);
double htnk_dfma_peak_tflops(cudaStream_t st, double ms);
/* integer issue ceiling with the generator's instruction mix, giga warp-instructions per second */
double htnk_int_issue_peak_gops(cudaStream_t st, double ms);

#ifdef __cplusplus
}
#endif
#endif
