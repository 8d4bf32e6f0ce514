/* Host layer internals (C99).  The host layer is plain C above a C ABI of kernel launchers
 * (../dev/launchers.h); it owns validation, workspace, ordering of launches and the error codes. */
#ifndef HTN_INTERNAL_H
#define HTN_INTERNAL_H

#include <stddef.h>
#include <stdint.h>

#include <cuda_runtime_api.h>

#include "../../../include/htnorm.h"
#include "../../../include/htnorm_batch.h"
#include "../dev/launchers.h"

#define HTN_NB 128 /* panel width of the blocked factorisation = N tile of the GEMM */

/* Layout default of THIS build of the library.  The reference selects the layout at compile time
 * (src/htnorm.c:5-10, -DHTNORM_COLMAJOR in src/Makevars:1 for the R binding); libhtnorm_b200_colmajor.so is this
 * library compiled the same way: every entry point then behaves as if HTN_FLAG_COLMAJOR were set. */
#ifdef HTNORM_COLMAJOR
#define HTN_DEFAULT_FLAGS HTN_FLAG_COLMAJOR
#else
#define HTN_DEFAULT_FLAGS 0u
#endif

typedef struct {
    int device;
    int prev_device;
    cudaStream_t stream;
    int own_stream;
    /* per-call options of the hyperplane path, set by the entry points (zero = defaults) */
    double* host_sink;    /* host-memory batch: caller's output array; finished sub-chunks are copied there directly */
    int host_sink_used;   /* out: the sink received the samples */
    int* async_info_out;  /* device-memory batch: per-sample info array filled on the device, no host synchronisation */
    int cov_failed;       /* out: chol(cov) failed, nothing was drawn */
} htn_ctx_t;

void htn_set_error(const char* fmt, ...);
/* bytes of normals drawn per chunk of a large batch: 3 GiB, or HTN_CHUNK_BYTES (tests: forces many small chunks) */
size_t htn_chunk_budget(void);

/* select the device, pick/create the stream.  host_mode: create a private non-blocking stream */
int htn_ctx_begin(htn_ctx_t* ctx, int device, void* stream, int host_mode);
/* synchronise (if sync), destroy a private stream, restore the previous device */
int htn_ctx_end(htn_ctx_t* ctx, int sync);

/* stream-ordered workspace (cudaMallocAsync on the device's default pool, kept cached) */
int htn_dalloc(htn_ctx_t* ctx, void** p, size_t bytes);
void htn_dfree(htn_ctx_t* ctx, void* p);

/* phase marks for the optional timing facility (context.c) */
void htn_prof_mark(htn_ctx_t* ctx, int slot);

#define HTN_TRY(expr)                 \
    do {                              \
        rc = (expr);                  \
        if (rc) goto done;            \
    } while (0)

#define HTN_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            htn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                          __FILE__, __LINE__);                                          \
            rc = (e__ == cudaErrorMemoryAllocation) ? HTNORM_ALLOC_ERROR : HTN_E_CUDA;  \
            goto done;                                                                  \
        }                                                                               \
    } while (0)

static inline size_t htn_round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

/* ---- dense building blocks (linalg.c) ---- */

/* Cholesky factor of an SPD matrix plus what the blocked triangular solves need */
typedef struct {
    int n;
    size_t ld;      /* leading dimension of f (and of u) */
    double* f;      /* n x ld: L in the lower triangle, zeros above */
    double* u;      /* n x ld: L^T (upper), zeros below; NULL until htn_factor_make_u */
    double* dinv;   /* nblk x 128 x 128: inverse of each diagonal block of L */
    double* dinvt;  /* transposes of the above */
    double* rdiag;  /* 1 / L[j][j] */
    double* pscr;   /* scratch of the fused panel kernel: 16 x 64 doubles (8 x 8 inverse factors) + 17 flags */
    double* linv;   /* n x ld: explicit L^-1 (lower), NULL until htn_factor_make_inverse */
    double* linvt;  /* its transpose */
    int inv_block;  /* 0: linv / linvt hold all of L^-1; W > 0: only its W x W diagonal blocks (block substitution) */
    int nblk;
    int solves;     /* the factor will be used by the blocked triangular solves (needs every inverse) */
} htn_factor_t;

/* allocate dinv/dinvt (+u if want_u) for an n x n factor living in f (caller-owned, n x ld) */
int htn_factor_init(htn_ctx_t* ctx, htn_factor_t* fa, double* f, size_t ld, int n, int want_u);
void htn_factor_free(htn_ctx_t* ctx, htn_factor_t* fa);
/* blocked right-looking Cholesky of the matrix in fa->f (lower triangle read).  *d_info (device int,
 * must hold 0) receives LAPACK's info.  Replaces dpotrf (reference src/htnorm_blas.h:111-112). */
int htn_potrf(htn_ctx_t* ctx, htn_factor_t* fa, int* d_info);
/* u = f^T */
int htn_factor_make_u(htn_ctx_t* ctx, htn_factor_t* fa);
/* X (m x n, ldx even, 16B-aligned) := X * L^-T   (solves  X_new L^T = X) */
int htn_trsm_right_lt(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m);
/* X := X * L^-1   (solves X_new L = X); needs fa->u */
int htn_trsm_right_l(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m);
int htn_factor_make_inverse(htn_ctx_t* ctx, htn_factor_t* fa);
/* only the inverses of the W x W diagonal blocks of L (W a multiple of 128): the solves below then substitute block by
 * block - cheaper than the whole inverse when few rows are solved against a large factor; needs fa->u */
int htn_factor_make_inverse_blocks(htn_ctx_t* ctx, htn_factor_t* fa, int w);
int htn_solve_l(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m);
int htn_solve_lt(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* y, size_t ldy, int m);
int htn_solve_spd(htn_ctx_t* ctx, const htn_factor_t* fa, const double* x, size_t ldx, double* t, double* y,
                  size_t ldy, int m);
/* X := X * (L L^T)^-1  (dpotrs for row vectors) */
int htn_potrs_rows(htn_ctx_t* ctx, const htn_factor_t* fa, double* x, size_t ldx, int m);

/* ---- generator helpers (rng.c) ---- */
/* 0 PCG64 / 1 XRS128P if rng was made by this library's constructors, -1 otherwise */
int htn_rng_kind(const rng_t* rng);

/* ---- samplers on device pointers (hyperplane.c / structured.c) ---- */
typedef struct {
    int gen;
    const uint64_t* seeds; /* device or NULL */
    uint64_t seed0;
    uint64_t* states;      /* device or NULL */
    const double* zpre;    /* device or NULL: pre-drawn normals of a foreign generator */
} htn_streams_t;
void htn_foreign_draw(rng_t* rng, size_t n, double* z);

int htn_ht_device(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, int independent, unsigned flags,
                  size_t k2, size_t k, const double* mean, const double* cov, const double* g,
                  const double* r, int diag, double* out, int* d_info_out, int* first_info);
int htn_ht_cov_info(htn_ctx_t* ctx, int k, const double* cov, unsigned flags, int* info);
int htn_sp_device(htn_ctx_t* ctx, const htn_streams_t* s, size_t nsamples, unsigned flags, size_t n, size_t p,
                  const double* mean, const double* a, const double* phi, const double* omega,
                  int struct_mean, int a_id, int o_id, double* out, int* d_info_out, int* first_info);

#endif
