/* htnorm-b200: public sampler API of the drop-in boundary (B200 / sm_100a implementation).
 *
 * Source-compatible with the reference's include/htnorm.h (zoj613/htnorm): same type names, field
 * order, enum values and function signatures, so a C caller (reference README.md:57-76), the Cython
 * wrapper (pyhtnorm/_htnorm.pyx:199-206, 296-304) or the R wrapper (src/htnorm_r_wrapper.c:55-98)
 * re-links against libhtnorm_b200.so unchanged.  Written from scratch; the algorithms are
 * Cong, Chen & Zhou (2017), "Fast simulation of hyperplane-truncated multivariate normal
 * distributions", Bayesian Analysis 12(4), Alg. 2 / Alg. 4 / Example 4.
 *
 *   type_t                         <-> reference include/htnorm.h:34
 *   ht_config_t / init_ht_config   <-> reference include/htnorm.h:36-56,  src/htnorm.c:17-28
 *   sp_config_t / init_sp_config   <-> reference include/htnorm.h:58-82,  src/htnorm.c:31-45
 *   htn_hyperplane_truncated_mvn   <-> reference include/htnorm.h:105,    src/htnorm.c:85-187
 *   htn_structured_precision_mvn   <-> reference include/htnorm.h:131,    src/htnorm.c:217-356
 *
 * All matrices are row-major (C order) doubles.  Symmetric inputs are read through their UPPER
 * triangle (row-major), which is what the reference's uplo='L' Fortran calls see.
 * Return codes: 0 success; >0 order of the leading minor that is not positive definite; -100
 * (HTNORM_ALLOC_ERROR) host or device allocation failure; other negative values: HTN_E_* in
 * htnorm_batch.h (CUDA runtime failure, bad argument).  NaN inputs give NaN outputs with code 0.
 */
#ifndef HTNORM_HTNORM_H
#define HTNORM_HTNORM_H

#include <stdbool.h>
#include <stddef.h>

#include "htnorm_rng.h"

#ifdef __cplusplus
#define HTN_RESTRICT __restrict__
extern "C" {
#else
#define HTN_RESTRICT restrict
#endif

#ifndef HTNORM_ALLOC_ERROR
#define HTNORM_ALLOC_ERROR -100 /* reference src/htnorm_distributions.h:19 */
#endif

/* structure of a precision matrix: dense, diagonal (only a[i][i] is read) or the identity */
typedef enum {NORMAL, DIAGONAL, IDENTITY} type_t;

/* N(mean, cov) truncated on {x : g x = r};  g is gnrow x gncol, cov is gncol x gncol */
typedef struct {
    size_t gnrow;
    size_t gncol;
    const double* mean;
    const double* cov;
    const double* g;
    const double* r;
    bool diag; /* cov is diagonal: only cov[i][i] is read */
} ht_config_t;

void init_ht_config(ht_config_t* conf, size_t gnrow, size_t gncol,
                    const double* mean, const double* cov, const double* g,
                    const double* r, bool diag);

/* N(mean, (a + phi^T omega phi)^-1), phi is pnrow x pncol, a is pncol x pncol, omega pnrow x pnrow.
 * With struct_mean the vector `mean` has length pnrow and plays the role of t in Example 4. */
typedef struct {
    type_t a_id;
    type_t o_id;
    size_t pnrow;
    size_t pncol;
    const double* mean;
    const double* a;
    const double* phi;
    const double* omega;
    bool struct_mean;
} sp_config_t;

void init_sp_config(sp_config_t* conf, size_t pnrow, size_t pncol, const double* mean,
                    const double* a, const double* phi, const double* omega,
                    bool struct_mean, type_t a_id, type_t o_id);

/* One draw; `out` has gncol doubles.  `rng` is advanced exactly as the reference advances it. */
int htn_hyperplane_truncated_mvn(rng_t* rng, const ht_config_t* conf, double* HTN_RESTRICT out);

/* One draw; `out` has pncol doubles. */
int htn_structured_precision_mvn(rng_t* rng, const sp_config_t* conf, double* HTN_RESTRICT out);

#ifdef __cplusplus
}
#endif
#endif
