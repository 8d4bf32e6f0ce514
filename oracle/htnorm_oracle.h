/* TEST INFRASTRUCTURE ONLY - CPU restatement ("oracle") of the reference's sampling path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this.  The product (htnorm_b200/, libhtnorm_b200.so) never includes, links or calls anything here.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (its tests pin properties only,
 * tests/test_pyhtnorm.py:34-124), so this restatement is pinned against outputs of the UNMODIFIED
 * reference compiled in this container (oracle/build_ref.sh -> oracle/_ref/libhtnorm_ref.so) and
 * against the fixtures generated from it (tests/golden/, tools/make_golden.py).
 */
#ifndef HTNORM_ORACLE_H
#define HTNORM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_GEN_PCG64 = 0, ORC_GEN_XRS128P = 1 };
enum { ORC_NORMAL = 0, ORC_DIAGONAL = 1, ORC_IDENTITY = 2 };

typedef struct {
    int kind;
    uint64_t s[4];   /* PCG64: gen0.state, gen0.inc, gen1.state, gen1.inc; XRS128P: s0, s1 */
    uint64_t ncalls; /* next_uint64 + next_double calls so far */
} orc_rng_t;

void orc_rng_seed(orc_rng_t* g, int kind, uint64_t seed);
uint64_t orc_next_u64(orc_rng_t* g);
double orc_next_double(orc_rng_t* g);
double orc_std_normal(orc_rng_t* g);

/* out[i*nwords + t] */
void orc_raw_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t nwords,
                  uint64_t* out);
/* out[i*n + (n-1-j)] = j-th normal of stream i; ndraws[i] = generator calls consumed (may be NULL) */
void orc_normal_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t n,
                     double* out, uint32_t* ndraws);

/* Lower Cholesky factor of the symmetric matrix read through the row-major UPPER triangle of a
 * (lda = n).  L is written row-major into l (strictly-upper part zeroed).  Returns LAPACK info. */
int orc_potrf(int n, const double* a, double* l);

/* single draws; rng advanced as the reference advances it; return = reference return code */
int orc_hyperplane(orc_rng_t* g, int k2, int k, const double* mean, const double* cov,
                   const double* gm, const double* r, int diag, double* out);
int orc_structured(orc_rng_t* g, int n, int p, const double* mean, const double* a,
                   const double* phi, const double* omega, int struct_mean, int a_id, int o_id,
                   int paper_sign, double* out);

/* batches: sample i uses a generator freshly seeded with seeds[i] (or seed0 + i).  independent != 0:
 * inputs hold nsamples problems back to back.  Shared inputs are factorised once (same arithmetic). */
int orc_hyperplane_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples,
                         int independent, int k2, int k, const double* mean, const double* cov,
                         const double* gm, const double* r, int diag, double* out, int* info);
int orc_structured_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples, int n,
                         int p, const double* mean, const double* a, const double* phi,
                         const double* omega, int struct_mean, int a_id, int o_id, int paper_sign,
                         double* out, int* info);

#ifdef __cplusplus
}
#endif
#endif
