/* htnorm-b200: batched entry points (NEW - no counterpart in the reference, which draws one sample
 * per call and refactorises on every call, src/htnorm_distributions.c:75-77).
 *
 * Parity contract.  Sample i of a batch is, by definition, what the reference returns from
 *
 *     rng_t* rng = rng_<gen>_new_seeded(seed_i);            // src/htnorm_rng.c:20-35 / 45-60
 *     htn_hyperplane_truncated_mvn(rng, conf_i, out_i);     // src/htnorm.c:85-187
 *       (or htn_structured_precision_mvn,                   // src/htnorm.c:217-356)
 *
 * with seed_i = seeds[i] (or seed0 + i when seeds == NULL) and conf_i = conf for shared inputs or
 * the i-th problem of the arrays conf points at for independent problems: raw 64-bit streams and
 * ziggurat accept/reject decisions bit-exact, samples within 1e-10 relative.
 *
 * Everything here is plain C ABI: pointers, sizes, integers.  `mem` says whether the config arrays,
 * `seeds`, `out` and `info` live in host memory (copied inside the call, call is synchronous) or are
 * CUDA device pointers on `device` (work is enqueued on `stream`, call returns without synchronising).
 */
#ifndef HTNORM_BATCH_H
#define HTNORM_BATCH_H

#include <stddef.h>
#include <stdint.h>

#include "htnorm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* generator family of the per-sample streams (reference src/htnorm_rng.c:20-35 / 45-60) */
typedef enum { HTN_GEN_PCG64 = 0, HTN_GEN_XRS128P = 1 } htn_gen_t;
typedef enum { HTN_MEM_HOST = 0, HTN_MEM_DEVICE = 1 } htn_mem_t;

/* negative return codes beyond HTNORM_ALLOC_ERROR (-100) */
#define HTN_E_CUDA        (-201) /* a CUDA runtime call or kernel launch failed; see htn_last_error() */
#define HTN_E_ARG         (-202) /* NULL / zero-sized / inconsistent argument */
#define HTN_E_NO_DEVICE   (-203) /* no usable sm_100 device: there is NO CPU fallback */
#define HTN_E_FOREIGN_RNG (-204) /* reserved: foreign rng_t objects are served through their callbacks */

/* flags */
#define HTN_FLAG_PAPER_SIGN 0x1u /* structured precision: use the sign of Cong et al. (2017), not the
                                    reference's (src/htnorm.c:229,320,336,343).  Default off. */
#define HTN_FLAG_COLMAJOR   0x2u /* g / phi are column-major, symmetric inputs read through the lower
                                    triangle: the reference's -DHTNORM_COLMAJOR build (src/Makevars:1) */

typedef struct {
    size_t nsamples;       /* B: number of samples to draw (= number of problems when independent) */
    htn_gen_t gen;         /* generator family of every stream */
    const uint64_t* seeds; /* nsamples seeds (same memory space as `out`), or NULL */
    uint64_t seed0;        /* used when seeds == NULL: seed_i = seed0 + i */
    int independent;       /* 0: all samples share conf's arrays (factorised once);
                              1: conf's arrays hold nsamples problems back to back
                                 (mean: B x k, cov: B x k x k, g: B x k2 x k, r: B x k2) */
    htn_mem_t mem;         /* memory space of conf arrays, seeds, out, info */
    int device;            /* CUDA device ordinal, -1 = current device */
    void* stream;          /* cudaStream_t for HTN_MEM_DEVICE (NULL = legacy default stream) */
    unsigned flags;        /* HTN_FLAG_* */
} htn_batch_t;

/* Fill *b with defaults: PCG64, seed0 = 0, shared inputs, host memory, current device. */
void htn_batch_init(htn_batch_t* b, size_t nsamples);

/* out: nsamples x gncol, row-major (sample i contiguous).  info: nsamples ints or NULL, per-sample
 * code with the meaning of the single-sample return value.  Return: <0 on a library failure,
 * otherwise (host memory) the first non-zero per-sample code or 0; (device memory) 0 once enqueued. */
int htn_hyperplane_truncated_mvn_batch(const htn_batch_t* b, const ht_config_t* conf,
                                       double* out, int* info);

/* out: nsamples x pncol.  Independent problems are not supported for this sampler (HTN_E_ARG). */
int htn_structured_precision_mvn_batch(const htn_batch_t* b, const sp_config_t* conf,
                                       double* out, int* info);

/* Raw streams: out[i * nwords + t] = t-th next_uint64() of stream i
 * (reference src/htnorm_pcg64.h:23-28, src/htnorm_xoroshiro128p.h:24-37). */
int htn_rng_raw_fill(const htn_batch_t* b, size_t nwords, uint64_t* out);

/* Ziggurat normals (reference src/htnorm_distributions.c:12-13, 18-54): out[i * n + (n-1-j)] is the
 * j-th accepted normal of stream i (the reference fills back to front).  ndraws (nsamples u32, or
 * NULL) receives the number of generator calls (next_uint64 + next_double) stream i consumed -
 * the record of its accept/reject decisions. */
int htn_std_normal_fill(const htn_batch_t* b, size_t n, double* out, uint32_t* ndraws);

/* diagnostics */
int htn_device_count(void);             /* number of CUDA devices, <0 on error */
const char* htn_last_error(void);       /* thread-local description of the last HTN_E_* */
unsigned long long htn_launch_count(void); /* kernels this library has launched in this process */
const char* htn_version(void);

/* Optional phase timing of the shared-covariance hyperplane path: when enabled, CUDA events are recorded
 * on the call's stream at the phase boundaries of every call (up to 256 calls are kept, no synchronisation);
 * htn_phase_timing_get synchronises on them and writes the AVERAGE milliseconds of [0] setup (factorisation
 * etc.), [1] exposed wait for the normals, [2] W z + coefficient solve, [3] the sampling GEMM (-1 where
 * unavailable; with several chunks per call the last chunk counts) and returns the number of calls averaged. */
void htn_phase_timing_enable(int on);
int htn_phase_timing_get(double* ms, int n);

/* Measured FP64 tensor-core (DMMA, mma.sync m8n8k4.f64) throughput of `device` in TFLOP/s, register
 * operands only, all SMs, timed with CUDA events over roughly `ms` milliseconds; <0 on error.  This is
 * the denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 entry). */
double htn_fp64_tensor_peak_tflops(int device, double ms);
/* Same for plain DFMA (FP64 FMA pipe), for comparison. */
double htn_fp64_fma_peak_tflops(int device, double ms);
/* Measured integer ISSUE throughput of `device` with the bit generators' own instruction mix (64-bit LCG steps,
 * xorshift / rotate outputs, no memory traffic), in giga warp-instructions per second over all SMs.  The ziggurat
 * generator is bound by issue slots, not by HBM: this is the denominator of its second roofline. */
double htn_int_issue_peak_gops(int device, double ms);

#ifdef __cplusplus
}
#endif
#endif
