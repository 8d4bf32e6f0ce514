/* TEST INFRASTRUCTURE ONLY - drives the UNMODIFIED reference library (oracle/_ref/libhtnorm_ref.so,
 * built by oracle/build_ref.sh from /root/reference/src in place) over batches, so that tests can pin
 * the oracle and the CUDA path against it and bench.py can time it on the host cores.
 *
 * It is compiled against THIS repo's include/htnorm.h - the drop-in headers - and linked against the
 * reference binary: that the two agree on every struct layout and signature is itself a parity check.
 * Sample i = rng_<gen>_new_seeded(seed_i); htn_*_mvn(rng, conf_i, out_i): the parity definition of
 * include/htnorm_batch.h.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "htnorm.h"

/* exported by the reference binary but declared only in its private header
 * (src/htnorm_distributions.h:83-84) */
int mvn_rand_cov(rng_t* rng, const double* mean, const double* mat, int nrow, bool diag,
                 double* restrict out);

static rng_t* make_rng(int kind, uint64_t seed)
{
    return kind == 0 ? rng_pcg64_new_seeded(seed) : rng_xrs128p_new_seeded(seed);
}

static uint64_t seed_of(const uint64_t* seeds, uint64_t seed0, size_t i)
{
    return seeds ? seeds[i] : seed0 + (uint64_t)i;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void refh_raw_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t nwords,
                   uint64_t* out)
{
    for (size_t i = 0; i < nstreams; i++) {
        rng_t* rng = make_rng(kind, seed_of(seeds, seed0, i));
        for (size_t t = 0; t < nwords; t++)
            out[i * nwords + t] = rng->next_uint64(rng->base);
        rng_free(rng);
    }
}

/* counting shim around a reference generator: records the accept/reject history as a call count */
typedef struct {
    rng_t* inner;
    uint64_t ncalls;
} counting_t;
static uint64_t counting_u64(void* b)
{
    counting_t* c = b;
    c->ncalls++;
    return c->inner->next_uint64(c->inner->base);
}
static double counting_dbl(void* b)
{
    counting_t* c = b;
    c->ncalls++;
    return c->inner->next_double(c->inner->base);
}

/* normals through the reference's own ziggurat: mvn_rand_cov with mean 0 and a diagonal unit
 * covariance returns 0 + sqrt(1) * z, i.e. the normals themselves, reverse-filled. */
int refh_normal_fill(int kind, const uint64_t* seeds, uint64_t seed0, size_t nstreams, size_t n,
                     double* out, uint32_t* ndraws)
{
    double* mean = calloc(n, sizeof(double));
    /* only the diagonal of cov is read when diag is set: point every row at a row of ones */
    double* cov = malloc(((n - 1) * n + n) * sizeof(double));
    if (!mean || !cov) {
        free(mean);
        free(cov);
        return -100;
    }
    for (size_t i = 0; i < (n - 1) * n + n; i++)
        cov[i] = 1.0;
    for (size_t i = 0; i < nstreams; i++) {
        rng_t* inner = make_rng(kind, seed_of(seeds, seed0, i));
        counting_t c = {inner, 0};
        rng_t shim = {&c, counting_u64, counting_dbl};
        mvn_rand_cov(&shim, mean, cov, (int)n, true, out + i * n);
        if (ndraws)
            ndraws[i] = (uint32_t)c.ncalls;
        rng_free(inner);
    }
    free(mean);
    free(cov);
    return 0;
}

typedef struct {
    int kind, sp;
    const uint64_t* seeds;
    uint64_t seed0;
    size_t lo, hi;
    int independent;
    ht_config_t ht;
    sp_config_t spc;
    double* out;
    int* info;
} job_t;

static void* worker(void* arg)
{
    job_t* j = arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        rng_t* rng = make_rng(j->kind, seed_of(j->seeds, j->seed0, i));
        int rc;
        if (j->sp) {
            rc = htn_structured_precision_mvn(rng, &j->spc, j->out + i * j->spc.pncol);
        } else {
            ht_config_t c = j->ht;
            if (j->independent) {
                c.mean += i * c.gncol;
                c.cov += i * c.gncol * c.gncol;
                c.g += i * c.gnrow * c.gncol;
                c.r += i * c.gnrow;
            }
            rc = htn_hyperplane_truncated_mvn(rng, &c, j->out + i * c.gncol);
        }
        if (j->info)
            j->info[i] = rc;
        rng_free(rng);
    }
    return NULL;
}

static int run(job_t* proto, size_t nsamples, int nthreads, double* seconds)
{
    if (nthreads < 1)
        nthreads = 1;
    if ((size_t)nthreads > nsamples)
        nthreads = (int)(nsamples ? nsamples : 1);
    pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
    job_t* jobs = malloc(sizeof(job_t) * nthreads);
    if (!th || !jobs) {
        free(th);
        free(jobs);
        return -100;
    }
    double t0 = now_s();
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *proto;
        jobs[t].lo = nsamples * t / nthreads;
        jobs[t].hi = nsamples * (t + 1) / nthreads;
        if (nthreads == 1)
            worker(&jobs[t]);
        else
            pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++)
            pthread_join(th[t], NULL);
    if (seconds)
        *seconds = now_s() - t0;
    free(th);
    free(jobs);
    return 0;
}

int refh_hyperplane_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples,
                          int independent, size_t k2, size_t k, const double* mean,
                          const double* cov, const double* g, const double* r, int diag,
                          double* out, int* info, int nthreads, double* seconds)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = kind; j.seeds = seeds; j.seed0 = seed0; j.independent = independent;
    init_ht_config(&j.ht, k2, k, mean, cov, g, r, diag != 0);
    j.out = out; j.info = info;
    return run(&j, nsamples, nthreads, seconds);
}

int refh_structured_batch(int kind, const uint64_t* seeds, uint64_t seed0, size_t nsamples, size_t n,
                          size_t p, const double* mean, const double* a, const double* phi,
                          const double* omega, int struct_mean, int a_id, int o_id, double* out,
                          int* info, int nthreads, double* seconds)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = kind; j.seeds = seeds; j.seed0 = seed0; j.sp = 1;
    init_sp_config(&j.spc, n, p, mean, a, phi, omega, struct_mean != 0, (type_t)a_id, (type_t)o_id);
    j.out = out; j.info = info;
    return run(&j, nsamples, nthreads, seconds);
}

/* struct layout probes, compared by tests with the product's own view of include/htnorm.h */
size_t refh_sizeof_ht_config(void) { return sizeof(ht_config_t); }
size_t refh_sizeof_sp_config(void) { return sizeof(sp_config_t); }
size_t refh_sizeof_rng(void) { return sizeof(rng_t); }
