/* Bit-generator objects of the public API (include/htnorm_rng.h).
 *
 * The constructors hand out rng_t objects whose function pointers are host implementations of the two
 * generators (so foreign code can call rng->next_uint64 / next_double exactly as with the reference,
 * src/htnorm_rng.c:20-67), and whose `base` is the same state layout the reference uses:
 *   PCG64   : {gen0.state, gen0.inc, gen1.state, gen1.inc}   (src/htnorm_pcg64.h:10)
 *   XRS128P : {s0, s1}                                       (src/htnorm_xoroshiro128p.h:14)
 * The samplers recognise these objects by function-pointer identity, upload the state, advance it on
 * the GPU (csrc/dev/rng.cuh) and write it back - the host functions below are never on the sampling path.
 */
#include "htn_internal.h"

#include <stdlib.h>
#include <time.h>

typedef struct { uint64_t s[4]; } pcg64_state_t;
typedef struct { uint64_t s[2]; } xrs_state_t;

/* Vigna's SplitMix64 seed expander (reference src/htnorm_splitmix64.h:17-24) */
static uint64_t splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* O'Neill's PCG32 XSH-RR 64/32 output + LCG step (reference src/htnorm_pcg32_minimal.h:19-29) */
static uint32_t pcg32_next(uint64_t* state, uint64_t inc)
{
    const uint64_t old = *state;
    *state = old * 6364136223846793005ULL + (inc | 1u);
    const uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    const uint32_t rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((32u - rot) & 31u));
}

static uint64_t pcg64_u64(void* base)
{
    pcg64_state_t* p = base;
    const uint64_t hi = pcg32_next(&p->s[0], p->s[1]);
    const uint64_t lo = pcg32_next(&p->s[2], p->s[3]);
    return (hi << 32) | lo;
}
static double pcg64_dbl(void* base) { return (double)(pcg64_u64(base) >> 11) * (1.0 / 9007199254740992.0); }

static uint64_t xrs_u64(void* base)
{
    xrs_state_t* x = base;
    const uint64_t s0 = x->s[0];
    uint64_t s1 = x->s[1];
    const uint64_t out = s0 + s1;
    s1 ^= s0;
    x->s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
    x->s[1] = (s1 << 37) | (s1 >> 27);
    return out;
}
static double xrs_dbl(void* base) { return (double)(xrs_u64(base) >> 11) * (1.0 / 9007199254740992.0); }

int htn_rng_kind(const rng_t* rng)
{
    if (!rng || !rng->base) return -1;
    if (rng->next_uint64 == pcg64_u64 && rng->next_double == pcg64_dbl) return HTN_GEN_PCG64;
    if (rng->next_uint64 == xrs_u64 && rng->next_double == xrs_dbl) return HTN_GEN_XRS128P;
    return -1;
}

void rng_free(rng_t* rng)
{
    if (!rng) return;
    free(rng->base);
    free(rng);
}

rng_t* rng_pcg64_new_seeded(uint64_t seed)
{
    rng_t* rng = malloc(sizeof(*rng));
    pcg64_state_t* st = malloc(sizeof(*st));
    if (!rng || !st) { /* unlike the reference never return a half-initialised object */
        free(rng);
        free(st);
        return NULL;
    }
    for (int i = 0; i < 4; i++) st->s[i] = splitmix64(&seed);
    rng->base = st;
    rng->next_uint64 = pcg64_u64;
    rng->next_double = pcg64_dbl;
    return rng;
}

rng_t* rng_pcg64_new(void) { return rng_pcg64_new_seeded((uint64_t)time(NULL)); }

rng_t* rng_xrs128p_new_seeded(uint64_t seed)
{
    rng_t* rng = malloc(sizeof(*rng));
    xrs_state_t* st = malloc(sizeof(*st));
    if (!rng || !st) {
        free(rng);
        free(st);
        return NULL;
    }
    st->s[0] = splitmix64(&seed);
    st->s[1] = splitmix64(&seed);
    rng->base = st;
    rng->next_uint64 = xrs_u64;
    rng->next_double = xrs_dbl;
    return rng;
}

rng_t* rng_xrs128p_new(void) { return rng_xrs128p_new_seeded((uint64_t)time(NULL)); }

void init_ht_config(ht_config_t* conf, size_t gnrow, size_t gncol, const double* mean, const double* cov,
                    const double* g, const double* r, bool diag)
{
    conf->gnrow = gnrow;
    conf->gncol = gncol;
    conf->mean = mean;
    conf->cov = cov;
    conf->g = g;
    conf->r = r;
    conf->diag = diag;
}

void init_sp_config(sp_config_t* conf, size_t pnrow, size_t pncol, const double* mean, const double* a,
                    const double* phi, const double* omega, bool struct_mean, type_t a_id, type_t o_id)
{
    conf->a_id = a_id;
    conf->o_id = o_id;
    conf->pnrow = pnrow;
    conf->pncol = pncol;
    conf->mean = mean;
    conf->a = a;
    conf->phi = phi;
    conf->omega = omega;
    conf->struct_mean = struct_mean;
}
