/* htnorm-b200: bit-generator interface of the drop-in boundary.
 *
 * Source-compatible with the reference's include/htnorm_rng.h:18-48 (same type name, same field
 * order and types, same five exported functions), written from scratch for this library.
 *
 *   rng_t                      <-> reference include/htnorm_rng.h:18-25
 *   rng_free                   <-> reference include/htnorm_rng.h:28,  src/htnorm_rng.c:12-17
 *   rng_pcg64_new[_seeded]     <-> reference include/htnorm_rng.h:41,43, src/htnorm_rng.c:20-42
 *   rng_xrs128p_new[_seeded]   <-> reference include/htnorm_rng.h:46,48, src/htnorm_rng.c:45-67
 *
 * "PCG64" here is the reference's generator: two PCG32 (XSH-RR 64/32) streams glued hi|lo
 * (src/htnorm_pcg64.h:23-28), seeded by four consecutive SplitMix64 outputs; it is NOT NumPy's
 * 128-bit PCG64. Generators created by the constructors below are recognised by the samplers
 * (function-pointer identity) and advanced ON THE GPU; their state is written back to `base`.
 */
#ifndef HTNORM_RNG_H
#define HTNORM_RNG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    /* opaque generator state (32 bytes for PCG64, 16 bytes for Xoroshiro128+) */
    void* base;
    /* next raw 64-bit word of the stream */
    uint64_t (*next_uint64)(void* base);
    /* next double in [0, 1): (next_uint64 >> 11) * 2^-53 */
    double (*next_double)(void* base);
} rng_t;

/* Release a generator returned by one of the constructors below. NULL is ignored. */
void rng_free(rng_t* rng);

/* PCG64 (2 x PCG32), seeded from time(NULL) / from `seed` through SplitMix64. NULL on allocation failure. */
rng_t* rng_pcg64_new(void);
rng_t* rng_pcg64_new_seeded(uint64_t seed);

/* Xoroshiro128+ (24,16,37), seeded from time(NULL) / from `seed` through SplitMix64. */
rng_t* rng_xrs128p_new(void);
rng_t* rng_xrs128p_new_seeded(uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
