/* Foreign generators.  An rng_t whose function pointers are not this library's (e.g. NumPy's
 * bit generator handed over by the reference's Cython wrapper, pyhtnorm/_htnorm.pyx:116-122) can only be
 * advanced on the host: its callbacks are host code.  For such generators - and only for them - the
 * normals of ONE draw are produced here, through the caller's callbacks, with exactly the reference's
 * ziggurat (src/htnorm_distributions.c:18-54), then shipped to the device where all the algebra runs.
 * This is an RNG-only host step; the library's own generators never come here. */
#include "htn_internal.h"

#include <math.h>
#include <string.h>

static const uint64_t zig_words[768] = {
#include "../common/zig_tables_u64.inc"
};

static inline double word_as_double(uint64_t w)
{
    double d;
    memcpy(&d, &w, 8);
    return d;
}

static double foreign_std_normal(rng_t* rng)
{
    static const double zr = 3.6541528853610087963519472518, zinv = 0.27366123732975827203338247596;
    for (;;) {
        uint64_t r = rng->next_uint64(rng->base);
        const int idx = (int)(r & 0xff);
        r >>= 8;
        const int sign = (int)(r & 1);
        const uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * word_as_double(zig_words[256 + idx]);
        if (sign) x = -x;
        if (rabs < zig_words[idx]) return x;
        if (idx == 0) {
            for (;;) {
                const double xx = -zinv * log(1.0 - rng->next_double(rng->base));
                const double yy = -log(1.0 - rng->next_double(rng->base));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -(zr + xx) : zr + xx;
            }
        } else {
            const double f1 = word_as_double(zig_words[512 + idx - 1]), f0 = word_as_double(zig_words[512 + idx]);
            if ((f1 - f0) * rng->next_double(rng->base) + f0 < exp(-0.5 * x * x)) return x;
        }
    }
}

/* n normals in DRAW order (the device scatters them back to front like std_normal_rand_fill) */
void htn_foreign_draw(rng_t* rng, size_t n, double* z)
{
    for (size_t j = 0; j < n; j++) z[j] = foreign_std_normal(rng);
}
