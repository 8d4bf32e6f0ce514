// Device-side bit generators and the 256-layer ziggurat (sm_100a).
//
// What is reproduced, bit for bit (reference paths relative to /root/reference):
//   SplitMix64 seeding          src/htnorm_splitmix64.h:17-24, src/htnorm_rng.c:20-35, 45-60
//   "PCG64" = 2 x PCG32 XSH-RR  src/htnorm_pcg32_minimal.h:19-29, src/htnorm_pcg64.h:23-37
//   Xoroshiro128+ (24,16,37)    src/htnorm_xoroshiro128p.h:24-45
//   ziggurat N(0,1)             src/htnorm_distributions.c:18-54 (tables: common/zig_tables_u64.inc)
//
// Floating point inside the ziggurat uses explicit round-to-nearest intrinsics (__dmul_rn/__dadd_rn)
// so nvcc cannot contract a*b+c into an FMA: x86-64 gcc -O2 emits none, and the accept/reject
// comparisons must see the same doubles.
#pragma once
#include <stdint.h>

namespace htn {

struct GenState {
    uint64_t s[4];  // PCG64: gen0.state, gen0.inc, gen1.state, gen1.inc ; XRS128P: s0, s1, -, -
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

template <int GEN>
__device__ __forceinline__ void gen_seed(GenState& g, uint64_t seed)
{
    g.s[0] = splitmix64(seed);
    g.s[1] = splitmix64(seed);
    if (GEN == 0) {
        g.s[2] = splitmix64(seed);
        g.s[3] = splitmix64(seed);
        // the LCG increment is used as (inc | 1) on every step: fold it in once
        g.s[1] |= 1ULL;
        g.s[3] |= 1ULL;
    } else {
        g.s[2] = g.s[3] = 0;
    }
}

__device__ __forceinline__ uint32_t pcg32_out(uint64_t old)
{
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return __funnelshift_r(xs, xs, rot);  // rotate right by rot (mod 32)
}

template <int GEN>
__device__ __forceinline__ uint64_t gen_next(GenState& g)
{
    if (GEN == 0) {
        uint64_t o0 = g.s[0], o1 = g.s[2];
        g.s[0] = o0 * 6364136223846793005ULL + g.s[1];
        g.s[2] = o1 * 6364136223846793005ULL + g.s[3];
        return ((uint64_t)pcg32_out(o0) << 32) | (uint64_t)pcg32_out(o1);
    } else {
        uint64_t s0 = g.s[0], s1 = g.s[1];
        uint64_t res = s0 + s1;
        s1 ^= s0;
        g.s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
        g.s[1] = (s1 << 37) | (s1 >> 27);
        return res;
    }
}

__device__ __forceinline__ double u64_to_unit(uint64_t u)
{
    return __dmul_rn((double)(u >> 11), 1.0 / 9007199254740992.0);
}

// ---- ziggurat ----
// tables live in global memory (768 words: ki | bits(wi) | bits(fi)); kernels stage ki/wi into
// shared memory (random per-lane index: constant memory would serialise), fi is read on the slow path.
static __device__ const uint64_t zig_table[768] = {
#include "../common/zig_tables_u64.inc"
};

// copy ki (u64), wi (f64) and fi (f64) into shared memory; call from all threads, then __syncthreads()
__device__ __forceinline__ void zig_stage_tables(uint64_t* s_ki, double* s_wi, double* s_fi)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_ki[i] = zig_table[i];
        s_wi[i] = __longlong_as_double((long long)zig_table[256 + i]);
        s_fi[i] = __longlong_as_double((long long)zig_table[512 + i]);
    }
}

#define HTN_ZIG_R 3.6541528853610087963519472518
#define HTN_ZIG_INV_R 0.27366123732975827203338247596

struct ZigDraw {
    uint64_t rabs;
    double x;
    int idx;
};

// decode one raw word; returns true when the fast path accepts (src/htnorm_distributions.c:26-36)
__device__ __forceinline__ bool zig_fast(uint64_t r, const uint64_t* __restrict__ s_ki,
                                         const double* __restrict__ s_wi, ZigDraw& d)
{
    d.idx = (int)(r & 0xff);
    r >>= 8;
    uint64_t sign = r & 1;
    d.rabs = (r >> 1) & 0x000fffffffffffffULL;
    // rabs < 2^52: (2^52 + rabs) is exactly representable, so the integer -> double conversion is one
    // OR + one subtraction instead of the multi-instruction I2F.F64.U64
    const double rd = __dadd_rn(__longlong_as_double((long long)(d.rabs | 0x4330000000000000ULL)),
                                -4503599627370496.0);
    const double x = __dmul_rn(rd, s_wi[d.idx]);
    // sign ? -x : x  as one integer XOR on the high word (identical bits, -0.0 included)
    d.x = __hiloint2double(__double2hiint(x) ^ (int)((uint32_t)sign << 31), __double2loint(x));
    return d.rabs < s_ki[d.idx];
}

// slow path for a rejected fast draw (src/htnorm_distributions.c:38-52).  Returns true and sets z
// when a normal is produced; false when the wedge test rejects (caller draws a fresh word).
template <int GEN>
__device__ __forceinline__ bool zig_slow(GenState& g, const ZigDraw& d, double& z, uint32_t& ncalls,
                                         const double* __restrict__ s_fi = nullptr)
{
    if (d.idx == 0) {
        for (;;) {
            double u1 = u64_to_unit(gen_next<GEN>(g));
            double u2 = u64_to_unit(gen_next<GEN>(g));
            ncalls += 2;
            double xx = __dmul_rn(-HTN_ZIG_INV_R, log(__dadd_rn(1.0, -u1)));
            double yy = -log(__dadd_rn(1.0, -u2));
            if (__dadd_rn(yy, yy) > __dmul_rn(xx, xx)) {
                double t = __dadd_rn(HTN_ZIG_R, xx);
                z = ((d.rabs >> 8) & 1) ? -t : t;
                return true;
            }
        }
    }
    double f1 = s_fi ? s_fi[d.idx - 1] : __longlong_as_double((long long)zig_table[512 + d.idx - 1]);
    double f0 = s_fi ? s_fi[d.idx] : __longlong_as_double((long long)zig_table[512 + d.idx]);
    double u = u64_to_unit(gen_next<GEN>(g));
    ncalls += 1;
    const double lhs = __dadd_rn(__dmul_rn(__dadd_rn(f1, -f0), u), f0);
    const double arg = __dmul_rn(__dmul_rn(-0.5, d.x), d.x);
    // The decision is lhs < exp(arg).  A single-precision exp2 brackets exp(arg) to ~1e-6 relative
    // (|arg| < 6.7: 4e-7 from rounding arg * log2(e) to float, 2.4e-7 from ex2.approx); outside a 1e-5 band the
    // comparison is already decided and the 45-instruction double exp() is skipped (it runs for ~0.2 % of the
    // wedge tests).  Inside the band the full-precision test below decides, exactly as before.
    const float e32 = exp2f((float)(arg * 1.4426950408889634));
    const double lo = (double)e32 * (1.0 - 1e-5), hi = (double)e32 * (1.0 + 1e-5);
    bool accept;
    if (lhs < lo) accept = true;
    else if (lhs > hi) accept = false;
    else accept = lhs < exp(arg);
    if (accept) z = d.x;
    return accept;
}

// load the initial state of stream b
template <int GEN>
__device__ __forceinline__ void stream_begin(GenState& g, const uint64_t* seeds, uint64_t seed0,
                                             const uint64_t* states, size_t b)
{
    if (states) {
        g.s[0] = states[4 * b + 0];
        g.s[1] = states[4 * b + 1];
        g.s[2] = states[4 * b + 2];
        g.s[3] = states[4 * b + 3];
        if (GEN == 0) {  // (inc | 1) folded; harmless for the write-back only if we restore - see stream_end
            g.s[1] |= 1ULL;
            g.s[3] |= 1ULL;
        }
    } else {
        gen_seed<GEN>(g, seeds ? seeds[b] : seed0 + (uint64_t)b);
    }
}

// write the advanced state back.  The PCG increments are never modified by the reference
// (src/htnorm_pcg32_minimal.h:24 only reads inc), so they are left untouched in `states`.
template <int GEN>
__device__ __forceinline__ void stream_end(const GenState& g, uint64_t* states, size_t b)
{
    if (!states) return;
    if (GEN == 0) {
        states[4 * b + 0] = g.s[0];
        states[4 * b + 2] = g.s[2];
    } else {
        states[4 * b + 0] = g.s[0];
        states[4 * b + 1] = g.s[1];
    }
}

}  // namespace htn
