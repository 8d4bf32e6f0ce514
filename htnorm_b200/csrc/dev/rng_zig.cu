// Generator kernels: raw 64-bit streams and ziggurat normals, one stream per thread.
//
// normal_fill is the Z producer of every sampler.  Layout problem: stream b writes element
// n-1-j of ITS row (reference fill order, src/htnorm_distributions.c:12-13), so a naive
// thread-per-stream store is strided by the row length.  Each warp therefore stages a 32-stream x
// 32-element tile in shared memory and writes it out row by row: 256 contiguous bytes per store.
//
// Ziggurat control flow: the 32 streams of a warp run the table-lookup fast path in lockstep (one
// normal per stream per iteration, conflict-free tile stores); rejected draws (1.5 %) are resolved by the
// warp together in one divergent region per iteration (exp / log only there).  The sequence of generator
// calls inside each stream is exactly the reference's (src/htnorm_distributions.c:18-54).
// [round-1 note] a first version parked rejected lanes to batch the slow path across iterations; the
// catch-up of parked lanes at tile boundaries cost 2.1x the iterations (ncu: 21 of 32 lanes active), so
// the simpler lockstep loop is both faster and smaller.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "launchers.h"
#include "rng.cuh"

namespace {

using namespace htn;

constexpr int kMaxBlocks = 128;      // overlapped launches leave SMs free for the factorisation kernels

template <int GEN>
__global__ void __launch_bounds__(128) raw_fill_kernel(const uint64_t* seeds, uint64_t seed0,
                                                       uint64_t* states, size_t nstreams,
                                                       size_t nwords, uint64_t* out)
{
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nstreams) return;
    GenState g;
    stream_begin<GEN>(g, seeds, seed0, states, b);
    for (size_t t = 0; t < nwords; t++) out[b * nwords + t] = gen_next<GEN>(g);
    stream_end<GEN>(g, states, b);
}

struct Seg {
    size_t n;
    double* out;
    size_t ld;
    const double* div;
    const double* mul;
    const double* add;
};

// kWarpsPerBlock warps per CTA (one CTA per SM), each staging a 32-stream x kTile-element tile
template <int GEN, int kWarpsPerBlock, int kTile>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kWarpsPerBlock >= 16 ? 1 : 16 / kWarpsPerBlock)
normal_fill_kernel(const uint64_t* seeds, uint64_t seed0, uint64_t* states, size_t nstreams, int nseg,
                   Seg seg0, Seg seg1, uint32_t* ndraws)
{
    extern __shared__ double s_dyn[];
    uint64_t* s_ki = reinterpret_cast<uint64_t*>(s_dyn);
    double* s_wi = s_dyn + 256;
    double* s_fi = s_dyn + 512;
    double(*s_tile)[32][kTile + 1] = reinterpret_cast<double(*)[32][kTile + 1]>(s_dyn + 768);
    zig_stage_tables(s_ki, s_wi, s_fi);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // persistent: warp-sized groups of 32 streams are dealt round-robin to the resident warps
    const size_t ngroups = (nstreams + 31) / 32;
    for (size_t grp = (size_t)blockIdx.x * kWarpsPerBlock + warp; grp < ngroups;
         grp += (size_t)gridDim.x * kWarpsPerBlock) {
    const size_t warp_base = grp * 32;
    const size_t b = warp_base + lane;
    const bool live = b < nstreams;
    double(*tile)[kTile + 1] = s_tile[warp];

    GenState g;
    uint32_t ncalls = 0;
    // lanes past the last stream replay the last one (their results are never stored): no per-draw liveness test
    stream_begin<GEN>(g, seeds, seed0, states, live ? b : nstreams - 1);

    for (int sgi = 0; sgi < nseg; sgi++) {
        const Seg sg = sgi == 0 ? seg0 : seg1;
        const bool plain = !sg.div && !sg.mul && !sg.add;
        for (size_t pos0 = 0; pos0 < sg.n; pos0 += kTile) {
            const int cnt = (int)((sg.n - pos0) < (size_t)kTile ? (sg.n - pos0) : (size_t)kTile);
            // Lanes advance in lockstep: iteration t produces the (pos0 + t)-th normal of EVERY stream of
            // the warp.  98.5 % of draws finish on the table-lookup fast path; when any lane's draw fails it,
            // the warp runs the wedge / tail test for those lanes right there (the other lanes idle for
            // that one divergent region) - the stream cannot advance past an unresolved test.
            for (int t = 0; t < cnt; t++) {
                double z;
                for (;;) {
                    ZigDraw d;
                    const uint64_t r = gen_next<GEN>(g);
                    ncalls++;
                    if (zig_fast(r, s_ki, s_wi, d)) { z = d.x; break; }
                    if (zig_slow<GEN>(g, d, z, ncalls, s_fi)) break;
                }
                tile[lane][t] = z;
            }
            __syncwarp();
            // write-out: row rr of the tile = stream warp_base + rr; lane (rsub, t) holds draw pos0 + t of the rows
            // rr = rsub, rsub + 32 / kTile, ...: every store instruction writes 32 / kTile runs of kTile doubles
            constexpr int kRowsPerStore = 32 / kTile;
            const int t = lane % kTile, rsub = lane / kTile;
            if (t < cnt) {
                const size_t i = sg.n - 1 - (pos0 + (size_t)t);
                const int nrows = (int)(nstreams - warp_base < 32 ? nstreams - warp_base : 32);
                double* op = sg.out + (warp_base + rsub) * sg.ld + i;
                const size_t ostep = (size_t)kRowsPerStore * sg.ld;
                if (plain) {  // the common case: 3 instructions per stored element
                    if (nrows == 32) {
#pragma unroll 8
                        for (int rr = rsub; rr < 32; rr += kRowsPerStore, op += ostep) *op = tile[rr][t];
                    } else {
                        for (int rr = rsub; rr < nrows; rr += kRowsPerStore, op += ostep) *op = tile[rr][t];
                    }
                } else {
                    const double dv = sg.div ? sg.div[i] : 1.0;
                    const double mv = sg.mul ? sg.mul[i] : 1.0;
                    const double av = sg.add ? sg.add[i] : 0.0;
                    for (int rr = rsub; rr < nrows; rr += kRowsPerStore, op += ostep) {
                        double v = tile[rr][t];
                        if (sg.div) v = __ddiv_rn(v, dv);
                        if (sg.mul) v = __dmul_rn(v, mv);
                        if (sg.add) v = __dadd_rn(av, v);
                        *op = v;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (live) {
        stream_end<GEN>(g, states, b);
        if (ndraws) ndraws[b] = ncalls;
    }
    __syncwarp();
    }  // groups
}

// ---- warp-cooperative SINGLE-STREAM ziggurat (few streams: structured precision with a small batch, the
// single-sample API) ----
// One warp draws ONE stream: lane l decodes the stream's word pos + l.  PCG64 here is two 64-bit LCGs
// (src/htnorm_pcg32_minimal.h:19-29), so lane l JUMPS AHEAD by l steps with a precomputed (multiplier, increment
// factor) pair: state_l = A_l * state + C_l * inc, A_l = M^l, C_l = 1 + M + ... + M^(l-1).  Xoroshiro128+ has no cheap
// jump: every lane steps the generator 32 times and keeps the word and the state of its own step.
// The 32 candidates run the table-lookup fast path together; a ballot finds the FIRST rejecting lane f (1.5 % of words):
// the normals of lanes < f stand, lane f's wedge / tail test is resolved by the whole warp on the (uniform) state after
// word f - it consumes further words exactly as src/htnorm_distributions.c:38-52 does - and the warp re-bases on the
// state it leaves.  The sequence of generator calls and every accept / reject decision are those of the reference's
// sequential loop (bit-exact, including the per-stream call count).
template <int GEN>
__device__ __forceinline__ void coop_words(const GenState& S, int lane, uint64_t jmul, uint64_t jinc, uint64_t& r, GenState& after)
{
    if (GEN == 0) {
        after.s[1] = S.s[1];
        after.s[3] = S.s[3];
        after.s[0] = jmul * S.s[0] + jinc * S.s[1];   // the state BEFORE word `lane`
        after.s[2] = jmul * S.s[2] + jinc * S.s[3];
        r = gen_next<0>(after);                        // word `lane`; `after` is now the state after it
    } else {
        GenState t = S;
        after = S;
        r = 0;
#pragma unroll 4
        for (int i = 0; i < 32; i++) {
            const uint64_t w = gen_next<1>(t);
            if (i == lane) { r = w; after = t; }
        }
    }
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src)
{
    return ((uint64_t)__shfl_sync(0xffffffffu, (unsigned)(v >> 32), src) << 32) | __shfl_sync(0xffffffffu, (unsigned)v, src);
}

template <int GEN>
__global__ void __launch_bounds__(128)
normal_fill_coop_kernel(const uint64_t* seeds, uint64_t seed0, uint64_t* states, size_t nstreams, int nseg, Seg seg0,
                        Seg seg1, uint32_t* ndraws)
{
    __shared__ uint64_t s_ki[256];
    __shared__ double s_wi[256], s_fi[256];
    zig_stage_tables(s_ki, s_wi, s_fi);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t b = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nstreams) return;
    // jump-ahead pair of this lane: (M^lane, 1 + M + ... + M^(lane-1)) mod 2^64
    uint64_t jmul = 1, jinc = 0;
    if (GEN == 0) {
        for (int i = 0; i < lane; i++) {
            jinc = jinc * 6364136223846793005ULL + 1;
            jmul *= 6364136223846793005ULL;
        }
    }
    GenState S;
    stream_begin<GEN>(S, seeds, seed0, states, b);
    uint32_t ncalls = 0;
    const size_t n0 = seg0.n, total = n0 + (nseg > 1 ? seg1.n : 0);
    size_t produced = 0;
    auto store = [&](size_t j, double z) {  // the j-th normal of the stream (draw order): reverse fill inside its segment
        const bool first_seg = j < n0;
        const Seg& sg = first_seg ? seg0 : seg1;
        const size_t i = sg.n - 1 - (first_seg ? j : j - n0);
        double v = z;
        if (sg.div) v = __ddiv_rn(v, sg.div[i]);
        if (sg.mul) v = __dmul_rn(v, sg.mul[i]);
        if (sg.add) v = __dadd_rn(sg.add[i], v);
        sg.out[b * sg.ld + i] = v;
    };
    while (produced < total) {
        uint64_t r;
        GenState after;
        coop_words<GEN>(S, lane, jmul, jinc, r, after);
        ZigDraw d;
        const bool ok = zig_fast(r, s_ki, s_wi, d);
        const unsigned rej = __ballot_sync(0xffffffffu, !ok);
        const int first = rej ? __ffs(rej) - 1 : 32;
        const size_t remaining = total - produced;
        const int m = (size_t)first < remaining ? first : (int)remaining;
        if (lane < m) store(produced + lane, d.x);
        if ((size_t)first >= remaining || first == 32) {
            // no rejecting word among the m >= 1 words used (all 32 accepted, or the run ends before the first rejection)
            S.s[0] = shfl_u64(after.s[0], m - 1);
            if (GEN == 0) S.s[2] = shfl_u64(after.s[2], m - 1);
            else S.s[1] = shfl_u64(after.s[1], m - 1);
            ncalls += (uint32_t)m;
            produced += (size_t)m;
            continue;
        }
        // word `first` failed the fast test: resolve it on the state after that word, every lane alike (uniform)
        S.s[0] = shfl_u64(after.s[0], first);
        if (GEN == 0) S.s[2] = shfl_u64(after.s[2], first);
        else S.s[1] = shfl_u64(after.s[1], first);
        const uint64_t rf = shfl_u64(r, first);
        ZigDraw df;
        zig_fast(rf, s_ki, s_wi, df);
        ncalls += (uint32_t)first + 1;
        double z = 0.0;
        const bool got = zig_slow<GEN>(S, df, z, ncalls, s_fi);
        if (got) {
            if (lane == 0) store(produced + (size_t)first, z);
            produced += (size_t)first + 1;
        } else {
            produced += (size_t)first;
        }
    }
    if (lane == 0) {
        stream_end<GEN>(S, states, b);
        if (ndraws) ndraws[b] = ncalls;
    }
}

// normals drawn elsewhere (foreign generator, host side): zpre[b * total + off + j] is the j-th draw of
// stream b; same placement and transforms as normal_fill_kernel's write-out
__global__ void __launch_bounds__(256)
prefilled_kernel(const double* __restrict__ zpre, size_t nstreams, size_t total, size_t off, Seg sg)
{
    const size_t n = sg.n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nstreams * n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / n, j = idx - b * n;
        const size_t i = n - 1 - j;
        double v = zpre[b * total + off + j];
        if (sg.div) v = __ddiv_rn(v, sg.div[i]);
        if (sg.mul) v = __dmul_rn(v, sg.mul[i]);
        if (sg.add) v = __dadd_rn(sg.add[i], v);
        sg.out[b * sg.ld + i] = v;
    }
}

}  // namespace

static unsigned long long g_launches = 0;
extern "C" void htnk_count_launch(void) { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
extern "C" unsigned long long htnk_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

static int launch_status(void)
{
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_raw_fill(cudaStream_t st, const htnk_streams_t* s, size_t nstreams, size_t nwords,
                             uint64_t* out)
{
    if (nstreams == 0 || nwords == 0) return 0;
    unsigned grid = (unsigned)((nstreams + 127) / 128);
    if (s->gen == 0)
        raw_fill_kernel<0><<<grid, 128, 0, st>>>(s->seeds, s->seed0, s->states, nstreams, nwords, out);
    else
        raw_fill_kernel<1><<<grid, 128, 0, st>>>(s->seeds, s->seed0, s->states, nstreams, nwords, out);
    return launch_status();
}

// one-shot hint for the NEXT htnk_normal_fill of this thread: it runs beside other kernels (side stream), so keep
// its persistent CTAs to kMaxBlocks SMs; without the hint the generator takes every SM of the device
static thread_local bool g_shared_next = false;
extern "C" void htnk_normal_fill_shares_device_next(void) { g_shared_next = true; }

static int sm_count(void)
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kMaxBlocks;
    }
    return n;
}

extern "C" int htnk_normal_fill(cudaStream_t st, const htnk_streams_t* s, size_t nstreams, int nseg,
                                const htnk_segment_t* seg, uint32_t* ndraws)
{
    const bool shared = g_shared_next;
    g_shared_next = false;
    if (nstreams == 0 || nseg <= 0) return 0;
    if (s->zpre) {  // pre-drawn normals: scatter only
        size_t total = 0, off = 0;
        for (int i = 0; i < nseg; i++) total += seg[i].n;
        for (int i = 0; i < nseg; i++) {
            if (seg[i].n == 0) continue;
            Seg sg{seg[i].n, seg[i].out, seg[i].ld, seg[i].div, seg[i].mul, seg[i].add};
            size_t work = nstreams * seg[i].n;
            unsigned grid = (unsigned)((work + 255) / 256 < 4096 ? (work + 255) / 256 : 4096);
            prefilled_kernel<<<grid, 256, 0, st>>>(s->zpre, nstreams, total, off, sg);
            if (launch_status()) return -201;
            off += seg[i].n;
        }
        return 0;
    }
    Seg a[2] = {{0, nullptr, 0, nullptr, nullptr, nullptr}, {0, nullptr, 0, nullptr, nullptr, nullptr}};
    for (int i = 0; i < nseg && i < 2; i++)
        a[i] = Seg{seg[i].n, seg[i].out, seg[i].ld, seg[i].div, seg[i].mul, seg[i].add};
    // few streams: one WARP per stream (jump-ahead / ballot / re-base), else one thread per stream.  HTN_RNG_COOP: the
    // largest stream count that takes the warp-cooperative kernel (0 = never)
    // (measured, B200: PCG64 256 x 9216 normals 1560 -> 166 us, 4096 x 2500 442 -> 112 us, 1 x 1000 94 -> 25 us, break-even
    // near 16 k streams; Xoroshiro128+ - 32 sequential steps per round - breaks even near 3 k streams)
    static long coop_max = -1;
    if (coop_max < 0) {
        const char* e = getenv("HTN_RNG_COOP");
        coop_max = e ? atol(e) : -2;
    }
    const long coop_lim = coop_max >= 0 ? coop_max : (s->gen == 0 ? 12288 : 2048);
    if ((long)nstreams <= coop_lim) {
        const unsigned grid = (unsigned)((nstreams + 3) / 4);
        if (s->gen == 0)
            normal_fill_coop_kernel<0><<<grid, 128, 0, st>>>(s->seeds, s->seed0, s->states, nstreams, nseg, a[0], a[1], ndraws);
        else
            normal_fill_coop_kernel<1><<<grid, 128, 0, st>>>(s->seeds, s->seed0, s->states, nstreams, nseg, a[0], a[1], ndraws);
        return launch_status();
    }
    static int cfg = -1;  // HTN_RNG_CFG (experiments): 0 = 16 warps x 32-wide tiles, 1 = 24 x 32, 2 = 32 x 16
    if (cfg < 0) {
        const char* e = getenv("HTN_RNG_CFG");
        cfg = e ? atoi(e) : 0;
    }
    static int shared_blocks = 0;  // HTN_RNG_SHARED_SMS (experiments): SMs the generator takes when it overlaps other kernels
    if (shared_blocks == 0) {
        const char* e = getenv("HTN_RNG_SHARED_SMS");
        shared_blocks = e && atoi(e) > 0 ? atoi(e) : kMaxBlocks;
    }
    const int max_blocks = shared ? shared_blocks : sm_count();
    auto go = [&](auto kern, int warps, int tile) {
        const size_t per_block = (size_t)warps * 32;
        size_t nblocks = (nstreams + per_block - 1) / per_block;
        const size_t cap = (size_t)max_blocks * (warps >= 16 ? 1 : 16 / warps);
        unsigned grid = (unsigned)(nblocks < cap ? nblocks : cap);
        const size_t smem = (768 + (size_t)warps * 32 * (tile + 1)) * sizeof(double);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -201;
        kern<<<grid, warps * 32, smem, st>>>(s->seeds, s->seed0, s->states, nstreams, nseg, a[0], a[1], ndraws);
        return launch_status();
    };
    // Beside the factorisation (side stream): pack the streams into 24-warp CTAs.  The generator's footprint is its
    // registers (one thread per stream for the whole launch), and an SM it sits on is lost to the panel-update /
    // panel-solve CTAs of the factorisation; 768-thread CTAs fill 86 SMs completely and leave 62 SMs entirely
    // free, instead of taking three quarters of 128 SMs (k = 1000 headline: setup 0.61 -> 0.53 ms).
    if (cfg == 1 || (cfg == 0 && shared && nstreams >= (size_t)kMaxBlocks * 32 * 16))
        return s->gen == 0 ? go(normal_fill_kernel<0, 24, 32>, 24, 32) : go(normal_fill_kernel<1, 24, 32>, 24, 32);
    if (cfg == 2) return s->gen == 0 ? go(normal_fill_kernel<0, 32, 16>, 32, 16) : go(normal_fill_kernel<1, 32, 16>, 32, 16);
    // few streams: small CTAs so that every SM still gets warps (a stream is one thread; nothing else to split)
    if (nstreams <= (size_t)kMaxBlocks * 32 * 4)
        return s->gen == 0 ? go(normal_fill_kernel<0, 2, 32>, 2, 32) : go(normal_fill_kernel<1, 2, 32>, 2, 32);
    if (nstreams < (size_t)kMaxBlocks * 32 * 16)
        return s->gen == 0 ? go(normal_fill_kernel<0, 4, 32>, 4, 32) : go(normal_fill_kernel<1, 4, 32>, 4, 32);
    return s->gen == 0 ? go(normal_fill_kernel<0, 16, 32>, 16, 32) : go(normal_fill_kernel<1, 16, 32>, 16, 32);
}
