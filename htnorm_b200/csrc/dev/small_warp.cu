// Many small independent problems, round-2 design: ONE WARP PER PROBLEM, the trailing matrix in REGISTERS, everything
// warp-synchronous, every address a compile-time constant (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems).
//
// Why (profiles/r1_small_batched_v7_summary.txt, profiles/r2_small_warp_*): the round-1 kernel (small_batched.cu: one
// 64-thread CTA per problem, factor warp + update warp, generic k) spent 12.8 k warp-instructions per problem - a quarter
// of them address arithmetic, 42 block barriers - and 2.9 k cycles of the SM's FP64 datapath on padded or redundant
// work; it reached 22 % of the HBM roofline.  A first warp-per-problem version that kept the trailing matrix in shared
// memory turned out bound by the shared-memory DATA PIPE (3.3 k wavefronts per problem, 66 % of its peak: 8-byte
// asynchronous copies scattered with bank conflicts, an LDS + STS round trip per tile update).  Here:
//   * the whole of src/htnorm.c:85-187 + mvn_rand_cov (src/htnorm_distributions.c:58-88) for one problem runs in one
//     warp: no block barrier, no role hand-off, 9 independent problems per SM;
//   * cov goes straight from HBM into REGISTERS as the 36 upper-triangle 8 x 8 tiles in the tensor-core accumulator
//     layout (one 16-byte load per tile and lane: the upper triangle of the row-major matrix is what the reference
//     reads) and the rank-8 trailing updates (DMMA m8n8k4) never touch memory;
//   * k is a template parameter (k = 8 NT): loops are fully unrolled, every shared-memory offset is an immediate;
//   * cov = U~^T D U~ (square-root free, U~ unit upper) on 8-row panels.  A finished panel is written to shared memory
//     once (it is needed there as tensor-core fragments anyway) and read back in the LANE = COLUMN layout: lane j holds
//     column j of the panel, the 8 pivots are eliminated by row operations whose only communication is a warp shuffle
//     of the pivot row - no 8 x 8 factor replicated in all lanes, no explicit inverse, no tensor-core panel solve, no
//     rsqrt on the chain (the update of a pivot uses 1 / d: reciprocal seed + two Newton steps);
//   * look-ahead inside the single instruction stream: the update of step p is split into the tile row the next
//     elimination needs and the rest, which is issued a few tiles at a time between the pivots of that elimination;
//   * 8 x 8 tiles are stored dense (64 doubles) with an XOR swizzle of the rows and a row swap in odd tile columns:
//     accumulator, fragment, transposed-fragment and lane = column accesses are all bank-conflict free without padding;
//   * the projection is reformulated around the factor so that cov G^T (2 k^2 k2 flops) is never formed:
//         V = sqrt(D) U~ G^T,  G cov G^T = V^T V,  G y = G mean + V^T z,  x = mean + U~^T sqrt(D) (z + V alpha)
//     (same algebra as the shared-covariance path's W = G L, hyperplane.c) - V on the tensor cores, x by column dot products.
// Error codes follow the reference: cov not PD -> pivot index, out untouched; G cov G^T not PD -> its
// pivot index and the unprojected y.  Handles k = 8, 16, ..., 64 (one instantiation per NT = k / 8; cfg2 is NT = 8), k2 <= 4,
// dense row-major cov, 16-byte aligned; every other shape keeps the generic kernel (small_batched.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "launchers.h"

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gsrc));
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double h = 0.5 * y;
        const double e = fma(-x * y, h, 0.5);
        y = fma(y, e, y);
    }
    return y;
}
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}

// Shared-memory map of one problem (doubles).  Tile (p, tj), p <= tj, of the factor is a dense block of 64 doubles at
// 64 * tid(p, tj); its element (r, c) sits at pr * 8 + (c ^ sw(pr)) with pr = r ^ (tj & 1) and sw(pr) = 4 * ((pr >> 1) & 1).
template <int NT>
struct Geo {
    static constexpr int K = 8 * NT;
    static constexpr int CS = (K + 31) / 32;  // columns per lane
    static __host__ __device__ constexpr int tid(int p, int tj) { return p * NT - p * (p - 1) / 2 + (tj - p); }
    static constexpr int NTILES = NT * (NT + 1) / 2;
    // position of tile (ti, tj) among the deferred tiles of step q (tile rows q + 2 .. NT - 1, row by row)
    static __host__ __device__ constexpr int didx(int q, int ti, int tj)
    {
        return (ti - q - 2) * NT - ((ti - 1) * ti - (q + 1) * (q + 2)) / 2 + (tj - ti);
    }
    static constexpr int GM = 64 * NTILES;  // [K][8]: G^T (cols 0-3) | mean (4) | z (5) | 0 0, later V in cols 0-3
    static constexpr int WV = GM + 8 * K;   // [K]    d, then sqrt(d), then sqrt(d) w
    static constexpr int SCR = WV + K;      // [2][64] two 8 x 8 result tiles
    static constexpr int TOTAL = SCR + 128;
};

// PROF (HTN_SMALL_PROF=1, development only): lane 0 of every warp adds the clock64() span of each phase to prof[]
#define PHASE(i)                                                                                              \
    if (PROF) {                                                                                               \
        const long long t_now = clock64();                                                                    \
        if (lane == 0) atomicAdd((unsigned long long*)&prof[i], (unsigned long long)(t_now - t_prev));        \
        t_prev = t_now;                                                                                       \
    }

template <int NT, bool PROF>
__global__ void __launch_bounds__(32, 8)
ht_warp_kernel(size_t nproblems, int k2, const double* __restrict__ mean, const double* __restrict__ cov,
               const double* __restrict__ gm, const double* __restrict__ rv, const double* __restrict__ zs,
               double* __restrict__ out, int* __restrict__ info, unsigned* __restrict__ next, long long* prof)
{
    using G = Geo<NT>;
    constexpr int K = G::K, CS = G::CS;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);

    // ---- lane constants (offsets in doubles inside a tile), for even [0] and odd [1] tile columns ----
    const int swg = 4 * ((g >> 1) & 1), swt = 4 * ((t4 >> 1) & 1);
    int cf[2], tf[2], rf0[2], rf1[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        cf[e] = (g ^ e) * 8 + ((2 * t4) ^ swg);        // accumulator pattern: (row g, columns 2 t4, 2 t4 + 1), 16 bytes
        tf[e] = (t4 ^ e) * 8 + (g ^ swt);              // transposed fragment: (row t4 + 4 s, column g); s = 1: + 32
        rf0[e] = (g ^ e) * 8 + (t4 ^ swg);             // row fragment: (row g, column t4 + 4 s)
        rf1[e] = (g ^ e) * 8 + ((t4 + 4) ^ swg);
    }
    // lane = column layout: column j = lane + 32 cs lives in tile column tj = j >> 3 (parity ej), column cj = j & 7.
    // Row c of that tile: (c ^ ej) * 8 + (cj ^ sw(c)); four lane constants per column set cover the eight rows:
    // colA[cs][2 * ((c >> 1) & 1) + (c & 1)] + 8 * (c & ~1)... kept simple: one offset per row, 8 registers per set.
    int col[CS][8];
#pragma unroll
    for (int cs = 0; cs < CS; cs++) {
        const int j = lane + 32 * cs, ej = (j >> 3) & 1, cj = j & 7;
#pragma unroll
        for (int c = 0; c < 8; c++) col[cs][c] = 64 * (j >> 3) + (c ^ ej) * 8 + (cj ^ (4 * ((c >> 1) & 1)));
    }
    // GM (row = variable i, 8 columns, swizzled by the row): lane = row i = lane + 32 rs: columns 0-3 at lo, 4-7 at hi
    int row_lo[CS], row_hi[CS];
#pragma unroll
    for (int rs = 0; rs < CS; rs++) {
        const int i = lane + 32 * rs;
        const int sw = 4 * ((i >> 1) & 1);
        row_lo[rs] = i * 8 + sw;
        row_hi[rs] = i * 8 + (4 - sw);
    }
    const int gtf = t4 * 8 + (g ^ swt);  // transposed fragment of GM rows (no tile parity there)
    const int gcf = g * 8 + ((2 * t4) ^ swg);
    // zero the region once: GM columns that are never written must read as zero for the tensor-core passes
    for (int idx = lane; idx < G::TOTAL; idx += 32) sm[idx] = 0.0;
    __syncwarp();

    long long t_prev = PROF ? clock64() : 0;
    // Dynamic distribution (a static deal would leave the less loaded scheduler partitions idle at the end), one problem
    // AHEAD: the index of the next problem is drawn while the current one is being loaded, so that the atomic's round trip
    // is never waited for and the next problem's covariance can be prefetched into L2 behind the current one's loads.
    unsigned nxt = 0;
    if (lane == 0) nxt = atomicAdd(next, 1u);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    for (;;) {
        const unsigned pbu = nxt;
        if ((size_t)pbu >= nproblems) break;
        const size_t pb = pbu;
        if (lane == 0) nxt = atomicAdd(next, 1u);  // consumed (shuffled) only after this problem's loads are in flight
        const double* covp = cov + pb * (size_t)K * K;
        const double* gp = gm + pb * (size_t)k2 * K;
        // ---- cov: the 36 upper tiles straight into registers, accumulator layout, 16 bytes per tile and lane ----
        double2 c[NT][NT];
        {
            const double* src = covp + g * K + 2 * t4;
#pragma unroll
            for (int ti = 0; ti < NT; ti++)
#pragma unroll
                for (int tj = ti; tj < NT; tj++) c[ti][tj] = __ldg(reinterpret_cast<const double2*>(src + 8 * ti * K + 8 * tj));
        }
        // G^T -> GM columns 0..k2-1, mean -> column 4, z -> column 5 (lane = row), asynchronous 8-byte copies
        double mreg[CS];
#pragma unroll
        for (int rs = 0; rs < CS; rs++) {
            const int i = lane + 32 * rs;
            mreg[rs] = 0.0;
            if (i < K) {
                const unsigned rowb = sbase + 8u * (unsigned)G::GM;
#pragma unroll
                for (int cc = 0; cc < 4; cc++)
                    if (cc < k2) cp_async8(rowb + 8u * (unsigned)(row_lo[rs] + cc), gp + (size_t)cc * K + i);
                cp_async8(rowb + 8u * (unsigned)row_hi[rs], mean + pb * (size_t)K + i);        // column 4
                cp_async8(rowb + 8u * (unsigned)(row_hi[rs] + 1), zs + pb * (size_t)K + i);    // column 5
                mreg[rs] = mean[pb * (size_t)K + i];
            }
        }
        double rreg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int cc = 0; cc < 4; cc++)
            if (cc < k2) rreg[cc] = rv[pb * (size_t)k2 + cc];
        asm volatile("cp.async.commit_group;\n" ::);
        // the next problem: its index, and its upper triangle into L2 (lane = rows lane and lane + 32, 128-byte lines)
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        if ((size_t)nxt < nproblems) {
            const char* nb = reinterpret_cast<const char*>(cov + (size_t)nxt * K * K);
#pragma unroll
            for (int rs = 0; rs < CS; rs++) {
                const int i = lane + 32 * rs;
                if (i < K) {
                    const char* rowp = nb + (size_t)i * K * 8;
#pragma unroll
                    for (int ln = 0; ln < K * 8 / 128; ln++)
                        if (128 * ln + 127 >= i * 8) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(rowp + 128 * ln));
                }
            }
        }
        PHASE(0)
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        PHASE(1)

        // ---- c0 = r - G mean: one tensor-core pass W^T W over W = GM (entries (a, 4) = G[a] . mean) ----
        double gmean[4];
        {
            double2 acc[4];
#pragma unroll
            for (int u = 0; u < 4; u++) acc[u] = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < K / 4; q++) {  // four independent accumulator chains
                const double f = sm[G::GM + 32 * q + gtf];
                dmma(acc[q & 3].x, acc[q & 3].y, f, f);
            }
            *reinterpret_cast<double2*>(&sm[G::SCR + gcf]) =
                make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 4; a++) gmean[a] = sm[G::SCR + a * 8 + (4 ^ (4 * ((a >> 1) & 1)))];
        }
        PHASE(2)

        // ---- cov = U~^T D U~, panel by panel ----
        int fail = 0;
        // transposed fragments of the tiles of the panel factorised last: fa = U~, fad = -D U~ (the rank-8 update is
        // C(ti, tj) -= U~(p, ti)^T D U~(p, tj))
        double fa[NT][2], fad[NT][2];
#pragma unroll
        for (int t = 0; t < NT; t++) fa[t][0] = fa[t][1] = fad[t][0] = fad[t][1] = 0.0;
#pragma unroll
        for (int p = 0; p < NT; p++) {
            // the panel's tiles leave the registers (accumulator layout) ...
#pragma unroll
            for (int tj = p; tj < NT; tj++)
                *reinterpret_cast<double2*>(&sm[64 * G::tid(p, tj) + cf[tj & 1]]) = c[p][tj];
            __syncwarp();
            // ... and come back in the lane = column layout: v[cs][r] = A[8 p + r][j]
            double v[CS][8];
#pragma unroll
            for (int cs = 0; cs < CS; cs++) {
#pragma unroll
                for (int r = 0; r < 8; r++) v[cs][r] = 0.0;
                if (32 * cs + 31 >= 8 * p) {  // compile-time: the column set reaches into the panel
                    const int j = lane + 32 * cs;
                    if (j >= 8 * p && j < K) {
#pragma unroll
                        for (int r = 0; r < 8; r++) v[cs][r] = sm[64 * (G::tid(p, p) - p) + col[cs][r]];
                    }
                }
            }
            const int csd = (8 * p) >> 5;  // column set that holds the diagonal tile (compile-time after unrolling)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // the chain first: pivot, its reciprocal, the update of the NEXT pivot's row in the diagonal columns
                // the pivot row of the diagonal tile (columns 8p + j .. 8p + 7, held by eight lanes) goes through shared
                // memory: one store, broadcast 16-byte loads - 5 wavefronts instead of the 2 (8 - j) of a shuffle each
                const int cc = (lane - 8 * p) & 31;
                double* bcast = &sm[G::SCR + 8 * (j & 1)];
                if (cc >= j && cc < 8) bcast[cc] = v[csd][j];
                __syncwarp();
                double bc[8];
#pragma unroll
                for (int r = j & ~1; r < 8; r += 2) {
                    const double2 t2 = *reinterpret_cast<const double2*>(&bcast[r]);
                    bc[r] = t2.x;
                    bc[r + 1] = t2.y;
                }
                const double ajj = bc[j];
                const double rcp = fast_rcp(ajj);
                double t[CS];
#pragma unroll
                for (int cs = 0; cs < CS; cs++) t[cs] = v[cs][j] * rcp;
                if (j + 1 < 8) v[csd][j + 1] = fma(-t[csd], bc[j + 1], v[csd][j + 1]);
#pragma unroll
                for (int cs = 0; cs < CS; cs++) {
                    if (32 * cs + 31 >= 8 * p) {
#pragma unroll
                        for (int r = j + 1; r < 8; r++)
                            if (!(cs == csd && r == j + 1)) v[cs][r] = fma(-t[cs], bc[r], v[cs][r]);
                    }
                }
                if (ajj <= 0.0 && fail == 0) fail = 8 * p + j + 1;  // NaN passes
                // square-root free: row j of the UNIT upper factor is a[j][.] / d_j = t, the pivot d_j goes to WV
#pragma unroll
                for (int cs = 0; cs < CS; cs++)
                    if (32 * cs + 31 >= 8 * p) v[cs][j] = t[cs];
                if (lane == ((8 * p + j) & 31)) sm[G::WV + 8 * p + j] = ajj;
                // deferred tiles of step p - 1 (tile rows >= p + 1), an eighth of them after every pivot: registers only
                if (p >= 1) {
                    const int q = p - 1;
                    const int nd = (NT - 2 - q) * (NT - 1 - q) / 2;
                    const int lo = j * nd / 8, hi = (j + 1) * nd / 8;
                    // the chunk's first k-steps, then its second ones: no DMMA waits for the one before it
#pragma unroll
                    for (int ti = q + 2; ti < NT; ti++) {
#pragma unroll
                        for (int tj = ti; tj < NT; tj++) {
                            const int idx = G::didx(q, ti, tj);
                            if (idx >= lo && idx < hi) dmma(c[ti][tj].x, c[ti][tj].y, fad[ti][0], fa[tj][0]);
                        }
                    }
#pragma unroll
                    for (int ti = q + 2; ti < NT; ti++) {
#pragma unroll
                        for (int tj = ti; tj < NT; tj++) {
                            const int idx = G::didx(q, ti, tj);
                            if (idx >= lo && idx < hi) dmma(c[ti][tj].x, c[ti][tj].y, fad[ti][1], fa[tj][1]);
                        }
                    }
                }
            }
            // columns of the diagonal tile: nothing below the diagonal (the products below read whole tiles)
#pragma unroll
            for (int cs = 0; cs < CS; cs++) {
                if (32 * cs + 31 >= 8 * p) {
                    const int j = lane + 32 * cs;
                    if (cs == csd) {  // compile-time: only this column set holds columns of the diagonal tile
                        const int cc = j - 8 * p;
#pragma unroll
                        for (int r = 1; r < 8; r++)
                            if (cc >= 0 && cc < r) v[cs][r] = 0.0;
                    }
                    if (j >= 8 * p && j < K) {
#pragma unroll
                        for (int r = 0; r < 8; r++) sm[64 * (G::tid(p, p) - p) + col[cs][r]] = v[cs][r];
                    }
                }
            }
            __syncwarp();
            PHASE(3)
            if (p + 1 < NT) {
                // fragments of panel p, then the urgent tile row p + 1 (all that the next elimination reads)
                const double ds0 = -sm[G::WV + 8 * p + t4], ds1 = -sm[G::WV + 8 * p + 4 + t4];  // -d of the fragment's rows
#pragma unroll
                for (int t = p + 1; t < NT; t++) {
                    fa[t][0] = sm[64 * G::tid(p, t) + tf[t & 1]];
                    fa[t][1] = sm[64 * G::tid(p, t) + 32 + tf[t & 1]];
                    fad[t][0] = fa[t][0] * ds0;
                    fad[t][1] = fa[t][1] * ds1;
                }
#pragma unroll
                for (int tj = p + 1; tj < NT; tj++) dmma(c[p + 1][tj].x, c[p + 1][tj].y, fad[p + 1][0], fa[tj][0]);
#pragma unroll
                for (int tj = p + 1; tj < NT; tj++) dmma(c[p + 1][tj].x, c[p + 1][tj].y, fad[p + 1][1], fa[tj][1]);
                PHASE(4)
            }
        }
        if (fail) {  // the reference returns before drawing (src/htnorm_distributions.c:78): out untouched
            if (lane == 0 && info) info[pb] = fail;
            continue;
        }

        // the Cholesky factor is sqrt(D) U~.  sd = sqrt(d), lane = row; WV then holds sd
        double sdreg[CS];
#pragma unroll
        for (int rs = 0; rs < CS; rs++) {
            const int i = lane + 32 * rs;
            const double dv = i < K ? sm[G::WV + i] : 1.0;
            sdreg[rs] = dv * fast_rsqrt(dv);  // sqrt(d), d > 0 here
        }
        __syncwarp();
#pragma unroll
        for (int rs = 0; rs < CS; rs++) {
            const int i = lane + 32 * rs;
            if (i < K) sm[G::WV + i] = sdreg[rs];
        }
        __syncwarp();
        // ---- V = sqrt(D) U~ G^T on the tensor cores: one accumulator per 8 rows of V, all chains interleaved; written
        // over G^T at the end ----
        {
            double2 va[NT];
#pragma unroll
            for (int tb = 0; tb < NT; tb++) va[tb] = make_double2(0.0, 0.0);
#pragma unroll
            for (int tj = 0; tj < NT; tj++) {
                // A[m][kk] = U~[8 tb + m][8 tj + kk] (row fragment of tile (tb, tj)), B[kk][n] = G[n][8 tj + kk]
                const double b0 = sm[G::GM + 64 * tj + gtf];
                const double b1 = sm[G::GM + 64 * tj + 32 + gtf];
                double a0[NT], a1[NT];
#pragma unroll
                for (int tb = 0; tb <= tj; tb++) {
                    a0[tb] = sm[64 * G::tid(tb, tj) + rf0[tj & 1]];
                    a1[tb] = sm[64 * G::tid(tb, tj) + rf1[tj & 1]];
                }
#pragma unroll
                for (int tb = 0; tb <= tj; tb++) dmma(va[tb].x, va[tb].y, a0[tb], b0);
#pragma unroll
                for (int tb = 0; tb <= tj; tb++) dmma(va[tb].x, va[tb].y, a1[tb], b1);
            }
            __syncwarp();  // every lane has read G^T
            if (t4 < 2) {
#pragma unroll
                for (int tb = 0; tb < NT; tb++) {
                    const double sdr = sm[G::WV + 8 * tb + g];
                    *reinterpret_cast<double2*>(&sm[G::GM + 64 * tb + gcf]) = make_double2(va[tb].x * sdr, va[tb].y * sdr);
                }
            }
            __syncwarp();
        }
        PHASE(5)
        // ---- S = V^T V and V^T z in one pass W^T W over W = [V | mean | z]: entries (a, b) and (a, 5) ----
        {
            double2 acc[4];
#pragma unroll
            for (int u = 0; u < 4; u++) acc[u] = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < K / 4; q++) {
                const double f = sm[G::GM + 32 * q + gtf];
                dmma(acc[q & 3].x, acc[q & 3].y, f, f);
            }
            *reinterpret_cast<double2*>(&sm[G::SCR + 64 + gcf]) =
                make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
            __syncwarp();
        }
        // ---- k2 x k2 solve (dposv, src/htnorm.c:169; k2 = 1 divides by the scalar, :77): every lane alike ----
        int f2 = 0;
        double al[4] = {0.0, 0.0, 0.0, 0.0};
        {
            double s[4][4], bv[4], ri[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int swi = 4 * ((i >> 1) & 1);
                const double vtz = sm[G::SCR + 64 + i * 8 + (5 ^ swi)];
                bv[i] = i < k2 ? rreg[i] - gmean[i] - vtz : 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j <= i) {
                        const double sv = sm[G::SCR + 64 + i * 8 + (j ^ swi)];
                        s[i][j] = (i < k2) ? sv : (i == j ? 1.0 : 0.0);
                    }
            }
            if (k2 == 1) {
                al[0] = bv[0] / s[0][0];
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double ajj = s[j][j];
                    if (ajj <= 0.0 && f2 == 0) f2 = j + 1;
                    const double rinv = fast_rsqrt(ajj);
                    ri[j] = rinv;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (i > j) s[i][j] *= rinv;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int cc = 0; cc < 4; cc++)
                            if (i > j && cc > j && cc <= i) s[i][cc] = fma(-s[i][j], s[cc][j], s[i][cc]);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    double t = bv[i];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u < i) t = fma(-s[i][u], bv[u], t);
                    bv[i] = t * ri[i];
                }
#pragma unroll
                for (int i = 3; i >= 0; i--) {
                    double t = bv[i];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u > i) t = fma(-s[u][i], bv[u], t);
                    bv[i] = t * ri[i];
                }
#pragma unroll
                for (int i = 0; i < 4; i++) al[i] = f2 ? 0.0 : bv[i];  // G cov G^T not PD: the unprojected y (src/htnorm.c:170)
            }
        }
        PHASE(6)
        // ---- w = sqrt(D) (z + V alpha) (lane = row) ----
#pragma unroll
        for (int rs = 0; rs < CS; rs++) {
            const int i = lane + 32 * rs;
            if (i < K) {
                const double2 v01 = *reinterpret_cast<const double2*>(&sm[G::GM + row_lo[rs]]);
                const double2 v23 = *reinterpret_cast<const double2*>(&sm[G::GM + row_lo[rs] + 2]);
                double w = sm[G::GM + row_hi[rs] + 1];  // z, column 5
                w = fma(v01.x, al[0], w);
                w = fma(v01.y, al[1], w);
                w = fma(v23.x, al[2], w);
                w = fma(v23.y, al[3], w);
                sm[G::WV + i] = w * sdreg[rs];
            }
        }
        __syncwarp();
        // ---- x = mean + U~^T w: column dot products, lane = column, panel by panel ----
        double acc[CS];
#pragma unroll
        for (int cs = 0; cs < CS; cs++) acc[cs] = mreg[cs];
#pragma unroll
        for (int p = 0; p < NT; p++) {
            const double2 w01 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p]);
            const double2 w23 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 2]);
            const double2 w45 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 4]);
            const double2 w67 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 6]);
            const double wv[8] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y};
#pragma unroll
            for (int cs = 0; cs < CS; cs++) {
                if (32 * cs + 31 >= 8 * p) {
                    const int j = lane + 32 * cs;
                    if (j >= 8 * p && j < K) {
                        double t = acc[cs];
#pragma unroll
                        for (int r = 0; r < 8; r++) t = fma(sm[64 * (G::tid(p, p) - p) + col[cs][r]], wv[r], t);
                        acc[cs] = t;
                    }
                }
            }
        }
#pragma unroll
        for (int cs = 0; cs < CS; cs++) {
            const int j = lane + 32 * cs;
            if (j < K) out[pb * (size_t)K + j] = acc[cs];
        }
        if (lane == 0 && info) info[pb] = f2;
        __syncwarp();  // the next problem's asynchronous copies overwrite what this one was still reading
        PHASE(7)
    }
}
#undef PHASE

}  // namespace

// returns 1 when the shape is not one this kernel is built for (the caller falls back to the generic kernel).
// counter: one zero-initialised unsigned on the device (the kernel's dynamic problem queue)
extern "C" int htnk_ht_small_warp(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                                  const double* cov, const double* g, const double* r, const double* zs, double* out,
                                  int* info, unsigned* counter)
{
    if (nproblems == 0) return 0;
    if (k < 8 || k > 64 || (k & 7) || k2 < 1 || k2 > 4 || nproblems > 0xfffffff0u || counter == nullptr) return 1;
    static const bool off = getenv("HTN_SMALL_IMPL") != nullptr && atoi(getenv("HTN_SMALL_IMPL")) == 0;
    if (off) return 1;
    if ((((uintptr_t)mean | (uintptr_t)g | (uintptr_t)r | (uintptr_t)zs | (uintptr_t)out) & 7) != 0 || ((uintptr_t)cov & 15) != 0)
        return 1;
    static const bool want_prof = getenv("HTN_SMALL_PROF") != nullptr;
    // one instantiation per k = 8 NT (every address in the kernel is a compile-time constant)
    void (*kern)(size_t, int, const double*, const double*, const double*, const double*, const double*, double*, int*,
                 unsigned*, long long*) = nullptr;
    size_t smem = 0;
    switch (k >> 3) {
        case 1: kern = ht_warp_kernel<1, false>; smem = Geo<1>::TOTAL; break;
        case 2: kern = ht_warp_kernel<2, false>; smem = Geo<2>::TOTAL; break;
        case 3: kern = ht_warp_kernel<3, false>; smem = Geo<3>::TOTAL; break;
        case 4: kern = ht_warp_kernel<4, false>; smem = Geo<4>::TOTAL; break;
        case 5: kern = ht_warp_kernel<5, false>; smem = Geo<5>::TOTAL; break;
        case 6: kern = ht_warp_kernel<6, false>; smem = Geo<6>::TOTAL; break;
        case 7: kern = ht_warp_kernel<7, false>; smem = Geo<7>::TOTAL; break;
        default: kern = want_prof ? ht_warp_kernel<8, true> : ht_warp_kernel<8, false>; smem = Geo<8>::TOTAL; break;
    }
    smem *= sizeof(double);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("HTN_SMALL_CTAS")) per_sm = atoi(e) < per_sm ? (atoi(e) > 0 ? atoi(e) : 1) : per_sm;
    size_t grid = (size_t)sms * per_sm;
    if (grid > nproblems) grid = nproblems;
    long long* prof = nullptr;
    if (want_prof && k == 64) {
        if (cudaMalloc(&prof, 16 * sizeof(long long)) != cudaSuccess) return -201;
        cudaMemsetAsync(prof, 0, 16 * sizeof(long long), st);
    }
    kern<<<(unsigned)grid, 32, smem, st>>>(nproblems, k2, mean, cov, g, r, zs, out, info, counter, prof);
    htnk_count_launch();
    if (prof) {  // development aid: synchronous on purpose
        long long hp[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(hp, prof, sizeof(hp), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        static const char* names[8] = {"load+stage", "wait", "c0", "panel store+elim(+deferred)", "frags+urgent", "sd+V", "S+solve", "w+x+out"};
        fprintf(stderr, "[ht_warp] %d warps/SM (%zu B smem), clk per problem:", per_sm, smem);
        double tot = 0;
        for (int i = 0; i < 8; i++) {
            fprintf(stderr, " %s %.0f", names[i], (double)hp[i] / (double)nproblems);
            tot += (double)hp[i] / (double)nproblems;
        }
        fprintf(stderr, " | total %.0f\n", tot);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
