// Many small independent problems, round-2 design: ONE WARP PER PROBLEM, everything warp-synchronous, every
// address a compile-time constant (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems).
//
// Why (profiles/r1_small_batched_v7_summary.txt): the round-1 kernel (small_batched.cu: one 64-thread CTA per problem,
// factor warp + update warp, generic k) spent 12.8 k warp-instructions per problem - a quarter of them address
// arithmetic (IMAD / LEA / ISETP), 42 block barriers - and 2.9 k cycles of the SM's FP64 datapath on padded or
// redundant work (all-lanes 8 x 8 factor, k2 and z padded to 8 on the tensor cores); it reached 22 % of the HBM
// roofline.  Here:
//   * the whole of src/htnorm.c:85-187 + mvn_rand_cov (src/htnorm_distributions.c:58-88) for one problem runs in one
//     warp out of 23 KB of shared memory: no block barrier, no role hand-off, 9 independent problems per SM;
//   * k is a template parameter (k = 8 NT): loops are fully unrolled and every shared-memory offset is an immediate;
//   * cov = L L^T in LOWER form on 8-wide column panels.  The panel step is done with the LANE = ROW layout: lane i
//     holds row i of the panel in registers, the 8 pivots of a panel are eliminated by row operations whose only
//     communication is a warp shuffle of the pivot row (no 8 x 8 factor replicated in all lanes, no explicit inverse,
//     no tensor-core panel solve); the update of a pivot uses 1 / a_jj (reciprocal seed + Newton), so the rsqrt that
//     scales the column is off the critical path;
//   * the rank-8 trailing update stays on the FP64 tensor cores (DMMA m8n8k4), fragments read conflict-free through
//     an XOR swizzle of the dense 8-double panel rows (no padding: 18 KB for the factor instead of 27);
//   * the projection is reformulated around the factor so that cov G^T (2 k^2 k2 flops) is never formed:
//         V = L^T G^T,  G cov G^T = V^T V,  G y = G mean + V^T z,  x = mean + L (z + V alpha)
//     (same algebra as the shared-covariance path's W = G L, hyperplane.c) - V on the tensor cores, x by row dot products.
// Error codes follow the reference: cov not PD -> pivot index, out untouched; G cov G^T not PD -> its
// pivot index and the unprojected y = mean + L z.
// Handles k in {64} (NT = 8), k2 <= 4, dense row-major cov (upper triangle read); every other shape keeps
// the generic kernel (small_batched.cu).
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
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Shared-memory map of one problem (doubles).  Panel p holds L[8p.., 8p..8p+7] as dense 8-double rows; element
// (row i, column c of the panel) sits at row * 8 + (c ^ sw(i)), sw(i) = 4 * ((i >> 1) & 1): with that XOR the
// tensor-core fragment patterns (row g / column t4, and the transposed one) touch 16 distinct 8-byte bank pairs per
// half-warp, and the 16-byte accumulator accesses 8 distinct 16-byte slots per quarter-warp.
template <int NT>
struct Geo {
    static constexpr int K = 8 * NT;
    static constexpr int RS = (K + 31) / 32;  // rows per lane
    static __host__ __device__ constexpr int off(int p) { return 8 * (K * p - 4 * p * (p - 1)); }
    static constexpr int GM = 8 * (K * NT - 4 * NT * (NT - 1));  // [K][8]: G^T (cols 0-3) | mean (4) | z (5) | 0 0, later V in 0-3
    static constexpr int WV = GM + 8 * K;                        // [K]   w = z + V alpha
    static constexpr int SCR = WV + K;                           // [2][64] two 8 x 8 result tiles
    static constexpr int TOTAL = SCR + 128;
};

// PROF (HTN_SMALL_PROF=1, development only): lane 0 of every warp adds the clock64() span of each phase to prof[]
#define PHASE(i)                                                                       \
    if (PROF) {                                                                        \
        const long long t_now = clock64();                                             \
        if (lane == 0) atomicAdd((unsigned long long*)&prof[i], (unsigned long long)(t_now - t_prev)); \
        t_prev = t_now;                                                                \
    }

template <int NT, bool PROF>
__global__ void __launch_bounds__(32, 9)
ht_warp_kernel(size_t nproblems, int k2, const double* __restrict__ mean, const double* __restrict__ cov,
               const double* __restrict__ gm, const double* __restrict__ rv, const double* __restrict__ zs,
               double* __restrict__ out, int* __restrict__ info, long long* prof)
{
    using G = Geo<NT>;
    constexpr int K = G::K, RS = G::RS;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);

    // ---- lane constants (byte offsets) ----
    // fragment "row pattern": element (row g, column t4 + 4 s) of an 8 x 8 tile
    const int swg = 4 * ((g >> 1) & 1);
    const unsigned rf0 = 8u * (unsigned)(g * 8 + (t4 ^ swg)), rf1 = 8u * (unsigned)(g * 8 + ((t4 + 4) ^ swg));
    // fragment "transposed pattern": element (row t4 + 4 s, column g); s = 1 is 4 rows = 256 bytes further
    const unsigned tf0 = 8u * (unsigned)(t4 * 8 + (g ^ (4 * ((t4 >> 1) & 1))));
    // accumulator pattern: elements (row g, columns 2 t4, 2 t4 + 1), one 16-byte access
    const unsigned cf = 8u * (unsigned)(g * 8 + ((2 * t4) ^ swg));
    // row pattern of the elimination: lane = row i = lane + 32 rs; columns 0-3 at lo, 4-7 at hi (swizzled halves)
    unsigned row_lo[RS], row_hi[RS];
#pragma unroll
    for (int rs = 0; rs < RS; rs++) {
        const int i = lane + 32 * rs;
        const int sw = 4 * ((i >> 1) & 1);
        row_lo[rs] = 8u * (unsigned)(i * 8 + sw);
        row_hi[rs] = 8u * (unsigned)(i * 8 + (4 - sw));
    }
    // zero the region once: GM columns that are never written must read as zero for the tensor-core passes
    for (int idx = lane; idx < G::TOTAL; idx += 32) sm[idx] = 0.0;
    __syncwarp();

    long long t_prev = PROF ? clock64() : 0;
    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)K * K;
        const double* gp = gm + pb * (size_t)k2 * K;
        // ---- stage: cov[i][j], j >= i (the upper triangle of the row-major matrix) lands at L position
        // (row j, column i): lane = j, one 8-byte asynchronous copy per element, 256 contiguous bytes per warp request
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int j = lane + 32 * rs;
            const double* src = covp + j;
#pragma unroll
            for (int i = 0; i < K; i++) {
                if (i <= 32 * rs + 31) {  // compile-time: this row set has elements j >= i
                    const int p = i >> 3, c = i & 7;
                    // (c ^ sw(j)) = c + sw for c < 4, c - sw for c >= 4
                    const unsigned dst = sbase + (c < 4 ? row_lo[rs] : row_hi[rs]) +
                                         8u * (unsigned)(G::off(p) - 64 * p + (c & 3));
                    if (j >= i && j < K) cp_async8(dst, src + (size_t)i * K);
                }
            }
        }
        // G^T -> GM columns 0..k2-1, mean -> column 4, z -> column 5 (lane = row)
        double mreg[RS];
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int i = lane + 32 * rs;
            mreg[rs] = 0.0;
            if (i < K) {
                const unsigned rowb = sbase + 8u * (unsigned)G::GM;
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < k2) cp_async8(rowb + row_lo[rs] + 8u * c, gp + (size_t)c * K + i);
                cp_async8(rowb + row_hi[rs], mean + pb * (size_t)K + i);          // column 4
                cp_async8(rowb + row_hi[rs] + 8u, zs + pb * (size_t)K + i);       // column 5
                mreg[rs] = mean[pb * (size_t)K + i];
            }
        }
        double rreg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < k2) rreg[c] = rv[pb * (size_t)k2 + c];
        asm volatile("cp.async.commit_group;\n" ::);
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
                const double f = sm[G::GM + 32 * q + (tf0 >> 3)];
                dmma(acc[q & 3].x, acc[q & 3].y, f, f);
            }
            *reinterpret_cast<double2*>(&sm[G::SCR + (cf >> 3)]) =
                make_double2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
            __syncwarp();
#pragma unroll
            for (int a = 0; a < 4; a++) gmean[a] = sm[G::SCR + a * 8 + (4 ^ (4 * ((a >> 1) & 1)))];
        }

        PHASE(2)
        // ---- Cholesky cov = L L^T, panel by panel, with look-ahead inside the single instruction stream: the
        // rank-8 update of step p is split into the URGENT tiles (block column p + 1: all that the elimination of panel
        // p + 1 reads) and the DEFERRED ones (block columns >= p + 2), which are issued a few at a time BETWEEN the
        // pivots of that elimination - independent tensor-core work in the stalls of the 8-pivot dependency chain ----
        int fail = 0;
        // fragments (row pattern) of the tiles of the panel factorised last: fa = L, fad = -L D (cov = L D L^T: the rank-8
        // update is C -= (L D) L^T)
        double fa[NT][2], fad[NT][2];
#pragma unroll
        for (int t = 0; t < NT; t++) fa[t][0] = fa[t][1] = fad[t][0] = fad[t][1] = 0.0;
        auto tile_update = [&](int ti, int tj) {
            double2* cp = reinterpret_cast<double2*>(&sm[G::off(tj) + 64 * (ti - tj) + (cf >> 3)]);
            double2 c = *cp;
            dmma(c.x, c.y, fad[ti][0], fa[tj][0]);
            dmma(c.x, c.y, fad[ti][1], fa[tj][1]);
            *cp = c;
        };
#pragma unroll
        for (int p = 0; p < NT; p++) {
            // elimination of panel p in the lane = row layout
            double v[RS][8];
#pragma unroll
            for (int rs = 0; rs < RS; rs++) {
                if (32 * rs + 31 >= 8 * p) {  // compile-time: the row set reaches into the panel
                    const int i = lane + 32 * rs;
                    const bool in = i >= 8 * p && i < K;
                    const int cb = G::off(p) - 64 * p;
                    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
                    if (in) {
                        a0 = *reinterpret_cast<const double2*>(&sm[cb + (row_lo[rs] >> 3)]);
                        a1 = *reinterpret_cast<const double2*>(&sm[cb + (row_lo[rs] >> 3) + 2]);
                        a2 = *reinterpret_cast<const double2*>(&sm[cb + (row_hi[rs] >> 3)]);
                        a3 = *reinterpret_cast<const double2*>(&sm[cb + (row_hi[rs] >> 3) + 2]);
                    }
                    v[rs][0] = a0.x; v[rs][1] = a0.y; v[rs][2] = a1.x; v[rs][3] = a1.y;
                    v[rs][4] = a2.x; v[rs][5] = a2.y; v[rs][6] = a3.x; v[rs][7] = a3.y;
                } else {
#pragma unroll
                    for (int c = 0; c < 8; c++) v[rs][c] = 0.0;
                }
            }
            const int rsd = (8 * p) >> 5;  // row set that holds the diagonal block (compile-time after unrolling)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // the chain first: pivot, its reciprocal, the update of the NEXT pivot's column in the diagonal rows
                const double ajj = shfl(v[rsd][j], (8 * p + j) & 31);
                double bc[8];
#pragma unroll
                for (int c = j + 1; c < 8; c++) bc[c] = shfl(v[rsd][j], (8 * p + c) & 31);  // a[c][j], unscaled
                const double rcp = fast_rcp(ajj);
                double t[RS];
#pragma unroll
                for (int rs = 0; rs < RS; rs++) t[rs] = v[rs][j] * rcp;
                if (j + 1 < 8) v[rsd][j + 1] = fma(-t[rsd], bc[j + 1], v[rsd][j + 1]);
#pragma unroll
                for (int rs = 0; rs < RS; rs++) {
                    if (32 * rs + 31 >= 8 * p) {
#pragma unroll
                        for (int c = j + 1; c < 8; c++)
                            if (!(rs == rsd && c == j + 1)) v[rs][c] = fma(-t[rs], bc[c], v[rs][c]);
                    }
                }
                if (ajj <= 0.0 && fail == 0) fail = 8 * p + j + 1;  // NaN passes
                // square-root free: column j of the UNIT lower factor is a[i][j] / d_j = t, the pivot d_j goes to DV
#pragma unroll
                for (int rs = 0; rs < RS; rs++)
                    if (32 * rs + 31 >= 8 * p) v[rs][j] = t[rs];
                if (lane == ((8 * p + j) & 31)) sm[G::WV + 8 * p + j] = ajj;
                // deferred tiles of step p - 1 (block columns >= p + 1), an eighth of them after every pivot
                if (p >= 1) {
                    const int q = p - 1;
                    const int nd = (NT - 2 - q) * (NT - 1 - q) / 2;
                    const int lo = j * nd / 8, hi = (j + 1) * nd / 8;
                    int idx = 0;
#pragma unroll
                    for (int tj = q + 2; tj < NT; tj++) {
#pragma unroll
                        for (int ti = tj; ti < NT; ti++) {
                            if (idx >= lo && idx < hi) tile_update(ti, tj);
                            idx++;
                        }
                    }
                }
            }
            // rows of the diagonal block: nothing above the diagonal (the products below read whole tiles)
#pragma unroll
            for (int rs = 0; rs < RS; rs++) {
                if (32 * rs + 31 >= 8 * p) {
                    const int i = lane + 32 * rs;
                    if (rs == rsd) {  // compile-time: only this row set holds rows of the diagonal block
                        const int r = i - 8 * p;
#pragma unroll
                        for (int c = 1; c < 8; c++)
                            if (r >= 0 && r < c) v[rs][c] = 0.0;
                    }
                    if (i >= 8 * p && i < K) {
                        const int cb = G::off(p) - 64 * p;
                        *reinterpret_cast<double2*>(&sm[cb + (row_lo[rs] >> 3)]) = make_double2(v[rs][0], v[rs][1]);
                        *reinterpret_cast<double2*>(&sm[cb + (row_lo[rs] >> 3) + 2]) = make_double2(v[rs][2], v[rs][3]);
                        *reinterpret_cast<double2*>(&sm[cb + (row_hi[rs] >> 3)]) = make_double2(v[rs][4], v[rs][5]);
                        *reinterpret_cast<double2*>(&sm[cb + (row_hi[rs] >> 3) + 2]) = make_double2(v[rs][6], v[rs][7]);
                    }
                }
            }
            __syncwarp();
            PHASE(3)
            if (p + 1 < NT) {
                // fragments of panel p, then the urgent tiles (block column p + 1): all loads, then two rounds of
                // independent DMMAs, then the stores
                const double ds0 = -sm[G::WV + 8 * p + t4], ds1 = -sm[G::WV + 8 * p + 4 + t4];  // -d of the fragment's columns
#pragma unroll
                for (int t = p + 1; t < NT; t++) {
                    fa[t][0] = sm[G::off(p) + 64 * (t - p) + (rf0 >> 3)];
                    fa[t][1] = sm[G::off(p) + 64 * (t - p) + (rf1 >> 3)];
                    fad[t][0] = fa[t][0] * ds0;
                    fad[t][1] = fa[t][1] * ds1;
                }
                double2 cu[NT];
#pragma unroll
                for (int ti = p + 1; ti < NT; ti++)
                    cu[ti] = *reinterpret_cast<const double2*>(&sm[G::off(p + 1) + 64 * (ti - p - 1) + (cf >> 3)]);
#pragma unroll
                for (int ti = p + 1; ti < NT; ti++) dmma(cu[ti].x, cu[ti].y, fad[ti][0], fa[p + 1][0]);
#pragma unroll
                for (int ti = p + 1; ti < NT; ti++) dmma(cu[ti].x, cu[ti].y, fad[ti][1], fa[p + 1][1]);
#pragma unroll
                for (int ti = p + 1; ti < NT; ti++)
                    *reinterpret_cast<double2*>(&sm[G::off(p + 1) + 64 * (ti - p - 1) + (cf >> 3)]) = cu[ti];
                __syncwarp();
                PHASE(4)
            }
        }
        if (fail) {  // the reference returns before drawing (src/htnorm_distributions.c:78): out untouched
            if (lane == 0 && info) info[pb] = fail;
            continue;
        }

        // cov = L D L^T with L unit lower: the Cholesky factor is L sqrt(D).  sd = sqrt(d), lane = row; DV then holds sd
        double sdreg[RS];
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int i = lane + 32 * rs;
            sdreg[rs] = i < K ? sqrt(sm[G::WV + i]) : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int i = lane + 32 * rs;
            if (i < K) sm[G::WV + i] = sdreg[rs];
        }
        __syncwarp();
        // ---- V = sqrt(D) L^T G^T on the tensor cores: one accumulator per 8 rows of V, all eight chains interleaved;
        // written over G^T at the end ----
        {
            double2 va[NT];
#pragma unroll
            for (int tb = 0; tb < NT; tb++) va[tb] = make_double2(0.0, 0.0);
#pragma unroll
            for (int ti = 0; ti < NT; ti++) {
                // A[m][kk] = L[8 ti + kk][8 tb + m] (transposed pattern on panel tb), B[kk][n] = G[n][8 ti + kk]
                const double b0 = sm[G::GM + 64 * ti + (tf0 >> 3)];
                const double b1 = sm[G::GM + 64 * ti + 32 + (tf0 >> 3)];
                double a0[NT], a1[NT];
#pragma unroll
                for (int tb = 0; tb <= ti; tb++) {
                    a0[tb] = sm[G::off(tb) + 64 * (ti - tb) + (tf0 >> 3)];
                    a1[tb] = sm[G::off(tb) + 64 * (ti - tb) + 32 + (tf0 >> 3)];
                }
#pragma unroll
                for (int tb = 0; tb <= ti; tb++) dmma(va[tb].x, va[tb].y, a0[tb], b0);
#pragma unroll
                for (int tb = 0; tb <= ti; tb++) dmma(va[tb].x, va[tb].y, a1[tb], b1);
            }
            __syncwarp();  // every lane has read G^T
            if (t4 < 2) {
#pragma unroll
                for (int tb = 0; tb < NT; tb++) {
                    const double sdr = sm[G::WV + 8 * tb + g];
                    *reinterpret_cast<double2*>(&sm[G::GM + 64 * tb + (cf >> 3)]) = make_double2(va[tb].x * sdr, va[tb].y * sdr);
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
                const double f = sm[G::GM + 32 * q + (tf0 >> 3)];
                dmma(acc[q & 3].x, acc[q & 3].y, f, f);
            }
            *reinterpret_cast<double2*>(&sm[G::SCR + 64 + (cf >> 3)]) =
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
                        for (int c = 0; c < 4; c++)
                            if (i > j && c > j && c <= i) s[i][c] = fma(-s[i][j], s[c][j], s[i][c]);
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
        // ---- w = z + V alpha (lane = row) ----
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int i = lane + 32 * rs;
            if (i < K) {
                const double2 v01 = *reinterpret_cast<const double2*>(&sm[G::GM + (row_lo[rs] >> 3)]);
                const double2 v23 = *reinterpret_cast<const double2*>(&sm[G::GM + (row_lo[rs] >> 3) + 2]);
                double w = sm[G::GM + (row_hi[rs] >> 3) + 1];  // z, column 5
                w = fma(v01.x, al[0], w);
                w = fma(v01.y, al[1], w);
                w = fma(v23.x, al[2], w);
                w = fma(v23.y, al[3], w);
                sm[G::WV + i] = w * sdreg[rs];  // Cholesky factor times w = L (sqrt(D) w)
            }
        }
        __syncwarp();
        // ---- x = mean + L (sqrt(D) w): row dot products, lane = row, panel by panel ----
        double acc[RS];
#pragma unroll
        for (int rs = 0; rs < RS; rs++) acc[rs] = mreg[rs];
#pragma unroll
        for (int p = 0; p < NT; p++) {
            const double2 w01 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p]);
            const double2 w23 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 2]);
            const double2 w45 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 4]);
            const double2 w67 = *reinterpret_cast<const double2*>(&sm[G::WV + 8 * p + 6]);
#pragma unroll
            for (int rs = 0; rs < RS; rs++) {
                if (32 * rs + 31 >= 8 * p) {
                    const int i = lane + 32 * rs;
                    if (i >= 8 * p && i < K) {
                        const int cb = G::off(p) - 64 * p;
                        const double2 a0 = *reinterpret_cast<const double2*>(&sm[cb + (row_lo[rs] >> 3)]);
                        const double2 a1 = *reinterpret_cast<const double2*>(&sm[cb + (row_lo[rs] >> 3) + 2]);
                        const double2 a2 = *reinterpret_cast<const double2*>(&sm[cb + (row_hi[rs] >> 3)]);
                        const double2 a3 = *reinterpret_cast<const double2*>(&sm[cb + (row_hi[rs] >> 3) + 2]);
                        double t = acc[rs];
                        t = fma(a0.x, w01.x, t); t = fma(a0.y, w01.y, t);
                        t = fma(a1.x, w23.x, t); t = fma(a1.y, w23.y, t);
                        t = fma(a2.x, w45.x, t); t = fma(a2.y, w45.y, t);
                        t = fma(a3.x, w67.x, t); t = fma(a3.y, w67.y, t);
                        acc[rs] = t;
                    }
                }
            }
        }
#pragma unroll
        for (int rs = 0; rs < RS; rs++) {
            const int i = lane + 32 * rs;
            if (i < K) out[pb * (size_t)K + i] = acc[rs];
        }
        if (lane == 0 && info) info[pb] = f2;
        __syncwarp();  // the next problem's asynchronous copies overwrite what this one was still reading
        PHASE(7)
    }
}
#undef PHASE

}  // namespace

// returns 1 when the shape is not one this kernel is built for (the caller falls back to the generic kernel)
extern "C" int htnk_ht_small_warp(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                                  const double* cov, const double* g, const double* r, const double* zs, double* out,
                                  int* info)
{
    if (nproblems == 0) return 0;
    if (k != 64 || k2 < 1 || k2 > 4) return 1;
    static const bool off = getenv("HTN_SMALL_IMPL") != nullptr && atoi(getenv("HTN_SMALL_IMPL")) == 0;
    if (off) return 1;
    if ((((uintptr_t)mean | (uintptr_t)cov | (uintptr_t)g | (uintptr_t)r | (uintptr_t)zs | (uintptr_t)out) & 7) != 0) return 1;
    static const bool want_prof = getenv("HTN_SMALL_PROF") != nullptr;
    auto kern = want_prof ? ht_warp_kernel<8, true> : ht_warp_kernel<8, false>;
    const size_t smem = (size_t)Geo<8>::TOTAL * sizeof(double);
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
    if (want_prof) {
        if (cudaMalloc(&prof, 16 * sizeof(long long)) != cudaSuccess) return -201;
        cudaMemsetAsync(prof, 0, 16 * sizeof(long long), st);
    }
    kern<<<(unsigned)grid, 32, smem, st>>>(nproblems, k2, mean, cov, g, r, zs, out, info, prof);
    htnk_count_launch();
    if (want_prof) {  // development aid: synchronous on purpose
        long long hp[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(hp, prof, sizeof(hp), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        static const char* names[8] = {"issue-stage", "wait-stage", "c0", "elim(+deferred)", "urgent", "V", "S+solve", "w+x+out"};
        fprintf(stderr, "[ht_warp] %d warps/SM (%zu B smem), clk per problem:", per_sm, smem);
        double tot = 0;
        for (int i = 0; i < 8; i++) { fprintf(stderr, " %s %.0f", names[i], (double)hp[i] / (double)nproblems); tot += (double)hp[i] / (double)nproblems; }
        fprintf(stderr, " | total %.0f\n", tot);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
