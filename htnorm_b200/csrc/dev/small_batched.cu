// Many small independent problems (BASELINE.json configs[1]: k = 64, k2 = 4, 2^20 problems, each with
// its own mean / cov / G / r and its own generator stream).
//
// One small CTA per problem (64 threads up to k = 64, 128 above) runs the WHOLE of the reference's
// htn_hyperplane_truncated_mvn (src/htnorm.c:85-187 + mvn_rand_cov, src/htnorm_distributions.c:58-88) out of
// shared memory, so every input byte is read from HBM at most once (35 872 B per problem at k = 64, k2 = 4)
// and only the k-vector result goes back: the roofline of this path is HBM, not the 127 kflop of arithmetic
// per problem.  What limits it in practice is LATENCY - a 64-step pivot chain per problem - so the design goal
// is to keep as many problems in flight per SM as shared memory allows: the covariance is held as its upper
// triangle only, packed by 8-row tile rows (20.5 KB instead of 36 KB at k = 64); 6 CTAs run per SM.
//
// The normals z of every problem are drawn beforehand by normal_fill_kernel (one stream per thread, fully
// parallel - a per-problem stream drawn inside this kernel would leave 63 threads waiting for one).
//
// "Upper world": the reference reads the UPPER triangle of a row-major cov, so the rows are staged
// as they lie in HBM (16-byte cp.async of the part right of the diagonal, nothing is symmetrised or transposed)
// and the factorisation is carried out as cov = U^T U on that triangle; U = L^T holds exactly the numbers of
// the reference's lower factor.  With HTN_FLAG_COLMAJOR the lower triangle is the valid one and g is column-major:
// both are transposed while staging.
//
//   cov * G^T on the tensor cores (DMMA), symmetric reads through the one triangle   dsymm,  src/htnorm.c:154
//   G cov G^T accumulated per warp from its own rows of cov G^T                      dgemm,  src/htnorm.c:164
//   Cholesky in 8-wide panels: local 8 x 8 factor + column solve per thread, DMMA rank-8 trailing update
//                                           dpotrf, distributions.c:77
//   y = mean + U^T z                        dtrmv + daxpy, distributions.c:81-83
//   r - G y, k2 x k2 solve                  dgemm :132, dposv :169
//   x = y + cov G^T alpha                   dgemv :176
//
// Error codes per problem follow the reference: cov not PD -> its pivot index and out untouched;
// G cov G^T not PD -> its pivot index with the unprojected y in out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "launchers.h"

namespace {

constexpr int MAXK2 = 16;
constexpr int MAXT = 16;  // tile rows (k <= 128)

__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}

// variants taking the 32-bit shared-window address directly (no generic -> shared conversion per copy)
__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8_s(unsigned smem_addr, const void* gsrc)
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

// Shared-memory geometry, the same on host and device.
// Tile row ti of the packed triangle holds rows 8 ti .. 8 ti + 7, columns 8 ti .. kpad - 1, with a leading
// dimension = 8 (mod 16) doubles: the 16-byte accumulator accesses of the trailing update are conflict-free.
struct SmallGeom {
    int kpad, ntiles, ldg, k2p, kc, nw;
    int a_doubles;
    size_t doubles;
};
__host__ __device__ inline int tile_ld(int ntiles, int ti)
{
    const int m = ntiles - ti;
    return 8 * m + ((m & 1) ? 0 : 8);
}
__host__ __device__ inline SmallGeom small_geom(int k, int k2, int nthreads)
{
    SmallGeom g;
    g.kpad = (k + 7) & ~7;
    g.ntiles = g.kpad >> 3;
    g.ldg = g.kpad + ((4 - g.kpad) & 15);  // 4 (mod 16): DMMA operand reads of G rows / the panel are conflict-free
    g.k2p = (k2 + 7) & ~7;
    g.kc = k2 <= 4 ? 4 : g.k2p;             // row length of cov G^T
    g.nw = nthreads / 32;
    g.a_doubles = 0;
    for (int t = 0; t < g.ntiles; t++) g.a_doubles += 8 * tile_ld(g.ntiles, t);
    g.doubles = (size_t)g.a_doubles + (size_t)k2 * g.ldg + (size_t)g.kpad * g.kc + 8 * (size_t)g.ldg +
                2 * (size_t)g.kpad + (size_t)g.nw * g.kc * g.kc + (size_t)g.nw * MAXK2 + MAXK2 + 2 * 8 * 12;
    return g;
}

// PROF (HTN_SMALL_PROF=1, development only): thread 0 of every CTA adds the clock64() span of each phase to prof[]
#define PHASE(i)                                               \
    if (PROF && tid == 0) {                                    \
        const long long t_now = clock64();                     \
        atomicAdd((unsigned long long*)&prof[i], (unsigned long long)(t_now - t_prev)); \
        t_prev = t_now;                                        \
    }

// 6 CTAs of 64 threads per SM: 170 registers per thread, enough for the 8 x 8 factor and its inverse without
// spills (7 would fit in shared memory but spill in the pivot chain; measured slower)
#ifndef HTN_SMALL_MINB
#define HTN_SMALL_MINB 6
#endif
template <int NTH, bool PROF>
__global__ void __launch_bounds__(NTH, NTH == 64 ? HTN_SMALL_MINB : 3)
ht_small_kernel(size_t nproblems, int k, int k2, const double* __restrict__ mean, const double* __restrict__ cov,
                const double* __restrict__ gm, const double* __restrict__ rv, const double* __restrict__ zs,
                int diag, int lower_src, int vec_ok, double* __restrict__ out, int* __restrict__ info, long long* prof)
{
    constexpr int NW = NTH / 32;
    extern __shared__ __align__(16) double sm[];
    long long t_prev = PROF ? clock64() : 0;
    const SmallGeom ge = small_geom(k, k2, NTH);
    const int kpad = ge.kpad, ntiles = ge.ntiles, ldg = ge.ldg, k2p = ge.k2p, kc = ge.kc;
    double* A = sm;                              // packed upper triangle: cov, then U in place
    double* G = A + ge.a_doubles;                // [k2][ldg]
    double* CG = G + (size_t)k2 * ldg;           // [kpad][kc]   row i = (cov * G^T)[i][:]
    double* P = CG + (size_t)kpad * kc;          // [8][ldg]     current row panel of U
    double* z = P + 8 * (size_t)ldg;             // [kpad]
    double* y = z + kpad;                        // [kpad]
    double* S2p = y + kpad;                      // [NW][kc][kc]    per-warp partial sums of G cov G^T
    double* gyp = S2p + (size_t)NW * kc * kc;    // [NW][MAXK2]     per-warp partial sums of G y
    double* alpha = gyp + NW * MAXK2;            // [MAXK2]         (k2 > 4 only)
    double* Tsm = alpha + MAXK2;                 // [2][8][12]      inverse of the diagonal block's factor (block b in half b & 1)
    __shared__ int s_fail, s_info2;
    __shared__ int toff[MAXT], tld[MAXT];
    // every tile (ti <= tj) of the triangle, row by row: {8 ti, 8 tj, A index of its element (0, 0), leading dimension}
    __shared__ int4 tiles[MAXT * (MAXT + 1) / 2];
    __shared__ int rowstart[MAXT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    // zero everything once: padding rows / columns must stay zero for the tensor-core tiles
    for (size_t idx = tid; idx < ge.doubles; idx += NTH) sm[idx] = 0.0;
    if (tid < ntiles) {
        int off = 0;
        for (int t = 0; t < tid; t++) off += 8 * tile_ld(ntiles, t);
        toff[tid] = off;
        tld[tid] = tile_ld(ntiles, tid);
        int first = 0;
        for (int t = 0; t < tid; t++) first += ntiles - t;
        rowstart[tid] = first;
        if (tid == ntiles - 1) rowstart[ntiles] = first + 1;
        for (int tj = tid; tj < ntiles; tj++)
            tiles[first + tj - tid] = make_int4(8 * tid, 8 * tj, off + (tj - tid) * 8, tile_ld(ntiles, tid));
    }
    __syncthreads();
    // this warp's tile rows in the tensor-core phases: warp, warp + NW, ... (at most 4: ntiles <= 4 NW)
    int trq[4], rbase[4];  // rbase: A index of (row, column 0) for the direct reads of that row
#pragma unroll
    for (int q = 0; q < 4; q++) {
        trq[q] = warp + NW * q;
        const int tr = trq[q] < ntiles ? trq[q] : 0;
        rbase[q] = toff[tr] + g * tld[tr] - 8 * tr;
    }

    for (size_t pb = blockIdx.x; pb < nproblems; pb += gridDim.x) {
        const double* covp = cov + pb * (size_t)k * k;
        const double* gp = gm + pb * (size_t)k2 * k;
        // ---- stage inputs: asynchronous copies global -> shared (LDGSTS), everything in flight at once ----
        if (diag) {
            for (int i = tid; i < k; i += NTH) cp_async8(&A[toff[i >> 3] + (i & 7) * tld[i >> 3] + (i & 7)], covp + (size_t)i * k + i);
        } else if (lower_src) {  // element (i, j <= i) is the valid one: park it at its mirror position (j, i)
            for (int i = warp; i < k; i += NW)
                for (int j = lane; j <= i; j += 32)
                    cp_async8(&A[toff[j >> 3] + (j & 7) * tld[j >> 3] + i - (j & ~7)], covp + (size_t)i * k + j);
        } else {
            // tile row by tile row, its 8 rows dealt to the warps.  Everything is carried along incrementally
            // (source pointer, 32-bit shared address): about six instructions per row of the matrix.
            const unsigned a_s = (unsigned)__cvta_generic_to_shared(A);
            if (vec_ok) {  // from the even column at or left of the diagonal, 16 bytes at a time
                // rows warp, warp + NW, ...: the row, its first column (row & ~1) and the destination all advance
                // by constants, so a row costs a handful of instructions
                const int wodd = warp & 1;
                const double* sp = covp + (size_t)warp * k + (warp - wodd) + 2 * lane;
                int col = (warp - wodd) + 2 * lane, i = warp;
                const size_t sstep = (size_t)NW * k + NW;
                for (int ti = 0; ti < ntiles; ti++) {
                    const int lt = tld[ti];
                    unsigned d = a_s + 8u * (unsigned)(toff[ti] + warp * lt + (warp - wodd) + 2 * lane);
                    const unsigned dstep = 8u * (unsigned)(NW * lt + NW);
#pragma unroll
                    for (int s8 = 0; s8 < 8 / NW; s8++) {
                        if (i < k && col < k) cp_async16_s(d, sp);
                        if (k > 64 && i < k && col + 64 < k) cp_async16_s(d + 512u, sp + 64);
                        d += dstep;
                        sp += sstep;
                        col += NW;
                        i += NW;
                    }
                }
            } else {
                for (int ti = 0; ti < ntiles; ti++) {
                    const int lt = tld[ti], c0 = 8 * ti, off = toff[ti];
                    for (int rr = warp; rr < 8; rr += NW) {
                        const int i = c0 + rr;
                        if (i < k)
                            for (int j = i + lane; j < k; j += 32)
                                cp_async8_s(a_s + 8u * (unsigned)(off + rr * lt + j - c0), covp + (size_t)i * k + j);
                    }
                }
            }
        }
        if (lower_src) {  // HTN_FLAG_COLMAJOR: g is k2 x k column-major, i.e. element (c, j) at j * k2 + c
            for (int idx = tid; idx < k2 * k; idx += NTH) {
                const int j = idx / k2, c = idx - j * k2;
                cp_async8(&G[c * ldg + j], gp + idx);
            }
        } else if (vec_ok) {
            for (int idx = tid; idx < k2 * (k >> 1); idx += NTH) {
                const int c = idx / (k >> 1), j = 2 * (idx - c * (k >> 1));
                cp_async16(&G[c * ldg + j], gp + (size_t)c * k + j);
            }
        } else {
            for (int c = warp; c < k2; c += NW)
                for (int j = lane; j < k; j += 32) cp_async8(&G[c * ldg + j], gp + (size_t)c * k + j);
        }
        if (vec_ok) {
            for (int j = 2 * tid; j < k; j += 2 * NTH) cp_async16(&z[j], zs + pb * (size_t)k + j);
        } else {
            for (int i = tid; i < k; i += NTH) cp_async8(&z[i], zs + pb * (size_t)k + i);
        }
        asm volatile("cp.async.commit_group;\n" ::);
        // mean and r are only needed at the very end of the problem: plain loads into registers, issued now
        double mreg[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int row = trq[q] * 8 + (lane & 7);
            mreg[q] = (lane < 8 && trq[q] < ntiles && row < k) ? mean[pb * (size_t)k + row] : 0.0;
        }
        double rreg[4] = {0.0, 0.0, 0.0, 0.0};
        if (k2 <= 4) {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < k2) rreg[c] = rv[pb * (size_t)k2 + c];
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
        if (tid == 0) { s_fail = 0; s_info2 = 0; }
        __syncthreads();
        PHASE(0)
        // ---- cov * G^T (before the factorisation overwrites the triangle), and this warp's share of G cov G^T ----
        if (diag) {
            for (int q = 0; q < 4; q++) {
                if (trq[q] >= ntiles) break;
                for (int e = lane; e < 8 * k2; e += 32) {
                    const int c = e >> 3, i = trq[q] * 8 + (e & 7);
                    CG[i * kc + c] = i < k ? A[toff[i >> 3] + (i & 7) * tld[i >> 3] + (i & 7)] * G[c * ldg + i] : 0.0;
                }
            }
        } else {
            // tensor cores: CG (k x k2) = cov (k x k) * G^T for the (up to) four 8-row tiles of this warp at once:
            // eight independent accumulator chains.  cov(row, col) lives at A(min, max): a transposed read left
            // of the diagonal, a direct read right of it - chosen per lane by selecting the ADDRESS, no branches.
            for (int tc = 0; tc < k2p / 8; tc++) {
                const bool grow = tc * 8 + g < k2;
                const double* gr = G + (grow ? (tc * 8 + g) * ldg : 0) + t4;
                double acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
#pragma unroll 2
                for (int kb = 0; kb < ntiles; kb++) {
                    const double b0 = grow ? gr[kb * 8] : 0.0, b1 = grow ? gr[kb * 8 + 4] : 0.0;
                    const int lk = tld[kb];
                    const int tb = toff[kb] + t4 * lk - 8 * kb;  // A index of (8 kb + t4, column 0)
                    const int c0 = kb * 8 + t4;                  // column of the first k half
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int tr = trq[q] < ntiles ? trq[q] : 0;  // out-of-range tiles recompute tile 0
                        const int row = tr * 8 + g;
                        const int i0 = c0 < row ? tb + row : rbase[q] + c0;
                        const int i1 = c0 + 4 < row ? tb + 4 * lk + row : rbase[q] + c0 + 4;
                        dmma_acc(acc[q][0], acc[q][1], A[i0], b0);
                        dmma_acc(acc[q][2], acc[q][3], A[i1], b1);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int col = tc * 8 + 2 * t4;
                    if (trq[q] < ntiles && col < kc) {  // kc is even: the pair is in or out together
                        const int row = trq[q] * 8 + g;
                        CG[row * kc + col] = acc[q][0] + acc[q][2];
                        CG[row * kc + col + 1] = acc[q][1] + acc[q][3];
                    }
                }
            }
        }
        __syncwarp();
        // S2p[warp] = G[:, rows of this warp] * CG[rows of this warp, :]
        for (int ta = 0; ta < k2p / 8; ta++)
            for (int tn = 0; tn < k2p / 8; tn++) {
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                const bool arow = ta * 8 + g < k2, bcol = tn * 8 + g < kc;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (trq[q] < ntiles) {
                        const int i0 = trq[q] * 8 + t4;
                        const double a0 = arow ? G[(ta * 8 + g) * ldg + i0] : 0.0;
                        const double a1 = arow ? G[(ta * 8 + g) * ldg + i0 + 4] : 0.0;
                        const double b0 = bcol ? CG[i0 * kc + tn * 8 + g] : 0.0;
                        const double b1 = bcol ? CG[(i0 + 4) * kc + tn * 8 + g] : 0.0;
                        dmma_acc(c0, c1, a0, b0);
                        dmma_acc(d0, d1, a1, b1);
                    }
                }
                const int sr = ta * 8 + g, sc = tn * 8 + 2 * t4;
                if (sr < kc && sc < kc) {  // kc is even: the pair is in or out together
                    double* sp = S2p + (size_t)warp * kc * kc + sr * kc + sc;
                    sp[0] = c0 + d0;
                    sp[1] = c1 + d1;
                }
            }
        __syncthreads();
        PHASE(2)
        // ---- Cholesky cov = U^T U in 8-row panels, factor warp + update warps ----
        //   warp 0 factors the 8 x 8 diagonal block in registers (all lanes alike) together with the inverse T of
        //   its factor, and lane 0 parks T in shared memory;
        //   the panel rows U = T * (rows of cov) are formed with DMMA, eight columns per tile, and the trailing update
        //   (rank 8, DMMA) follows; both are split by role: warp 0 forms only the panel tile right of the diagonal,
        //   updates just the NEXT diagonal block with it and factors that at once, while the other warps form the rest of
        //   the panel row and update every other tile - the 8-pivot dependency chain of the next block runs beside all
        //   the bulk work of the pass instead of after it (look-ahead), one block barrier per pass.
        if (!diag) {
            auto factor_block = [&](int bb) {  // warp 0 only
                const double* Ab = A + toff[bb];
                const int lb = tld[bb], j0 = 8 * bb;
                double l[8][8];  // l[r][c] = U[j0 + c][j0 + r] (c <= r)
                double w[8];     // column (lane & 7) of the inverse of the lower factor: forward substitution on e_c
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        l[r][c] = (c <= r) ? Ab[c * lb + r] : 0.0;
                    }
#pragma unroll
                for (int r = 0; r < 8; r++) w[r] = (r == (lane & 7)) ? 1.0 : 0.0;
                int bad = -1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    double ajj = l[j][j];
                    if (j0 + j >= k) ajj = 1.0;  // padding columns of the last panel
                    if (ajj <= 0.0 && bad < 0) bad = j;  // NaN passes
                    const double rinv = fast_rsqrt(ajj);
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (r > j) l[r][j] *= rinv;
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++)
                            if (r > j && c > j && c <= r) l[r][c] = fma(-l[r][j], l[c][j], l[r][c]);
                    // forward substitution on this lane's unit vector (off the critical path): 36 FP64 instructions for
                    // the whole inverse instead of 120 when every lane carried all of it
                    w[j] *= rinv;
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (r > j) w[r] = fma(-l[r][j], w[j], w[r]);
                }
                if (lane == 0 && bad >= 0) s_fail = j0 + bad + 1;
                if (lane < 8) {  // lane c publishes column c of T (zeros above the diagonal come out by themselves)
#pragma unroll
                    for (int r = 0; r < 8; r++) Tsm[(bb & 1) * 96 + r * 12 + lane] = w[r];
                }
            };
            if (warp == 0) factor_block(0);
            __syncthreads();
            // One pass per 8-row panel, one block barrier per pass.  The pivot chain runs on warp 0 alone:
            //   factor(b) -> panel tile (b, b+1) -> update of the diagonal block b+1 -> factor(b+1),
            // while the other warps form the rest of panel row b and apply the rank-8 update to every other tile.  The
            // only hand-off inside a pass is tile (b, b+1) of the panel row (named barrier 1: warp 0 arrives, the others wait).
            for (int b = 0; b < ntiles && !s_fail; b++) {
                const int j0 = 8 * b;
                double* Ab = A + toff[b];
                const int lb = tld[b];
                // (T is double-buffered: warp 0 parks T of block b+1 during this pass, whenever it gets there)
                const double ta0 = Tsm[(b & 1) * 96 + g * 12 + t4], ta1 = Tsm[(b & 1) * 96 + g * 12 + 4 + t4];
                double* pd = P + g * ldg + j0 + 2 * t4;
                const double* bs = Ab + t4 * lb + g;
                double* cd = Ab + g * lb + 2 * t4;
                if (warp == 0) {
                    if (b + 1 < ntiles) {
                        {  // panel tile (b, b+1)
                            const double x0 = bs[8], x1 = bs[4 * lb + 8];
                            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                            dmma_acc(c0, c1, ta0, x0);
                            dmma_acc(e0, e1, ta1, x1);
                            const double2 v = make_double2(c0 + e0, c1 + e1);
                            *reinterpret_cast<double2*>(cd + 8) = v;
                            *reinterpret_cast<double2*>(pd + 8) = v;
                        }
                        __threadfence_block();
                        __syncwarp();
                        asm volatile("bar.arrive 1, %0;\n" ::"n"(NTH) : "memory");
                        // next diagonal block: update, then factor
                        const int ti = b + 1;
                        const double a0 = -P[t4 * ldg + ti * 8 + g], a1 = -P[(4 + t4) * ldg + ti * 8 + g];
                        const double b0 = P[t4 * ldg + ti * 8 + g], b1 = P[(4 + t4) * ldg + ti * 8 + g];
                        double2* cp = reinterpret_cast<double2*>(A + toff[ti] + g * tld[ti] + 2 * t4);
                        double2 c = *cp;
                        double e0 = 0.0, e1 = 0.0;
                        dmma_acc(c.x, c.y, a0, b0);
                        dmma_acc(e0, e1, a1, b1);
                        *cp = make_double2(c.x + e0, c.y + e1);
                        __syncwarp();
                        factor_block(ti);
                    }
                } else {
                    // panel rows: U[j0 .. j0+7][columns of tile tj] = T * cov-rows
                    if (warp == 1) {  // the diagonal block is symmetric: element (row, col) lives at (min, max)
                        const double b0 = g < t4 ? Ab[g * lb + t4] : Ab[t4 * lb + g];
                        const double b1 = g < 4 + t4 ? Ab[g * lb + 4 + t4] : Ab[(4 + t4) * lb + g];
                        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                        dmma_acc(c0, c1, ta0, b0);
                        dmma_acc(e0, e1, ta1, b1);
                        const double2 v = make_double2(c0 + e0, c1 + e1);
                        *reinterpret_cast<double2*>(Ab + g * lb + 2 * t4) = v;
                        *reinterpret_cast<double2*>(pd) = v;
                    }
                    // tiles b+2 .. of the row, dealt to the update warps, two per pass
                    constexpr int NU = NW - 1;
                    int tj = b + 2 + (warp - 1);
                    for (; tj + NU < ntiles; tj += 2 * NU) {
                        const int ca = (tj - b) * 8, cb2 = ca + 8 * NU;
                        const double x0 = bs[ca], x1 = bs[4 * lb + ca], y0 = bs[cb2], y1 = bs[4 * lb + cb2];
                        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, h0 = 0.0, h1 = 0.0;
                        dmma_acc(c0, c1, ta0, x0);
                        dmma_acc(f0, f1, ta0, y0);
                        dmma_acc(e0, e1, ta1, x1);
                        dmma_acc(h0, h1, ta1, y1);
                        const double2 v = make_double2(c0 + e0, c1 + e1), w2 = make_double2(f0 + h0, f1 + h1);
                        *reinterpret_cast<double2*>(cd + ca) = v;
                        *reinterpret_cast<double2*>(pd + ca) = v;
                        *reinterpret_cast<double2*>(cd + cb2) = w2;
                        *reinterpret_cast<double2*>(pd + cb2) = w2;
                    }
                    if (tj < ntiles) {
                        const int ca = (tj - b) * 8;
                        const double x0 = bs[ca], x1 = bs[4 * lb + ca];
                        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                        dmma_acc(c0, c1, ta0, x0);
                        dmma_acc(e0, e1, ta1, x1);
                        const double2 v = make_double2(c0 + e0, c1 + e1);
                        *reinterpret_cast<double2*>(cd + ca) = v;
                        *reinterpret_cast<double2*>(pd + ca) = v;
                    }
                    if (b + 1 < ntiles) {
                        // the whole panel row (warp 0's tile included) must be in P before the update reads it
                        asm volatile("bar.sync 1, %0;\n" ::"n"(NTH) : "memory");
                        // every tile (ti <= tj) beyond the panel except (b+1, b+1): the flat tile list from the entry
                        // after (b+1, b+1) to its end, four tiles per pass dealt to the update warps; all operands of a
                        // pass are loaded before its (independent) DMMAs are issued
                        const double* pb0 = P + t4 * ldg + g;
                        const double* pb1 = pb0 + 4 * ldg;
                        const int tend = rowstart[ntiles];
                        for (int e0i = rowstart[b + 1] + 1 + 4 * (warp - 1); e0i < tend; e0i += 4 * (NW - 1)) {
                            int4 te[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) te[u] = tiles[min(e0i + u, tend - 1)];
                            double a0[4], a1[4], b0[4], b1[4];
                            double2 c[4];
                            double* cp[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                a0[u] = -pb0[te[u].x];
                                a1[u] = -pb1[te[u].x];
                                b0[u] = pb0[te[u].y];
                                b1[u] = pb1[te[u].y];
                                cp[u] = A + te[u].z + g * te[u].w + 2 * t4;
                                c[u] = *reinterpret_cast<const double2*>(cp[u]);
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                double e0 = 0.0, e1 = 0.0;  // second, independent chain
                                dmma_acc(c[u].x, c[u].y, a0[u], b0[u]);
                                dmma_acc(e0, e1, a1[u], b1[u]);
                                c[u].x += e0;
                                c[u].y += e1;
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                if (e0i + u < tend) *reinterpret_cast<double2*>(cp[u]) = c[u];
                        }
                    }
                }
                __syncthreads();
                PHASE(4)
            }
            if (s_fail) {  // reference returns before drawing (src/htnorm_distributions.c:78); out untouched
                if (tid == 0 && info) info[pb] = s_fail;
                __syncthreads();
                continue;
            }
        }
        // ---- y = mean + U^T z   /   mean + sqrt(cov_ii) z, for this warp's rows ----
        if (diag) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int row = trq[q] * 8 + (lane & 7);
                if (lane < 8 && trq[q] < ntiles && row < k)
                    y[row] = __dadd_rn(mreg[q], __dmul_rn(sqrt(A[toff[row >> 3] + (row & 7) * tld[row >> 3] + (row & 7)]), z[row]));
            }
        } else {
            // triangular product on the tensor cores: z is column 0 of a (k x 8) right-hand side; the operand
            // is U transposed, so row i of the tile reads column i of U; entries below the diagonal of the
            // diagonal block are masked
            const double* zg = z + t4;
            double acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
            int trmax = warp;
#pragma unroll
            for (int q = 1; q < 4; q++)
                if (trq[q] < ntiles) trmax = trq[q];
#pragma unroll 2
            for (int kb = 0; kb <= trmax; kb++) {
                const double z0 = g == 0 ? zg[kb * 8] : 0.0, z1 = g == 0 ? zg[kb * 8 + 4] : 0.0;
                const int lk = tld[kb];
                const int tb = toff[kb] + t4 * lk - 8 * kb;
                const int c0 = kb * 8 + t4;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int tr = trq[q] < ntiles ? trq[q] : 0;
                    if (kb <= tr) {  // warp-uniform
                        const int row = tr * 8 + g;
                        // U(c, row) for c <= row, zero below the diagonal of the diagonal block
                        const double a0 = c0 <= row ? A[tb + row] : 0.0;
                        const double a1 = c0 + 4 <= row ? A[tb + 4 * lk + row] : 0.0;
                        dmma_acc(acc[q][0], acc[q][1], a0, z0);
                        dmma_acc(acc[q][2], acc[q][3], a1, z1);
                    }
                }
            }
            // column 0 of row g sits in the t4 == 0 lane; hand it to lane g (the lane that loaded the mean)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double v = __shfl_sync(0xffffffffu, acc[q][0] + acc[q][2], (lane & 7) * 4);
                const int row = trq[q] * 8 + (lane & 7);
                if (lane < 8 && trq[q] < ntiles && row < k) y[row] = v + mreg[q];
            }
        }
        __syncwarp();
        // this warp's share of G y: lane -> one of its (up to 32) rows, butterfly sum per constraint
        {
            const int q = lane >> 3, row = (warp + NW * q) * 8 + (lane & 7);
            const bool ok = (warp + NW * q) < ntiles && row < k;
            const double yv = ok ? y[row] : 0.0;
            for (int c = 0; c < k2; c++) {
                double p = ok ? G[c * ldg + row] * yv : 0.0;
#pragma unroll
                for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) gyp[warp * MAXK2 + c] = p;
            }
        }
        __syncthreads();
        PHASE(5)
        // ---- k2 x k2 solve (dposv, src/htnorm.c:169) ----
        int f2 = 0;
        double al[4] = {0.0, 0.0, 0.0, 0.0};
        if (k2 <= 4) {
            // every thread solves the tiny system in registers (broadcast reads): no hand-off, no barrier.
            // Unused rows / columns are an identity block.
            double s[4][4], bv[4], ri[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double gsum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; w++) gsum += gyp[w * MAXK2 + i];
                bv[i] = i < k2 ? rreg[i] - gsum : 0.0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j <= i) {
                        double v = 0.0;
#pragma unroll
                        for (int w = 0; w < NW; w++) v += S2p[(size_t)w * kc * kc + i * kc + j];
                        s[i][j] = (i < k2) ? v : (i == j ? 1.0 : 0.0);
                    }
            }
            if (k2 == 1) {
                al[0] = bv[0] / s[0][0];  // the reference divides by the scalar (src/htnorm.c:77)
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
                for (int i = 0; i < 4; i++) {  // forward: L w = b
                    double v = bv[i];
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < i) v = fma(-s[i][t], bv[t], v);
                    bv[i] = v * ri[i];
                }
#pragma unroll
                for (int i = 3; i >= 0; i--) {  // backward: L^T alpha = w
                    double v = bv[i];
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t > i) v = fma(-s[t][i], bv[t], v);
                    bv[i] = v * ri[i];
                }
#pragma unroll
                for (int i = 0; i < 4; i++) al[i] = bv[i];
            }
            if (tid == 0 && info) info[pb] = f2;
        } else {
            if (tid == 0) {
                double* S = S2p;  // sum the partials into warp 0's block, factor there
                for (int i = 0; i < k2; i++) {
                    double gsum = 0.0;
                    for (int w = 0; w < NW; w++) gsum += gyp[w * MAXK2 + i];
                    alpha[i] = rv[pb * (size_t)k2 + i] - gsum;
                    for (int j = 0; j <= i; j++) {
                        double v = 0.0;
                        for (int w = 0; w < NW; w++) v += S2p[(size_t)w * kc * kc + i * kc + j];
                        S[i * kc + j] = v;
                    }
                }
                for (int j = 0; j < k2 && !f2; j++) {
                    double ajj = S[j * kc + j];
                    for (int t = 0; t < j; t++) ajj -= S[j * kc + t] * S[j * kc + t];
                    if (ajj <= 0.0) { f2 = j + 1; break; }
                    const double rinv = fast_rsqrt(ajj);
                    S[j * kc + j] = rinv;  // the reciprocal of the pivot
                    for (int i = j + 1; i < k2; i++) {
                        double sv = S[i * kc + j];
                        for (int t = 0; t < j; t++) sv -= S[i * kc + t] * S[j * kc + t];
                        S[i * kc + j] = sv * rinv;
                    }
                }
                if (!f2) {
                    for (int i = 0; i < k2; i++) {
                        double sv = alpha[i];
                        for (int t = 0; t < i; t++) sv -= S[i * kc + t] * alpha[t];
                        alpha[i] = sv * S[i * kc + i];
                    }
                    for (int i = k2 - 1; i >= 0; i--) {
                        double sv = alpha[i];
                        for (int t = i + 1; t < k2; t++) sv -= S[t * kc + i] * alpha[t];
                        alpha[i] = sv * S[i * kc + i];
                    }
                }
                s_info2 = f2;
                if (info) info[pb] = f2;
            }
            __syncthreads();
            f2 = s_info2;
        }
        // ---- x = y + cov G^T alpha ----
        for (int i = tid; i < k; i += NTH) {
            double acc = y[i];
            if (!f2) {
                double p = 0.0;
                if (k2 <= 4) {
#pragma unroll
                    for (int c = 0; c < 4; c++) p = fma(CG[i * kc + c], al[c], p);  // columns >= k2 are zero
                } else {
                    for (int c = 0; c < k2; c++) p = fma(CG[i * kc + c], alpha[c], p);
                }
                acc += p;
            }
            out[pb * (size_t)k + i] = acc;
        }
        __syncthreads();
        PHASE(7)
    }
}
#undef PHASE

template <int NTH>
int launch_small(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean, const double* cov,
                 const double* g, const double* r, const double* zs, int diag, int lower_src, double* out, int* info)
{
    const SmallGeom ge = small_geom(k, k2, NTH);
    const size_t smem = ge.doubles * 8;
    // 16-byte copies need an even row length and 16-byte aligned arrays (torch / cudaMalloc give 256)
    const int vec_ok = (k % 2 == 0) && ((((uintptr_t)cov | (uintptr_t)g | (uintptr_t)zs) & 15) == 0);
    static const bool want_prof = getenv("HTN_SMALL_PROF") != nullptr;
    auto kern = want_prof ? ht_small_kernel<NTH, true> : ht_small_kernel<NTH, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTH, smem);
    if (per_sm < 1) per_sm = 1;
    if (const char* e = getenv("HTN_SMALL_CTAS")) per_sm = atoi(e) < per_sm ? (atoi(e) > 0 ? atoi(e) : 1) : per_sm;  // experiments
    size_t grid = (size_t)sms * per_sm;  // persistent: every resident slot loops over problems
    if (grid > nproblems) grid = nproblems;
    long long* prof = nullptr;
    if (want_prof) {
        if (cudaMalloc(&prof, 16 * sizeof(long long)) != cudaSuccess) return -201;
        cudaMemsetAsync(prof, 0, 16 * sizeof(long long), st);
    }
    kern<<<(unsigned)grid, NTH, smem, st>>>(nproblems, k, k2, mean, cov, g, r, zs, diag, lower_src, vec_ok, out, info,
                                            prof);
    htnk_count_launch();
    if (want_prof) {  // development aid: synchronous on purpose
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        static const char* names[8] = {"stage", "-", "covG+S2", "panel", "update+factor", "trmv+Gy", "-", "solve+out"};
        fprintf(stderr, "[ht_small] %d threads, %d CTAs/SM (%zu B smem), clk per problem:", NTH, per_sm, smem);
        for (int i = 0; i < 8; i++) fprintf(stderr, " %s %.0f", names[i], (double)h[i] / (double)nproblems);
        fprintf(stderr, "\n");
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

}  // namespace

// zs: nproblems x k normals already drawn (reverse-fill order), one stream per problem
extern "C" int htnk_ht_small_batched(cudaStream_t st, size_t nproblems, int k, int k2, const double* mean,
                                     const double* cov, const double* g, const double* r, const double* zs,
                                     int diag, int lower_src, double* out, int* info)
{
    if (nproblems == 0) return 0;
    if (k < 1 || k > 128 || k2 < 1 || k2 > MAXK2) return -202;
    if (k <= 64) return launch_small<64>(st, nproblems, k, k2, mean, cov, g, r, zs, diag, lower_src, out, info);
    return launch_small<128>(st, nproblems, k, k2, mean, cov, g, r, zs, diag, lower_src, out, info);
}
