// Batched MID-SIZE independent problems (128 < k <= 1024, one sample per problem, each with its own mean / cov / G / r).
// Round 1 ran one full shared-path call per problem in a host loop (>= 0.5 ms each); here ALL problems go through one launch
// sequence (hyperplane.c ht_independent_mid), a LEFT-LOOKING blocked Cholesky with 64-column panels:
//     chol_rows_kernel<true>    diagonal block of the panel: cov block - L(rows, :j0) L(rows, :j0)^T            (this file)
//     potf64_kernel             its factorisation, ONE WARP per problem (the scheme of small_warp.cu)          (this file)
//     chol_rows_kernel<false>   rows below: (cov chunk - L(chunk, :j0) L(panel, :j0)^T) L_jj^-T, stored as L   (this file)
// and at the end ht_mid_project_kernel, one CTA per problem: the projection of src/htnorm.c:85-187 formed around the factor
// in two passes over L (V = L^T G^T, G cov G^T = V^T V, x = mean + L (z + V alpha)).
// Left-looking because the workspace is then written once per entry: the first version was right-looking (a batched
// rank-64 update of the whole trailing matrix per panel through the general GEMM kernel: 7.2 ms for 4096 x k = 256, 4.8 ms
// with a dedicated 64 x 64-tile update kernel) and spent its time re-reading and re-writing the trailing matrices; this
// form reads each covariance once, writes each factor once, and runs the same batch in 1.9 ms.
// Reference semantics per problem: src/htnorm.c:85-187, src/htnorm_distributions.c:58-88; error codes SURVEY Q5 / Q6.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "launchers.h"

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
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

// shared-memory geometry of small_warp.cu for an 8 x 8 grid of tiles (see there): dense swizzled 8 x 8 tiles
struct G {
    static constexpr int NT = 8, K = 64, CS = 2;
    static __host__ __device__ constexpr int tid(int p, int tj) { return p * NT - p * (p - 1) / 2 + (tj - p); }
    static constexpr int NTILES = NT * (NT + 1) / 2;
    static __host__ __device__ constexpr int didx(int q, int ti, int tj)
    {
        return (ti - q - 2) * NT - ((ti - 1) * ti - (q + 1) * (q + 2)) / 2 + (tj - ti);
    }
    static constexpr int WV = 64 * NTILES;  // [64] the pivots d
    static constexpr int SCR = WV + 64;     // [16] pivot-row broadcast, double-buffered
    static constexpr int TOTAL = SCR + 16;
};

// One warp factorises the nb x nb (nb <= 64) diagonal block at (j0, j0) of one problem's matrix (LOWER storage, row-major):
// the block is loaded as the 36 upper 8 x 8 tiles of the symmetric block in the tensor-core accumulator layout (element
// (r, c), r <= c, is stored at [c][r]; rows / columns beyond nb are padded with the identity), factorised exactly as in
// small_warp.cu, and the Cholesky factor is written back row by row as each 8-row panel finishes.
__global__ void __launch_bounds__(32, 8)
potf64_kernel(double* __restrict__ F, size_t ld, size_t sF, int j0, int nb, double* __restrict__ rdiag, size_t sr,
              int* __restrict__ info, size_t nbatch, unsigned* __restrict__ next, const double* __restrict__ src,
              size_t ssrc, size_t src_sc, size_t src_sr)
{
    constexpr int NT = G::NT, K = G::K, CS = G::CS;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int g = lane >> 2, t4 = lane & 3;
    const int swg = 4 * ((g >> 1) & 1), swt = 4 * ((t4 >> 1) & 1);
    int cf[2], tf[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        cf[e] = (g ^ e) * 8 + ((2 * t4) ^ swg);
        tf[e] = (t4 ^ e) * 8 + (g ^ swt);
    }
    int col[CS][8];
#pragma unroll
    for (int cs = 0; cs < CS; cs++) {
        const int j = lane + 32 * cs, ej = (j >> 3) & 1, cj = j & 7;
#pragma unroll
        for (int c = 0; c < 8; c++) col[cs][c] = 64 * (j >> 3) + (c ^ ej) * 8 + (cj ^ (4 * ((c >> 1) & 1)));
    }
    unsigned nxt = 0;
    if (lane == 0) nxt = atomicAdd(next, 1u);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    for (;;) {
        const size_t pb = nxt;
        if (pb >= nbatch) break;
        if (lane == 0) nxt = atomicAdd(next, 1u);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        if (info[pb] != 0) continue;  // an earlier panel of this problem failed: nothing to do (uniform)
        double* Fp = F + pb * sF;
        double* rd = rdiag + pb * sr;
        double2 c[NT][NT];
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
            for (int tj = ti; tj < NT; tj++) {
                // element (r, c), r <= c, of the block: from the lower storage of F ([c][r]), or - first panel - straight from
                // the caller's matrix (src: its valid triangle through the strides src_sc, src_sr)
                const int r = 8 * ti + g, c0 = 8 * tj + 2 * t4;
                const double* e0 = src ? src + pb * ssrc + (size_t)c0 * src_sc + (size_t)r * src_sr
                                       : Fp + (size_t)(j0 + c0) * ld + j0 + r;
                const size_t step = src ? src_sc : ld;
                c[ti][tj].x = (r < nb && c0 < nb) ? __ldg(e0) : (r == c0 ? 1.0 : 0.0);
                c[ti][tj].y = (r < nb && c0 + 1 < nb) ? __ldg(e0 + step) : (r == c0 + 1 ? 1.0 : 0.0);
            }
        // ---- A_bb = U~^T D U~, panel by panel (the scheme of small_warp.cu) ----
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
            double dpiv[8];
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
                if (ajj <= 0.0 && fail == 0) fail = j0 + 8 * p + j + 1;  // NaN passes
                dpiv[j] = ajj;
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
                    // the Cholesky factor L = (sqrt(D) U~)^T leaves for global memory: row j0 + j, columns j0 + 8p .. + 7
                    if (j >= 8 * p && j < nb) {
                        double2* lp = reinterpret_cast<double2*>(Fp + (size_t)(j0 + j) * ld + j0 + 8 * p);
#pragma unroll
                        for (int q2 = 0; q2 < 4; q2++) {
                            const double s0 = dpiv[2 * q2] * fast_rsqrt(dpiv[2 * q2]), s1 = dpiv[2 * q2 + 1] * fast_rsqrt(dpiv[2 * q2 + 1]);
                            if (8 * p + 2 * q2 + 1 < nb) lp[q2] = make_double2(v[cs][2 * q2] * s0, v[cs][2 * q2 + 1] * s1);
                            else if (8 * p + 2 * q2 < nb) reinterpret_cast<double*>(lp + q2)[0] = v[cs][2 * q2] * s0;
                        }
                    }
                }
            }
            if (lane < 8 && 8 * p + lane < nb) {  // 1 / L[j][j] for the panel solve
                double dv = dpiv[0];
#pragma unroll
                for (int r = 1; r < 8; r++)
                    if (lane == r) dv = dpiv[r];
                rd[j0 + 8 * p + lane] = fast_rsqrt(dv);
            }
            __syncwarp();
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
            }
        }
        if (fail && lane == 0) info[pb] = fail;  // cov not PD at that pivot (src/htnorm_distributions.c:78)
        __syncwarp();
    }
}

// ---- LEFT-LOOKING panel kernel: one 64-row chunk of one problem's 64-column panel per CTA ----
// chunk(i, c) = cov(i, c) - sum_{t < j0} L(i, t) L(c, t)   for the rows i of the chunk and the panel's columns c = j0 .. j0+63,
// read straight from the caller's covariance (through its UPPER triangle for the row-major build): the factor
// workspace is written once per entry and never updated in place, so a problem's HBM traffic is cov once + L once + the
// left-looking re-reads of L (against a read-modify-write of the whole trailing matrix per panel for the right-looking form).
//   DIAG:   the chunk is the diagonal block (rows j0 .. j0+63): its lower triangle is stored for potf64_kernel.
//   !DIAG:  the chunk lies below: it is solved against the factored diagonal block, X = chunk L_jj^-T, and stored as L.
// Eight warps, each 8 rows x 64 columns as eight accumulator tiles.  The panel's own rows of L (the B operand, shared by all
// warps) are staged through shared memory 64 columns of K at a time (cp.async, double-buffered); the chunk's rows (the A
// operand, private to a warp) come straight from global memory, 32 contiguous bytes per lane.  K is contracted in a permuted
// order (the lane's four consecutive values feed four successive DMMAs, the same permutation on both operands), which is what
// makes 16-byte loads possible without any shuffle; the same trick feeds accumulator tiles back as A operands in the solve.
constexpr int CR_KS = 32;               // K columns per stage
constexpr int CR_LDS = 34;              // stage row stride: 68 words = 4 mod 32, conflict-free for the 16-byte fragment reads
constexpr int CR_TILE = 64 * CR_LDS;    // one operand of one stage
constexpr int CR_STAGE = 2 * CR_TILE;   // [B rows | A rows]
constexpr int CR_NSTAGE = 2;
constexpr int CR_SMEM_DOUBLES = CR_NSTAGE * CR_STAGE;
static_assert(CR_STAGE >= 36 * 64, "the packed diagonal block reuses one buffer of the ring");

__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// one K stage of a warp's NT accumulator tiles (a branch per tile count: a predicated-off DMMA still occupies the pipe)
template <int NT>
__device__ __forceinline__ void cr_stage_mma(double (&acc)[8][2], const double* As, const double* Bs)
{
#pragma unroll
    for (int u = 0; u < CR_KS / 16; u++) {
        const double2 a01 = *reinterpret_cast<const double2*>(As + 16 * u);
        const double2 a23 = *reinterpret_cast<const double2*>(As + 16 * u + 2);
        const double a0 = -a01.x, a1 = -a01.y, a2 = -a23.x, a3 = -a23.y;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double2 b01 = *reinterpret_cast<const double2*>(Bs + 8 * n * CR_LDS + 16 * u);
            const double2 b23 = *reinterpret_cast<const double2*>(Bs + 8 * n * CR_LDS + 16 * u + 2);
            dmma(acc[n][0], acc[n][1], a0, b01.x);
            dmma(acc[n][0], acc[n][1], a1, b01.y);
            dmma(acc[n][0], acc[n][1], a2, b23.x);
            dmma(acc[n][0], acc[n][1], a3, b23.y);
        }
    }
}

// offset of the packed 8 x 8 block (tc, b), tc >= b, of the factored diagonal block
__device__ __forceinline__ int cr_blk(int tc, int b) { return (tc * (tc + 1) / 2 + b) * 64; }

template <bool DIAG>
__device__ __forceinline__ void cr_chunk(double* sm, const double* __restrict__ cp, int lower_src, double* Fp, size_t ld, int k,
                                         int j0, int chunk, const double* __restrict__ rd)
{
    const int tid = threadIdx.x, lane = tid & 31;
    // DIAG: row block rb needs rb + 1 tiles; warps w and w + 4 share a scheduler partition, so they take the row blocks
    // (w, 7 - w): nine tiles per partition
    const int warp = (DIAG && (tid >> 5) >= 4) ? 11 - (tid >> 5) : (tid >> 5);
    const int g = lane >> 2, t4 = lane & 3;
    const int nb = k - j0 < 64 ? k - j0 : 64;
    const int r0 = j0 + 64 * (DIAG ? 0 : chunk + 1);
    const int row = r0 + 8 * warp + g;
    const bool live = row < k;
    const int nstage = j0 / CR_KS;

    // stage s: columns 32 s .. 32 s + 31 of the panel's rows j0 .. j0+63 (B) and of the chunk's rows r0 .. r0+63 (A; the same
    // rows when DIAG); rows beyond k are zero-filled
    auto load_stage = [&](int s, int buf) {
        double* base = sm + buf * CR_STAGE;
#pragma unroll
        for (int it = 0; it < (DIAG ? 4 : 8); it++) {
            const int idx = tid + 256 * it, op = idx >> 10, r = (idx >> 4) & 63, c2 = (idx & 15) * 2;
            const int grow = (op ? r0 : j0) + r;
            double* dst = base + op * CR_TILE + r * CR_LDS + c2;
            if (grow < k) cp_async16(dst, Fp + (size_t)grow * ld + CR_KS * s + c2);
            else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
        }
        cp_async_commit();
    };
    // the factored diagonal block L_jj as 36 packed 8 x 8 blocks (cr_blk), into the ring buffer the K loop no longer needs:
    // issued before the last stage is computed, so its latency hides behind that stage (or behind the covariance loads)
    double* const Dm = sm + (nstage & 1) * CR_STAGE;
    auto load_diag = [&]() {
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int idx = tid + 256 * it, r = idx >> 5, c2 = (idx & 31) * 2;
            if ((c2 >> 3) <= (r >> 3)) {
                double* dst = Dm + cr_blk(r >> 3, c2 >> 3) + (r & 7) * 8 + (c2 & 7);
                if (r < nb && c2 <= r) cp_async16(dst, Fp + (size_t)(j0 + r) * ld + j0 + c2);
                else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
            }
        }
        cp_async_commit();
    };
    if (nstage > 0) load_stage(0, 0);
    if (!DIAG && nstage <= 1) load_diag();

    double acc[8][2];
#pragma unroll
    for (int n = 0; n < 8; n++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = j0 + 8 * n + 2 * t4 + e;
            const bool ok = live && c < k && (!DIAG || (n <= warp && c <= row));
            acc[n][e] = ok ? (lower_src ? cp[(size_t)row * k + c] : cp[(size_t)c * k + row]) : 0.0;
        }

    for (int s = 0; s < nstage; s++) {
        if (s + 1 < nstage) {
            load_stage(s + 1, (s + 1) & 1);
            cp_async_wait<1>();
        } else if (!DIAG && nstage > 1) {
            load_diag();  // the other buffer was released by the barrier that ended the previous iteration
            cp_async_wait<1>();
        } else if (!DIAG) {
            cp_async_wait<1>();  // nstage == 1: the diagonal block was requested after stage 0 and may still be in flight
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double* Bs = sm + (s & 1) * CR_STAGE + g * CR_LDS + 4 * t4;
        const double* As = Bs + (DIAG ? 0 : CR_TILE) + 8 * warp * CR_LDS;
        if (DIAG) {
            switch (warp) {  // warp-uniform: only the tiles on and below the diagonal
            case 0: cr_stage_mma<1>(acc, As, Bs); break;
            case 1: cr_stage_mma<2>(acc, As, Bs); break;
            case 2: cr_stage_mma<3>(acc, As, Bs); break;
            case 3: cr_stage_mma<4>(acc, As, Bs); break;
            case 4: cr_stage_mma<5>(acc, As, Bs); break;
            case 5: cr_stage_mma<6>(acc, As, Bs); break;
            case 6: cr_stage_mma<7>(acc, As, Bs); break;
            default: cr_stage_mma<8>(acc, As, Bs); break;
            }
        } else {
            cr_stage_mma<8>(acc, As, Bs);
        }
        __syncthreads();  // the buffer is refilled by the load issued in the next iteration
    }

    if (DIAG) {
        if (live) {
#pragma unroll
            for (int n = 0; n < 8; n++) {
                if (n > warp) continue;
                const int c = j0 + 8 * n + 2 * t4;
                if (c <= row && c < k) Fp[(size_t)row * ld + c] = acc[n][0];
                if (c + 1 <= row && c + 1 < k) Fp[(size_t)row * ld + c + 1] = acc[n][1];
            }
        }
        return;
    }

    // ---- X = chunk L_jj^-T ----
    cp_async_wait<0>();
    __syncthreads();
    if (8 * warp < nb) {  // warp b inverts the diagonal 8 x 8 block b IN PLACE: lane c < 8 owns column c of the inverse
        double* Db = Dm + cr_blk(warp, warp);
        rd += 8 * warp;
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) {  // right-looking: one multiply and one fused multiply-add on the chain per step
            v[t] *= (8 * warp + t < nb) ? rd[t] : 1.0;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i > t) v[i] = fma(-Db[i * 8 + t], v[t], v[i]);  // broadcast reads
        }
        __syncwarp();  // every lane has read the block before it is overwritten
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; i++) Db[i * 8 + lane] = v[i];
        }
    }
    __syncthreads();
    double* xr = Fp + (size_t)(live ? row : j0) * ld + j0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        if (8 * b < nb) {  // uniform
            const double2 iv = *reinterpret_cast<const double2*>(Dm + cr_blk(b, b) + g * 8 + 2 * t4);
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, acc[b][0], iv.x);
            dmma(x0, x1, acc[b][1], iv.y);
            const int c = 8 * b + 2 * t4;
            if (live && c + 1 < nb) *reinterpret_cast<double2*>(xr + c) = make_double2(x0, x1);
            else if (live && c < nb) xr[c] = x0;
            const double m0 = -x0, m1 = -x1;
#pragma unroll
            for (int tc = b + 1; tc < 8; tc++) {
                if (8 * tc < nb) {  // uniform
                    const double2 lv = *reinterpret_cast<const double2*>(Dm + cr_blk(tc, b) + g * 8 + 2 * t4);
                    dmma(acc[tc][0], acc[tc][1], m0, lv.x);
                    dmma(acc[tc][0], acc[tc][1], m1, lv.y);
                }
            }
        }
    }
}

// (Forming the NEXT panel's diagonal block at the end of the first chunk's CTA - its rows are complete by then - saves a launch
// per panel but was measured slower: 1103 us against 1062 us for the three panels of k = 256, and it lengthens the critical
// path of small batches.  The diagonal block keeps its own launch.)
template <bool DIAG>
__global__ void __launch_bounds__(256, 3)
chol_rows_kernel(const double* __restrict__ cov, size_t scov, int lower_src, double* F, size_t ld, size_t sF, int k, int j0,
                 const double* __restrict__ rdiag, size_t sr, const int* __restrict__ info)
{
    extern __shared__ __align__(16) double sm[];
    const size_t pb = blockIdx.y;
    if (info[pb] != 0) return;  // an earlier panel of this problem failed (uniform)
    const double* cp = cov + pb * scov;
    double* Fp = F + pb * sF;
    if (DIAG) {
        cr_chunk<true>(sm, cp, lower_src, Fp, ld, k, j0, 0, nullptr);
    } else {
        cr_chunk<false>(sm, cp, lower_src, Fp, ld, k, j0, (int)blockIdx.x, rdiag + pb * sr + j0);
    }
}

// ---- projection, one CTA per problem: L streamed through shared memory by cp.async ----
// y = mean + L z, V = L^T G^T (k x k2), S = V^T V = G cov G^T, G y = G mean + V^T z, alpha = S^-1 (r - G y),
// x = mean + L (z + V alpha): two passes over the lower factor.  The rows of L go through a ring of PJ_NBUF 16 KB row blocks
// filled by cp.async (PJ_NBUF - 1 blocks in flight per CTA, the first blocks of pass 2 already during the k2 x k2 solve), and
// the arithmetic reads shared memory:
//   pass 1 (rows ascending): thread t owns the columns t + 256 m of V and adds row i's contribution L[i][j] G[:, i] - no
//           reduction across threads at all, V never leaves the registers;
//   pass 2 (rows descending, so the tail of pass 1 is still in L2): one row per warp (k <= 256) or per 2 / 4 warps.
// Variants measured and dropped (4096 x k = 256, this kernel 0.565 ms of the 1.9 ms call): a first form with warps reading
// rows straight from global memory (0.75 ms: every warp waited for every row); single bulk copies per row counted by
// full / empty mbarrier pairs with row-length-packed blocks (0.69 ms: a quarter of the instructions, but short rows make
// small TMA requests and the ring advances at the pace of the slowest warp); a ring of four blocks (two CTAs per SM: 0.63 ms);
// seven rows per block so that four CTAs fit an SM (64 registers, 55 KB: no gain over three, 1.87 vs 1.85 ms per call, while
// one / two CTAs per SM cost 2.08 / 2.01 ms: the kernel is latency-bound up to three).
// KP = 256 MC is the row pitch of a block, RB = 2048 / KP its rows.
constexpr int PT = 256;    // threads of a projection CTA
constexpr int PJ_NBUF = 3;  // ring depth of the projection's row blocks (PJ_NBUF - 1 blocks in flight)
template <int MC>
__global__ void __launch_bounds__(PT)
ht_mid_project_kernel(size_t nbatch, int k, int k2, const double* __restrict__ F, size_t ld, size_t sF,
                       const double* __restrict__ mean, const double* __restrict__ gm, const double* __restrict__ rv,
                       const double* __restrict__ zs, int g_colmajor, double* __restrict__ out, int* __restrict__ info)
{
    constexpr int KP = 256 * MC, RB = 2048 / KP, NSEG = 8 / RB, LOGH = (MC == 1 ? 7 : MC == 2 ? 8 : 9);
    extern __shared__ __align__(16) double ps[];
    const int kq = (k + 3) & ~3;
    double* w = ps;             // [kq]
    double* Gt = w + kq;        // [kq][4]: G transposed, rows beyond k2 zero
    double* red = Gt + 4 * kq;  // [8 warps][24]
    double* pp = red + 192;     // [8]: partial row sums of pass 2
    double* bufs = pp + 8;      // [PJ_NBUF][2048]
    __shared__ double s_al[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (k + RB - 1) / RB;
    for (size_t pb = blockIdx.x; pb < nbatch; pb += gridDim.x) {
        if (info[pb] != 0) continue;  // cov not PD: out untouched (uniform across the CTA)
        const double* Lp = F + pb * sF;
        const double* mp = mean + pb * (size_t)k;
        const double* zp = zs + pb * (size_t)k;
        auto issue = [&](int q) {  // block q of the stream: pass 1 blocks ascending, then pass 2 blocks descending
            if (q < 2 * nblk) {
                const int i0 = (q < nblk ? q : 2 * nblk - 1 - q) * RB;
                // thread t copies the 16-byte piece (t mod KP/2) of the rows (t div KP/2) + (512 / KP) it; a warp whose first
                // piece lies beyond the block's last row has nothing to copy (the triangle: half of all warp-blocks)
                const int c2 = (tid & (KP / 2 - 1)) * 2, r0 = tid >> LOGH;
                if (MC == 4) {  // 512 pieces per row: two per thread and row
                    double* dst = bufs + (q % PJ_NBUF) * 2048;
#pragma unroll
                    for (int it = 0; it < 4; it++) {
                        const int r = it >> 1, c4 = (tid + 256 * (it & 1)) * 2, i = i0 + r;
                        if (i < k && c4 <= i) cp_async16(dst + r * KP + c4, Lp + (size_t)i * ld + c4);
                    }
                } else if (((tid & ~31) & (KP / 2 - 1)) * 2 <= i0 + RB - 1) {
                    double* dst = bufs + (q % PJ_NBUF) * 2048 + r0 * KP + c2;
                    const double* src = Lp + (size_t)(i0 + r0) * ld + c2;
#pragma unroll
                    for (int it = 0; it < 4; it++) {
                        const int i = i0 + r0 + (512 / KP) * it;
                        if (i < k && c2 <= i) cp_async16(dst + (512 / KP) * it * KP, src + (size_t)(512 / KP) * it * ld);
                    }
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int q = 0; q < PJ_NBUF - 1; q++) issue(q);
        for (int c = 0; c < 4; c++)
            for (int i = tid; i < k; i += PT) {
                double v = 0.0;
                if (c < k2) v = g_colmajor ? gm[pb * (size_t)k2 * k + (size_t)i * k2 + c] : gm[pb * (size_t)k2 * k + (size_t)c * k + i];
                Gt[4 * i + c] = v;
            }
        double acc[MC][4];
#pragma unroll
        for (int m = 0; m < MC; m++) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.0;
        for (int q = 0; q < nblk; q++) {
            cp_async_wait<PJ_NBUF - 2>();
            __syncthreads();
            issue(q + PJ_NBUF - 1);
            const double* B = bufs + (q % PJ_NBUF) * 2048;
            const int i0 = q * RB;
            if (32 * warp <= i0 + RB - 1)  // the block's rows hold none of this warp's columns otherwise (m = 0 is its first)
#pragma unroll
            for (int r = 0; r < RB; r++) {
                const int i = i0 + r;
                if (i < k) {  // uniform
                    const double2 g01 = *reinterpret_cast<const double2*>(Gt + 4 * i);
                    const double2 g23 = *reinterpret_cast<const double2*>(Gt + 4 * i + 2);
#pragma unroll
                    for (int m = 0; m < MC; m++) {
                        const int j = tid + 256 * m;
                        if (j <= i) {
                            const double l = B[r * KP + j];
                            acc[m][0] = fma(l, g01.x, acc[m][0]);
                            acc[m][1] = fma(l, g01.y, acc[m][1]);
                            acc[m][2] = fma(l, g23.x, acc[m][2]);
                            acc[m][3] = fma(l, g23.y, acc[m][3]);
                        }
                    }
                }
            }
        }
        // S (10 values), V^T z (4), G mean (4) over the thread's own columns, fixed-order reduction
        double part[18], zr[MC];
#pragma unroll
        for (int q = 0; q < 18; q++) part[q] = 0.0;
#pragma unroll
        for (int m = 0; m < MC; m++) {
            const int j = tid + 256 * m;
            zr[m] = 0.0;
            if (j < k) {
                int q = 0;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b <= a; b++) part[q++] += acc[m][a] * acc[m][b];
                const double zj = zp[j], mj = mp[j];
                zr[m] = zj;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    part[10 + a] += acc[m][a] * zj;
                    part[14 + a] += Gt[4 * j + a] * mj;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 18; q++) {
            double v = part[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp * 24 + q] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double t[18];
            for (int q = 0; q < 18; q++) {
                double v = 0.0;
                for (int ww = 0; ww < PT / 32; ww++) v += red[ww * 24 + q];
                t[q] = v;
            }
            double sq[4][4], bv[4], ri[4];
            int q = 0, f2 = 0;
            for (int a = 0; a < 4; a++)
                for (int b = 0; b <= a; b++, q++) sq[a][b] = a < k2 ? t[q] : (a == b ? 1.0 : 0.0);
            for (int a = 0; a < 4; a++) bv[a] = a < k2 ? rv[pb * (size_t)k2 + a] - t[14 + a] - t[10 + a] : 0.0;
            if (k2 == 1) {
                bv[0] = bv[0] / sq[0][0];  // the reference divides by the scalar (src/htnorm.c:77)
            } else {
                for (int j = 0; j < 4; j++) {
                    const double ajj = sq[j][j];
                    if (ajj <= 0.0 && f2 == 0) f2 = j + 1;
                    const double rinv = fast_rsqrt(ajj);
                    ri[j] = rinv;
                    for (int i = j + 1; i < 4; i++) sq[i][j] *= rinv;
                    for (int i = j + 1; i < 4; i++)
                        for (int cc = j + 1; cc <= i; cc++) sq[i][cc] = fma(-sq[i][j], sq[cc][j], sq[i][cc]);
                }
                for (int i = 0; i < 4; i++) {
                    double v = bv[i];
                    for (int u = 0; u < i; u++) v = fma(-sq[i][u], bv[u], v);
                    bv[i] = v * ri[i];
                }
                for (int i = 3; i >= 0; i--) {
                    double v = bv[i];
                    for (int u = i + 1; u < 4; u++) v = fma(-sq[u][i], bv[u], v);
                    bv[i] = v * ri[i];
                }
            }
            for (int a = 0; a < 4; a++) s_al[a] = f2 ? 0.0 : bv[a];  // G cov G^T not PD: the unprojected y (src/htnorm.c:170)
            info[pb] = f2;
        }
        __syncthreads();
        {
            const double a0 = s_al[0], a1 = s_al[1], a2 = s_al[2], a3 = s_al[3];
#pragma unroll
            for (int m = 0; m < MC; m++) {
                const int j = tid + 256 * m;
                if (j < k) w[j] = zr[m] + ((acc[m][0] * a0 + acc[m][1] * a1) + (acc[m][2] * a2 + acc[m][3] * a3));
            }
        }
        // pass 2: x[i] = mean[i] + sum_{j <= i} L[i][j] w[j]
        for (int q = nblk; q < 2 * nblk; q++) {
            cp_async_wait<PJ_NBUF - 2>();
            __syncthreads();  // also orders the writes of w before the first block
            issue(q + PJ_NBUF - 1);
            const double* B = bufs + (q % PJ_NBUF) * 2048;
            const int i0 = (2 * nblk - 1 - q) * RB;
            const int r = warp / NSEG, seg = warp % NSEG, i = i0 + r;
            double v = 0.0;
            if (i < k)
                for (int j = seg * 32 + lane; j <= i; j += 32 * NSEG) v = fma(B[r * KP + j], w[j], v);
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (NSEG == 1) {
                if (lane == 0 && i < k) out[pb * (size_t)k + i] = mp[i] + v;
            } else {
                if (lane == 0) pp[warp] = v;
                __syncthreads();
                if (tid < RB && i0 + tid < k) {
                    double sacc = 0.0;
#pragma unroll
                    for (int sg = 0; sg < NSEG; sg++) sacc += pp[tid * NSEG + sg];
                    out[pb * (size_t)k + i0 + tid] = mp[i0 + tid] + sacc;
                }
            }
        }
        __syncthreads();  // the ring, w and Gt are rewritten for the next problem
    }
}

}  // namespace

extern "C" int htnk_potf64_batched(cudaStream_t st, double* F, size_t ld, size_t sF, int j0, int nb, double* rdiag, size_t sr,
                                   int* info, size_t nbatch, unsigned* counter, const double* src, size_t ssrc, size_t src_ld,
                                   int src_lower)
{
    if (src && j0 != 0) return -202;  /* only the first block can come from the caller's matrix: the others need the update */
    if (nbatch == 0 || nb <= 0) return 0;
    if (nb > 64 || (ld & 1) || (j0 & 1) || (sF & 1) || nbatch > 0xfffffff0u || (reinterpret_cast<uintptr_t>(F) & 15)) return -202;
    const size_t smem = (size_t)G::TOTAL * sizeof(double);
    if (cudaFuncSetAttribute(potf64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -201;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, potf64_kernel, 32, smem);
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)sms * per_sm;
    if (grid > nbatch) grid = nbatch;
    /* element (r, c), r <= c: [c][r] of a lower-valid (column-major build) matrix, [r][c] of an upper-valid (row-major) one */
    potf64_kernel<<<(unsigned)grid, 32, smem, st>>>(F, ld, sF, j0, nb, rdiag, sr, info, nbatch, counter, src, ssrc,
                                                    src_lower ? src_ld : 1, src_lower ? 1 : src_ld);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_ht_mid_project(cudaStream_t st, size_t nbatch, int k, int k2, const double* F, size_t ld, size_t sF,
                                   const double* mean, const double* g, const double* r, const double* zs, int g_colmajor,
                                   double* out, int* info)
{
    if (nbatch == 0) return 0;
    if (k < 1 || k > 4 * PT || k2 < 1 || k2 > 4) return -202;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((ld & 1) || (sF & 1) || (reinterpret_cast<uintptr_t>(F) & 15)) return -202;
    const size_t kq = ((size_t)k + 3) & ~(size_t)3;
    const size_t smem = (5 * kq + 192 + 8 + PJ_NBUF * 2048) * sizeof(double);
    auto kern = k <= 256 ? ht_mid_project_kernel<1> : k <= 512 ? ht_mid_project_kernel<2> : ht_mid_project_kernel<4>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -201;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT, smem);
    if (per_sm < 1) per_sm = 1;
    size_t grid = (size_t)sms * per_sm;
    if (grid > nbatch) grid = nbatch;
    kern<<<(unsigned)grid, PT, smem, st>>>(nbatch, k, k2, F, ld, sF, mean, g, r, zs, g_colmajor, out, info);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

/* one panel step of the left-looking batched factorisation: diag != 0 forms the diagonal block (rows j0 .. j0+63) for
 * htnk_potf64_batched, diag == 0 forms and solves the rows below it against the factored block. */
extern "C" int htnk_chol_rows_batched(cudaStream_t st, const double* cov, size_t scov, int lower_src, double* F, size_t ld,
                                      size_t sF, int k, int j0, const double* rdiag, size_t sr, const int* info, int diag,
                                      size_t nbatch)
{
    if (nbatch == 0 || k <= 0 || j0 >= k) return 0;
    if ((ld & 1) || (sF & 1) || (j0 & 63) || (reinterpret_cast<uintptr_t>(F) & 15)) return -202;
    const int nb = k - j0 < 64 ? k - j0 : 64;
    const int nchunk = diag ? 1 : (k - j0 - nb + 63) / 64;
    if (nchunk <= 0) return 0;
    const size_t smem = (size_t)CR_SMEM_DOUBLES * sizeof(double);
    if (cudaFuncSetAttribute(chol_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(chol_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    for (size_t b0 = 0; b0 < nbatch; b0 += 65535) {
        const unsigned nbt = (unsigned)(nbatch - b0 < 65535 ? nbatch - b0 : 65535);
        const dim3 grid((unsigned)nchunk, nbt);
        if (diag)
            chol_rows_kernel<true><<<grid, 256, smem, st>>>(cov + b0 * scov, scov, lower_src, F + b0 * sF, ld, sF, k, j0,
                                                            rdiag + b0 * sr, sr, info + b0);
        else
            chol_rows_kernel<false><<<grid, 256, smem, st>>>(cov + b0 * scov, scov, lower_src, F + b0 * sF, ld, sF, k, j0,
                                                             rdiag + b0 * sr, sr, info + b0);
        htnk_count_launch();
        if (cudaGetLastError() != cudaSuccess) return -201;
    }
    return 0;
}
