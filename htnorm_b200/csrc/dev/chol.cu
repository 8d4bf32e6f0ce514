// Shared-memory / register kernels of the blocked Cholesky and of the blocked triangular solves:
//   potf2_kernel         factor one nb x nb (nb <= 128) diagonal block (dpotrf's unblocked base case;
//                        reference call sites src/htnorm_distributions.c:77,115 via src/htnorm_blas.h:111)
//   trsm_rows_kernel     panel solve L21 = A21 L11^-T on the tensor cores, rows independent
//   trtri_batched_kernel inverse of every diagonal block at once, so that the many-right-hand-side
//                        solves of the samplers (dtrsv / dposv / dpotri's solve phases,
//                        src/htnorm_distributions.c:120, src/htnorm.c:169,210,330) become DMMA GEMMs
// Pivot rule: fail on `ajj <= 0` only - NaN passes and propagates with info = 0, as the reference +
// OpenBLAS do (reference tests/test_pyhtnorm.py:46-48, 91-95).
//
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "launchers.h"

namespace {

constexpr int NB = 128;
constexpr int LDB = NB + 4;  // 132
constexpr int NT = 512;

// 1/sqrt(x): MUFU seed (rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps in FP64 (~1 ulp).  About a third
// of the latency of the library rsqrt()/sqrt()+div pair, and it sits on the critical path of EVERY column.
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double h = 0.5 * y;
        const double e = fma(-x * y, h, 0.5);  // 0.5 - 0.5 x y^2
        y = fma(y, e, y);
    }
    return y;
}

// full 128 x 128 block global -> shared by 16-byte LDGSTS (every element in flight at once); partial blocks
// (nb < 128) take the masked scalar path so that the padding stays zero
template <int NTHREADS = NT>
__device__ __forceinline__ void load_block(double* S, const double* __restrict__ a, size_t ld, int nb, int tid,
                                           bool mask_upper)
{
    if (nb == NB && (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
        for (int idx = tid; idx < NB * (NB / 2); idx += NTHREADS) {
            const int r = idx >> 6, ch = idx & 63;
            const unsigned d = (unsigned)__cvta_generic_to_shared(S + r * LDB + 2 * ch);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(a + (size_t)r * ld + 2 * ch));
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
    } else {
        for (int idx = tid; idx < NB * NB; idx += NTHREADS) {
            const int r = idx >> 7, c = idx & (NB - 1);
            S[r * LDB + c] = (r < nb && c < nb && (!mask_upper || c <= r)) ? a[(size_t)r * ld + c] : 0.0;
        }
    }
}

// potf2: the 128 x 128 block is factorised in 16 column panels of width 8, with a FACTOR WARP and look-ahead:
//   factor  : warp 0 factors the 8 x 8 diagonal sub-block in registers (every lane alike: no communication on
//             the 8-pivot chain) together with T = inverse of its factor; lane 0 parks T in shared memory;
//   panel   : L[:, panel] = A[:, panel] * T^T on the tensor cores, one 8-row tile per warp (2 DMMAs);
//   update  : rank-8 update of the trailing lower triangle (2 DMMAs per 8 x 8 tile).  Warp 0 takes only the NEXT
//             diagonal tile and goes straight on to factor it; warps 1..15 share every other tile (a flat tile
//             list, four tiles per pass, all operands loaded before the DMMAs are issued).
// The 128-step pivot chain (rsqrt + two dependent FP64 operations per column) is what remains serial; everything
// else runs beside it.  Two barriers per panel.
__device__ __forceinline__ void dmma_sub(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int LDP = 12;      // panel row stride (doubles): 8 + 4 keeps the DMMA fragment reads conflict-free
constexpr int NPAN = NB / 8;  // 16 panels
constexpr int NTP = 512;      // potf2: 16 warps, one 8-row tile of the panel each

// ---- fused panel: the row blocks under the diagonal block, solved in the same launch (CTAs 1, 2, ...) ----
// CTA 0 (the potf2 role below) publishes, per 8-column sub-panel b, the inverse T_b of the 8 x 8 factor (scratch
// `tscr`) and - by storing its tiles of L straight into the matrix - the sub-panel's columns of L11; then it
// releases flags[b] = epoch.  A row-block CTA acquires the flag, stages T_b and L11[8(b+1).., 8b..8b+7], forms
// its X_b = A_b T_b^T and updates its later columns, all on DMMA (32 rows per CTA): it trails the factorisation by about one
// sub-panel, so the panel solve (13 us as a kernel of its own) disappears behind the 128-pivot chain.
// No cluster is needed: the only producer is CTA 0, which never waits, so co-residency is irrelevant for
// progress; flags carry a launch-unique (even) epoch and are never reset; epoch | 1 means "not PD, stop".
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

constexpr int RB = 32;  // rows per row-block CTA: 32 x 128 x 128 FMAs (8 k cycles of one SM's FP64 pipe) keep pace with
                        // the 128-pivot chain of CTA 0; a 128-row block (33 k cycles) would not

__device__ void panel_rows_role(double* __restrict__ xr, size_t ld, int nrows, const double* __restrict__ adiag,
                                const double* __restrict__ tscr, const unsigned* __restrict__ flags, unsigned epoch,
                                double* sm)
{
    // Warps 12..15 are LOADERS: they acquire flags[b+1] and stage T_{b+1} and L11's sub-panel b+1 into the other
    // half of the double buffers while warps 0..11 compute sub-panel b (4 row tiles x 3 column groups).  One CTA
    // barrier per sub-panel; the compute warps meet once more, among themselves, between the solve and the update.
    double* S = sm;                       // [RB][LDB]     this CTA's rows of the panel
    double* Pb = sm + RB * LDB;           // [2][NB][LDP]  L11[:, 8b .. 8b+7] (rows > 8b+7 used), double-buffered
    double* Px = Pb + 2 * NB * LDP;       // [RB][LDP]     this CTA's X_b (A operand of its update)
    double* Tl = Px + RB * LDP;           // [2][8][12]    T_b, double-buffered
    __shared__ int s_stop[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    constexpr int NLOAD = 128, NCOMP = NTP - NLOAD;
    const bool loader = tid >= NCOMP;
    const int ltid = tid - NCOMP;
    for (int idx = tid; idx < RB * (NB / 2); idx += NTP) {
        const int r = idx >> 6, c = 2 * (idx & 63);
        if (r < nrows) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(S + r * LDB + c);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(xr + (size_t)r * ld + c));
        } else {
            *reinterpret_cast<double2*>(S + r * LDB + c) = make_double2(0.0, 0.0);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);

    auto stage = [&](int bb) {  // loaders only
        const int buf = bb & 1;
        if (ltid == 0) {
            unsigned v;
            do v = ld_acquire_u32(flags + bb); while ((v & ~1u) != epoch);  // plain spin: one L2 round trip per poll
            s_stop[buf] = (int)(v & 1u);                                     // epoch | 1: not positive definite, stop
        }
        asm volatile("bar.sync 2, %0;\n" ::"n"(NLOAD) : "memory");           // the loaders: flag seen
        if (s_stop[buf]) return;
        double* Tb = Tl + buf * 96;
        double* Pd = Pb + buf * NB * LDP;
        if (ltid < 64) Tb[(ltid >> 3) * 12 + (ltid & 7)] = __ldcg(tscr + bb * 64 + ltid);
        for (int idx = ltid; idx < (NB - 8 * (bb + 1)) * 8; idx += NLOAD) {
            const int r = 8 * (bb + 1) + (idx >> 3), c = idx & 7;
            Pd[r * LDP + c] = __ldcg(adiag + (size_t)r * ld + 8 * bb + c);
        }
    };

    if (loader) stage(0);
    __syncthreads();
    const int rt = warp & 3, cgp = warp >> 2;  // compute warps: row tile, column group (0..2)
    const int row = rt * 8 + g;
    for (int b = 0; b < NPAN; b++) {
        const int buf = b & 1;
        if (s_stop[buf]) break;
        if (loader) {
            if (b + 1 < NPAN) stage(b + 1);
        } else {
            const double* Tb = Tl + buf * 96;
            const double* Pd = Pb + buf * NB * LDP;
            if (cgp == 0) {  // X_b tile of row tile rt
                const double* ar = S + row * LDB + 8 * b;
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                dmma_sub(c0, c1, ar[t4], Tb[g * 12 + t4]);
                dmma_sub(e0, e1, ar[4 + t4], Tb[g * 12 + 4 + t4]);
                const double2 v = make_double2(c0 + e0, c1 + e1);
                *reinterpret_cast<double2*>(Px + row * LDP + 2 * t4) = v;
                if (row < nrows) *reinterpret_cast<double2*>(xr + (size_t)row * ld + 8 * b + 2 * t4) = v;  // final
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(NCOMP) : "memory");  // the compute warps: X_b is in Px
            // later columns tc = b+1+cgp, +3, ...: C[row tile][tc] -= X_b * L11[tc][b]^T, at most five tiles: one pass
            const double a0 = -Px[row * LDP + t4], a1 = -Px[row * LDP + 4 + t4];
            double* crow = S + row * LDB + 2 * t4;
            const int tc0 = b + 1 + cgp;
            double2 c[5];
            double b0[5], b1[5];
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int tc = min(tc0 + 3 * u, NPAN - 1);
                b0[u] = Pd[(tc * 8 + g) * LDP + t4];
                b1[u] = Pd[(tc * 8 + g) * LDP + 4 + t4];
                c[u] = *reinterpret_cast<const double2*>(crow + tc * 8);
            }
#pragma unroll
            for (int u = 0; u < 5; u++) {
                double e0 = 0.0, e1 = 0.0;
                dmma_sub(c[u].x, c[u].y, a0, b0[u]);
                dmma_sub(e0, e1, a1, b1[u]);
                c[u].x += e0;
                c[u].y += e1;
            }
#pragma unroll
            for (int u = 0; u < 5; u++)
                if (tc0 + 3 * u < NPAN) *reinterpret_cast<double2*>(crow + (tc0 + 3 * u) * 8) = c[u];
        }
        __syncthreads();  // sub-panel b consumed, sub-panel b+1 staged
    }
}

__global__ void __launch_bounds__(NTP, 1)
potf2_kernel(double* __restrict__ a, size_t ld, int nb, int col0, double* __restrict__ rdiag_out,
             int* __restrict__ info, int nrows_below, double* __restrict__ tscr, unsigned* __restrict__ flags,
             unsigned epoch)
{
    extern __shared__ double sm[];
    if (blockIdx.x > 0) {  // a row block under the diagonal block
        const int r0 = ((int)blockIdx.x - 1) * RB;
        panel_rows_role(a + (size_t)(NB + r0) * ld, ld, min(RB, nrows_below - r0), a, tscr, flags, epoch, sm);
        return;
    }
    double* S = sm;                  // [NB][LDB]  the block (lower triangle meaningful)
    double* P = sm + NB * LDB;       // [NB][LDP]  current panel: L[:, 8b .. 8b+7]
    double* rdiag = P + NB * LDP;    // [NB]       1 / L[j][j]
    double* Tsm = rdiag + NB;        // [8][12]    inverse of the current 8 x 8 factor
    __shared__ int s_fail;
    __shared__ short2 tiles[NPAN * (NPAN + 1) / 2];  // (8 tr, 8 tc), tr >= tc, column by column
    __shared__ int colstart[NPAN + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;

    // the strictly upper part of a full block may hold junk (diagonal tiles of earlier trailing updates): it is
    // never used - sub-block loads read c <= r, diagonal tiles mirror, the write-back masks too
    // Full, aligned blocks take the streaming path: only the lower triangle is loaded (16-byte LDGSTS), the zeros
    // above the diagonal go to global memory right away (fire-and-forget), and every 8 x 8 tile of L is stored the
    // moment its panel is final - there is no write-back phase at the end.
    const bool fastio = nb == NB && (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    if (fastio) {
        for (int idx = tid; idx < NB * (NB / 2); idx += NTP) {
            const int r = idx >> 6, c = 2 * (idx & 63);
            if (c <= r) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(S + r * LDB + c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(a + (size_t)r * ld + c));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
        for (int idx = tid; idx < NB * (NB / 2); idx += NTP) {
            const int r = idx >> 6, c = 2 * (idx & 63);
            if (c > r) *reinterpret_cast<double2*>(a + (size_t)r * ld + c) = make_double2(0.0, 0.0);
        }
    } else {
        load_block<NTP>(S, a, ld, nb, tid, true);
    }
    // set-up that does not depend on the block runs while the LDGSTS of the fast path are still in flight
    for (int idx = tid; idx < NB * LDP; idx += NTP) P[idx] = 0.0;
    for (int idx = tid; idx < 8 * 12; idx += NTP) Tsm[idx] = 0.0;
    const int npanels = (nb + 7) >> 3;
    if (tid < npanels) {
        const int first = tid * npanels - tid * (tid - 1) / 2;  // tiles of the columns before this one
        colstart[tid] = first;
        if (tid == npanels - 1) colstart[npanels] = first + 1;
        for (int tr = tid; tr < npanels; tr++) tiles[first + tr - tid] = make_short2((short)(8 * tr), (short)(8 * tid));
    }
    if (tid == 0) s_fail = 0;
    if (fastio) asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();

    auto factor_block = [&](int bb) {  // warp 0 only
        const int j0 = 8 * bb;
        const double* Sb = S + j0 * LDB + j0;
        double l[8][8];  // the sub-block / its factor (c <= r)
        double w[8];     // column (lane & 7) of the inverse of the factor: forward substitution on e_c
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) {
                l[r][c] = (c <= r) ? Sb[r * LDB + c] : 0.0;
            }
#pragma unroll
        for (int r = 0; r < 8; r++) w[r] = (r == (lane & 7)) ? 1.0 : 0.0;
        int bad = -1;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double ajj = l[j][j];
            // columns beyond nb (partial last panel) are padding: treat them as the identity
            if (j0 + j >= nb) ajj = 1.0;
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
        if (lane == 0 && bad >= 0) {
            atomicCAS(info, 0, col0 + j0 + bad + 1);
            s_fail = j0 + bad + 1;
        }
        if (lane < 8) {  // lane c publishes column c of T (zeros above the diagonal come out by themselves)
#pragma unroll
            for (int r = 0; r < 8; r++) {
                Tsm[r * 12 + lane] = w[r];
                if (tscr != nullptr) tscr[bb * 64 + r * 8 + lane] = w[r];  // for the row-block CTAs
                if (lane == r && j0 + r < nb) rdiag[j0 + r] = w[r];  // T[r][r] = 1 / L[r][r]
            }
        }
    };

    if (warp == 0) factor_block(0);
    __syncthreads();
    int nfact = nb;
    for (int b = 0; b < npanels; b++) {
        if (s_fail) {
            nfact = s_fail - 1;
            if (flags != nullptr && tid == 0) {  // wake every waiting row block and tell it to stop
                __threadfence();
                for (int q = b; q < NPAN; q++) st_release_u32(flags + q, epoch | 1u);
            }
            break;
        }
        const int j0 = 8 * b;
        // ---- panel: 8-row tile tr = b + warp, X = A_tile * T^T (the diagonal tile is read symmetrically) ----
        for (int tr = b + warp; tr < npanels; tr += NTP / 32) {
            {
                const double tb0 = Tsm[g * 12 + t4], tb1 = Tsm[g * 12 + 4 + t4];  // B[k][n] = T[n][k]
                const double* ar = S + (tr * 8 + g) * LDB + j0;
                double a0, a1;
                if (tr == b) {  // symmetric block, lower triangle valid: element (g, c) lives at (max, min)
                    a0 = t4 <= g ? ar[t4] : S[(j0 + t4) * LDB + j0 + g];
                    a1 = 4 + t4 <= g ? ar[4 + t4] : S[(j0 + 4 + t4) * LDB + j0 + g];
                } else {
                    a0 = ar[t4];
                    a1 = ar[4 + t4];
                }
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
                dmma_sub(c0, c1, a0, tb0);
                dmma_sub(e0, e1, a1, tb1);
                double2 v = make_double2(c0 + e0, c1 + e1);
                const int row = tr * 8 + g;
                if (tr == b) {  // nothing above the diagonal
                    if (2 * t4 > g) v.x = 0.0;
                    if (2 * t4 + 1 > g) v.y = 0.0;
                }
                if (row >= nb) v = make_double2(0.0, 0.0);
                if (j0 + 2 * t4 >= nb) v.x = 0.0;
                if (j0 + 2 * t4 + 1 >= nb) v.y = 0.0;
                *reinterpret_cast<double2*>(S + row * LDB + j0 + 2 * t4) = v;
                *reinterpret_cast<double2*>(P + row * LDP + 2 * t4) = v;
                if (fastio) *reinterpret_cast<double2*>(a + (size_t)row * ld + j0 + 2 * t4) = v;  // final: out it goes
            }
        }
        __syncthreads();
        // sub-panel b of L11 and T_b are in global memory: release the row blocks.  Done by a thread of warp 4, which
        // has no part in the update phase - the fence must not sit on the factor warp's chain
        if (flags != nullptr && tid == 128) {
            __threadfence();
            st_release_u32(flags + b, epoch);
        }
        // ---- rank-8 update of the trailing lower triangle on the tensor cores, with look-ahead ----
        if (b + 1 < npanels) {
            if (warp == 0) {  // next diagonal tile: update it, then factor it
                const int r0 = (b + 1) * 8 + g;
                const double pa0 = P[r0 * LDP + t4], pa1 = P[r0 * LDP + 4 + t4];
                double2* cp = reinterpret_cast<double2*>(&S[r0 * LDB + (b + 1) * 8 + 2 * t4]);
                double2 c = *cp;
                double e0 = 0.0, e1 = 0.0;
                dmma_sub(c.x, c.y, -pa0, pa0);
                dmma_sub(e0, e1, -pa1, pa1);
                *cp = make_double2(c.x + e0, c.y + e1);
                __syncwarp();
                factor_block(b + 1);
            } else if ((warp & 3) != 0) {
                // DMMA executes on the FP64 pipe of the issuing warp's scheduler partition (warp id mod 4), the same
                // pipe the pivot chain of warp 0 lives on: the update is left to the warps of the OTHER three
                // partitions, so the chain runs undisturbed (clock64 probes: 1.2 k instead of 1.7 k cycles per 8 pivots)
                constexpr int NUW = 3 * (NTP / 128);
                const int uw = warp - 1 - (warp >> 2);
                const int tend = colstart[npanels];
                for (int e0i = colstart[b + 1] + 1 + 4 * uw; e0i < tend; e0i += 4 * NUW) {
                    short2 te[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) te[u] = tiles[min(e0i + u, tend - 1)];
                    double a0[4], a1[4], b0[4], b1[4];
                    double2 c[4];
                    double2* cp[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double* pa = P + (te[u].x + g) * LDP + t4;
                        const double* pb = P + (te[u].y + g) * LDP + t4;
                        a0[u] = -pa[0];
                        a1[u] = -pa[4];
                        b0[u] = pb[0];
                        b1[u] = pb[4];
                        cp[u] = reinterpret_cast<double2*>(&S[(te[u].x + g) * LDB + te[u].y + 2 * t4]);
                        c[u] = *cp[u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        double e0 = 0.0, e1 = 0.0;  // second, independent chain
                        dmma_sub(c[u].x, c[u].y, a0[u], b0[u]);
                        dmma_sub(e0, e1, a1[u], b1[u]);
                        c[u].x += e0;
                        c[u].y += e1;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (e0i + u < tend) *cp[u] = c[u];
                }
            }
        }
        __syncthreads();
    }
    const bool failed = nfact < nb;
    // write L back: lower triangle, zeros above; nothing beyond a failed column is meaningful
    const int cmax = (nfact + 7) & ~7;
    if (fastio) {
        // everything is out already; after a failure the columns from the failing panel on are zeroed, as the
        // write-back of the general path does
        if (failed)
            for (int idx = tid; idx < NB * (NB / 2); idx += NTP) {
                const int r = idx >> 6, c = 2 * (idx & 63);
                if (c >= cmax && c <= r) *reinterpret_cast<double2*>(a + (size_t)r * ld + c) = make_double2(0.0, 0.0);
            }
    } else {
        for (int idx = tid; idx < NB * NB; idx += NTP) {
            const int r = idx >> 7, c = idx & (NB - 1);
            if (r < nb && c < nb) a[(size_t)r * ld + c] = (c <= r && c < cmax) ? S[r * LDB + c] : 0.0;
        }
    }
    if (rdiag_out != nullptr && !failed)
        for (int c = tid; c < nb; c += NTP) rdiag_out[c] = rdiag[c];
}

// ---- inverse of diagonal blocks, batched: one CTA per block, all blocks in parallel ----
// inv(L) by forward elimination on the identity held in registers (thread (i, cq) owns columns
// c == cq mod 4 of row i); row t of the inverse is final after step t and is published by 128 threads
// (coalesced) while the others update.  Only launched for factors that feed the blocked TRSMs.
__global__ void __launch_bounds__(NT, 1)
trtri_batched_kernel(const double* __restrict__ f, size_t ld, int n, double* __restrict__ dinv_all,
                     double* __restrict__ dinvt_all)
{
    extern __shared__ double sm[];
    double* S = sm;                  // [NB][LDB]
    double* xrow = sm + NB * LDB;    // [2][NB]
    double* rdiag = xrow + 2 * NB;   // [NB]
    const int tid = threadIdx.x;
    const int i = tid >> 2, cq = tid & 3;
    const int j0 = blockIdx.x * NB;
    const int nb = min(NB, n - j0);
    const double* a = f + (size_t)j0 * ld + j0;
    double* dinv = dinv_all + (size_t)blockIdx.x * NB * NB;
    double* dinvt = dinvt_all + (size_t)blockIdx.x * NB * NB;

    for (int idx = tid; idx < NB * NB; idx += NT) {
        int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    if (tid < nb) rdiag[tid] = 1.0 / S[tid * LDB + tid];
    double R[32];
#pragma unroll
    for (int jj = 0; jj < 32; jj++) R[jj] = (cq + 4 * jj == i) ? 1.0 : 0.0;
    __syncthreads();
    for (int t = 0; t < nb; t++) {
        double* xr = xrow + (t & 1) * NB;
        if (i == t) {  // raw row t of the right-hand side; scaling by 1/L[t][t] is folded into its readers
#pragma unroll
            for (int jj = 0; jj < 32; jj++) xr[cq + 4 * jj] = R[jj];
        }
        __syncthreads();
        const double rinv = rdiag[t];
        if (tid < nb) {
            const double x = xr[tid] * rinv;
            dinv[(size_t)t * NB + tid] = x;
            dinvt[(size_t)tid * NB + t] = x;
        }
        if (i > t && i < nb) {
            const double lit = S[i * LDB + t] * rinv;
            if (t < 64) {  // uniform: columns beyond t are still zero
#pragma unroll
                for (int jj = 0; jj < 16; jj++) R[jj] = fma(-lit, xr[cq + 4 * jj], R[jj]);
            } else {
#pragma unroll
                for (int jj = 0; jj < 32; jj++) R[jj] = fma(-lit, xr[cq + 4 * jj], R[jj]);
            }
        }
        // xrow is double-buffered: step t+1 writes the other half, and the write of step t+2 is
        // ordered after this step's reads by the barrier of step t+1
    }
}

// ---- inverse of diagonal blocks, second form: recursive doubling on the tensor cores ----
// The forward elimination above is a chain of 128 block-wide steps (75-100 us per launch, on the critical path of every
// structured-precision call: twice in cfg3).  Here inv(L) is built bottom-up and IN PLACE in one shared-memory copy of the
// block: the sixteen 8 x 8 diagonal inverses first (one warp each, lane = column), then for h = 8, 16, 32, 64 every
// 2h-block gets its off-diagonal part
//     X21 = -inv(L22) (L21 inv(L11))
// as two DMMA products (zero halves of the triangular factors skipped).  A warp owns an 8-row block of a product: T = L21
// inv(L11) overwrites L21 (row block i of T depends on row block i of L21 only); X21 is formed in registers and overwrites
// T after a barrier.
template <int H>
__device__ __forceinline__ void trtri_level(double* S, int warp, int lane)
{
    constexpr int NTL = H / 8;  // 8-column tiles per row block (and row blocks per pair)
    const int g = lane >> 2, t4 = lane & 3;
    const int pair = warp / NTL, rb = warp % NTL;  // eight tasks per level: warps 0 .. 7
    const int o = 2 * H * pair;
    double acc[NTL][2];
    if (warp < 8) {  // T = L21 inv(L11), in place
        const double* arow = S + (o + H + 8 * rb + g) * LDB + o + t4;
#pragma unroll
        for (int nt = 0; nt < NTL; nt++) {
            acc[nt][0] = acc[nt][1] = 0.0;
            const double* bcol = S + (size_t)(o + t4) * LDB + o + 8 * nt + g;
#pragma unroll
            for (int k0 = 8 * nt; k0 < H; k0 += 4)  // inv(L11)[k][n] = 0 for k < n
                dmma_sub(acc[nt][0], acc[nt][1], arow[k0], bcol[k0 * LDB]);
        }
        __syncwarp();
        double* trow = S + (o + H + 8 * rb + g) * LDB + o + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < NTL; nt++) *reinterpret_cast<double2*>(trow + 8 * nt) = make_double2(acc[nt][0], acc[nt][1]);
    }
    __syncthreads();
    if (warp < 8) {  // X21 = -inv(L22) T
        const double* arow = S + (o + H + 8 * rb + g) * LDB + o + H + t4;
#pragma unroll
        for (int nt = 0; nt < NTL; nt++) {
            acc[nt][0] = acc[nt][1] = 0.0;
            const double* bcol = S + (size_t)(o + H + t4) * LDB + o + 8 * nt + g;
#pragma unroll 2
            for (int k0 = 0; k0 < 8 * (rb + 1); k0 += 4)  // inv(L22)[r][k] = 0 for k > r
                dmma_sub(acc[nt][0], acc[nt][1], arow[k0], bcol[k0 * LDB]);
        }
    }
    __syncthreads();  // every row block of T has been read
    if (warp < 8) {
        double* xrow = S + (o + H + 8 * rb + g) * LDB + o + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < NTL; nt++) *reinterpret_cast<double2*>(xrow + 8 * nt) = make_double2(-acc[nt][0], -acc[nt][1]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT, 1)
trtri_dmma_kernel(const double* __restrict__ f, size_t ld, int n, double* __restrict__ dinv_all, double* __restrict__ dinvt_all)
{
    extern __shared__ double sm[];
    double* S = sm;  // [NB][LDB]: L (lower, zero above), turned into inv(L) block by block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = blockIdx.x * NB;
    const int nb = min(NB, n - j0);
    const double* a = f + (size_t)j0 * ld + j0;
    double* dinv = dinv_all + (size_t)blockIdx.x * NB * NB;
    double* dinvt = dinvt_all + (size_t)blockIdx.x * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += NT) {
        const int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDB + c] = (r < nb && c <= r) ? a[(size_t)r * ld + c] : (r == c ? 1.0 : 0.0);  // identity beyond a partial block
    }
    __syncthreads();
    {  // the 8 x 8 diagonal inverses, in place: warp b, lane c < 8 owns column c (right-looking: one multiply + one FMA per step)
        double* Sb = S + (8 * warp) * LDB + 8 * warp;
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            v[t] *= 1.0 / Sb[t * LDB + t];
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i > t) v[i] = fma(-Sb[i * LDB + t], v[t], v[i]);
        }
        __syncwarp();  // every lane has read the block
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; i++) Sb[i * LDB + lane] = v[i];
        }
    }
    __syncthreads();
    trtri_level<8>(S, warp, lane);
    trtri_level<16>(S, warp, lane);
    trtri_level<32>(S, warp, lane);
    trtri_level<64>(S, warp, lane);
    for (int idx = tid; idx < nb * NB; idx += NT) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c < nb) {
            dinv[(size_t)r * NB + c] = S[r * LDB + c];
            dinvt[(size_t)r * NB + c] = S[c * LDB + r];
        }
    }
}

// ---- panel solve of the blocked Cholesky: X := X * L^-T for the rows of a panel, on the tensor cores ----
// A warp owns 8 rows and keeps their 128 entries as sixteen 8 x 8 DMMA accumulator tiles.  For each 8-column
// block b: X_b = Acc_b * inv(L_bb)^T (two DMMAs with the 8 x 8 inverse computed once per CTA), X_b is final
// and stored; then Acc_c -= X_b * L[c-block, b-block]^T for every later block c (two DMMAs each, fragments of
// L straight from shared memory).  Accumulator -> A-fragment conversions are quad-local shuffles.  Rows never
// interact: no block-level synchronisation after the load.  128 rows per CTA.
constexpr int TR_ROWS = 128;
constexpr int LDI = 12;
__device__ __forceinline__ void c_to_a(double c0, double c1, int lane, double& a0, double& a1)
{
    const int gq = lane & ~3, t4 = lane & 3;
    const int s0 = gq + (t4 >> 1), s1 = gq + 2 + (t4 >> 1);
    const double v0 = __shfl_sync(0xffffffffu, c0, s0), v1 = __shfl_sync(0xffffffffu, c1, s0);
    const double w0 = __shfl_sync(0xffffffffu, c0, s1), w1 = __shfl_sync(0xffffffffu, c1, s1);
    a0 = (t4 & 1) ? v1 : v0;  // element (row g, column t4)
    a1 = (t4 & 1) ? w1 : w0;  // element (row g, column 4 + t4)
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR, 1)
trsm_rows_kernel(double* __restrict__ x, size_t ldx, int nrows, const double* __restrict__ l, size_t ldl,
                 int nb, const double* __restrict__ rdiag_g, size_t sx, size_t sl, size_t sr)
{
    // problem blockIdx.y of a batched launch (strides in doubles; all zero for a single problem)
    x += (size_t)blockIdx.y * sx;
    l += (size_t)blockIdx.y * sl;
    if (rdiag_g) rdiag_g += (size_t)blockIdx.y * sr;
    extern __shared__ double sm[];
    double* S = sm;                                    // [nb rounded up to 8][LDB] : L (lower, zero-padded)
    double* I8 = sm + ((nb + 7) & ~7) * LDB;           // [16][8][LDI] : inverses of the 8 x 8 diagonal blocks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    if (nb == NB) {
        load_block<NTHR>(S, l, ldl, nb, tid, true);  // potf2 left zeros above the diagonal
    } else {
        // partial / narrow blocks (the 64-wide panels of the batched mid-size problems): only the rows and columns that are
        // read - the lower triangle up to the next multiple of 8, zero-padded
        const int nbp = (nb + 7) & ~7;
        for (int idx = tid; idx < nbp * nbp; idx += NTHR) {
            const int r = idx / nbp, c = idx - r * nbp;
            S[r * LDB + c] = (r < nb && c <= r) ? l[(size_t)r * ldl + c] : 0.0;
        }
    }
    __syncthreads();
    // the warps invert the 8 x 8 diagonal blocks: lane c < 8 solves L_bb y = e_c; the reciprocal pivots come from
    // potf2 (rdiag), so there is no division on the chain
    for (int blk = warp; blk < 16 && 8 * blk < nb; blk += NTHR / 32) {  // only the blocks that exist (S holds no more rows)
        const int j0 = 8 * blk;
        if (lane < 8) {
            double y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double v = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int t = 0; t < 8; t++)
                    if (t < i) v = fma(-S[(j0 + i) * LDB + j0 + t], y[t], v);
                double rd = 1.0;
                if (j0 + i < nb) rd = rdiag_g ? rdiag_g[j0 + i] : 1.0 / S[(j0 + i) * LDB + j0 + i];
                y[i] = v * rd;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) I8[(blk * 8 + i) * LDI + lane] = y[i];
        }
    }
    __syncthreads();

    const size_t r0 = (size_t)blockIdx.x * (NTHR / 4) + warp * 8;
    const size_t row = r0 + g;
    const bool live = row < (size_t)nrows;
    double* xr = x + (live ? row : 0) * ldx;
    double acc[16][2];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const int c = 8 * t + 2 * t4;
        acc[t][0] = (live && c < nb) ? xr[c] : 0.0;
        acc[t][1] = (live && c + 1 < nb) ? xr[c + 1] : 0.0;
    }
#pragma unroll
    for (int b = 0; b < 16; b++) {
        if (8 * b < nb) {  // uniform
            double a0, a1;
            c_to_a(acc[b][0], acc[b][1], lane, a0, a1);
            double x0 = 0.0, x1 = 0.0;
            dmma_sub(x0, x1, a0, I8[(b * 8 + g) * LDI + t4]);
            dmma_sub(x0, x1, a1, I8[(b * 8 + g) * LDI + 4 + t4]);
            const int c = 8 * b + 2 * t4;
            if (live && c < nb) xr[c] = x0;
            if (live && c + 1 < nb) xr[c + 1] = x1;
            double xa0, xa1;
            c_to_a(x0, x1, lane, xa0, xa1);
            xa0 = -xa0;
            xa1 = -xa1;
#pragma unroll
            for (int tc = b + 1; tc < 16; tc++) {
                if (8 * tc < nb) {  // uniform
                    const double* lr = S + (tc * 8 + g) * LDB + 8 * b + t4;
                    dmma_sub(acc[tc][0], acc[tc][1], xa0, lr[0]);
                    dmma_sub(acc[tc][0], acc[tc][1], xa1, lr[4]);
                }
            }
        }
    }
}


// ---- look-ahead update of the NEXT block column: C(rows x nb2) -= X(rows x 128) X(0:nb2, :)^T ----
// The one GEMM that stands between two fused panel kernels in the look-ahead factorisation (linalg.c): rows x 128 x 128,
// far too small for the pipelined GEMM kernels (15 us, most of it prologue and epilogue latency).  Here a CTA takes 32 rows
// x 64 columns with the WHOLE K = 128 of both operands staged at once by cp.async (100 KB, two CTAs per SM), four warps of
// 8 rows x 64 columns, K contracted in the permuted order that makes every fragment a 16-byte load (see mid_batched.cu):
// about 7 us for n = 2000 ... 4096.
constexpr int PU_LD = 130;  // 260 words = 4 mod 32: conflict-free 16-byte fragment reads
__global__ void __launch_bounds__(128)
panel_update_kernel(double* __restrict__ c, size_t ldc, const double* __restrict__ x, size_t ldx, int rows, int nb2)
{
    extern __shared__ __align__(16) double pu[];
    double* As = pu;               // [32][PU_LD]: this CTA's rows of X
    double* Bs = pu + 32 * PU_LD;  // [64][PU_LD]: rows c0 .. c0+63 of X (the columns of C)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int r0 = 32 * (int)blockIdx.x, c0 = 64 * (int)blockIdx.y;
#pragma unroll
    for (int it = 0; it < 48; it++) {  // 96 rows x 64 pieces of 16 bytes
        const int idx = tid + 128 * it, r = idx >> 6, c2 = (idx & 63) * 2;
        const int grow = r < 32 ? r0 + r : c0 + (r - 32);
        const bool ok = r < 32 ? grow < rows : grow < nb2;
        double* dst = (r < 32 ? As + r * PU_LD : Bs + (r - 32) * PU_LD) + c2;
        if (ok) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(x + (size_t)grow * ldx + c2));
        } else {
            *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    const int row = r0 + 8 * warp + g;
    const bool live = row < rows;
    double* crow = c + (size_t)(live ? row : 0) * ldc + c0 + 2 * t4;
    double acc[8][2];
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const int cc = c0 + 8 * n + 2 * t4;
        acc[n][0] = (live && cc < nb2) ? crow[8 * n] : 0.0;
        acc[n][1] = (live && cc + 1 < nb2) ? crow[8 * n + 1] : 0.0;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    const double* ar = As + (8 * warp + g) * PU_LD + 4 * t4;
    const double* br = Bs + g * PU_LD + 4 * t4;
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const double2 a01 = *reinterpret_cast<const double2*>(ar + 16 * u);
        const double2 a23 = *reinterpret_cast<const double2*>(ar + 16 * u + 2);
        const double a0 = -a01.x, a1 = -a01.y, a2 = -a23.x, a3 = -a23.y;
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const double2 b01 = *reinterpret_cast<const double2*>(br + 8 * n * PU_LD + 16 * u);
            const double2 b23 = *reinterpret_cast<const double2*>(br + 8 * n * PU_LD + 16 * u + 2);
            dmma_sub(acc[n][0], acc[n][1], a0, b01.x);
            dmma_sub(acc[n][0], acc[n][1], a1, b01.y);
            dmma_sub(acc[n][0], acc[n][1], a2, b23.x);
            dmma_sub(acc[n][0], acc[n][1], a3, b23.y);
        }
    }
    if (live) {
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const int cc = c0 + 8 * n + 2 * t4;
            if (cc < nb2) crow[8 * n] = acc[n][0];
            if (cc + 1 < nb2) crow[8 * n + 1] = acc[n][1];
        }
    }
}

}  // namespace

static int potf2_launch(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info,
                        int nrows_below, double* tscr, unsigned* flags, int* rows_done)
{
    if (rows_done) *rows_done = 0;
    if (nb <= 0) return 0;
    if (nb > NB) return -202;
    // max over the two roles: [NB][LDB] + [NB][LDP] + NB + 96  |  [RB][LDB] + 2 [NB][LDP] + [RB][LDP] + 2 * 96
    const size_t smem = (size_t)(NB * LDB + 2 * NB * LDP + NB + 2 * 8 * 12) * sizeof(double);
    if (cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    // the fused panel needs the streaming path of the kernel (full, aligned block) and its scratch
    // (very tall panels keep the separate row-solve kernel: its 128-row CTAs amortise the load of L11 better than
    // hundreds of 32-row followers; measured break-even between 8 k and 16 k rows)
    const bool fused = nrows_below > 0 && nrows_below <= 8192 && tscr != nullptr && flags != nullptr && nb == NB &&
                       (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    unsigned grid = 1, epoch = 0;
    if (fused) {
        static unsigned seq = 0;
        do epoch = __atomic_add_fetch(&seq, 1u, __ATOMIC_RELAXED) << 1; while (epoch == 0);  // even, unique per launch, never 0
        grid = 1u + (unsigned)((nrows_below + RB - 1) / RB);
        if (rows_done) *rows_done = 1;
    }
    potf2_kernel<<<grid, NTP, smem, st>>>(a, ld, nb, col0, rdiag, info, fused ? nrows_below : 0, fused ? tscr : nullptr,
                                          fused ? flags : nullptr, epoch);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_potf2(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info)
{
    return potf2_launch(st, a, ld, nb, col0, rdiag, info, 0, nullptr, nullptr, nullptr);
}

// diagonal block + the nrows_below rows under it (L21 = A21 L11^-T) in one launch when the block is full and aligned
// (*rows_done = 1); otherwise only the diagonal block is factorised (*rows_done = 0: the caller runs htnk_trsm_rows).
// tscr: 16 * 64 doubles, flags: 17 unsigned (zero-initialised once, never reset), both private to the factorisation.
extern "C" int htnk_potf2_panel(cudaStream_t st, double* a, size_t ld, int nb, int col0, double* rdiag, int* info,
                                int nrows_below, double* tscr, unsigned* flags, int* rows_done)
{
    return potf2_launch(st, a, ld, nb, col0, rdiag, info, nrows_below, tscr, flags, rows_done);
}

extern "C" int htnk_trtri_batched(cudaStream_t st, const double* f, size_t ld, int n, double* dinv, double* dinvt)
{
    if (n <= 0) return 0;
    static int old_form = -1;  // HTN_TRTRI=0: the forward-elimination kernel
    if (old_form < 0) {
        const char* e = getenv("HTN_TRTRI");
        old_form = (e && *e == '0') ? 1 : 0;
    }
    if (!old_form) {
        const size_t smem2 = (size_t)NB * LDB * sizeof(double);
        if (cudaFuncSetAttribute(trtri_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) return -201;
        trtri_dmma_kernel<<<(n + NB - 1) / NB, NT, smem2, st>>>(f, ld, n, dinv, dinvt);
        htnk_count_launch();
        return cudaGetLastError() == cudaSuccess ? 0 : -201;
    }
    const size_t smem = (size_t)(NB * LDB + 3 * NB) * sizeof(double);
    if (cudaFuncSetAttribute(trtri_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return -201;
    trtri_batched_kernel<<<(n + NB - 1) / NB, NT, smem, st>>>(f, ld, n, dinv, dinvt);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

extern "C" int htnk_trsm_rows(cudaStream_t st, double* x, size_t ldx, int nrows, const double* l, size_t ldl, int nb,
                              const double* rdiag)
{
    if (nrows <= 0 || nb <= 0) return 0;
    if (nb > NB) return -202;
    const size_t smem = (size_t)(NB * LDB + 16 * 8 * LDI) * sizeof(double);
    // rows are independent: small CTAs (32 rows) until the 128-row CTAs would fill most of the device by themselves
    if ((nrows + TR_ROWS - 1) / TR_ROWS >= 96) {
        if (cudaFuncSetAttribute(trsm_rows_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -201;
        trsm_rows_kernel<512><<<(nrows + TR_ROWS - 1) / TR_ROWS, 512, smem, st>>>(x, ldx, nrows, l, ldl, nb, rdiag, 0, 0, 0);
    } else {
        if (cudaFuncSetAttribute(trsm_rows_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -201;
        trsm_rows_kernel<128><<<(nrows + 31) / 32, 128, smem, st>>>(x, ldx, nrows, l, ldl, nb, rdiag, 0, 0, 0);
    }
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

// the same panel solve for nbatch independent problems in one launch (gridDim.y = problem; strides in doubles)

/* C (rows x nb2, leading dimension ldc) -= X (rows x 128, leading dimension ldx) * X[0:nb2, :]^T: the next block column's
 * update between two fused panels of the look-ahead factorisation.  Returns 1 when the operands are not 16-byte aligned
 * (the caller then takes the general GEMM). */
extern "C" int htnk_panel_update(cudaStream_t st, double* c, size_t ldc, const double* x, size_t ldx, int rows, int nb2)
{
    if (rows <= 0 || nb2 <= 0) return 0;
    if (nb2 > NB || (ldx & 1) || (reinterpret_cast<uintptr_t>(x) & 15)) return 1;
    const size_t smem = (size_t)96 * PU_LD * sizeof(double);
    if (cudaFuncSetAttribute(panel_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -201;
    panel_update_kernel<<<dim3((unsigned)((rows + 31) / 32), (unsigned)((nb2 + 63) / 64)), 128, smem, st>>>(c, ldc, x, ldx, rows, nb2);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
