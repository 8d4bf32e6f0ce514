// FP64 tensor-core GEMM for sm_100a:  C = alpha * A * B^T + beta * C + bias   ("TN": both operands
// K-contiguous, row-major), built on DMMA  mma.sync.aligned.m8n8k4.f64  (tcgen05 has no f64 kind,
// so the FP64 tensor path on Blackwell is the warp-level DMMA instruction).
//
// This one kernel replaces every dense Level-3 call site of the reference (src/htnorm_blas.h:84-107:
// dgemm NN/NT/TN, dsymm, dsyrk) and, through blocked drivers in the host layer, dtrmm / dtrsm /
// the trailing updates of dpotrf:
//   * sampling   X = Z~ * L~^T + mean     (mode B_LOWER: K range clipped to the non-zero part of L)
//   * Cholesky   A22 -= L21 * L21^T       (mode C_LOWER: only tiles touching the lower triangle)
//   * TRSM       X_j = (B_j - X_<j L_j<^T) * inv(L_jj)^T
//
// Tiling: CTA tile 128 x BN, BK = 16, 3-stage cp.async pipeline, warp tile 64 x 32 (32 independent DMMAs
// and 12 LDS.64 per k4 step) for the wide configurations:
//   BN = 64  : 4 warps (2 x 2), 128 threads, 92 KB smem -> TWO CTAs per SM, so one CTA's barrier /
//              prologue / epilogue bubbles are filled by the other's DMMAs (the DMMA pipe is the bound)
//   BN = 128 : 8 warps (2 x 4), 1 CTA per SM (kept for in-place panel updates that need one N tile)
//   BN = 32/8: 8 warps (8 x 1) of 16 x BN, for skinny N and for small problems that need more CTAs
// Shared-memory rows are padded to 20 doubles: the DMMA fragment pattern (8 rows x 4 consecutive
// doubles per half-warp pair) then hits 16 distinct 8-byte bank pairs -> conflict-free LDS.64.
// Tile order: n fastest and descending, so the CTAs that share one 128-row slab of A run together
// (A streams from HBM once, B stays in L2) and, for B_LOWER, the longest K loops start first.
// Wasted work is trimmed at 8-column granularity: in the diagonal K blocks of a lower-triangular B, in
// the N tail and in partial K chunks the DMMAs that would only multiply zeros are not issued.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "launchers.h"

namespace {

constexpr int BM = 128;

struct GemmParams {
    int M, N, K;
    double alpha, beta;
    const double* A;
    const double* B;
    double* C;
    const double* bias;
    size_t lda, ldb, ldc;
    int mode, ktri, kext0;
    int tiles_m, tiles_n;
    const int* skip;  // device flag: non-zero turns the launch into a no-op
    size_t sA, sB, sC;  // batched launches (gridDim.y problems): strides between consecutive problems, in doubles
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// rows x BK tile (row-major source, leading dimension ld) -> smem [rows][LDS]; elements outside
// [0, nrows_total) x [k0, kend) are zero-filled by cp.async's src-size operand.
template <int ROWS, int NTHREADS, int BK>
__device__ __forceinline__ void load_tile(double* s, const double* g, size_t ld, int row0, int nrows_total,
                                          int k0, int kend, int tid)
{
    constexpr int LDS = BK + 4;
    constexpr int CPR = BK / 2;              // 16-byte chunks per row
    constexpr int CHUNKS = ROWS * CPR;
#pragma unroll
    for (int it = 0; it < (CHUNKS + NTHREADS - 1) / NTHREADS; it++) {
        int id = tid + it * NTHREADS;
        if (CHUNKS % NTHREADS != 0 && id >= CHUNKS) break;
        int r = id / CPR, ch = id % CPR;
        int grow = row0 + r;
        int kk = k0 + ch * 2;
        int rem = kend - kk;
        int bytes = (grow < nrows_total && rem > 0) ? (rem >= 2 ? 16 : 8) : 0;
        const double* src = bytes ? g + (size_t)grow * ld + kk : g;
        cp_async16(s + r * LDS + ch * 2, src, bytes);
    }
}

template <int BN, int WARPS_M, int WARPS_N, int WM, int WN, int MINB, int BK, int STAGES>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N, MINB) gemm_tn_kernel(const GemmParams p)
{
    constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
    constexpr int LDS = BK + 4;  // padded row (doubles): 4 mod 16 -> the DMMA fragment pattern is conflict-free
    static_assert(WARPS_M * WM * 8 == BM, "BM");
    static_assert(WARPS_N * WN * 8 == BN, "BN");
    extern __shared__ __align__(16) double smem[];
    constexpr int A_STAGE = BM * LDS;
    constexpr int B_STAGE = BN * LDS;
    double* sA = smem;
    double* sB = smem + STAGES * A_STAGE;

    if (p.skip != nullptr && *p.skip != 0) return;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp / WARPS_N;
    const int wn = warp % WARPS_N;
    const double* const Ag = p.A + (size_t)blockIdx.y * p.sA;  // problem blockIdx.y of a batched launch
    const double* const Bg = p.B + (size_t)blockIdx.y * p.sB;
    double* const Cg = p.C + (size_t)blockIdx.y * p.sC;

    const int bid = blockIdx.x;
    const int tn = p.tiles_n - 1 - (bid % p.tiles_n);
    const int tm = bid / p.tiles_n;
    const int m0 = tm * BM;
    const int n0 = tn * BN;
    if (p.mode == HTNK_GEMM_C_LOWER && m0 + BM - 1 < n0) return;  // tile entirely above the diagonal

    // K schedule: chunks of BK from segment 1 [0, kend1) then segment 2 [kext0, K)
    int kbeg1 = 0, kend1 = p.K, kbeg2 = p.K;
    const int kend2 = p.K;
    if (p.mode == HTNK_GEMM_B_LOWER) {
        int n1 = min(p.N, n0 + BN);
        kend1 = min(n1, p.ktri);
        kbeg2 = min(p.kext0, p.K);
    } else if (p.mode == HTNK_GEMM_B_UPPER) {
        kbeg1 = min((n0 / BK) * BK, p.K);  // B[n][k] == 0 for k < n: skip the chunks left of the tile
    }
    const int nc1 = (kend1 - kbeg1 + BK - 1) / BK;
    const int nc2 = (kend2 - kbeg2 + BK - 1) / BK;
    const int nchunks = nc1 + nc2;

    auto chunk_range = [&](int c, int& k0, int& kend) {
        if (c < nc1) { k0 = kbeg1 + c * BK; kend = kend1; }
        else { k0 = kbeg2 + (c - nc1) * BK; kend = kend2; }
    };
    auto issue = [&](int c) {
        if (c < nchunks) {
            int k0, kend;
            chunk_range(c, k0, kend);
            int s = c % STAGES;
            load_tile<BM, NTHREADS, BK>(sA + s * A_STAGE, Ag, p.lda, m0, p.M, k0, kend, tid);
            load_tile<BN, NTHREADS, BK>(sB + s * B_STAGE, Bg, p.ldb, n0, p.N, k0, kend, tid);
        }
        cp_async_commit();
    };

    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; i++)
#pragma unroll
        for (int j = 0; j < WN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue(s);

    const int g = lane >> 2, t4 = lane & 3;
    const int nw = n0 + wn * WN * 8;                       // first column of this warp
    const int jhi = min(WN, (p.N - nw + 7) >> 3);          // 8-column sub-tiles that exist (N tail)
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(c + STAGES - 1);
        int k0, kend;
        chunk_range(c, k0, kend);
        const int klim = min(BK, kend - k0);
        // K order inside a 16-wide chunk: k4 step s multiplies k = {2s, 2s+1, 8+2s, 9+2s} - the SAME order as the
        // TMA kernel (gemm_tma.cu), so that both kernels sum every output element in the identical sequence and a
        // result never depends on which of them a batch / shard / chunk size selects (bit-identical shards)
        const int tp = (t4 & 1) + 8 * (t4 >> 1);
        const double* a_s = sA + (c % STAGES) * A_STAGE + (wm * WM * 8 + g) * LDS + tp;
        const double* b_s = sB + (c % STAGES) * B_STAGE + (wn * WN * 8 + g) * LDS + tp;
        // does this chunk need trimming for this warp?  (all conditions are warp-uniform)
        const bool diag = (p.mode == HTNK_GEMM_B_LOWER) && (c < nc1) && (k0 + BK > nw);
        if (!diag && jhi == WN && klim == BK) {
#pragma unroll
            for (int s4 = 0; s4 < BK / 4; s4++) {
                const int kk = 16 * (s4 >> 2) + 2 * (s4 & 3);
                double af[WM], bf[WN];
#pragma unroll
                for (int i = 0; i < WM; i++) af[i] = a_s[i * 8 * LDS + kk];
#pragma unroll
                for (int j = 0; j < WN; j++) bf[j] = b_s[j * 8 * LDS + kk];
#pragma unroll
                for (int i = 0; i < WM; i++)
#pragma unroll
                    for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        } else {
#pragma unroll
            for (int s4 = 0; s4 < BK / 4; s4++) {
                const int kk = 16 * (s4 >> 2) + 2 * (s4 & 3);  // the step's smallest k is k0 + kk
                if (kk < klim) {
                    // sub-tile j (columns nw + 8j .. +7) of a lower-triangular B is non-zero only for
                    // k <= nw + 8j + 7: the first sub-tile that still needs k = k0 + kk
                    int jlo = 0;
                    if (diag) jlo = max(0, (k0 + kk - nw) >> 3);
                    if (jlo < jhi) {
                        double af[WM];
#pragma unroll
                        for (int i = 0; i < WM; i++) af[i] = a_s[i * 8 * LDS + kk];
#pragma unroll
                        for (int j = 0; j < WN; j++) {
                            if (j >= jlo && j < jhi) {
                                const double bfj = b_s[j * 8 * LDS + kk];
#pragma unroll
                                for (int i = 0; i < WM; i++) dmma(acc[i][j][0], acc[i][j][1], af[i], bfj);
                            }
                        }
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[m][n], C[m][n+1] with m = .. + g, n = .. + 2*t4
    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cg) & 15) == 0);
#pragma unroll
    for (int i = 0; i < WM; i++) {
        const int m = m0 + wm * WM * 8 + i * 8 + g;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < WN; j++) {
            const int n = n0 + wn * WN * 8 + j * 8 + 2 * t4;
            if (n >= p.N) continue;
            double* cp = Cg + (size_t)m * p.ldc + n;
            const bool two = (n + 1 < p.N);
            double v0, v1;
            if (p.bias) {  // explicitly fused, as in gemm_tma.cu: the two kernels must round alike
                v0 = fma(p.alpha, acc[i][j][0], p.bias[n]);
                v1 = two ? fma(p.alpha, acc[i][j][1], p.bias[n + 1]) : 0.0;
            } else {
                v0 = p.alpha * acc[i][j][0];
                v1 = p.alpha * acc[i][j][1];
            }
            if (vec_ok && two) {
                if (p.beta != 0.0) {
                    double2 o = *reinterpret_cast<const double2*>(cp);
                    v0 += p.beta * o.x;
                    v1 += p.beta * o.y;
                }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
            } else {
                if (p.beta != 0.0) {
                    v0 += p.beta * cp[0];
                    if (two) v1 += p.beta * cp[1];
                }
                cp[0] = v0;
                if (two) cp[1] = v1;
            }
        }
    }
}

template <int BN, int WARPS_M, int WARPS_N, int WM, int WN, int MINB, int BK = 16, int STAGES = 3>
int launch(cudaStream_t st, const GemmParams& p, unsigned nbatch = 1)
{
    constexpr size_t smem = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
    auto kern = gemm_tn_kernel<BN, WARPS_M, WARPS_N, WM, WN, MINB, BK, STAGES>;
    // per-device attribute: set on every launch (cheap) so that any current device is covered
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    const dim3 grid((unsigned)p.tiles_m * (unsigned)p.tiles_n, nbatch);
    kern<<<grid, 32 * WARPS_M * WARPS_N, smem, st>>>(p);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}

}  // namespace

// one-shot guard for the NEXT htnk_gemm_tn launch of this thread: the kernel returns at once when *flag != 0
static thread_local const int* g_next_skip = nullptr;
extern "C" void htnk_gemm_guard_next(const int* flag) { g_next_skip = flag; }

// wide-N configuration: 64 (default) or an experimental one selected with HTN_GEMM_WIDE
static int wide_config()
{
    static int cfg = 0;
    if (cfg == 0) {
        const char* e = getenv("HTN_GEMM_WIDE");
        cfg = e ? atoi(e) : 70;
        if (cfg < 64 || cfg > 129) cfg = 64;
    }
    return cfg;
}

extern "C" int htnk_gemm_tn(cudaStream_t st, int M, int N, int K, double alpha, const double* A, size_t lda,
                            const double* B, size_t ldb, double beta, double* C, size_t ldc,
                            const double* bias, int mode, int ktri, int kext0, int bn)
{
    // the one-shot guard belongs to THIS call whatever happens next: consume it before any early return
    const int* const skip = g_next_skip;
    g_next_skip = nullptr;
    if (M <= 0 || N <= 0) return 0;
    if (K < 0) return -202;
    if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return -202;
    const int tiles_m = (M + BM - 1) / BM;
    if (bn == 0) {
        if (N <= 8) bn = 8;
        else if (N <= 32) bn = 32;
        else {
            // wide N: 128 x 64 tiles, unless that leaves most of the 148 SMs (x2 CTAs) without a tile
            long long ctas64 = (long long)tiles_m * ((N + 63) / 64);
            static int min64 = -1;  // HTN_GEMM_MIN64 (experiments): fewest 128 x 64 tiles that still take the wide kernel
            if (min64 < 0) {
                const char* e = getenv("HTN_GEMM_MIN64");
                min64 = e ? atoi(e) : 148;
            }
            // ... and unless K is long: a 64 x 64 tile of the TMA kernel covers the same area as a 128 x 32 tile
            // here, so the narrow kernel only wins where its short prologue matters (the K = 128 panel updates)
            static int kwide = -1;  // HTN_GEMM_KWIDE (experiments): shortest K that takes the wide kernel regardless
            if (kwide < 0) {
                const char* e = getenv("HTN_GEMM_KWIDE");
                kwide = e ? atoi(e) : 512;
            }
            // (up to 64 tiles the narrow kernel stays - measured on the products of the recursive inversion, K <= 2048 - unless K is
            // very long: cfg4's b = Phi y1 + y2, 256 x 1024 x 8192, takes the 32 x 64 tiles of the TMA kernel)
            bn = (ctas64 < min64 && (K < kwide || (ctas64 < 64 && K < 4096))) ? 32 : wide_config();
        }
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.alpha = alpha; p.beta = beta;
    p.A = A; p.B = B; p.C = C; p.bias = bias;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.mode = mode; p.ktri = ktri; p.kext0 = kext0;
    p.skip = skip;
    p.sA = p.sB = p.sC = 0;
    p.tiles_m = tiles_m;
    const int bn_cols = bn >= 128 ? 128 : (bn >= 64 ? 64 : bn);
    p.tiles_n = (N + bn_cols - 1) / bn_cols;
    if ((long long)p.tiles_m * p.tiles_n > 0x7fffffffLL) return -202;
    switch (bn) {
        case 128: return launch<128, 2, 4, 8, 4, 1>(st, p);
        case 70: {  // TMA + mbarrier ring (gemm_tma.cu); falls back if a tensor map cannot be built
            int r = htnk_gemm_tn_tma(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, mode, ktri, kext0, p.skip);
            if (r != 1) return r;
            return launch<64, 2, 2, 8, 4, 2>(st, p);
        }
        case 64: return launch<64, 2, 2, 8, 4, 2>(st, p);
        case 65: return launch<64, 4, 2, 4, 4, 2>(st, p);    // 16 warps / SM, warp tile 32 x 32
        case 66: return launch<64, 2, 2, 8, 4, 2, 32, 2>(st, p);  // BK = 32, two stages: half the barriers
        case 67: return launch<64, 2, 2, 8, 4, 2, 16, 4>(st, p);  // four stages (1 CTA/SM fits only: smem)
        case 129: return launch<128, 4, 4, 4, 4, 1>(st, p);  // 16 warps / SM in one CTA
        case 32: return launch<32, 8, 1, 2, 4, 2>(st, p);  // small accumulators: two CTAs per SM, one wave for the panel updates
        case 8: return launch<8, 8, 1, 2, 1, 1>(st, p);
        default: return -202;
    }
}
