// TMA-fed variant of the FP64 tensor-core GEMM (same contract as gemm_f64.cu: C = alpha A B^T + beta C + bias,
// "TN", modes FULL / B_LOWER / C_LOWER), used for the wide-N sampling GEMM and trailing updates.
//
// What changes against the cp.async kernel:
//   * operand tiles arrive by TMA (cp.async.bulk.tensor.2d, one elected thread, 2 instructions per K chunk)
//     instead of ~100 integer/address instructions per warp per chunk; out-of-range rows / columns are
//     zero-filled by the TMA unit (no predicates);
//   * a 3-deep ring of stages guarded by mbarrier full/empty pairs replaces __syncthreads(): warps of a
//     CTA are never forced to the same point, a warp only ever waits for DATA;
//   * tiles are dense 128-byte rows with the hardware 128B swizzle.  The DMMA m8n8k4 fragment pattern would
//     bank-conflict on such rows, so the K order INSIDE a 16-wide chunk is permuted identically for A and B:
//     k4 step s uses k = {2s, 2s+1, 8+2s, 9+2s}; a half-warp then touches 16 distinct 8-byte bank pairs.
// CTA tile 128 x 64, 4 warps of 64 x 32, three CTAs per SM (3 x 73 KB of smem, 168 registers per thread).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <stdlib.h>

#include "launchers.h"

namespace {

// three stages of 24 KB: three CTAs (12 warps) per SM fit in shared memory, and in registers at 168 per thread
constexpr int BN = 64, BK = 16, STAGES = 3, NT = 128;
constexpr int WN = 4;  // warp tile (8 WM) x 32 in m8n8 units; WM = 8 (CTA tile 128 x 64) or 4 (64 x 64) is a template parameter
constexpr int B_BYTES = BN * BK * 8;
constexpr int AHEAD = 1;  // chunks issued ahead of the one being consumed

struct TmaGemmParams {
    int M, N, K;
    double alpha, beta;
    double* C;
    const double* bias;
    size_t ldc;
    int mode, ktri, kext0;
    int tiles_m, tiles_n;
    const int* skip;  // device flag: a non-zero value turns the launch into a no-op (failed factorisation)
    int paired;  // triangular B: one CTA runs the N tiles (tiles_n-1-q, q) of a row slab back to back
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// per-tile K schedule: chunks of BK from segment 1 [kbeg1, kend1) then segment 2 [kbeg2, K)
struct TileK {
    int n0, kbeg1, kend1, kbeg2, nc1, nchunks;
};

// MODE >= 0: compile-time mode; MODE < 0: the mode comes from the parameters at run time
template <int MODE>
__device__ __forceinline__ int mode_of(const TmaGemmParams& p) { return MODE < 0 ? p.mode : MODE; }

template <int MODE>
__device__ __forceinline__ TileK make_tile(const TmaGemmParams& p, int tn)
{
    TileK t;
    t.n0 = tn * BN;
    t.kbeg1 = 0;
    t.kend1 = p.K;
    t.kbeg2 = p.K;
    if (mode_of<MODE>(p) == HTNK_GEMM_B_LOWER) {
        int n1 = min(p.N, t.n0 + BN);
        t.kend1 = min(n1, p.ktri);
        t.kbeg2 = min(p.kext0, p.K);
    } else if (mode_of<MODE>(p) == HTNK_GEMM_B_UPPER) {
        t.kbeg1 = min((t.n0 / BK) * BK, p.K);  // B[n][k] == 0 for k < n: skip the chunks left of the tile
    }
    t.nc1 = (t.kend1 - t.kbeg1 + BK - 1) / BK;
    t.nchunks = t.nc1 + (p.K - t.kbeg2 + BK - 1) / BK;
    return t;
}

__device__ __forceinline__ void tile_chunk(const TmaGemmParams& p, const TileK& t, int c, int& k0, int& kend)
{
    if (c < t.nc1) { k0 = t.kbeg1 + c * BK; kend = t.kend1; }
    else { k0 = t.kbeg2 + (c - t.nc1) * BK; kend = p.K; }
}

// A CTA owns ONE output tile, or (p.paired: lower-triangular B) the PAIR of N tiles (tiles_n-1-q, q) of one
// 128-row slab: every CTA then runs the same number of K chunks (the longest and the shortest loop together,
// as in causal-attention schedulers), and the TMA ring keeps streaming across the tile boundary, so the
// second tile's first chunks are already in flight while the first tile's epilogue is being written.
template <int WM, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TmaGemmParams p)
{
    constexpr int BM = 16 * WM, A_BYTES = BM * BK * 8;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle pattern (the swizzle XORs address bits [4:6] with [7:9])
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;
    unsigned char* sB = base + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;

    if (p.skip != nullptr && *p.skip != 0) return;  // uniform: the factor this GEMM would read is invalid
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int bid = blockIdx.x;
    // the CTA's tiles as two tile indices and one chunk count: the descriptors are recomputed where they are needed
    // (a dozen integer instructions) instead of being carried in registers next to the 128 accumulators
    int ntl = 1, tn0, tn1, m0;
    if ((mode_of<MODE>(p) == HTNK_GEMM_B_LOWER || mode_of<MODE>(p) == HTNK_GEMM_B_UPPER) && p.paired) {
        const int npairs = (p.tiles_n + 1) >> 1;
        const int q = bid % npairs;
        m0 = (bid / npairs) * BM;
        tn0 = p.tiles_n - 1 - q;
        tn1 = q;
        ntl = (tn0 != tn1) ? 2 : 1;
    } else {
        tn0 = tn1 = p.tiles_n - 1 - (bid % p.tiles_n);
        m0 = (bid / p.tiles_n) * BM;
        if (mode_of<MODE>(p) == HTNK_GEMM_C_LOWER && m0 + BM - 1 < tn0 * BN) return;
    }
    const int nch0 = make_tile<MODE>(p, tn0).nchunks;
    const int total_chunks = nch0 + (ntl == 2 ? make_tile<MODE>(p, tn1).nchunks : 0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    // one elected thread feeds the ring; global chunk cg (over the CTA's tiles) lands in stage cg % STAGES
    auto issue = [&](int cg) {
        if (cg >= total_chunks) return;
        const int s = cg % STAGES, round = cg / STAGES;
        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
        const bool first = cg < nch0;
        const TileK tt = make_tile<MODE>(p, first ? tn0 : tn1);
        int k0, kend;
        tile_chunk(p, tt, first ? cg : cg - nch0, k0, kend);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        tma_load_2d(sA + s * A_BYTES, &tmA, k0, m0, &full[s]);
        tma_load_2d(sB + s * B_BYTES, &tmB, k0, tt.n0, &full[s]);
    };
    if (tid == 0) {
        for (int c = 0; c < AHEAD; c++) issue(c);
    }

    const int g = lane >> 2, t4 = lane & 3;
    // swizzled fragment addressing: row r (r & 7 == g), 16-byte chunk ((t4 >> 1) * 4 + s) ^ g, half t4 & 1
    const int a_off = (wm * WM * 8 + g) * 128 + (t4 & 1) * 8;
    const int b_off = (wn * WN * 8 + g) * 128 + (t4 & 1) * 8;
    int xs[4];
#pragma unroll
    for (int s = 0; s < 4; s++) xs[s] = ((((t4 >> 1) * 4 + s) ^ g) << 4);
    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);

    int cg = 0;
    for (int tl = 0; tl < ntl; tl++) {
        const TileK tk = make_tile<MODE>(p, tl == 0 ? tn0 : tn1);
        const int n0 = tk.n0;
        const int nw = n0 + wn * WN * 8;
        const int jhi = min(WN, (p.N - nw + 7) >> 3);
        double acc[WM][WN][2];
#pragma unroll
        for (int i = 0; i < WM; i++)
#pragma unroll
            for (int j = 0; j < WN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int c = 0; c < tk.nchunks; c++, cg++) {
            if (tid == 0) issue(cg + AHEAD);
            const int st = cg % STAGES;
            mbar_wait(&full[st], (cg / STAGES) & 1);
            int k0, kend;
            tile_chunk(p, tk, c, k0, kend);
            const unsigned char* a_s = sA + st * A_BYTES + a_off;
            const unsigned char* b_s = sB + st * B_BYTES + b_off;
            const bool diag = (mode_of<MODE>(p) == HTNK_GEMM_B_LOWER) && (c < tk.nc1) && (k0 + BK > nw);
            if (!diag && jhi == WN && kend - k0 >= BK) {
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    double af[WM], bf[WN];
#pragma unroll
                    for (int i = 0; i < WM; i++) af[i] = *reinterpret_cast<const double*>(a_s + i * 1024 + xs[s]);
#pragma unroll
                    for (int j = 0; j < WN; j++) bf[j] = *reinterpret_cast<const double*>(b_s + j * 1024 + xs[s]);
#pragma unroll
                    for (int i = 0; i < WM; i++)
#pragma unroll
                        for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else {
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    // step s multiplies k = k0 + {2s, 2s+1, 8+2s, 9+2s}; its smallest k decides whether any
                    // non-zero element can be involved
                    const int kmin = k0 + 2 * s;
                    if (kmin < kend) {
                        int jlo = 0;
                        if (diag) jlo = max(0, (kmin - nw) >> 3);
                        if (jlo < jhi) {
                            double af[WM];
#pragma unroll
                            for (int i = 0; i < WM; i++)
                                af[i] = *reinterpret_cast<const double*>(a_s + i * 1024 + xs[s]);
                            // The live fragment range [jlo, jhi) is warp-uniform and takes few values (jlo is 0 or
                            // 2 with 16-wide chunks, jhi < WN only on the last N tile): dispatch with REAL branches to
                            // bodies with compile-time bounds.  (As predicated code the masked DMMAs were still issued:
                            // 1.393e8 DMMA instructions instead of the 1.311e8 the tile geometry needs - ncu.)
                            auto frag = [&](auto jl, auto jh) {
#pragma unroll
                                for (int j = decltype(jl)::value; j < decltype(jh)::value; j++) {
                                    const double bfj = *reinterpret_cast<const double*>(b_s + j * 1024 + xs[s]);
#pragma unroll
                                    for (int i = 0; i < WM; i++) dmma(acc[i][j][0], acc[i][j][1], af[i], bfj);
                                }
                            };
                            using I0 = std::integral_constant<int, 0>;
                            using I1 = std::integral_constant<int, 1>;
                            using I2 = std::integral_constant<int, 2>;
                            using I3 = std::integral_constant<int, 3>;
                            using I4 = std::integral_constant<int, 4>;
                            if (jlo == 0) {
                                if (jhi == 4) frag(I0{}, I4{});
                                else if (jhi == 1) frag(I0{}, I1{});
                                else if (jhi == 2) frag(I0{}, I2{});
                                else frag(I0{}, I3{});
                            } else if (jlo == 2) {
                                if (jhi == 4) frag(I2{}, I4{});
                                else frag(I2{}, I3{});
                            } else if (jlo == 1) {
                                if (jhi == 4) frag(I1{}, I4{});
                                else if (jhi == 3) frag(I1{}, I3{});
                                else frag(I1{}, I2{});
                            } else {
                                frag(I3{}, I4{});
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);  // this warp has read everything it needs from the stage
        }

        // epilogue of this tile.  Common case first (interior tile, beta = 0, 16-byte aligned C): column pair by column
        // pair - the two bias values are fetched once for the eight rows of the pair - and nothing but FMA + STG.128
        if (vec_ok && p.beta == 0.0 && n0 + BN <= p.N && m0 + BM <= p.M) {
            double* crow = p.C + (size_t)(m0 + wm * WM * 8 + g) * p.ldc + nw + 2 * t4;
#pragma unroll
            for (int j = 0; j < WN; j++) {
                if (p.bias) {
                    const double bz0 = p.bias[nw + j * 8 + 2 * t4], bz1 = p.bias[nw + j * 8 + 2 * t4 + 1];
#pragma unroll
                    for (int i = 0; i < WM; i++)
                        *reinterpret_cast<double2*>(crow + (size_t)i * 8 * p.ldc + j * 8) =
                            make_double2(fma(p.alpha, acc[i][j][0], bz0), fma(p.alpha, acc[i][j][1], bz1));
                } else {
#pragma unroll
                    for (int i = 0; i < WM; i++)
                        *reinterpret_cast<double2*>(crow + (size_t)i * 8 * p.ldc + j * 8) =
                            make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
                }
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < WM; i++) {
            const int m = m0 + wm * WM * 8 + i * 8 + g;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < WN; j++) {
                const int n = n0 + wn * WN * 8 + j * 8 + 2 * t4;
                if (n >= p.N) continue;
                double* cp = p.C + (size_t)m * p.ldc + n;
                const bool two = (n + 1 < p.N);
                double v0, v1;
                if (p.bias) {  // the same fused form as the interior-tile path above: one result whatever the tile
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
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn_t lookup_encode()
{
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<encode_fn_t>(ptr);
    return nullptr;
}

encode_fn_t get_encode()
{
    static const encode_fn_t fn = lookup_encode();  // initialised once, thread-safe (C++11)
    return fn;
}

// rows x cols row-major matrix (leading dimension ld doubles), box = 16 columns x box_rows rows, 128B swizzle
bool make_map(CUtensorMap* map, const double* ptr, size_t ld, int rows, int cols, int box_rows)
{
    encode_fn_t enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// returns 1 if the TMA path cannot be used for these operands (caller falls back to the cp.async kernel)
extern "C" int htnk_gemm_tn_tma(cudaStream_t st, int M, int N, int K, double alpha, const double* A, size_t lda,
                                const double* B, size_t ldb, double beta, double* C, size_t ldc,
                                const double* bias, int mode, int ktri, int kext0, const int* skip)
{
    if (M <= 0 || N <= 0) return 0;
    if (K <= 0 || (lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) ||
        (reinterpret_cast<uintptr_t>(B) & 15))
        return 1;
    CUtensorMap tmA, tmB;
    // CTA tile 128 x 64 (3 CTAs/SM) or 64 x 64 (4 CTAs/SM): the same speed on a full device (measured), so the small
    // tile is taken when the large one would leave SMs without work.  HTN_TMA_WM = 8 / 4 forces one (experiments).
    static int wm_env = -1;
    if (wm_env < 0) {
        const char* e = getenv("HTN_TMA_WM");
        wm_env = e ? atoi(e) : 0;
    }
    int wm_cfg = wm_env == 2 || wm_env == 4 || wm_env == 8 ? wm_env : 8;
    if (wm_env != 2 && wm_env != 4 && wm_env != 8) {
        const long long t128 = (long long)((M + 127) / 128) * ((N + BN - 1) / BN);
        const long long work = (mode == HTNK_GEMM_B_LOWER || mode == HTNK_GEMM_C_LOWER || mode == HTNK_GEMM_B_UPPER) ? t128 / 2 : t128;
        if (2 * work < 148 && K >= 2048) {
            // not even the 64 x 64 tiles give every SM a CTA (few rows: small batches against a large factor): 32 x 64 tiles,
            // six CTAs per SM.  Only with a long K: the small products of the recursive inversion are faster on 64 x 64 tiles
            // (cfg3 2.87 -> 2.92 ms without this condition)
            wm_cfg = 2;
        } else if (work < 3 * 148) {
            wm_cfg = 4;
        } else {
            // enough large tiles to fill the device once - but a last, mostly empty wave is expensive (a lone CTA of four
            // warps runs at half the SM's rate): compare how full the waves are with 3 x 148 slots of 128 x 64 tiles and with
            // 4 x 148 slots of 64 x 64 tiles (cfg3's Y1 = Z1 L^-1: 512 CTA pairs on 444 slots against 1024 on 592)
            const long long c8 = work, c4 = 2 * work;
            const double e8 = (double)c8 / (double)(((c8 + 443) / 444) * 444);
            const double e4 = (double)c4 / (double)(((c4 + 591) / 592) * 592);
            if (e8 < 0.80 && e4 > e8 + 0.05) wm_cfg = 4;
        }
    }
    const int BM = 16 * wm_cfg, A_BYTES = BM * BK * 8;
    if (!make_map(&tmA, A, lda, M, K, BM) || !make_map(&tmB, B, ldb, N, K, BN)) return 1;
    TmaGemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.alpha = alpha; p.beta = beta;
    p.C = C; p.bias = bias; p.ldc = ldc;
    p.mode = mode; p.ktri = ktri; p.kext0 = kext0;
    p.skip = skip;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    if ((long long)p.tiles_m * p.tiles_n > 0x7fffffffLL) return -202;
    // triangular B: K length grows (B_LOWER) or shrinks (B_UPPER) linearly with the N tile - pair the ends
    p.paired = ((mode == HTNK_GEMM_B_LOWER || mode == HTNK_GEMM_B_UPPER) && p.tiles_n >= 2) ? 1 : 0;
    const unsigned grid = p.paired ? (unsigned)p.tiles_m * (unsigned)((p.tiles_n + 1) / 2)
                                   : (unsigned)p.tiles_m * (unsigned)p.tiles_n;
    const size_t smem = (size_t)STAGES * (A_BYTES + B_BYTES) + 2 * STAGES * sizeof(uint64_t) + 1024;
    void (*kern)(const CUtensorMap, const CUtensorMap, const TmaGemmParams) = nullptr;
    switch (mode) {  // the mode is a template parameter: each instance carries only its own tile logic
        case HTNK_GEMM_B_LOWER: kern = wm_cfg == 2 ? gemm_tma_kernel<2, 6, HTNK_GEMM_B_LOWER> : wm_cfg == 4 ? gemm_tma_kernel<4, 4, HTNK_GEMM_B_LOWER> : gemm_tma_kernel<8, 3, HTNK_GEMM_B_LOWER>; break;
        case HTNK_GEMM_C_LOWER: kern = wm_cfg == 2 ? gemm_tma_kernel<2, 6, HTNK_GEMM_C_LOWER> : wm_cfg == 4 ? gemm_tma_kernel<4, 4, HTNK_GEMM_C_LOWER> : gemm_tma_kernel<8, 3, HTNK_GEMM_C_LOWER>; break;
        case HTNK_GEMM_B_UPPER: kern = wm_cfg == 2 ? gemm_tma_kernel<2, 6, HTNK_GEMM_B_UPPER> : wm_cfg == 4 ? gemm_tma_kernel<4, 4, HTNK_GEMM_B_UPPER> : gemm_tma_kernel<8, 3, HTNK_GEMM_B_UPPER>; break;
        default: kern = wm_cfg == 2 ? gemm_tma_kernel<2, 6, HTNK_GEMM_FULL> : wm_cfg == 4 ? gemm_tma_kernel<4, 4, HTNK_GEMM_FULL> : gemm_tma_kernel<8, 3, HTNK_GEMM_FULL>; break;
    }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -201;
    kern<<<grid, NT, smem, st>>>(tmA, tmB, p);
    htnk_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -201;
}
