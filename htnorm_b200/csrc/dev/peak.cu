// Roofline denominators measured in place: FP64 tensor-core (DMMA) and FP64 FMA-pipe throughput.
// MEASURED_PEAKS.json carries HBM and bf16 only, so the FP64 peak the GEMM / Cholesky fractions are
// quoted against is this microbenchmark: register operands only, 8 independent accumulator chains per
// warp (DMMA) / per thread (DFMA), every SM saturated with 8 CTAs x 256 threads, CUDA-event timed.
#include <cuda_runtime.h>

#include "launchers.h"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* sink, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) sink[0] = s;  // never true: keeps the chain alive
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9, b = 1e-12 * threadIdx.x;
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    if (s == 123.456) sink[0] = s;
}

// Integer-issue peak: the generator's own instruction mix - eight independent PCG32 streams per thread (64-bit LCG step
// = IMAD.WIDE + 2 IMAD, the XSH-RR output = shifts, LOP3, SHF), no memory traffic.  Reported as warp-instructions per
// second from the EXECUTED instruction count of the unrolled loop body (counted from SASS once: kIntInstrPerIter), i.e. the
// issue-slot ceiling the ziggurat kernel is bound by (profiles/: normal_fill is issue-bound, not HBM-bound).
__global__ void __launch_bounds__(256) int_issue_peak_kernel(double* sink, int iters, double seed)
{
    unsigned long long st[8], acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = (unsigned long long)(seed * 1e6) + threadIdx.x * 977u + i * 7919u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const unsigned long long old = st[i];
            st[i] = old * 6364136223846793005ULL + (2 * i + 1);
            const unsigned xs = (unsigned)(((old >> 18) ^ old) >> 27);
            acc += __funnelshift_r(xs, xs, (unsigned)(old >> 59));
        }
    }
    if (acc == 0x123456789abcdefULL) sink[0] = (double)acc;
}

template <typename K>
double time_kernel(cudaStream_t st, K kern, double flops_per_thread_iter_x32, double ms_target)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 8;
    int iters = 2000;
    double best = -1.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, st);
        kern<<<grid, 256, 0, st>>>(sink, iters, 1.0);
        htnk_count_launch();
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        // flops: warps * iters * per-warp-iteration flops
        double flops = (double)grid * 8.0 * iters * flops_per_thread_iter_x32;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
        if (rep < 2 && ms > 0) {  // calibrate the loop length to about ms_target
            double scale = ms_target / ms;
            if (scale > 64) scale = 64;
            if (scale < 1.0 / 64) scale = 1.0 / 64;
            iters = (int)(iters * scale);
            if (iters < 100) iters = 100;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace

extern "C" double htnk_dmma_peak_tflops(cudaStream_t st, double ms)
{
    // per warp per iteration: 8 DMMA m8n8k4 = 8 * 8*8*4*2 flops
    return time_kernel(st, dmma_peak_kernel, 8.0 * 512.0, ms);
}

extern "C" double htnk_dfma_peak_tflops(cudaStream_t st, double ms)
{
    // per warp per iteration: 32 lanes * 8 FMA * 2 flops
    return time_kernel(st, dfma_peak_kernel, 32.0 * 8.0 * 2.0, ms);
}

// giga warp-instructions per second over all SMs.  kIntInstrPerIter: SASS instructions of one loop iteration of
// int_issue_peak_kernel (8 streams x [IMAD.WIDE.U32, 2 IMAD, IMAD.IADD, 2-3 SHF, LOP3, SHF.R.W, IADD3 + carry] + register moves + loop overhead),
// counted with `cuobjdump -sass` (tools/count_int_peak_sass.py re-derives it; the count is asserted there).
#ifndef HTN_INT_INSTR_PER_ITER
#define HTN_INT_INSTR_PER_ITER 103.0
#endif
extern "C" double htnk_int_issue_peak_gops(cudaStream_t st, double ms)
{
    // time_kernel returns (grid * 8 warps * iters * x) / time / 1e12: with x = instructions per iteration * 1e3 -> G/s
    return time_kernel(st, int_issue_peak_kernel, HTN_INT_INSTR_PER_ITER * 1e3, ms);
}

// ---- latency probes (one warp / one CTA): calibrates the cycle models in DESIGN.md ----
namespace {
__global__ void latency_kernel(double* out, double seed)
{
    __shared__ double sh[64];
    const int n = 512;
    double x = seed + threadIdx.x * 1e-9, y = 1.0 - 1e-9;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = fma(x, y, 1e-9);
    t1 = clock64();
    double r_dfma = double(t1 - t0) / n;
    // dependent DMMA chain (accumulator dependency)
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    t1 = clock64();
    double r_dmma = double(t1 - t0) / n;
    // dependent library rsqrt chain
    double z = 2.0 + x * 1e-30;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) z = rsqrt(z) + 1.5;
    t1 = clock64();
    double r_rsqrt = double(t1 - t0) / 64;
    // dependent sqrt + divide chain
    double w = 2.0 + x * 1e-30;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 64; i++) w = 1.0 / sqrt(w) + 1.5;
    t1 = clock64();
    double r_sqrtdiv = double(t1 - t0) / 64;
    // dependent LDS chain (pointer chasing through shared memory)
    if (threadIdx.x < 64) sh[threadIdx.x] = (double)((threadIdx.x * 7 + 1) & 63);
    __syncthreads();
    int idx = threadIdx.x & 63;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; i++) idx = (int)sh[idx];
    t1 = clock64();
    double r_lds = double(t1 - t0) / 256;
    // back-to-back barriers with the whole CTA
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; i++) __syncthreads();
    t1 = clock64();
    double r_bar = double(t1 - t0) / 256;
    // shuffle chain
    double sv = x;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; i++) sv = __shfl_sync(0xffffffffu, sv, (threadIdx.x + 1) & 31) + 1.0;
    t1 = clock64();
    double r_shfl = double(t1 - t0) / 256;
    if (threadIdx.x == 0) {
        out[0] = r_dfma; out[1] = r_dmma; out[2] = r_rsqrt; out[3] = r_sqrtdiv; out[4] = r_lds;
        out[5] = r_bar; out[6] = r_shfl; out[7] = x + c0 + c1 + z + w + idx + sv;
    }
}
}  // namespace

// out[0..6] = cycles per dependent DFMA, DMMA, rsqrt(), 1/sqrt(), LDS (incl. I2F), __syncthreads (nthreads),
// shuffle+DADD
extern "C" int htnk_latency_probe(cudaStream_t st, int nthreads, double* host_out)
{
    double* d = nullptr;
    if (cudaMalloc(&d, 8 * sizeof(double)) != cudaSuccess) return -201;
    latency_kernel<<<1, nthreads, 0, st>>>(d, 1.0);
    htnk_count_launch();
    cudaError_t e = cudaMemcpyAsync(host_out, d, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    return e == cudaSuccess ? 0 : -201;
}

// ---- DMMA issue sweep: throughput vs resident warps and independent accumulators per warp ----
namespace {
template <int ILP>
__global__ void dmma_sweep_kernel(double* sink, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12 * threadIdx.x;
    double c[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) c[i] = i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) s += c[i];
    if (s == 123.456) sink[0] = s;
}
}  // namespace

// TFLOP/s with `warps_per_sm` resident warps (one CTA per SM) and ilp in {8, 32} accumulators per warp
extern "C" double htnk_dmma_sweep(cudaStream_t st, int warps_per_sm, int ilp)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    double best = -1;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, st);
        if (ilp == 32) dmma_sweep_kernel<32><<<sms, warps_per_sm * 32, 0, st>>>(sink, iters / 4, 1.0);
        else dmma_sweep_kernel<8><<<sms, warps_per_sm * 32, 0, st>>>(sink, iters, 1.0);
        htnk_count_launch();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = (double)sms * warps_per_sm * (double)iters * 8.0 * 512.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

// ---- synthetic inner loop of the GEMM: fragments from shared memory + 32 DMMAs per k4 step, no global
// loads, no barriers.  Isolates "operand delivery" from "pipeline / synchronisation" losses. ----
namespace {
__global__ void __launch_bounds__(128, 2) dmma_tile_kernel(double* sink, int iters)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (128 + 64) * 20; i += 128) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int g = lane >> 2, t4 = lane & 3;
    const double* a_s = sm + ((warp >> 1) * 64 + g) * 20 + t4;
    const double* b_s = sm + 128 * 20 + ((warp & 1) * 32 + g) * 20 + t4;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = a_s[i * 8 * 20 + kk];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = b_s[j * 8 * 20 + kk];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(af[i]), "d"(bf[j]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) sink[0] = s;
}
}  // namespace

extern "C" double htnk_dmma_tile_probe(cudaStream_t st, int ctas_per_sm)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
    const size_t smem = ctas_per_sm >= 2 ? 92160 : 180000;  // forces the requested residency
    cudaFuncSetAttribute(dmma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000;
    double best = -1;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, st);
        dmma_tile_kernel<<<sms * ctas_per_sm, 128, smem, st>>>(sink, iters);
        htnk_count_launch();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = (double)sms * ctas_per_sm * 4.0 * iters * 128.0 * 512.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}
