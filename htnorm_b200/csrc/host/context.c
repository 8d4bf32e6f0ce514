/* Device context, stream and workspace plumbing of the host layer. */
#include "htn_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static _Thread_local char g_err[512] = "";

void htn_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char* htn_last_error(void) { return g_err; }

size_t htn_chunk_budget(void)
{
    const char* e = getenv("HTN_CHUNK_BYTES");
    if (e && atoll(e) >= 4096) return (size_t)atoll(e);
    return (size_t)3 << 30;
}
const char* htn_version(void)
{
#ifdef HTNORM_COLMAJOR
    return "htnorm-b200 0.2 (sm_100a, column-major build)";
#else
    return "htnorm-b200 0.2 (sm_100a)";
#endif
}
unsigned long long htn_launch_count(void) { return htnk_launch_count(); }

int htn_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        htn_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return HTN_E_NO_DEVICE;
    }
    return n;
}

void htn_batch_init(htn_batch_t* b, size_t nsamples)
{
    memset(b, 0, sizeof(*b));
    b->nsamples = nsamples;
    b->gen = HTN_GEN_PCG64;
    b->mem = HTN_MEM_HOST;
    b->device = -1;
}

static int pool_configured[64];

int htn_ctx_begin(htn_ctx_t* ctx, int device, void* stream, int host_mode)
{
    int rc = 0;
    memset(ctx, 0, sizeof(*ctx));
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        /* there is no CPU fallback: fail loudly */
        htn_set_error("no CUDA device available (%s); htnorm-b200 has no CPU path",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return HTN_E_NO_DEVICE;
    }
    HTN_CUDA(cudaGetDevice(&ctx->prev_device));
    ctx->device = device < 0 ? ctx->prev_device : device;
    if (ctx->device >= n) {
        htn_set_error("device %d out of range (%d devices)", ctx->device, n);
        return HTN_E_ARG;
    }
    if (ctx->device != ctx->prev_device) HTN_CUDA(cudaSetDevice(ctx->device));
    if (ctx->device < 64 && !__atomic_load_n(&pool_configured[ctx->device], __ATOMIC_ACQUIRE)) {
        cudaMemPool_t pool;
        unsigned long long keep = ~0ULL; /* never trim: workspaces are reused call after call */
        HTN_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        HTN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        __atomic_store_n(&pool_configured[ctx->device], 1, __ATOMIC_RELEASE); /* idempotent: a racing thread repeats it */
    }
    if (host_mode) {
        /* the library's own stream for host-memory calls: one per host thread and device, created once and kept
         * (the call still synchronises on it before returning) */
        static _Thread_local struct { int device; cudaStream_t stream; } own = {-1, NULL};
        if (own.device != ctx->device) {
            if (own.device >= 0) cudaStreamDestroy(own.stream);
            own.device = -1;
            HTN_CUDA(cudaStreamCreateWithFlags(&own.stream, cudaStreamNonBlocking));
            own.device = ctx->device;
        }
        ctx->stream = own.stream;
        ctx->own_stream = 1;
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
done:
    return rc;
}

int htn_ctx_end(htn_ctx_t* ctx, int sync)
{
    int rc = 0;
    if (sync || ctx->own_stream) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            htn_set_error("stream synchronize: %s", cudaGetErrorString(e));
            rc = HTN_E_CUDA;
        }
    }
    if (ctx->device != ctx->prev_device) cudaSetDevice(ctx->prev_device);
    return rc;
}

int htn_dalloc(htn_ctx_t* ctx, void** p, size_t bytes)
{
    *p = NULL;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e != cudaSuccess) {
        htn_set_error("cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
        *p = NULL;
        return e == cudaErrorMemoryAllocation ? HTNORM_ALLOC_ERROR : HTN_E_CUDA;
    }
    return 0;
}

void htn_dfree(htn_ctx_t* ctx, void* p)
{
    if (p) cudaFreeAsync(p, ctx->stream);
}

/* ---- optional phase timing (CUDA events on the caller's stream; off by default) ----
 * Every call of the shared-covariance hyperplane path records one SET of marks; up to HTN_PROF_SETS sets are
 * kept since the facility was enabled, so a benchmark loop can run without synchronising per step and read
 * the per-phase averages afterwards.
 * The state is PER HOST THREAD (thread-local): a thread sees only the marks of its own calls, enabling the
 * facility in one thread neither slows nor races with sampler calls of another - the library keeps no mutable
 * state that is shared between threads (SURVEY 8b threading contract). */
#define HTN_PROF_SLOTS 5
#define HTN_PROF_SETS 256
static _Thread_local int g_prof_on = 0;
static _Thread_local cudaEvent_t g_prof_ev[HTN_PROF_SETS][HTN_PROF_SLOTS];
static _Thread_local unsigned char g_prof_set[HTN_PROF_SETS][HTN_PROF_SLOTS];
static _Thread_local unsigned char g_prof_have[HTN_PROF_SETS];
static _Thread_local int g_prof_dev = -1, g_prof_cur = -1;

void htn_phase_timing_enable(int on)
{
    g_prof_on = on;
    g_prof_cur = -1;
    memset(g_prof_set, 0, sizeof(g_prof_set));
}

void htn_prof_mark(htn_ctx_t* ctx, int slot)
{
    if (!g_prof_on || slot < 0 || slot >= HTN_PROF_SLOTS) return;
    if (g_prof_dev != ctx->device) { /* events belong to a device: forget the ones of the device used before */
        for (int s = 0; s < HTN_PROF_SETS; s++)
            if (g_prof_have[s])
                for (int i = 0; i < HTN_PROF_SLOTS; i++) cudaEventDestroy(g_prof_ev[s][i]);
        memset(g_prof_have, 0, sizeof(g_prof_have));
        memset(g_prof_set, 0, sizeof(g_prof_set));
        g_prof_dev = ctx->device;
        g_prof_cur = -1;
    }
    if (slot == 0) { /* a new call: next set (the ring overwrites the oldest) */
        g_prof_cur = (g_prof_cur + 1) % HTN_PROF_SETS;
        memset(g_prof_set[g_prof_cur], 0, HTN_PROF_SLOTS);
    }
    if (g_prof_cur < 0) return;
    if (!g_prof_have[g_prof_cur]) { /* a set's five events are created the first time the ring reaches it */
        for (int i = 0; i < HTN_PROF_SLOTS; i++) cudaEventCreate(&g_prof_ev[g_prof_cur][i]);
        g_prof_have[g_prof_cur] = 1;
    }
    if (cudaEventRecord(g_prof_ev[g_prof_cur][slot], ctx->stream) == cudaSuccess) g_prof_set[g_prof_cur][slot] = 1;
}

/* ms[i] = AVERAGE over the recorded calls of the time between mark i and mark i+1 (after synchronising on
 * the events); -1 where no call has both marks.  Returns the number of calls averaged. */
int htn_phase_timing_get(double* ms, int n)
{
    int ncalls = 0;
    double sum[HTN_PROF_SLOTS - 1];
    int cnt[HTN_PROF_SLOTS - 1];
    for (int i = 0; i < HTN_PROF_SLOTS - 1; i++) { sum[i] = 0.0; cnt[i] = 0; }
    for (int s = 0; s < HTN_PROF_SETS; s++) {
        int any = 0;
        for (int i = 0; i + 1 < HTN_PROF_SLOTS; i++) {
            if (g_prof_set[s][i] && g_prof_set[s][i + 1]) {
                float t = 0;
                if (cudaEventSynchronize(g_prof_ev[s][i + 1]) == cudaSuccess &&
                    cudaEventElapsedTime(&t, g_prof_ev[s][i], g_prof_ev[s][i + 1]) == cudaSuccess) {
                    sum[i] += t;
                    cnt[i]++;
                    any = 1;
                }
            }
        }
        ncalls += any;
    }
    for (int i = 0; i < n; i++) ms[i] = (i < HTN_PROF_SLOTS - 1 && cnt[i]) ? sum[i] / cnt[i] : -1.0;
    return ncalls;
}
