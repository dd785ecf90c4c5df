// Context management, error reporting and the on-box peak probes of the physicl_b200 C ABI.
// Stands in for cl.create_some_context()/cl.CommandQueue()/get_device_info()
// (reference physicl/__init__.py:428-429, :470-499).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "pcl_common.cuh"

static thread_local char g_tls_err[512] = "";

void pcl_set_error(pcl_ctx *ctx, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_tls_err, sizeof(g_tls_err), fmt, ap);
    va_end(ap);
    if (ctx) {
        strncpy(ctx->err, g_tls_err, sizeof(ctx->err) - 1);
        ctx->err[sizeof(ctx->err) - 1] = 0;
    }
}

void pcl_hostpipe_destroy(pcl_ctx *ctx);  // hostpipe.cu

extern "C" int pcl_abi_version(void) { return PCL_ABI_VERSION; }

extern "C" const char *pcl_last_error(pcl_ctx *ctx) { return ctx ? ctx->err : g_tls_err; }

extern "C" int pcl_init(int device, pcl_ctx **out) {
    if (!out) {
        pcl_set_error(nullptr, "pcl_init: out is null");
        return -2;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        pcl_set_error(nullptr, "pcl_init: no CUDA device (%s); there is no CPU fallback",
                      cudaGetErrorString(e));
        return -3;
    }
    if (device < 0 || device >= count) {
        pcl_set_error(nullptr, "pcl_init: device %d out of range [0,%d)", device, count);
        return -4;
    }
    pcl_ctx *ctx = (pcl_ctx *)calloc(1, sizeof(pcl_ctx));
    if (!ctx) return -5;
    ctx->device = device;
    PCL_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    PCL_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pcl_set_error(nullptr, "pcl_init: device %d is sm_%d%d; this library ships sm_100a code only",
                      device, prop.major, prop.minor);
        free(ctx);
        return -6;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->hbm_bytes = prop.totalGlobalMem;
    snprintf(ctx->name, sizeof(ctx->name), "%.120s", prop.name);
    {  // direction table of the photon kernels: (sin, cos)(2 pi k / 512) in double, rounded once to binary32
        float host[2 * PCL_TRIG_N];  // sin table, then cos table
        for (int k = 0; k < PCL_TRIG_N; ++k) {
            const double a = 2.0 * 3.14159265358979323846 * (double)k / (double)PCL_TRIG_N;
            host[k] = (float)sin(a);
            host[PCL_TRIG_N + k] = (float)cos(a);
        }
        cudaError_t e2 = cudaMalloc(&ctx->trig, sizeof(host));
        if (e2 == cudaSuccess) e2 = cudaMemcpy(ctx->trig, host, sizeof(host), cudaMemcpyHostToDevice);
        if (e2 != cudaSuccess) {
            pcl_set_error(nullptr, "pcl_init: direction table: %s", cudaGetErrorString(e2));
            if (ctx->trig) cudaFree(ctx->trig);
            free(ctx);
            return -7;
        }
    }
    *out = ctx;
    return 0;
}

extern "C" int pcl_destroy(pcl_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    pcl_hostpipe_destroy(ctx);
    if (ctx->scan_buf) cudaFree(ctx->scan_buf);
    if (ctx->grav_part) cudaFree(ctx->grav_part);
    if (ctx->grav_pairs) cudaFree(ctx->grav_pairs);
    if (ctx->trig) cudaFree(ctx->trig);
    free(ctx);
    return 0;
}

extern "C" int pcl_device_info(pcl_ctx *ctx, char *name, int name_len, int *sm_count,
                               uint64_t *hbm_bytes, uint64_t *l2_bytes) {
    PCL_ENTER(ctx);
    if (name && name_len > 0) {
        strncpy(name, ctx->name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = ctx->sm_count;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    return 0;
}

extern "C" uint64_t pcl_launch_count(pcl_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int pcl_stream_sync(pcl_ctx *ctx, uintptr_t stream) {
    PCL_ENTER(ctx);
    PCL_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Stream gate: a one-thread kernel that holds `stream` until a word in mapped host memory becomes
// non-zero.  A caller queues a whole timed region behind it (event, launches, event) and then opens
// the gate, so no host-side launch latency lies between the two events.  The wait gives up after
// timeout_ms, so a host that dies (or blocks on the stream) cannot hang the GPU.
// ---------------------------------------------------------------------------------------------
__global__ void pcl_k_gate(const volatile uint32_t *flag, unsigned long long timeout_ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag == 0u) {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) break;
    }
}

extern "C" int pcl_stream_gate(pcl_ctx *ctx, uintptr_t stream, const uint32_t *flag_host, uint32_t timeout_ms) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, flag_host != nullptr, "null flag");
    cudaPointerAttributes a;
    PCL_CUDA(ctx, cudaPointerGetAttributes(&a, flag_host));
    PCL_REQUIRE(ctx, a.type == cudaMemoryTypeHost && a.devicePointer != nullptr,
                "the gate flag must live in page-locked, mapped host memory");
    if (timeout_ms == 0 || timeout_ms > 10000) timeout_ms = 10000;
    pcl_k_gate<<<1, 1, 0, (cudaStream_t)stream>>>((const volatile uint32_t *)a.devicePointer,
                                                  (unsigned long long)timeout_ms * 1000000ull);
    PCL_CUDA(ctx, cudaGetLastError());  // not counted in pcl_launch_count: it does no particle work
    return 0;
}

extern "C" int pcl_host_register(pcl_ctx *ctx, void *ptr, uint64_t bytes) {
    PCL_ENTER(ctx);
    PCL_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}
extern "C" int pcl_host_unregister(pcl_ctx *ctx, void *ptr) {
    PCL_ENTER(ctx);
    PCL_CUDA(ctx, cudaHostUnregister(ptr));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Roofline denominators that MEASURED_PEAKS.json does not carry (FP32 FMA rate) or that we want to
// re-check on the very box the bench runs on (copy bandwidth).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pcl_k_ffma_peak(float *out, int iters, float a, float b) {
    // 16 independent accumulators per thread: enough ILP to saturate both FMA pipes
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456f) out[0] = s;  // keep the loop alive
}

__global__ void __launch_bounds__(256) pcl_k_ffma2_peak(float *out, int iters, float a, float b) {
    unsigned long long r[8], pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = (float)(threadIdx.x + i) * 1e-3f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(v));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(pa), "l"(pb));
    }
    unsigned long long x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= r[i];
    if (x == 0x123456789abcdefull) out[0] = 1.f;
}

// packed (FFMA2) rate: same units as pcl_measure_fp32_peak (2 flop per FMA lane-element)
extern "C" int pcl_measure_fp32x2_peak(pcl_ctx *ctx, double *tflops) {
    PCL_ENTER(ctx);
    float *d = nullptr;
    PCL_CUDA(ctx, cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    PCL_CUDA(ctx, cudaEventCreate(&e0));
    PCL_CUDA(ctx, cudaEventCreate(&e1));
    const int iters = 4096;
    const int blocks = ctx->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        PCL_CUDA(ctx, cudaEventRecord(e0, 0));
        pcl_k_ffma2_peak<<<blocks, 256>>>(d, iters, 0.999f, 1e-4f);
        PCL_LAUNCHED(ctx);
        PCL_CUDA(ctx, cudaEventRecord(e1, 0));
        PCL_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        PCL_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (tflops) *tflops = best;
    return 0;
}

extern "C" int pcl_measure_fp32_peak(pcl_ctx *ctx, double *tflops) {
    PCL_ENTER(ctx);
    float *d = nullptr;
    PCL_CUDA(ctx, cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    PCL_CUDA(ctx, cudaEventCreate(&e0));
    PCL_CUDA(ctx, cudaEventCreate(&e1));
    const int iters = 4096;
    const int blocks = ctx->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        PCL_CUDA(ctx, cudaEventRecord(e0, 0));
        pcl_k_ffma_peak<<<blocks, 256>>>(d, iters, 0.999f, 1e-4f);
        PCL_LAUNCHED(ctx);
        PCL_CUDA(ctx, cudaEventRecord(e1, 0));
        PCL_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        PCL_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (tflops) *tflops = best;
    return 0;
}

__global__ void __launch_bounds__(256) pcl_k_copy(const float *__restrict__ src, float *dst,
                                                  uint64_t nvec) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (uint64_t)gridDim.x * blockDim.x)
        pcl_st4(dst + 4 * i, pcl_ld4(src + 4 * i));
}

extern "C" int pcl_measure_copy_peak(pcl_ctx *ctx, uint64_t bytes, double *gbs) {
    PCL_ENTER(ctx);
    if (bytes < (1u << 20)) bytes = 1u << 20;
    bytes &= ~(uint64_t)15;
    float *a = nullptr, *b = nullptr;
    PCL_CUDA(ctx, cudaMalloc(&a, bytes));
    PCL_CUDA(ctx, cudaMalloc(&b, bytes));
    PCL_CUDA(ctx, cudaMemset(a, 0, bytes));
    cudaEvent_t e0, e1;
    PCL_CUDA(ctx, cudaEventCreate(&e0));
    PCL_CUDA(ctx, cudaEventCreate(&e1));
    uint64_t nvec = bytes / 16;
    unsigned grid = pcl_stream_grid(ctx, nvec, 256, 16);
    double best = 0.0;
    for (int rep = 0; rep < 8; ++rep) {
        PCL_CUDA(ctx, cudaEventRecord(e0, 0));
        pcl_k_copy<<<grid, 256>>>(a, b, nvec);
        PCL_LAUNCHED(ctx);
        PCL_CUDA(ctx, cudaEventRecord(e1, 0));
        PCL_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        PCL_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 1 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    if (gbs) *gbs = best;
    return 0;
}
