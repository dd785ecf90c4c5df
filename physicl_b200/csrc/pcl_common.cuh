// Shared device/host helpers for the physicl_b200 kernels (sm_100a only).
// Arithmetic contract: this library is compiled with -fmad=false, so `a*b+c` is never contracted;
// every fused multiply-add is an explicit fmaf().  The CPU oracle twin (oracle/c/oracle_f32.c) uses
// the same sequence of IEEE-754 binary32 operations (mul, add, fma, sqrt, rint), which is what
// makes the integer tallies bit-exact between the two.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/physicl_b200.h"

#ifndef PCL_BLOCK
#define PCL_BLOCK 256
#endif
// minimum resident CTAs per SM requested from ptxas for the fused photon kernels (register cap)
#ifndef PCL_PHOTON_MINB
#define PCL_PHOTON_MINB 3
#endif
#define PCL_WARPS (PCL_BLOCK / 32)

struct pcl_graph_key {
    pcl_soa p;
    float dt;
    int accel;
    float a[3];
    uint32_t nsteps;
    uintptr_t stream;
};

struct pcl_hostpipe;  // hostpipe.cu

struct pcl_ctx {
    int device;
    int sm_count;
    size_t l2_bytes;
    size_t hbm_bytes;
    char name[128];
    char err[512];
    uint64_t launches;
    // compaction scratch: per-block live counts / offsets
    uint32_t *scan_buf;
    size_t scan_cap;
    // gravity: partial accelerations of the j splits
    float *grav_part;
    size_t grav_cap;
    // cached CUDA graph of a multi-step kinematics loop
    cudaGraphExec_t kin_graph;
    pcl_graph_key kin_key;
    // cached CUDA graph of a multi-step photon loop is rebuilt per call (step index changes)
    pcl_hostpipe *pipe;
};

void pcl_set_error(pcl_ctx *ctx, const char *fmt, ...);

#define PCL_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            pcl_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(e__));                                          \
            return -(int)e__ - 1000;                                                         \
        }                                                                                    \
    } while (0)

#define PCL_REQUIRE(ctx, cond, msg)                                                          \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            pcl_set_error(ctx, "%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);            \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

#define PCL_ENTER(ctx)                                                                       \
    do {                                                                                     \
        if (!(ctx)) {                                                                        \
            pcl_set_error(nullptr, "%s: null context", __func__);                            \
            return -2;                                                                       \
        }                                                                                    \
        PCL_CUDA(ctx, cudaSetDevice((ctx)->device));                                         \
    } while (0)

#define PCL_LAUNCHED(ctx)                                                                    \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        PCL_CUDA(ctx, cudaGetLastError());                                                   \
    } while (0)

static inline bool pcl_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// persistent-style grid for streaming kernels: a multiple of the SM count, capped by the work
static inline unsigned pcl_stream_grid(const pcl_ctx *ctx, uint64_t work_items, int per_block,
                                       int blocks_per_sm) {
    uint64_t need = (work_items + (uint64_t)per_block - 1) / (uint64_t)per_block;
    uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)blocks_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (id_lo, id_hi, step, stream), key = seed.
// ---------------------------------------------------------------------------------------------
#define PCL_PHILOX_M0 0xD2511F53u
#define PCL_PHILOX_M1 0xCD9E8D57u
#define PCL_PHILOX_W0 0x9E3779B9u
#define PCL_PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ uint4 pcl_philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(PCL_PHILOX_M0, c.x), lo0 = PCL_PHILOX_M0 * c.x;
        uint32_t hi1 = __umulhi(PCL_PHILOX_M1, c.z), lo1 = PCL_PHILOX_M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += PCL_PHILOX_W0;
        k.y += PCL_PHILOX_W1;
    }
    return c;
}

// 24-bit uniform in [0,1): exact in binary32
__device__ __forceinline__ float pcl_u01(uint32_t r) { return (float)(r >> 8) * 0x1p-24f; }

// ---------------------------------------------------------------------------------------------
// sin(pi t), cos(pi t) for t in [0, 2], built only from rintf / fmaf / mul so that the CPU twin
// reproduces it bit for bit.  Quadrant reduction is exact; the kernels are odd/even Taylor
// polynomials on |r| <= 1/4 (truncation error < 2.5e-9).
// ---------------------------------------------------------------------------------------------
#define PCL_S0 0x1.921fb6p+1f    /*  pi           */
#define PCL_S1 -0x1.4abbcep+2f  /* -pi^3/3!      */
#define PCL_S2 0x1.466bc6p+1f    /*  pi^5/5!      */
#define PCL_S3 -0x1.32d2ccp-1f  /* -pi^7/7!      */
#define PCL_S4 0x1.507834p-4f    /*  pi^9/9!      */
#define PCL_C1 -0x1.3bd3ccp+2f  /* -pi^2/2!      */
#define PCL_C2 0x1.03c1fp+2f     /*  pi^4/4!      */
#define PCL_C3 -0x1.55d3c8p+0f  /* -pi^6/6!      */
#define PCL_C4 0x1.e1f506p-3f    /*  pi^8/8!      */
#define PCL_C5 -0x1.a6d1f2p-6f  /* -pi^10/10!    */

__device__ __forceinline__ void pcl_sincospi(float t, float &s, float &c) {
    float q = rintf(t + t);
    float r = fmaf(q, -0.5f, t);
    int qi = (int)q;
    float r2 = r * r;
    float ps = fmaf(r2, PCL_S4, PCL_S3);
    ps = fmaf(r2, ps, PCL_S2);
    ps = fmaf(r2, ps, PCL_S1);
    ps = fmaf(r2, ps, PCL_S0);
    float sr = r * ps;
    float pc = fmaf(r2, PCL_C5, PCL_C4);
    pc = fmaf(r2, pc, PCL_C3);
    pc = fmaf(r2, pc, PCL_C2);
    pc = fmaf(r2, pc, PCL_C1);
    float cr = fmaf(r2, pc, 1.0f);
    float a = (qi & 1) ? cr : sr;
    float b = (qi & 1) ? sr : cr;
    s = (qi & 2) ? -a : a;
    c = ((qi + 1) & 2) ? -b : b;
}

// ---------------------------------------------------------------------------------------------
// global memory access: 128-bit, L1 no-allocate (every plane is touched once per step)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 pcl_ld4(const float *p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void pcl_st4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 pcl_ld4u(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void pcl_st4u(uint32_t *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float &pcl_f4(float4 &v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ uint32_t &pcl_u4(uint4 &v, int i) { return (&v.x)[i]; }

#endif  // __CUDACC__
