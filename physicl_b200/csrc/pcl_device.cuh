// Device-only helpers shared by the hand-written kernels and by run-time compiled (NVRTC) kernels:
// Philox4x32-10, mantissa uniforms and the reproducible sin(pi t)/cos(pi t).  No host code, no
// standard headers needed under NVRTC.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

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
