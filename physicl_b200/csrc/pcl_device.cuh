// Device-only helpers shared by the hand-written kernels and by run-time compiled (NVRTC) kernels:
// Philox (4x32 for the emission sampler, 2x32 for the photon steps), mantissa uniforms and the
// reproducible table-driven sin/cos of the scattering angles.  No host code, no standard headers
// needed under NVRTC.
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
// Philox2x32-10 (same paper): ONE 32x32 multiply per round, 64 random bits per block -- exactly what a
// photon needs per timestep (24 bits for the collision test, 24 + 16 for the two angles), at half the
// instructions of Philox4x32.  counter = (global id, step); the ten round keys key + r*W are formed on
// the host and arrive as kernel parameters, i.e. as constant-bank operands of the XOR.
// ---------------------------------------------------------------------------------------------
#define PCL_PHILOX2_M 0xD256D193u
#define PCL_PHILOX2_W 0x9E3779B9u

__device__ __forceinline__ uint2 pcl_philox2x32_10(uint32_t c0, uint32_t c1, const uint32_t (&rk)[10]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi = __umulhi(PCL_PHILOX2_M, c0), lo = PCL_PHILOX2_M * c0;
        c0 = hi ^ rk[r] ^ c1;
        c1 = lo;
    }
    return make_uint2(c0, c1);
}

// ---------------------------------------------------------------------------------------------
// Packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2 = PTX fma/add/sub/mul.rn.f32x2): one instruction, the same IEEE
// operation on the two halves of a 64-bit register pair.  Same flops as two scalar instructions, half the issue slots:
// the photon kernels are bound by instruction issue, so their per-photon arithmetic runs on PAIRS of photons.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---------------------------------------------------------------------------------------------
// sin / cos of the scattering angles, reproducible bit for bit on a CPU: 512-entry tables of
// sin(2 pi k / 512) and cos(2 pi k / 512), rounded from double, and the angle addition theorem for the remainder
//   sin(a + b) = sin a cos b + cos a sin b,   cos(a + b) = cos a cos b - sin a sin b,   0 <= b < 2 pi / 256
// with cos b = 1 - b^2/2 and sin b = b - b^3/6 (truncation < 1.6e-8 and < 8e-11; the table entries carry the
// usual 6e-8 rounding).  mul / fmaf only, no quadrant logic, no conversions on the slow pipe.
// The tables sit in shared memory (sin: 2 KB, then cos: 2 KB, per CTA); `at` is the BYTE offset of the entry
// inside either table.  Operation order (the CPU twin repeats it):
//   b2 = b*b; cb = fma(b2, -1/2, 1); sb = fma(b*b2, -1/6, b); s = fma(sa, cb, ca*sb); c = fma(ca, cb, -(sa*sb))
// ---------------------------------------------------------------------------------------------
#define PCL_TRIG_N 512
#define PCL_TRIG_BYTES (PCL_TRIG_N * 8)
#define PCL_TRIG_COS (PCL_TRIG_N * 4) /* byte offset of the cos table */
#define PCL_TWO_PI_256 0x1.921fb6p-6f /* 2 pi / 256 : theta = (k8 + f) * this  */
#define PCL_PI_256 0x1.921fb6p-7f     /* pi / 256   : phi   = (k8 + f) * this  */
#define PCL_MSIXTH -0x1.555556p-3f

__device__ __forceinline__ void pcl_sincos_tab(const unsigned char *tab, uint32_t at, float b, float &s, float &c) {
    const float sa = *reinterpret_cast<const float *>(tab + at);
    const float ca = *reinterpret_cast<const float *>(tab + PCL_TRIG_COS + at);
    const float b2 = b * b;
    const float cb = fmaf(b2, -0.5f, 1.0f);
    const float sb = fmaf(b * b2, PCL_MSIXTH, b);
    s = fmaf(sa, cb, ca * sb);
    c = fmaf(ca, cb, -(sa * sb));
}

// the same for two photons at once (entries at0 / at1, remainders in the halves of b)
__device__ __forceinline__ void pcl_sincos_tab2(const unsigned char *tab, uint32_t at0, uint32_t at1, f32x2 b, f32x2 &s, f32x2 &c) {
    const f32x2 sa = pk(*reinterpret_cast<const float *>(tab + at0), *reinterpret_cast<const float *>(tab + at1));
    const f32x2 ca = pk(*reinterpret_cast<const float *>(tab + PCL_TRIG_COS + at0), *reinterpret_cast<const float *>(tab + PCL_TRIG_COS + at1));
    const f32x2 b2 = mul2(b, b);
    const f32x2 cb = fma2(b2, pk(-0.5f, -0.5f), pk(1.0f, 1.0f));
    const f32x2 sb = fma2(mul2(b, b2), pk(PCL_MSIXTH, PCL_MSIXTH), b);
    s = fma2(sa, cb, mul2(ca, sb));
    // -(sa*sb) as 0 - sa*sb: the packed PTX has no negation, and ptxas is free to contract the explicitly rounded packed
    // mul + sub into one FFMA2 (it does; scalar mul + sub it leaves alone): written this way the value is
    // fma(ca, cb, -round(sa*sb)) whether or not it contracts, exactly what the scalar form and the CPU twin compute
    c = fma2(ca, cb, sub2(pk(0.f, 0.f), mul2(sa, sb)));
}

// PCL_SCATTER_SFU: sin and cos of two angles (radians) on the special-function unit (MUFU.SIN / MUFU.COS after the
// range-reduction multiply): 2^-20.9 absolute error on [-pi, pi], a little more up to 2 pi.
__device__ __forceinline__ void pcl_sincos_sfu2(f32x2 ang, f32x2 &s, f32x2 &c) {
    float a0, a1;
    upk(ang, a0, a1);
    s = pk(__sinf(a0), __sinf(a1));
    c = pk(__cosf(a0), __cosf(a1));
}

// One photon's random numbers for one timestep, in the form the step body consumes them.
struct pcl_draw3 {
    float ur;         // U[0,1) of the collision test (24 bits)
    uint32_t at, ap;  // byte offsets of the table entries below theta and below phi
    float bt, bp;     // remainders: theta = entry angle + bt (bt < 2 pi/256), phi = entry angle + bp (bp < pi/256)
    float tf, pf;     // RAW draws only: the whole 24-bit theta field and 16-bit phi field as floats (PCL_SCATTER_SFU)
};

// from one Philox2x32 block (w0, w1):  ur = w0[31:8];  theta = 2 pi * w1[31:8] / 2^24;  phi = pi * (w1[7:0] : w0[7:0]) / 2^16.
// RAW form: ur, bt, bp are the integer fields converted to float, NOT yet scaled; the packed step body applies the three
// power-of-two-exact scalings to two photons per instruction (and folds 2^-24 into 1/k).  pcl_draw_finish scales one draw.
#define PCL_BT_SCALE (PCL_TWO_PI_256 * 0x1p-16f)
#define PCL_BP_SCALE (PCL_PI_256 * 0x1p-8f)
#define PCL_SFU_T_SCALE (PCL_TWO_PI_256 * 0x1p-16f) /* 2 pi / 2^24: theta from the whole 24-bit field */
#define PCL_SFU_P_SCALE (PCL_PI_256 * 0x1p-8f)      /*   pi / 2^16: phi from the whole 16-bit field   */
__device__ __forceinline__ pcl_draw3 pcl_draw_bits_raw(uint32_t w0, uint32_t w1) {
    pcl_draw3 d;
    d.ur = (float)(w0 >> 8);
    d.at = (w1 >> 21) & 0x7f8u;  // k8 = w1[31:24], entry 2*k8, 4 bytes each
    d.bt = (float)((w1 >> 8) & 0xffffu);
    d.ap = (w1 & 0xffu) << 2;  // k8 = w1[7:0], entry k8
    d.bp = (float)(w0 & 0xffu);
    d.tf = (float)(w1 >> 8);                                // dead code unless the SFU form is instantiated
    d.pf = (float)(((w1 & 0xffu) << 8) | (w0 & 0xffu));
    return d;
}
__device__ __forceinline__ pcl_draw3 pcl_draw_finish(pcl_draw3 d) {
    d.ur = d.ur * 0x1p-24f;
    d.bt = d.bt * PCL_BT_SCALE;
    d.bp = d.bp * PCL_BP_SCALE;
    return d;
}
__device__ __forceinline__ pcl_draw3 pcl_draw_bits(uint32_t w0, uint32_t w1) { return pcl_draw_finish(pcl_draw_bits_raw(w0, w1)); }

// from injected uniforms (the reference's host draws rtheta = 2 pi u, rphi = pi u, rand; light.py:285)
__device__ __forceinline__ pcl_draw3 pcl_draw_floats(float ut, float up, float ur) {
    pcl_draw3 d;
    d.ur = ur;
    const float tt = ut * 256.0f, kt = floorf(tt);
    d.at = ((uint32_t)(int)kt & 0xffu) << 3;
    d.bt = (tt - kt) * PCL_TWO_PI_256;
    const float tp = up * 256.0f, kp = floorf(tp);
    d.ap = ((uint32_t)(int)kp & 0xffu) << 2;
    d.bp = (tp - kp) * PCL_PI_256;
    d.tf = d.pf = 0.f;
    return d;
}

// copy the tables (global, built once per context: sin[512] then cos[512]) into this CTA's shared memory; ends with
// a CTA barrier
__device__ __forceinline__ void pcl_trig_to_shared(unsigned char *s_tab, const float *g_tab) {
    for (uint32_t q = threadIdx.x; q < 2 * PCL_TRIG_N; q += blockDim.x) reinterpret_cast<float *>(s_tab)[q] = g_tab[q];
    __syncthreads();
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

// ---------------------------------------------------------------------------------------------
// TMA engine, 1-D form: bulk copies global -> shared that complete on an mbarrier (cp.async.bulk ...
// mbarrier::complete_tx), used by the TMA-staged photon step and by the all-pairs gravity kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pcl_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pcl_mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pcl_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void pcl_mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pcl_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pcl_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pcl_smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(pcl_smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void pcl_mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PCL_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PCL_DONE_%=;\n"
        "bra PCL_WAIT_%=;\n"
        "PCL_DONE_%=:\n"
        "}\n" ::"r"(pcl_smem_u32(b)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void pcl_mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pcl_smem_u32(b)) : "memory");
}

// per-thread asynchronous 16-byte copies global -> shared (LDGSTS), grouped and awaited by the issuing thread
__device__ __forceinline__ void pcl_cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void pcl_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void pcl_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ float &pcl_f4(float4 &v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ uint32_t &pcl_u4(uint4 &v, int i) { return (&v.x)[i]; }
