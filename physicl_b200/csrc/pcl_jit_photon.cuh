// Photon kernels that only exist at run time: ScatterIsotropicStep(variable_n=True) splices a user
// expression for the number density n(r) into the kernel text (physicl/light.py:295-299).  This file
// is compiled by NVRTC (jit.cu) for sm_100a with the expression in PCL_USER_N_EXPR; it reuses the
// hand-written step body (pcl_photon_body.cuh), so the kinematics, direction law, Philox streams,
// escape sphere and tallies are the very same code as in the pre-compiled kernels.
//
// Defines expected from the generated translation unit:
//   PCL_USER_N_EXPR   the user's expression text: OpenCL-C over r0[gid], r1[gid], r2[gid], E[gid],
//                     d0[gid]..d2[gid], norm, A, n with double arithmetic (pow, exp, sqrt, ...)
//   PCL_JIT_WAVE      1: pcoll *= (h c / E)^-4 (light.py:300-301)
//   PCL_JIT_DEL       1: scattered photons are removed
#pragma once
#include "physicl_b200.h"
#include "pcl_device.cuh"
#include "pcl_photon_body.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_F
#define M_PI_F 3.14159274101257f
#endif

// Fused timestep (kinematics -> scatter with n(r) -> escape -> tallies), in place, four photons per
// thread with 128-bit accesses when the planes allow it.
template <bool INJ>
__device__ __forceinline__ void pcl_jit_step(const pcl_soa &p, const StepK &K, int64_t *row, int aligned) {
    constexpr bool WAVE = PCL_JIT_WAVE != 0, DEL = PCL_JIT_DEL != 0;
    __shared__ __align__(16) unsigned char s_tab[PCL_TRIG_BYTES];
    __shared__ unsigned int s_acc[C_N];
    if (threadIdx.x < C_N) s_acc[threadIdx.x] = 0u;
    pcl_trig_to_shared(s_tab, K.trig);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (aligned) {
        const uint64_t nvec = pcl_valid_slots(p) / 4;
        const bool has_e = p.e != nullptr, has_id = p.id != nullptr;
        const uint64_t lane = threadIdx.x & 31u;
        // whole warps enter the loop body together (warp-level tally hand-over)
        for (uint64_t g0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); g0 < nvec; g0 += stride) {
            const uint64_t g = g0 + lane;
            const bool in = g < nvec;
            const uint64_t i = (in ? g : g0) * 4;
            float4 x = pcl_ld4(p.x + i), y = pcl_ld4(p.y + i), z = pcl_ld4(p.z + i);
            float4 vx = pcl_ld4(p.vx + i), vy = pcl_ld4(p.vy + i), vz = pcl_ld4(p.vz + i);
            float4 e = make_float4(1.f, 1.f, 1.f, 1.f);
            if (WAVE || has_e) e = pcl_ld4(p.e + i);
            uint4 id = make_uint4((uint32_t)i, (uint32_t)i + 1u, (uint32_t)i + 2u, (uint32_t)i + 3u);
            if (has_id) id = pcl_ld4u(p.id + i);
            float4 ut4 = make_float4(0.f, 0.f, 0.f, 0.f), up4 = ut4, ur4 = ut4;
            if (INJ) {
                ut4 = pcl_ld4(K.u_theta + i);
                up4 = pcl_ld4(K.u_phi + i);
                ur4 = pcl_ld4(K.u_rand + i);
            }
            uint4 nsc = make_uint4(0u, 0u, 0u, 0u);
            if (p.nscat) nsc = pcl_ld4u(p.nscat + i);
            if (!in) x.x = x.y = x.z = x.w = __int_as_float(0x7fc00000);
            pcl_step_group4_masked<WAVE, DEL, INJ, true, true>(p, K, s_tab, i, x, y, z, vx, vy, vz, e, id, nsc, ut4, up4, ur4, s_acc, in);
        }
        pcl_step_tail<WAVE, DEL, INJ, true, true>(p, K, s_tab, nvec * 4, s_acc);
    } else {
        const uint64_t end = pcl_valid_slots(p);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride)
            pcl_step_scalar<WAVE, DEL, INJ, true, true>(p, K, s_tab, i, s_acc);
    }
    __syncthreads();
    pcl_row_to_global(s_acc, row);
}

extern "C" __global__ void __launch_bounds__(256) pcl_jit_photon_step(pcl_soa p, StepK K, int64_t *row, int aligned) {
    if (K.u_rand)
        pcl_jit_step<true>(p, K, row, aligned);
    else
        pcl_jit_step<false>(p, K, row, aligned);
}

// Stand-alone scatter on the dr planes the kinematics step wrote: what CLProgram.run launches
// (physicl/__init__.py:656) plus the host write-back loop (light.py:325-331), as pcl_k_scatter does
// for the fixed laws.  n(r) is evaluated at the current position (kinematics has already run).
template <bool INJ>
__device__ __forceinline__ void pcl_jit_scatter_body(const pcl_soa &p, const StepK &K, int32_t *flags, int64_t *row,
                                                     uint64_t n) {
    constexpr bool WAVE = PCL_JIT_WAVE != 0, DEL = PCL_JIT_DEL != 0;
    __shared__ __align__(16) unsigned char s_tab[PCL_TRIG_BYTES];
    pcl_trig_to_shared(s_tab, K.trig);
    uint32_t cnt[C_PLANE0];
#pragma unroll
    for (int q = 0; q < C_PLANE0; ++q) cnt[q] = 0u;
    const float qnan = __int_as_float(0x7fc00000);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float xx = p.x[i];
        int32_t flag = 0;
        if (xx == xx) {
            cnt[C_LIVEIN] += 1u;
            const pcl_draw3 d = INJ ? pcl_draw_floats(K.u_theta[i], K.u_phi[i], K.u_rand[i])
                                    : pcl_draw_at(K, K.step, (uint32_t)p.id_base + (p.id ? p.id[i] : (uint32_t)i));
            float vx = 0.f, vy = 0.f, vz = 0.f;
            const float e = p.e ? p.e[i] : 1.f;
            const float dx = p.dx[i], dy = p.dy[i], dz = p.dz[i];
            const double kn = K.kd * pcl_user_n((double)xx, (double)p.y[i], (double)p.z[i], (double)e * K.e0, (double)dx,
                                                (double)dy, (double)dz, (double)pcl_norm3(dx, dy, dz), K.a_slot, K.n_slot);
            const bool hit = pcl_scatter_one<WAVE, DEL, true>(true, dx, dy, dz, e, d, K, s_tab, vx, vy, vz, kn);
            if (hit) {
                flag = 1;
                cnt[C_SCAT] += 1u;
                if (DEL) {
                    cnt[C_ABS] += 1u;
                    p.x[i] = qnan;
                } else {
                    p.vx[i] = vx;
                    p.vy[i] = vy;
                    p.vz[i] = vz;
                    if (p.nscat) p.nscat[i] += 1u;
                }
            }
            if (!(DEL && hit)) cnt[C_ALIVE] += 1u;
        }
        if (flags) flags[i] = flag;
    }
    if (row) pcl_flush_tally(cnt, row, 0u);
}

extern "C" __global__ void __launch_bounds__(256) pcl_jit_scatter(pcl_soa p, StepK K, int32_t *flags, int64_t *row,
                                                                  uint64_t n) {
    if (K.u_rand)
        pcl_jit_scatter_body<true>(p, K, flags, row, n);
    else
        pcl_jit_scatter_body<false>(p, K, flags, row, n);
}
