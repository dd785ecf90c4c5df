// Newtonian kinematics step: the device form of NewtonianKinematicsStep.run
// (reference physicl/newton.py:14-16):   obj.dr = obj.v * sim.dt ; obj.r += obj.dr
//
// HBM-bound streaming kernel.  Algorithmic traffic (FP32 SoA):
//   reference law        read r,v (24 B) + write r,dr (24 B)              = 48 B / particle-step
//   accel, a planes      read r,v,a (36 B) + write r,v,dr (36 B)          = 72 B / particle-step
//   accel, uniform a     read r,v (24 B) + write r,v,dr (36 B)            = 60 B / particle-step
// One thread moves 4 consecutive particles per plane with 128-bit accesses; a warp therefore
// touches 512 contiguous bytes of each plane.  The grid is persistent (a multiple of the SM
// count) and walks tiles in grid-stride order.
#include <stdlib.h>

#include "pcl_common.cuh"

// ACCEL: 0 = reference law, 1 = per-particle a planes, 2 = uniform a
template <int ACCEL, bool WRITE_DR>
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_kinematics(pcl_soa p, float dt, float aux, float auy, float auz, uint64_t nvec, uint32_t nsteps) {
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; g < nvec; g += stride) {
        const uint64_t i = g * 4;
        float4 x = pcl_ld4(p.x + i), y = pcl_ld4(p.y + i), z = pcl_ld4(p.z + i);
        float4 vx = pcl_ld4(p.vx + i), vy = pcl_ld4(p.vy + i), vz = pcl_ld4(p.vz + i);
        float4 ax, ay, az;
        if (ACCEL == 1) {
            ax = pcl_ld4(p.ax + i);
            ay = pcl_ld4(p.ay + i);
            az = pcl_ld4(p.az + i);
        }
        float4 dx = make_float4(0.f, 0.f, 0.f, 0.f), dy = dx, dz = dx;
        // nsteps timesteps in registers: the same binary32 operations in the same order as nsteps
        // launches, so the result is bit-identical; only the last dr is observable
        for (uint32_t st = 0; st < nsteps; ++st)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (ACCEL == 1) {
                pcl_f4(vx, l) = pcl_f4(vx, l) + pcl_f4(ax, l) * dt;
                pcl_f4(vy, l) = pcl_f4(vy, l) + pcl_f4(ay, l) * dt;
                pcl_f4(vz, l) = pcl_f4(vz, l) + pcl_f4(az, l) * dt;
            } else if (ACCEL == 2) {
                pcl_f4(vx, l) = pcl_f4(vx, l) + aux * dt;
                pcl_f4(vy, l) = pcl_f4(vy, l) + auy * dt;
                pcl_f4(vz, l) = pcl_f4(vz, l) + auz * dt;
            }
            pcl_f4(dx, l) = pcl_f4(vx, l) * dt;
            pcl_f4(dy, l) = pcl_f4(vy, l) * dt;
            pcl_f4(dz, l) = pcl_f4(vz, l) * dt;
            pcl_f4(x, l) = pcl_f4(x, l) + pcl_f4(dx, l);
            pcl_f4(y, l) = pcl_f4(y, l) + pcl_f4(dy, l);
            pcl_f4(z, l) = pcl_f4(z, l) + pcl_f4(dz, l);
        }
        pcl_st4(p.x + i, x);
        pcl_st4(p.y + i, y);
        pcl_st4(p.z + i, z);
        if (ACCEL != 0) {
            pcl_st4(p.vx + i, vx);
            pcl_st4(p.vy + i, vy);
            pcl_st4(p.vz + i, vz);
        }
        if (WRITE_DR) {
            pcl_st4(p.dx + i, dx);
            pcl_st4(p.dy + i, dy);
            pcl_st4(p.dz + i, dz);
        }
    }
}

// scalar tail (n % 4 particles) and unaligned fallback
template <int ACCEL, bool WRITE_DR>
__global__ void pcl_k_kinematics_tail(pcl_soa p, float dt, float aux, float auy, float auz,
                                      uint64_t begin, uint64_t end, uint32_t nsteps) {
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float x = p.x[i], y = p.y[i], z = p.z[i];
    float vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
    float ax = 0.f, ay = 0.f, az = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    if (ACCEL == 1) {
        ax = p.ax[i];
        ay = p.ay[i];
        az = p.az[i];
    }
    for (uint32_t st = 0; st < nsteps; ++st) {
        if (ACCEL == 1) {
            vx = vx + ax * dt;
            vy = vy + ay * dt;
            vz = vz + az * dt;
        } else if (ACCEL == 2) {
            vx = vx + aux * dt;
            vy = vy + auy * dt;
            vz = vz + auz * dt;
        }
        dx = vx * dt;
        dy = vy * dt;
        dz = vz * dt;
        x = x + dx;
        y = y + dy;
        z = z + dz;
    }
    p.x[i] = x;
    p.y[i] = y;
    p.z[i] = z;
    if (ACCEL != 0) {
        p.vx[i] = vx;
        p.vy[i] = vy;
        p.vz[i] = vz;
    }
    if (WRITE_DR) {
        p.dx[i] = dx;
        p.dy[i] = dy;
        p.dz[i] = dz;
    }
}

template <int ACCEL, bool WRITE_DR>
static int launch_kin(pcl_ctx *ctx, cudaStream_t st, const pcl_soa &p, float dt, const float *au, uint32_t nsteps) {
    bool aligned = pcl_aligned16(p.x) && pcl_aligned16(p.y) && pcl_aligned16(p.z) &&
                   pcl_aligned16(p.vx) && pcl_aligned16(p.vy) && pcl_aligned16(p.vz);
    if (WRITE_DR) aligned = aligned && pcl_aligned16(p.dx) && pcl_aligned16(p.dy) && pcl_aligned16(p.dz);
    if (ACCEL == 1) aligned = aligned && pcl_aligned16(p.ax) && pcl_aligned16(p.ay) && pcl_aligned16(p.az);
    uint64_t nvec = aligned ? p.n / 4 : 0;
    if (nvec) {
        unsigned grid = pcl_stream_grid(ctx, nvec, PCL_BLOCK, 8);
        pcl_k_kinematics<ACCEL, WRITE_DR><<<grid, PCL_BLOCK, 0, st>>>(p, dt, au[0], au[1], au[2], nvec, nsteps);
        PCL_LAUNCHED(ctx);
    }
    uint64_t begin = nvec * 4;
    if (begin < p.n) {
        uint64_t rem = p.n - begin;
        unsigned grid = (unsigned)((rem + 255) / 256);
        pcl_k_kinematics_tail<ACCEL, WRITE_DR><<<grid, 256, 0, st>>>(p, dt, au[0], au[1], au[2], begin, p.n, nsteps);
        PCL_LAUNCHED(ctx);
    }
    return 0;
}

static int kin_dispatch(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, float dt, int accel,
                        const float *a_uniform, uint32_t nsteps) {
    PCL_REQUIRE(ctx, p != nullptr, "null particle view");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    if (p->n == 0) return 0;
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->vx && p->vy && p->vz, "r and v planes are required");
    const bool dr = p->dx != nullptr;
    if (dr) PCL_REQUIRE(ctx, p->dy && p->dz, "dr planes must come as a triple");
    static const float zero3[3] = {0.f, 0.f, 0.f};
    const float *au = a_uniform ? a_uniform : zero3;
    int mode = 0;
    if (accel) {
        if (p->ax) {
            PCL_REQUIRE(ctx, p->ay && p->az, "a planes must come as a triple");
            mode = 1;
        } else {
            PCL_REQUIRE(ctx, a_uniform != nullptr, "accel requested without a planes or a_uniform");
            mode = 2;
        }
    }
    switch (mode * 2 + (dr ? 1 : 0)) {
        case 0: return launch_kin<0, false>(ctx, st, *p, dt, au, nsteps);
        case 1: return launch_kin<0, true>(ctx, st, *p, dt, au, nsteps);
        case 2: return launch_kin<1, false>(ctx, st, *p, dt, au, nsteps);
        case 3: return launch_kin<1, true>(ctx, st, *p, dt, au, nsteps);
        case 4: return launch_kin<2, false>(ctx, st, *p, dt, au, nsteps);
        default: return launch_kin<2, true>(ctx, st, *p, dt, au, nsteps);
    }
}

// for hostpipe.cu: the same dispatch on a stream of the host pipeline
int pcl_kinematics_impl(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, float dt, int accel, const float *a_uniform,
                        uint32_t nsteps) {
    return kin_dispatch(ctx, st, p, dt, accel, a_uniform, nsteps);
}

extern "C" int pcl_kinematics(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt, int accel,
                              const float *a_uniform) {
    PCL_ENTER(ctx);
    return kin_dispatch(ctx, (cudaStream_t)stream, p, dt, accel, a_uniform, 1);
}

// nsteps timesteps of equal dt.  Particles do not interact, so a thread keeps its four particles in
// registers and applies the step nsteps times before writing back: ONE HBM round trip per launch
// instead of one per timestep, bit-identical to nsteps calls of pcl_kinematics (same binary32
// operations in the same order; only the last dr is observable, as with the reference's loop,
// physicl/newton.py:14-16).  PCL_KIN_FUSE caps the timesteps per launch (1 = one launch per timestep,
// the form the HBM-roofline figures are quoted on).
extern "C" int pcl_kinematics_steps(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                                    int accel, const float *a_uniform, uint32_t nsteps) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p != nullptr, "null particle view");
    if (nsteps == 0 || p->n == 0) return 0;
    static int fuse = -1;
    if (fuse < 0) {
        const char *e = getenv("PCL_KIN_FUSE");
        fuse = e ? atoi(e) : 4096;
        if (fuse < 1) fuse = 1;
    }
    for (uint32_t s = 0; s < nsteps;) {
        const uint32_t run = nsteps - s < (uint32_t)fuse ? nsteps - s : (uint32_t)fuse;
        int rc = kin_dispatch(ctx, (cudaStream_t)stream, p, dt, accel, a_uniform, run);
        if (rc) return rc;
        s += run;
    }
    return 0;
}
