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
#include "pcl_common.cuh"

// ACCEL: 0 = reference law, 1 = per-particle a planes, 2 = uniform a
template <int ACCEL, bool WRITE_DR>
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_kinematics(pcl_soa p, float dt, float aux, float auy, float auz, uint64_t nvec) {
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
        float4 dx, dy, dz;
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
                                      uint64_t begin, uint64_t end) {
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
    if (ACCEL == 1) {
        vx = vx + p.ax[i] * dt;
        vy = vy + p.ay[i] * dt;
        vz = vz + p.az[i] * dt;
    } else if (ACCEL == 2) {
        vx = vx + aux * dt;
        vy = vy + auy * dt;
        vz = vz + auz * dt;
    }
    float dx = vx * dt, dy = vy * dt, dz = vz * dt;
    p.x[i] = p.x[i] + dx;
    p.y[i] = p.y[i] + dy;
    p.z[i] = p.z[i] + dz;
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
static int launch_kin(pcl_ctx *ctx, cudaStream_t st, const pcl_soa &p, float dt, const float *au) {
    bool aligned = pcl_aligned16(p.x) && pcl_aligned16(p.y) && pcl_aligned16(p.z) &&
                   pcl_aligned16(p.vx) && pcl_aligned16(p.vy) && pcl_aligned16(p.vz);
    if (WRITE_DR) aligned = aligned && pcl_aligned16(p.dx) && pcl_aligned16(p.dy) && pcl_aligned16(p.dz);
    if (ACCEL == 1) aligned = aligned && pcl_aligned16(p.ax) && pcl_aligned16(p.ay) && pcl_aligned16(p.az);
    uint64_t nvec = aligned ? p.n / 4 : 0;
    if (nvec) {
        unsigned grid = pcl_stream_grid(ctx, nvec, PCL_BLOCK, 8);
        pcl_k_kinematics<ACCEL, WRITE_DR><<<grid, PCL_BLOCK, 0, st>>>(p, dt, au[0], au[1], au[2], nvec);
        PCL_LAUNCHED(ctx);
    }
    uint64_t begin = nvec * 4;
    if (begin < p.n) {
        uint64_t rem = p.n - begin;
        unsigned grid = (unsigned)((rem + 255) / 256);
        pcl_k_kinematics_tail<ACCEL, WRITE_DR><<<grid, 256, 0, st>>>(p, dt, au[0], au[1], au[2], begin, p.n);
        PCL_LAUNCHED(ctx);
    }
    return 0;
}

static int kin_dispatch(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, float dt, int accel,
                        const float *a_uniform) {
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
        case 0: return launch_kin<0, false>(ctx, st, *p, dt, au);
        case 1: return launch_kin<0, true>(ctx, st, *p, dt, au);
        case 2: return launch_kin<1, false>(ctx, st, *p, dt, au);
        case 3: return launch_kin<1, true>(ctx, st, *p, dt, au);
        case 4: return launch_kin<2, false>(ctx, st, *p, dt, au);
        default: return launch_kin<2, true>(ctx, st, *p, dt, au);
    }
}

extern "C" int pcl_kinematics(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt, int accel,
                              const float *a_uniform) {
    PCL_ENTER(ctx);
    return kin_dispatch(ctx, (cudaStream_t)stream, p, dt, accel, a_uniform);
}

// nsteps steps as one CUDA graph launch (the graph holds one kernel node per step, so the
// accounting stays "one HBM round trip per step").  Needed where a step is ~10 us (1M particles):
// launch gaps would otherwise dominate.
extern "C" int pcl_kinematics_steps(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                                    int accel, const float *a_uniform, uint32_t nsteps) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p != nullptr, "null particle view");
    if (nsteps == 0 || p->n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    pcl_graph_key key;
    memset(&key, 0, sizeof(key));
    key.p = *p;
    key.dt = dt;
    key.accel = accel;
    if (a_uniform) memcpy(key.a, a_uniform, sizeof(key.a));
    key.nsteps = nsteps;
    key.stream = stream;
    if (!ctx->kin_graph || memcmp(&key, &ctx->kin_key, sizeof(key)) != 0) {
        if (ctx->kin_graph) {
            cudaGraphExecDestroy(ctx->kin_graph);
            ctx->kin_graph = nullptr;
        }
        // capture on a private stream so a legacy default stream argument is acceptable
        cudaStream_t cap;
        PCL_CUDA(ctx, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        PCL_CUDA(ctx, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        uint64_t before = ctx->launches;
        for (uint32_t s = 0; s < nsteps && rc == 0; ++s) rc = kin_dispatch(ctx, cap, p, dt, accel, a_uniform);
        ctx->launches = before;  // counted at replay time below
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cap, &g);
        cudaStreamDestroy(cap);
        if (rc != 0) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        PCL_CUDA(ctx, ce);
        PCL_CUDA(ctx, cudaGraphInstantiate(&ctx->kin_graph, g, 0));
        cudaGraphDestroy(g);
        ctx->kin_key = key;
    }
    PCL_CUDA(ctx, cudaGraphLaunch(ctx->kin_graph, st));
    const uint64_t per_step = (p->n / 4 ? 1 : 0) + (p->n % 4 ? 1 : 0);
    ctx->launches += (uint64_t)nsteps * per_step;
    return 0;
}
