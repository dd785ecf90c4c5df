// All-pairs Newtonian gravity (softened):  a_i = G * sum_j m_j (r_j - r_i) / (|r_ij|^2 + eps2)^(3/2)
// Not in the reference (no force step exists there); it sits behind the
// reference's Step API as NewtonianGravityStep.  FP32 FMA-pipe bound, no tensor cores: this is not a
// dense contraction.
//
// Work decomposition: a CTA of GRAV_THREADS threads owns GRAV_THREADS*IB i-bodies (IB per thread,
// in registers); j-bodies stream through shared memory in tiles of GRAV_JT float4 (x,y,z,m),
// double-buffered with cp.async so the next tile lands while the current one is consumed.  Every
// lane reads the same j-body (LDS.128 broadcast, conflict-free).
// Per interaction: 3 FADD + 3 FFMA + 1 MUFU.RSQ + 3 FMUL + 3 FFMA = 12 FP32-pipe ops + 1 SFU op,
// counted as 20 FLOP (the customary all-pairs figure).
#include <stdlib.h>

#include "pcl_common.cuh"

#define GRAV_JT 256

__device__ __forceinline__ float pcl_rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void pcl_cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void pcl_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void pcl_cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

template <int IB, int GRAV_THREADS>
__global__ void __launch_bounds__(GRAV_THREADS)
pcl_k_gravity(const float4 *__restrict__ pi, uint64_t n_local, const float4 *__restrict__ pj, uint64_t n_total,
              float G, float eps2, float *ax, float *ay, float *az, int accumulate) {
    __shared__ float4 s_j[2][GRAV_JT];
    const uint64_t i0 = (uint64_t)blockIdx.x * (GRAV_THREADS * IB) + threadIdx.x;
    float xi[IB], yi[IB], zi[IB], axi[IB], ayi[IB], azi[IB];
#pragma unroll
    for (int m = 0; m < IB; ++m) {
        uint64_t i = i0 + (uint64_t)m * GRAV_THREADS;
        float4 b = (i < n_local) ? pi[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        xi[m] = b.x;
        yi[m] = b.y;
        zi[m] = b.z;
        axi[m] = ayi[m] = azi[m] = 0.f;
    }
    const uint64_t ntile = (n_total + GRAV_JT - 1) / GRAV_JT;
    auto issue = [&](uint64_t t, int buf) {
        for (int q = threadIdx.x; q < GRAV_JT; q += GRAV_THREADS) {
            uint64_t j = t * GRAV_JT + q;
            if (j < n_total) pcl_cp_async16(&s_j[buf][q], pj + j);
            else s_j[buf][q] = make_float4(0.f, 0.f, 0.f, 0.f);  // zero mass: contributes nothing
        }
        pcl_cp_async_commit();
    };
    issue(0, 0);
    for (uint64_t t = 0; t < ntile; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < ntile) {
            issue(t + 1, buf ^ 1);
            pcl_cp_async_wait<1>();
        } else {
            pcl_cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll 8
        for (int q = 0; q < GRAV_JT; ++q) {
            const float4 b = s_j[buf][q];
#pragma unroll
            for (int m = 0; m < IB; ++m) {
                float dx = b.x - xi[m], dy = b.y - yi[m], dz = b.z - zi[m];
                float r2 = fmaf(dx, dx, eps2);
                r2 = fmaf(dy, dy, r2);
                r2 = fmaf(dz, dz, r2);
                float rinv = pcl_rsqrt_approx(r2);
                float rinv2 = rinv * rinv;
                float s = b.w * rinv;
                s = s * rinv2;
                axi[m] = fmaf(s, dx, axi[m]);
                ayi[m] = fmaf(s, dy, ayi[m]);
                azi[m] = fmaf(s, dz, azi[m]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < IB; ++m) {
        uint64_t i = i0 + (uint64_t)m * GRAV_THREADS;
        if (i < n_local) {
            float gx = G * axi[m], gy = G * ayi[m], gz = G * azi[m];
            if (accumulate) {
                gx += ax[i];
                gy += ay[i];
                gz += az[i];
            }
            ax[i] = gx;
            ay[i] = gy;
            az[i] = gz;
        }
    }
}

extern "C" int pcl_gravity_accel(pcl_ctx *ctx, uintptr_t stream, const float *posm_local, uint64_t n_local,
                                 const float *posm_all, uint64_t n_total, float G, float eps2, float *ax, float *ay,
                                 float *az, int accumulate) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, posm_local && posm_all && ax && ay && az, "null argument");
    PCL_REQUIRE(ctx, pcl_aligned16(posm_local) && pcl_aligned16(posm_all), "posm arrays must be 16-byte aligned");
    PCL_REQUIRE(ctx, eps2 > 0.f, "softening eps2 must be > 0 (the i == j term relies on it)");
    if (n_local == 0 || n_total == 0) return 0;
    // Tile shape: IB i-bodies per thread x T threads per CTA.  Small CTAs keep the number of CTAs per
    // SM nearly uniform (262144 bodies -> 1024 CTAs of 256 bodies: 6.9 per SM), which matters because
    // every CTA does the same amount of work and the kernel is issue-bound.
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("PCL_GRAV_VARIANT");  // tuning aid; default chosen from ncu/bench data
        variant = e ? atoi(e) : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
#define PCL_GRAV(IB, T)                                                                                       \
    pcl_k_gravity<IB, T><<<(unsigned)((n_local + (T) * (IB)-1) / ((T) * (IB))), T, 0, st>>>(                  \
        (const float4 *)posm_local, n_local, (const float4 *)posm_all, n_total, G, eps2, ax, ay, az, accumulate)
    switch (variant) {
        case 1: PCL_GRAV(4, 128); break;
        case 2: PCL_GRAV(2, 128); break;
        case 3: PCL_GRAV(2, 64); break;
        case 4: PCL_GRAV(8, 64); break;
        case 5: PCL_GRAV(8, 32); break;
        case 6: PCL_GRAV(4, 32); break;
        default: PCL_GRAV(4, 64); break;
    }
#undef PCL_GRAV
    PCL_LAUNCHED(ctx);
    return 0;
}

// kick-drift (semi-implicit Euler, same ordering as the kinematics law: v first, then r)
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_kick_drift(uint64_t n, float4 *posm, float *vx, float *vy, float *vz, const float *ax, const float *ay,
                 const float *az, float dt, float *x, float *y, float *z) {
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n; i += stride) {
        float4 b = posm[i];
        float u = vx[i] + ax[i] * dt, v = vy[i] + ay[i] * dt, w = vz[i] + az[i] * dt;
        b.x = b.x + u * dt;
        b.y = b.y + v * dt;
        b.z = b.z + w * dt;
        vx[i] = u;
        vy[i] = v;
        vz[i] = w;
        posm[i] = b;
        if (x) {  // keep the SoA position planes of the store current
            x[i] = b.x;
            y[i] = b.y;
            z[i] = b.z;
        }
    }
}

extern "C" int pcl_gravity_kick_drift(pcl_ctx *ctx, uintptr_t stream, uint64_t n, float *posm, float *vx, float *vy,
                                      float *vz, const float *ax, const float *ay, const float *az, float dt,
                                      float *x, float *y, float *z) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, posm && vx && vy && vz && ax && ay && az, "null argument");
    PCL_REQUIRE(ctx, pcl_aligned16(posm), "posm must be 16-byte aligned");
    PCL_REQUIRE(ctx, (x && y && z) || (!x && !y && !z), "x, y, z planes come as a triple or not at all");
    if (n == 0) return 0;
    unsigned grid = pcl_stream_grid(ctx, n, PCL_BLOCK, 8);
    pcl_k_kick_drift<<<grid, PCL_BLOCK, 0, (cudaStream_t)stream>>>(n, (float4 *)posm, vx, vy, vz, ax, ay, az, dt, x, y, z);
    PCL_LAUNCHED(ctx);
    return 0;
}
