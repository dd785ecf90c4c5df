// Device forms of the reference's list-producing measure steps:
//   ScatterMeasureStep(measure_E=True)  physicl/light.py:380-402  -> energies of the photons that crossed a plane
//   TracePathMeasureStep                physicl/light.py:447-460  -> every object's position at every timestep
// Both are HBM-bound gathers: the first appends (id, e) pairs through one atomic range reservation per
// warp, the second scatters r by particle id into the timestep's slab of a trajectory buffer.
#include "pcl_common.cuh"

// crossing rule of light.py:385-399, identical to the tally kernels
__device__ __forceinline__ bool pcl_crossed(float r, float d, float loc) {
    float prev = r - d;
    return (prev <= loc && loc <= r) || (prev >= loc && loc >= r);
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_plane_crossers(pcl_soa p, pcl_planes pl, uint32_t *out_id, float *out_e, unsigned long long *counts, uint64_t cap) {
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    const uint64_t n_round = (p.n + 31) / 32 * 32;  // whole warps stay in the loop: ballots need them
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < p.n;
        float x = in ? p.x[i] : __int_as_float(0x7fc00000);
        const bool live = x == x;
        float y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, e = 1.f;
        uint32_t id = (uint32_t)i;
        if (live) {
            y = p.y[i];
            z = p.z[i];
            dx = p.dx[i];
            dy = p.dy[i];
            dz = p.dz[i];
            if (p.e) e = p.e[i];
            if (p.id) id = p.id[i];
        }
        for (uint32_t q = 0; q < pl.count; ++q) {
            const uint32_t ax = pl.axis[q];
            const float r = ax == 0 ? x : (ax == 1 ? y : z), d = ax == 0 ? dx : (ax == 1 ? dy : dz);
            const bool hit = live && pcl_crossed(r, d, pl.loc[q]);
            const uint32_t b = __ballot_sync(0xffffffffu, hit);
            if (!b) continue;
            const uint32_t lane = threadIdx.x & 31u;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&counts[q], (unsigned long long)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned long long o = base + __popc(b & ((1u << lane) - 1u));
                if (o < cap) {
                    out_id[(uint64_t)q * cap + o] = id;
                    out_e[(uint64_t)q * cap + o] = e;
                }
            }
        }
    }
}

extern "C" int pcl_plane_crossers(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_planes *planes,
                                  uint32_t *out_id, float *out_e, uint64_t *counts_dev, uint64_t cap) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p && planes && out_id && out_e && counts_dev, "null argument");
    PCL_REQUIRE(ctx, planes->count >= 1 && planes->count <= PCL_MAX_PLANES, "1..8 planes");
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->dx && p->dy && p->dz, "r and dr planes are required");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    cudaStream_t st = (cudaStream_t)stream;
    PCL_CUDA(ctx, cudaMemsetAsync(counts_dev, 0, planes->count * sizeof(uint64_t), st));
    if (p->n == 0) return 0;
    unsigned grid = pcl_stream_grid(ctx, p->n, PCL_BLOCK, 8);
    pcl_k_plane_crossers<<<grid, PCL_BLOCK, 0, st>>>(*p, *planes, out_id, out_e, (unsigned long long *)counts_dev, cap);
    PCL_LAUNCHED(ctx);
    return 0;
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_trace(pcl_soa p, float *slab, uint64_t n_ids, uint32_t *nscat_by_id) {
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < p.n; i += stride) {
        const float x = p.x[i];
        if (x != x) continue;  // retired: the slab keeps its NaN fill ("object does not exist", light.py:435)
        const uint64_t id = p.id ? (uint64_t)p.id[i] : i;
        if (id >= n_ids) continue;
        slab[id] = x;
        slab[n_ids + id] = p.y[i];
        slab[2 * n_ids + id] = p.z[i];
        if (nscat_by_id && p.nscat) nscat_by_id[id] = p.nscat[i];  // survives the photon's retirement
    }
}

extern "C" int pcl_trace_positions(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float *slab, uint64_t n_ids,
                                   uint32_t *nscat_by_id) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p && slab, "null argument");
    PCL_REQUIRE(ctx, p->x && p->y && p->z, "r planes are required");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    if (p->n == 0) return 0;
    unsigned grid = pcl_stream_grid(ctx, p->n, PCL_BLOCK, 8);
    pcl_k_trace<<<grid, PCL_BLOCK, 0, (cudaStream_t)stream>>>(*p, slab, n_ids, nscat_by_id);
    PCL_LAUNCHED(ctx);
    return 0;
}
