// Stable stream compaction of live photons: the device form of Simulation.remove_obj
// (reference physicl/__init__.py:455-459, called once per flagged photon from light.py:203-205 and
// :258-260 -- an O(N^2) list.remove loop there).
//
// Three launches, all HBM-bound:
//   1. per-tile live counts            reads x                        4 B / slot
//   2. exclusive scan of tile counts   one block, ntiles * 4 B
//   3. scatter                         re-reads x, moves every plane  4 B / slot + 2 * B_state / live
// Ranks inside a tile come from warp ballots + popc, so a warp with no live lane does no stores,
// and the order of survivors is preserved (results never depend on block scheduling).
#include "pcl_common.cuh"

#define PCL_TILE (PCL_BLOCK * 4)

__device__ __forceinline__ uint32_t pcl_live_mask4(const float *x, uint64_t i, uint64_t n) {
    uint32_t m = 0u;
    if (i + 3 < n) {
        float4 v = pcl_ld4(x + i);
        m = (v.x == v.x ? 1u : 0u) | (v.y == v.y ? 2u : 0u) | (v.z == v.z ? 4u : 0u) | (v.w == v.w ? 8u : 0u);
    } else {
        for (int l = 0; l < 4; ++l)
            if (i + l < n) {
                float v = x[i + l];
                m |= (v == v) ? (1u << l) : 0u;
            }
    }
    return m;
}

__global__ void __launch_bounds__(PCL_BLOCK) pcl_k_count_live(const float *x, uint64_t n, uint32_t *counts) {
    __shared__ uint32_t s_w[PCL_WARPS];
    uint64_t i = ((uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x) * 4;
    uint32_t c = (i < n) ? __popc(pcl_live_mask4(x, i, n)) : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < PCL_WARPS; ++w) t += s_w[w];
        counts[blockIdx.x] = t;
    }
}

// one block: counts[0..m) -> exclusive offsets in place, total -> *total_out
__global__ void __launch_bounds__(1024) pcl_k_scan_tiles(uint32_t *counts, uint32_t m, uint64_t *total_out) {
    __shared__ uint64_t s_sum[1024];
    const uint32_t per = (m + 1023u) / 1024u;
    const uint32_t b = threadIdx.x * per;
    const uint32_t e = min(b + per, m);
    uint64_t s = 0;
    for (uint32_t i = b; i < e; ++i) s += counts[i];
    s_sum[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partial sums
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint64_t v = (threadIdx.x >= off) ? s_sum[threadIdx.x - off] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = s_sum[threadIdx.x] - s;
    for (uint32_t i = b; i < e; ++i) {
        uint32_t c = counts[i];
        counts[i] = (uint32_t)run;
        run += c;
    }
    if (threadIdx.x == 1023) *total_out = s_sum[1023];
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_compact(pcl_soa s, pcl_soa d, const uint32_t *offsets) {
    __shared__ uint32_t s_w[PCL_WARPS];
    const uint64_t i = ((uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x) * 4;
    const uint32_t m = (i < s.n) ? pcl_live_mask4(s.x, i, s.n) : 0u;
    const uint32_t c = __popc(m);
    // exclusive rank of this thread's first survivor inside the tile
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += v;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < PCL_WARPS; ++w)
        if (w < (int)(threadIdx.x >> 5)) wbase += s_w[w];
    if (__ballot_sync(0xffffffffu, m != 0u) == 0u) return;  // whole warp retired: nothing to move
    if (!m) return;
    uint64_t o = (uint64_t)offsets[blockIdx.x] + wbase + (inc - c);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        if (!(m & (1u << l))) continue;
        const uint64_t a = i + l;
        d.x[o] = s.x[a];
        d.y[o] = s.y[a];
        d.z[o] = s.z[a];
        d.vx[o] = s.vx[a];
        d.vy[o] = s.vy[a];
        d.vz[o] = s.vz[a];
        if (s.dx && d.dx) {
            d.dx[o] = s.dx[a];
            d.dy[o] = s.dy[a];
            d.dz[o] = s.dz[a];
        }
        if (s.ax && d.ax) {
            d.ax[o] = s.ax[a];
            d.ay[o] = s.ay[a];
            d.az[o] = s.az[a];
        }
        if (s.e && d.e) d.e[o] = s.e[a];
        if (s.nscat && d.nscat) d.nscat[o] = s.nscat[a];
        d.id[o] = s.id ? s.id[a] : (uint32_t)a;
        ++o;
    }
}

extern "C" int pcl_compact(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *src, const pcl_soa *dst,
                           uint64_t *n_live_dev) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, src && dst && n_live_dev, "null argument");
    PCL_REQUIRE(ctx, src->x && src->y && src->z && src->vx && src->vy && src->vz, "src r and v planes are required");
    PCL_REQUIRE(ctx, dst->x && dst->y && dst->z && dst->vx && dst->vy && dst->vz && dst->id,
                "dst needs r, v and id planes");
    PCL_REQUIRE(ctx, dst->x != src->x, "compaction is out of place");
    PCL_REQUIRE(ctx, src->n_dev == nullptr, "compaction needs the exact slot count on the host (n_dev must be null)");
    PCL_REQUIRE(ctx, src->n < (1ull << 32), "a shard holds fewer than 2^32 slots");
    PCL_REQUIRE(ctx, pcl_aligned16(src->x), "src x plane must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (src->n == 0) {
        PCL_CUDA(ctx, cudaMemsetAsync(n_live_dev, 0, sizeof(uint64_t), st));
        return 0;
    }
    const uint64_t ntiles = (src->n + PCL_TILE - 1) / PCL_TILE;
    if (ctx->scan_cap < ntiles) {
        if (ctx->scan_buf) PCL_CUDA(ctx, cudaFree(ctx->scan_buf));
        ctx->scan_buf = nullptr;
        ctx->scan_cap = 0;
        size_t cap = (size_t)ntiles + (size_t)ntiles / 4 + 1024;
        PCL_CUDA(ctx, cudaMalloc(&ctx->scan_buf, cap * sizeof(uint32_t)));
        ctx->scan_cap = cap;
    }
    pcl_k_count_live<<<(unsigned)ntiles, PCL_BLOCK, 0, st>>>(src->x, src->n, ctx->scan_buf);
    PCL_LAUNCHED(ctx);
    pcl_k_scan_tiles<<<1, 1024, 0, st>>>(ctx->scan_buf, (uint32_t)ntiles, n_live_dev);
    PCL_LAUNCHED(ctx);
    pcl_k_compact<<<(unsigned)ntiles, PCL_BLOCK, 0, st>>>(*src, *dst, ctx->scan_buf);
    PCL_LAUNCHED(ctx);
    return 0;
}
