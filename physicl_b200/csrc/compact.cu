// Stable stream compaction of live photons: the device form of Simulation.remove_obj
// (reference physicl/__init__.py:455-459, called once per flagged photon from light.py:203-205 and
// :258-260 -- an O(N^2) list.remove loop there).
//
// Three launches, all HBM-bound:
//   1. per-tile live counts            reads x                        4 B / slot
//   2. exclusive scan of tile counts   one block, ntiles * 4 B
//   3. move                            re-reads x, moves every plane  (4 + B_state) B / slot + B_state / live
// Ranks inside a tile come from 4-bit live masks + a shuffle scan, survivors are staged in shared
// memory and written as contiguous runs; the order of survivors is preserved (results never depend on
// block scheduling).
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

// One tile (1024 slots) per CTA.  Every plane is read with 128-bit loads, its survivors are staged in
// shared memory at their rank inside the tile and then written out as one contiguous, coalesced run
// (128-byte aligned lines, as in the retire-and-compact photon kernel), plane after plane.
__device__ __forceinline__ float4 pcl_ld4_guarded(const float *p, uint64_t i, uint64_t n) {
    if (i + 3 < n) return pcl_ld4(p + i);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < 4; ++l)
        if (i + l < n) pcl_f4(v, l) = p[i + l];
    return v;
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_compact(pcl_soa s, pcl_soa d, const uint32_t *offsets) {
    __shared__ uint32_t s_w[PCL_WARPS];
    __shared__ float s_stage[PCL_TILE];
    const uint64_t i = ((uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x) * 4;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t m = (i < s.n) ? pcl_live_mask4(s.x, i, s.n) : 0u;
    const uint32_t c = __popc(m);
    // exclusive rank of this thread's first survivor inside the tile
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PCL_WARPS; ++w) {
        const uint32_t t = s_w[w];
        wbase += (w < (int)wid) ? t : 0u;
        tot += t;
    }
    if (tot == 0) return;  // whole tile retired: nothing to move
    const uint32_t rank = wbase + (inc - c);
    const uint64_t base = offsets[blockIdx.x];
    const uint32_t skew = (uint32_t)(base & 31ull);
    const float *src[15] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.dx, s.dy, s.dz, s.ax, s.ay, s.az, s.e, (const float *)s.nscat,
                            (const float *)s.id};
    float *dst[15] = {d.x, d.y, d.z, d.vx, d.vy, d.vz, d.dx, d.dy, d.dz, d.ax, d.ay, d.az, d.e, (float *)d.nscat, (float *)d.id};
#pragma unroll
    for (int q = 0; q < 15; ++q) {
        const bool is_id = q == 14;
        if (!dst[q] || (!src[q] && !is_id)) continue;  // a plane moves when both sides have it; ids always do
        float4 v;
        if (src[q]) {
            v = (i < s.n) ? pcl_ld4_guarded(src[q], i, s.n) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {  // no id plane yet: the id of a slot is its index
            v = make_float4(__uint_as_float((uint32_t)i), __uint_as_float((uint32_t)i + 1u), __uint_as_float((uint32_t)i + 2u),
                            __uint_as_float((uint32_t)i + 3u));
        }
        uint32_t rk = rank;
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (m & (1u << l)) s_stage[rk++] = pcl_f4(v, l);
        __syncthreads();
        for (uint32_t qq = threadIdx.x; qq < tot + skew; qq += PCL_BLOCK)
            if (qq >= skew) dst[q][base + (qq - skew)] = s_stage[qq - skew];
        __syncthreads();
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
