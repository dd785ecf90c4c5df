// Stable stream compaction of live photons: the device form of Simulation.remove_obj
// (reference physicl/__init__.py:455-459, called once per flagged photon from light.py:203-205 and
// :258-260 -- an O(N^2) list.remove loop there).
//
// Three launches, all HBM-bound:
//   1. per-tile live counts            reads x (four tiles per CTA)   4 B / slot
//   2. exclusive scan of tile counts   one block, ntiles * 4 B (16-byte loads, shuffle scans)
//   3. move                            reads every plane once         B_state / slot + B_state / live
// Ranks inside a tile come from 4-bit live masks + a shuffle scan, survivors are staged in shared
// memory and written as contiguous, 16-byte vectorised runs; the order of survivors is preserved (results
// never depend on block scheduling).  16 Mi slots of (r, v) + generated ids: 173 us all live (0.83 of the HBM copy
// bandwidth in real traffic), 123 us with 17 % live (0.68; five resident CTAs per SM); the first version took 214 / 158 us.  A single-launch form
// with a chained scan (tickets + decoupled look-back) was built and measured: 332 / 279 us -- with 1024-slot tiles and
// four resident CTAs per SM the look-back latency sits on every tile's critical path -- and dropped.
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

// live slots of PCL_COUNT_TILES consecutive tiles per CTA: every thread has that many independent 16-byte loads in flight
#define PCL_COUNT_TILES 4
__global__ void __launch_bounds__(PCL_BLOCK) pcl_k_count_live(const float *x, uint64_t n, uint32_t ntiles, uint32_t *counts) {
    __shared__ uint32_t s_w[PCL_COUNT_TILES][PCL_WARPS];
    uint32_t c[PCL_COUNT_TILES];
#pragma unroll
    for (int k = 0; k < PCL_COUNT_TILES; ++k) {
        const uint64_t i = (((uint64_t)blockIdx.x * PCL_COUNT_TILES + k) * PCL_BLOCK + threadIdx.x) * 4;
        c[k] = (i < n) ? __popc(pcl_live_mask4(x, i, n)) : 0u;
    }
#pragma unroll
    for (int k = 0; k < PCL_COUNT_TILES; ++k) {
        c[k] = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0) s_w[k][threadIdx.x >> 5] = c[k];
    }
    __syncthreads();
    if (threadIdx.x < PCL_COUNT_TILES) {
        const uint32_t tile = blockIdx.x * PCL_COUNT_TILES + threadIdx.x;
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < PCL_WARPS; ++w) t += s_w[threadIdx.x][w];
        if (tile < ntiles) counts[tile] = t;
    }
}

// one block: counts[0..m) -> exclusive offsets in place, total -> *total_out.  Thread t owns the `per` consecutive
// counts from t * per (loaded up front when per <= 16: m <= 16384 tiles = 16 Mi slots), warp-shuffle scan of the
// thread sums, one pass through shared memory for the 32 warp totals.
__global__ void __launch_bounds__(1024) pcl_k_scan_tiles(uint32_t *counts, uint32_t m, uint64_t *total_out) {
    __shared__ uint64_t s_warp[32];
    const uint32_t per = (m + 1023u) / 1024u;
    const uint32_t b = threadIdx.x * per;
    const uint32_t e = min(b + per, m);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t held[16];
    const bool small = per <= 16u;
    uint64_t s = 0;
    if (small && per == 16u && b + 16u <= m) {  // 16 Mi slots: four 16-byte loads per thread
        const uint4 *c4 = reinterpret_cast<const uint4 *>(counts + b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = c4[k];
            held[4 * k] = v.x, held[4 * k + 1] = v.y, held[4 * k + 2] = v.z, held[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s += held[k];
    } else if (small) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            held[k] = (b + k < e) ? counts[b + k] : 0u;
            s += held[k];
        }
    } else {
        for (uint32_t i = b; i < e; ++i) s += counts[i];
    }
    uint64_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31u) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint64_t wt = s_warp[lane];
        uint64_t winc = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += v;
        }
        s_warp[lane] = winc - wt;  // exclusive offset of warp `lane`
        if (lane == 31u) *total_out = winc;
    }
    __syncthreads();
    uint64_t run = s_warp[wid] + (inc - s);
    if (small && per == 16u && b + 16u <= m) {
        uint32_t outv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            outv[k] = (uint32_t)run;
            run += held[k];
        }
        uint4 *c4 = reinterpret_cast<uint4 *>(counts + b);
#pragma unroll
        for (int k = 0; k < 4; ++k) c4[k] = make_uint4(outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]);
    } else if (small) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (b + k < e) counts[b + k] = (uint32_t)run;
            run += held[k];
        }
    } else {
        for (uint32_t i = b; i < e; ++i) {
            const uint32_t c = counts[i];
            counts[i] = (uint32_t)run;
            run += c;
        }
    }
}

// One tile (1024 slots) per CTA.  Every plane that moves is read with ONE 128-bit load per thread, all of them issued
// before anything else (7-15 independent loads in flight per thread; x doubles as the live mask, so it is read once
// here, not twice).  Survivors are then staged in shared memory at their rank inside the tile PLUS skew = base mod 32,
// four planes at a time, so that stage index q and output slot (base - skew) + q have the same alignment: the copy-out
// moves 16 bytes per thread and plane (LDS.128 -> STG.128) in whole 128-byte lines, as in the retire-and-compact
// photon kernel; only the threads on the two edges of the tile's output range store scalars.
// EXT: the group carries dx dy dz / ax ay az planes too (kept out of the common instantiation's registers).
#define PCL_CB 4  // planes per staging batch

__device__ __forceinline__ float4 pcl_ld4_guarded(const float *p, uint64_t i, uint64_t n, float fill) {
    if (i + 3 < n) return pcl_ld4(p + i);
    float4 v = make_float4(fill, fill, fill, fill);
    for (int l = 0; l < 4; ++l)
        if (i + l < n) pcl_f4(v, l) = p[i + l];
    return v;
}

template <bool EXT>
__global__ void __launch_bounds__(PCL_BLOCK, EXT ? 3 : 5)
pcl_k_compact(pcl_soa s, pcl_soa d, const uint32_t *offsets, int vec_ok) {
    constexpr int NP = EXT ? 15 : 9;
    __shared__ uint32_t s_w[PCL_WARPS];
    __shared__ __align__(16) float s_stage[PCL_CB][PCL_TILE + 32];
    const uint64_t i = ((uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x) * 4;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t base = __ldg(offsets + blockIdx.x);  // asked for first: its latency hides behind the plane loads
    // common planes first: a photon group (r, v, id [, e] [, nscat]) fills two batches
    const float *src[15] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, (const float *)s.id, s.e, (const float *)s.nscat,
                            s.dx, s.dy, s.dz, s.ax, s.ay, s.az};
    float *dst[15] = {d.x, d.y, d.z, d.vx, d.vy, d.vz, (float *)d.id, d.e, (float *)d.nscat, d.dx, d.dy, d.dz, d.ax, d.ay, d.az};
    const float qnan = __int_as_float(0x7fc00000);
    float4 v[NP];
    bool moves[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const bool is_id = q == 6;
        moves[q] = dst[q] != nullptr && (src[q] != nullptr || is_id);  // a plane moves when both sides have it; ids always do
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!moves[q]) continue;
        if (src[q]) {
            v[q] = pcl_ld4_guarded(src[q], i, s.n, q == 0 ? qnan : 0.f);  // slots past the end count as retired
        } else {  // no id plane yet: the id of a slot is its index
            v[q] = make_float4(__uint_as_float((uint32_t)i), __uint_as_float((uint32_t)i + 1u), __uint_as_float((uint32_t)i + 2u),
                               __uint_as_float((uint32_t)i + 3u));
        }
    }
    const uint32_t m = (v[0].x == v[0].x ? 1u : 0u) | (v[0].y == v[0].y ? 2u : 0u) | (v[0].z == v[0].z ? 4u : 0u) |
                       (v[0].w == v[0].w ? 8u : 0u);
    const uint32_t c = __popc(m);
    // exclusive rank of this thread's first survivor inside the tile
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
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
    const uint32_t skew = (uint32_t)(base & 31ull);
    const uint32_t rank = skew + wbase + (inc - c);
    const uint32_t end = skew + tot;
    const uint64_t obase = base - skew;  // a multiple of 32 slots
#pragma unroll
    for (int b0 = 0; b0 < NP; b0 += PCL_CB) {
        bool any = false;
#pragma unroll
        for (int p = 0; p < PCL_CB; ++p)
            if (b0 + p < NP) any = any || moves[b0 + p];
        if (!any) continue;  // uniform: the plane pointers are kernel arguments
        uint32_t rk = rank;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (!(m & (1u << l))) continue;
#pragma unroll
            for (int p = 0; p < PCL_CB; ++p)
                if (b0 + p < NP && moves[b0 + p]) s_stage[p][rk] = pcl_f4(v[b0 + p], l);
            ++rk;
        }
        __syncthreads();
        for (uint32_t q4 = threadIdx.x * 4; q4 < end; q4 += PCL_BLOCK * 4) {
            const uint64_t o = obase + q4;
            if (vec_ok && q4 >= skew && q4 + 3 < end) {
#pragma unroll
                for (int p = 0; p < PCL_CB; ++p)
                    if (b0 + p < NP && moves[b0 + p]) pcl_st4(dst[b0 + p] + o, *reinterpret_cast<const float4 *>(&s_stage[p][q4]));
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t q = q4 + l;
                    if (q < skew || q >= end) continue;
#pragma unroll
                    for (int p = 0; p < PCL_CB; ++p)
                        if (b0 + p < NP && moves[b0 + p]) dst[b0 + p][o + l] = s_stage[p][q];
                }
            }
        }
        __syncthreads();  // the stage is rewritten by the next batch
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
    pcl_k_count_live<<<(unsigned)((ntiles + PCL_COUNT_TILES - 1) / PCL_COUNT_TILES), PCL_BLOCK, 0, st>>>(src->x, src->n, (uint32_t)ntiles,
                                                                                                           ctx->scan_buf);
    PCL_LAUNCHED(ctx);
    pcl_k_scan_tiles<<<1, 1024, 0, st>>>(ctx->scan_buf, (uint32_t)ntiles, n_live_dev);
    PCL_LAUNCHED(ctx);
    const void *outp[15] = {dst->x, dst->y, dst->z, dst->vx, dst->vy, dst->vz, dst->id, dst->e, dst->nscat,
                            dst->dx, dst->dy, dst->dz, dst->ax, dst->ay, dst->az};
    int vec_ok = 1;  // 16-byte stores need 16-byte aligned output planes (absent planes are null: aligned)
    for (int q = 0; q < 15; ++q) vec_ok = vec_ok && pcl_aligned16(outp[q]);
    const bool ext = (dst->dx && src->dx) || (dst->dy && src->dy) || (dst->dz && src->dz) || (dst->ax && src->ax) ||
                     (dst->ay && src->ay) || (dst->az && src->az);
    if (ext)
        pcl_k_compact<true><<<(unsigned)ntiles, PCL_BLOCK, 0, st>>>(*src, *dst, ctx->scan_buf, vec_ok);
    else
        pcl_k_compact<false><<<(unsigned)ntiles, PCL_BLOCK, 0, st>>>(*src, *dst, ctx->scan_buf, vec_ok);
    PCL_LAUNCHED(ctx);
    return 0;
}
