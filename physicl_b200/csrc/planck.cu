// Photon emission: device form of planck_phot_distribution (reference physicl/light.py:73-104).
//
// The reference bins [E_min, E_max] on linspace(E_min, E_max, bins) (:82), integrates its density
// over each of the bins-1 intervals (:85-86), normalises (:88-89), accumulates a CDF (:91-93) and,
// per photon, draws rand ~ U[0,1) (:101) and scans for the first x >= 1 with
// cdf[x-1] <= rand <= cdf[x] (:102-104), returning the GRID energy E[x].  When rand < cdf[0] the
// loop finds nothing and the function returns None.
//
// Here the host builds the float64 CDF once (closed form of the bin integrals) and the device looks it up.
//
// Uniforms: photon `gid` takes word gid & 3 of the Philox4x32-10 block with counter (gid >> 2, 0, stream 1), so a
// thread that owns four consecutive photons draws ONE block for them (the generator was 3/5 of the instructions when
// every photon drew its own block and dropped three words).  u = (word >> 8) / 2^24 = m / 2^24.
//
// Look-up: the answer is lo(m) = #{idx : cdf[idx] < m / 2^24} (first idx with cdf[idx] >= u).  With
// S[idx] = floor(cdf[idx] 2^24) + 1 (exact in binary64) this is #{idx : S[idx] <= m}: an INTEGER problem, so the
// float64 table is not needed per photon.  A guide table over the top 16 bits of m, guide[g] = lo(256 g), brackets the
// answer, and inside bucket g only the low byte of S - 1 matters:  lo(m) = guide[g] + #{idx in [guide[g], guide[g+1]) :
// s8[idx] < (m & 255)},  s8[idx] = floor(cdf[idx] 2^24) & 255.  For tables of up to 65535 entries (the reference's
// examples use 200 ... 50 000 bins) guide (uint16, 128 KB) and s8 (1 B per entry) fit the SM's shared memory: a small
// kernel builds their image once, one persistent 1024-thread CTA per SM copies it (16-byte loads) and then serves every
// photon from shared memory: two guide reads and two unconditional byte probes per photon; only buckets with more than
// two thresholds (the tails of the law: 4 % of the buckets at 50 000 bins) take a byte-wise binary search.  Larger
// tables (and small draws, where 148 table copies would cost more than the draw) use the guide in global memory and
// probe the float64 table itself.  Both forms give the same bin for the same m: the bin index is an integer result,
// bit-exact against the oracle's linear scan of the reference's rule.
// Traffic: 4 B written per photon (+4 B with bin_out).  Measured (64 Mi photons, 50 000 bins): 218 us = 308 G photons/s
// (one Philox block per photon and float64 probes through L1/L2: 450 us); shared-memory wavefronts 63 %, issue 67 %.
// Also measured: serving the crowded-bucket photons warp-wide (owner's range broadcast, one entry per lane, ballot
// count) instead of per-lane searches: 251 us -- the shuffles and votes cost more than the divergent loops.
#include "pcl_common.cuh"

#define PCL_GUIDE_BITS 16
#define PCL_GUIDE_N (1u << PCL_GUIDE_BITS)
#define PCL_PLANCK_THREADS 1024
#define PCL_PLANCK_NQ 2  // Philox blocks (four photons each) per thread and loop iteration: two independent chains in flight
#define PCL_PLANCK_S8_OFF ((2u * (PCL_GUIDE_N + 1u) + 15u) & ~15u)  // byte offset of s8 behind the uint16 guide
#define PCL_PLANCK_SMEM_MAX_NCDF 65535u
#define PCL_PLANCK_SMEM_MIN_N (1u << 18)  // below this the 148 table copies cost more than they save
// context scratch (uint32 units): guide[2^16 + 1], then (16-byte aligned) the shared-memory image of the tables
#define PCL_PLANCK_IMG_OFF_U32 (PCL_GUIDE_N + 4u)
__host__ __device__ static inline uint32_t pcl_planck_img_bytes(uint32_t ncdf) {
    return PCL_PLANCK_S8_OFF + ((ncdf + 2u + 15u) & ~15u);
}

// lower bound: first idx in [lo, hi) with cdf[idx] >= u, else hi
__device__ __forceinline__ uint32_t pcl_cdf_lower_bound(const double *__restrict__ cdf, uint32_t lo, uint32_t hi, double u) {
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) >= u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Tables of one draw: guide[g] = first idx with cdf[idx] >= g / 2^16 (uint32, global-memory form) and, when the table
// fits (img != nullptr), the image the sampling CTAs copy into shared memory: the same guide as uint16, then
// s8[idx] = floor(cdf[idx] 2^24) & 255 (255 for entries >= 1, which no m < 2^24 passes, and for the pad bytes that let
// every photon read s8[lo] and s8[lo + 1] unconditionally).
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_planck_guide(const double *__restrict__ cdf, uint32_t ncdf, uint32_t *guide, unsigned char *img) {
    const uint32_t g = blockIdx.x * PCL_BLOCK + threadIdx.x;
    if (g <= PCL_GUIDE_N) {
        // g / 2^16 is exact in binary64; guide[2^16] = ncdf closes the last bucket
        const uint32_t v = g == PCL_GUIDE_N ? ncdf : pcl_cdf_lower_bound(cdf, 0u, ncdf, (double)g * (1.0 / PCL_GUIDE_N));
        guide[g] = v;
        if (img) reinterpret_cast<uint16_t *>(img)[g] = (uint16_t)v;
    } else if (img && g < PCL_PLANCK_S8_OFF / 2u) {
        reinterpret_cast<uint16_t *>(img)[g] = 0;  // padding between the two tables
    }
    if (img && g < pcl_planck_img_bytes(ncdf) - PCL_PLANCK_S8_OFF) {
        const double c = g < ncdf ? __ldg(cdf + g) * 16777216.0 : 16777216.0;  // exact scaling
        img[PCL_PLANCK_S8_OFF + g] = c >= 16777216.0 ? (uint8_t)255 : (uint8_t)((uint32_t)__double2ll_rd(c) & 255u);
    }
}

// lo(m) -> bin of the reference's rule (light.py:102-104: x >= 1 with cdf[x-1] <= u <= cdf[x]), without branches.
// m0: the one m with m / 2^24 == cdf[0] exactly (0xffffffff if there is none, or if the table has a single entry):
// for lo = 0 the rule needs cdf[x-1] <= rand with x >= 1, i.e. equality with cdf[0].
__device__ __forceinline__ int32_t pcl_planck_bin(uint32_t lo, uint32_t ncdf, uint32_t m, uint32_t m0) {
    int32_t bin = lo >= ncdf ? -1 : (int32_t)lo;  // u above the last cumulative value (rounding of the table's tail)
    return lo == 0u ? (m == m0 ? 1 : -1) : bin;
}

__device__ __forceinline__ uint32_t pcl_planck_m0(const double *__restrict__ cdf, uint32_t ncdf) {
    const double c = __ldg(cdf) * 16777216.0;  // exact scaling
    return (ncdf > 1u && c >= 0.0 && c < 16777216.0 && c == (double)__double2ll_rd(c)) ? (uint32_t)__double2ll_rd(c) : 0xffffffffu;
}

// One thread = PCL_PLANCK_NQ consecutive Philox blocks per iteration, four photons each (global ids 4 q .. 4 q + 3).
// SMEM: tables in shared memory (img: the image built by pcl_k_planck_guide).
template <bool SMEM>
__global__ void __launch_bounds__(SMEM ? PCL_PLANCK_THREADS : PCL_BLOCK)
pcl_k_planck(uint64_t n, uint64_t id_base, uint32_t seed_lo, uint32_t seed_hi, const double *__restrict__ cdf,
             uint32_t ncdf, const uint32_t *__restrict__ guide, const unsigned char *__restrict__ img, float e_lo,
             float e_step, float *e_out, int32_t *bin_out) {
    constexpr int NQ = PCL_PLANCK_NQ, NP = 4 * NQ;
    extern __shared__ __align__(16) unsigned char pcl_planck_sm[];
    const uint16_t *s_guide = reinterpret_cast<const uint16_t *>(pcl_planck_sm);
    const uint8_t *s_s8 = pcl_planck_sm + PCL_PLANCK_S8_OFF;
    if (SMEM) {
        const uint32_t nvec = pcl_planck_img_bytes(ncdf) / 16u;
        const uint4 *src = reinterpret_cast<const uint4 *>(img);
        uint4 *dst = reinterpret_cast<uint4 *>(pcl_planck_sm);
#pragma unroll 4
        for (uint32_t k = threadIdx.x; k < nvec; k += PCL_PLANCK_THREADS) dst[k] = __ldg(src + k);
        __syncthreads();
    }
    const uint32_t m0 = pcl_planck_m0(cdf, ncdf);
    const uint64_t q0 = id_base >> 2;
    const uint64_t nq = ((id_base + n - 1) >> 2) - q0 + 1;
    const bool vec_ok = (id_base & 3u) == 0 && ((uintptr_t)e_out & 15u) == 0 && (!bin_out || ((uintptr_t)bin_out & 15u) == 0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * NQ;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * NQ; w < nq; w += stride) {
        uint32_t word[NP], lo[NP], hi[NP], m[NP];
#pragma unroll
        for (int u = 0; u < NQ; ++u) {  // a block past the end of the shard is drawn and dropped
            const uint64_t q = q0 + w + u;
            const uint4 r = pcl_philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), 0u, 1u), make_uint2(seed_lo, seed_hi));
            word[4 * u] = r.x, word[4 * u + 1] = r.y, word[4 * u + 2] = r.z, word[4 * u + 3] = r.w;
        }
#pragma unroll
        for (int l = 0; l < NP; ++l) {  // independent look-ups: the loads of all of them are in flight together
            m[l] = word[l] >> 8;
            const uint32_t g = word[l] >> 16;
            lo[l] = SMEM ? (uint32_t)s_guide[g] : __ldg(guide + g);
            hi[l] = SMEM ? (uint32_t)s_guide[g + 1] : __ldg(guide + g + 1);
        }
        if (SMEM) {
            uint32_t crowded = 0u;
#pragma unroll
            for (int l = 0; l < NP; ++l) {  // the common cases (0, 1 or 2 thresholds in the bucket) without a branch
                const uint32_t cnt = hi[l] - lo[l], ml = m[l] & 255u;
                const uint32_t b0 = s_s8[lo[l]], b1 = s_s8[lo[l] + 1u];  // unconditional (pad bytes follow the table)
                const uint32_t c0 = (cnt > 0u) & (b0 < ml);
                const uint32_t c1 = (cnt > 1u) & (b1 < ml);  // s8 is sorted inside a bucket: c1 implies c0
                crowded |= (cnt > 2u) & c1;
                hi[l] = lo[l] + c0 + c1;
            }
            if (crowded) {  // a bucket in a tail of the law: first k in [lo + 2, hi) with s8[k] >= ml
#pragma unroll
                for (int l = 0; l < NP; ++l) {
                    const uint32_t ml = m[l] & 255u;
                    uint32_t a = lo[l] + 2u, b = (uint32_t)s_guide[(word[l] >> 16) + 1u];
                    if (hi[l] == a) {  // both probes passed (b <= a: the bucket had exactly two entries, nothing to search)
                        while (a < b) {
                            const uint32_t mid = (a + b) >> 1;
                            if ((uint32_t)s_s8[mid] >= ml) b = mid; else a = mid + 1u;
                        }
                        hi[l] = a;
                    }
                }
            }
#pragma unroll
            for (int l = 0; l < NP; ++l) lo[l] = hi[l];
        } else {
#pragma unroll
            for (int l = 0; l < NP; ++l) lo[l] = pcl_cdf_lower_bound(cdf, lo[l], hi[l], (double)m[l] * 0x1p-24);
        }
        int32_t bin[NP];
        float e[NP];
#pragma unroll
        for (int l = 0; l < NP; ++l) {
            bin[l] = pcl_planck_bin(lo[l], ncdf, m[l], m0);
            e[l] = (bin[l] < 0) ? __int_as_float(0x7fc00000) : fmaf((float)bin[l], e_step, e_lo);
        }
        const int64_t i0 = (int64_t)(4u * (q0 + w) - id_base);  // index of the first block's first photon (may lie before 0)
        if (vec_ok && i0 + NP - 1 < (int64_t)n) {
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                *reinterpret_cast<float4 *>(e_out + i0 + 4 * u) = make_float4(e[4 * u], e[4 * u + 1], e[4 * u + 2], e[4 * u + 3]);
                if (bin_out)
                    *reinterpret_cast<int4 *>(bin_out + i0 + 4 * u) = make_int4(bin[4 * u], bin[4 * u + 1], bin[4 * u + 2], bin[4 * u + 3]);
            }
        } else {
#pragma unroll
            for (int l = 0; l < NP; ++l) {
                const int64_t i = i0 + l;
                if (i >= 0 && i < (int64_t)n) {
                    e_out[i] = e[l];
                    if (bin_out) bin_out[i] = bin[l];
                }
            }
        }
    }
}

extern "C" int pcl_planck_sample(pcl_ctx *ctx, uintptr_t stream, uint64_t n, uint64_t id_base, uint64_t seed,
                                 const double *cdf, uint32_t ncdf, float e_lo, float e_step, float *e_out,
                                 int32_t *bin_out) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, cdf != nullptr && ncdf >= 1 && e_out != nullptr, "cdf table and e_out are required");
    if (n == 0) return 0;
    PCL_REQUIRE(ctx, id_base + n >= id_base, "global photon ids must not wrap 2^64");
    const bool smem_form = ncdf <= PCL_PLANCK_SMEM_MAX_NCDF && n >= PCL_PLANCK_SMEM_MIN_N;
    // context scratch: the guide table and, for the shared-memory form, the image of both tables
    const size_t need = (size_t)PCL_PLANCK_IMG_OFF_U32 + (smem_form ? pcl_planck_img_bytes(ncdf) / 4u : 0u);
    if (ctx->scan_cap < need) {
        if (ctx->scan_buf) PCL_CUDA(ctx, cudaFree(ctx->scan_buf));
        ctx->scan_buf = nullptr;
        ctx->scan_cap = 0;
        PCL_CUDA(ctx, cudaMalloc(&ctx->scan_buf, need * sizeof(uint32_t)));
        ctx->scan_cap = need;
    }
    uint32_t *guide = ctx->scan_buf;
    unsigned char *img = smem_form ? reinterpret_cast<unsigned char *>(ctx->scan_buf + PCL_PLANCK_IMG_OFF_U32) : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t prep = PCL_GUIDE_N + 1u;  // threads of the table kernel: one per guide entry, pad word and s8 byte
    if (smem_form && PCL_PLANCK_S8_OFF / 2u > prep) prep = PCL_PLANCK_S8_OFF / 2u;
    if (smem_form && pcl_planck_img_bytes(ncdf) - PCL_PLANCK_S8_OFF > prep) prep = pcl_planck_img_bytes(ncdf) - PCL_PLANCK_S8_OFF;
    pcl_k_planck_guide<<<(prep + PCL_BLOCK - 1) / PCL_BLOCK, PCL_BLOCK, 0, st>>>(cdf, ncdf, guide, img);
    PCL_LAUNCHED(ctx);
    const uint64_t nq = ((id_base + n - 1) >> 2) - (id_base >> 2) + 1;  // Philox blocks
    const uint64_t nw = (nq + PCL_PLANCK_NQ - 1) / PCL_PLANCK_NQ;       // threads' work items
    const uint32_t sl = (uint32_t)seed, sh = (uint32_t)(seed >> 32);
    if (smem_form) {
        PCL_CUDA(ctx, cudaFuncSetAttribute(pcl_k_planck<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)pcl_planck_img_bytes(PCL_PLANCK_SMEM_MAX_NCDF)));
        unsigned grid = pcl_stream_grid(ctx, nw, PCL_PLANCK_THREADS, 1);
        pcl_k_planck<true><<<grid, PCL_PLANCK_THREADS, pcl_planck_img_bytes(ncdf), st>>>(n, id_base, sl, sh, cdf, ncdf, guide, img,
                                                                                         e_lo, e_step, e_out, bin_out);
    } else {
        unsigned grid = pcl_stream_grid(ctx, nw, PCL_BLOCK, 8);
        pcl_k_planck<false><<<grid, PCL_BLOCK, 0, st>>>(n, id_base, sl, sh, cdf, ncdf, guide, nullptr, e_lo, e_step, e_out,
                                                         bin_out);
    }
    PCL_LAUNCHED(ctx);
    return 0;
}
