// Photon emission: device form of planck_phot_distribution (reference physicl/light.py:73-104).
//
// The reference bins [E_min, E_max] on linspace(E_min, E_max, bins) (:82), integrates its density
// over each of the bins-1 intervals (:85-86), normalises (:88-89), accumulates a CDF (:91-93) and,
// per photon, draws rand ~ U[0,1) (:101) and scans for the first x >= 1 with
// cdf[x-1] <= rand <= cdf[x] (:102-104), returning the GRID energy E[x].  When rand < cdf[0] the
// loop finds nothing and the function returns None.
//
// Here the host builds the float64 CDF once (closed form of the bin integrals), and one thread per
// photon draws its uniform from Philox (stream 1) and looks the table up.  The uniform is a 24-bit
// fraction m / 2^24, so a guide table over its top 16 bits (guide[g] = first idx with
// cdf[idx] >= g / 2^16, built on the device per call) brackets the answer: the search runs over
// [guide[g], guide[g+1]], usually one or two entries, instead of 16 dependent steps over the whole
// table.  The bin index is an integer result: bit-exact against the oracle for the same table.
// Traffic: 4 B written per photon (+4 B with bin_out); the tables (<= 400 KB + 256 KB) live in L2.
#include "pcl_common.cuh"

#define PCL_GUIDE_BITS 16
#define PCL_GUIDE_N (1u << PCL_GUIDE_BITS)

// lower bound: first idx in [lo, hi) with cdf[idx] >= u, else hi
__device__ __forceinline__ uint32_t pcl_cdf_lower_bound(const double *__restrict__ cdf, uint32_t lo, uint32_t hi, double u) {
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) >= u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_planck_guide(const double *__restrict__ cdf, uint32_t ncdf, uint32_t *guide) {
    const uint32_t g = blockIdx.x * PCL_BLOCK + threadIdx.x;
    if (g > PCL_GUIDE_N) return;
    // g / 2^16 is exact in binary64; guide[2^16] = ncdf closes the last bucket
    guide[g] = g == PCL_GUIDE_N ? ncdf : pcl_cdf_lower_bound(cdf, 0u, ncdf, (double)g * (1.0 / PCL_GUIDE_N));
}

__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_planck(uint64_t n, uint64_t id_base, uint32_t seed_lo, uint32_t seed_hi, const double *__restrict__ cdf,
             uint32_t ncdf, const uint32_t *__restrict__ guide, float e_lo, float e_step, float *e_out, int32_t *bin_out) {
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n; i += stride) {
        const uint64_t gid = id_base + i;
        uint4 r = pcl_philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), 0u, 1u),
                                    make_uint2(seed_lo, seed_hi));
        const double u = (double)pcl_u01(r.x);
        // u = (r.x >> 8) / 2^24 lies in bucket g = r.x >> 16: g/2^16 <= u < (g+1)/2^16, so the first idx with
        // cdf[idx] >= u lies in [guide[g], guide[g+1]]
        const uint32_t g = r.x >> (32 - PCL_GUIDE_BITS);
        const uint32_t lo = pcl_cdf_lower_bound(cdf, __ldg(guide + g), __ldg(guide + g + 1), u);
        int32_t bin;
        if (lo >= ncdf) {
            bin = -1;  // u above the last cumulative value (rounding of the table's tail)
        } else if (lo == 0) {
            bin = (ncdf > 1 && u == __ldg(cdf)) ? 1 : -1;  // light.py:102 needs cdf[x-1] <= rand with x >= 1
        } else {
            bin = (int32_t)lo;
        }
        float e = (bin < 0) ? __int_as_float(0x7fc00000) : fmaf((float)bin, e_step, e_lo);
        e_out[i] = e;
        if (bin_out) bin_out[i] = bin;
    }
}

extern "C" int pcl_planck_sample(pcl_ctx *ctx, uintptr_t stream, uint64_t n, uint64_t id_base, uint64_t seed,
                                 const double *cdf, uint32_t ncdf, float e_lo, float e_step, float *e_out,
                                 int32_t *bin_out) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, cdf != nullptr && ncdf >= 1 && e_out != nullptr, "cdf table and e_out are required");
    if (n == 0) return 0;
    const size_t need = (size_t)PCL_GUIDE_N + 1;  // the guide table lives in the context's scratch
    if (ctx->scan_cap < need) {
        if (ctx->scan_buf) PCL_CUDA(ctx, cudaFree(ctx->scan_buf));
        ctx->scan_buf = nullptr;
        ctx->scan_cap = 0;
        PCL_CUDA(ctx, cudaMalloc(&ctx->scan_buf, need * sizeof(uint32_t)));
        ctx->scan_cap = need;
    }
    uint32_t *guide = ctx->scan_buf;
    cudaStream_t st = (cudaStream_t)stream;
    pcl_k_planck_guide<<<(PCL_GUIDE_N + PCL_BLOCK) / PCL_BLOCK, PCL_BLOCK, 0, st>>>(cdf, ncdf, guide);
    PCL_LAUNCHED(ctx);
    unsigned grid = pcl_stream_grid(ctx, n, PCL_BLOCK, 8);
    pcl_k_planck<<<grid, PCL_BLOCK, 0, st>>>(n, id_base, (uint32_t)seed, (uint32_t)(seed >> 32), cdf, ncdf, guide, e_lo,
                                             e_step, e_out, bin_out);
    PCL_LAUNCHED(ctx);
    return 0;
}
