// Photon steps: scatter (isotropic / wavelength-dependent / delete), the fused whole-timestep
// kernel, and the integer tallies of the reference's measure steps.
//
// Reference semantics restated (physicl/light.py):
//   :305  norm  = sqrt(d0^2 + d1^2 + d2^2)                (dr left by newton.py:15)
//   :306  pcoll = A * n * norm [ * ((h*c)/E)^-4 ]          (:300-301)
//   :307  if (pcoll >= rand)                               rand ~ U[0,1)      (:285)
//   :309-311  v = c*sin(rtheta)*cos(rphi), c*sin(rtheta)*sin(rphi), c*cos(rtheta)
//             rtheta = U*2*pi, rphi = U*pi                 (:285)
//   :146-158 / :239-249  delete flavour: result = pcoll >= rand, flagged photons are removed
//   :414-431  sign tally: N, #(v_x>0), #(v_y>0), #(v_z>0)
//   :385-399  plane tally: loc in the closed interval between r-dr and r
//
// FP32 re-expression: host folds k = A*n (and (E0/(h c))^4 for the wavelength law) in float64 and
// photons carry e = E/E0, so pcoll = (k*norm)*(e^2)^2 stays in binary32 range for the reference's
// own Rayleigh constants (A ~ 4e-56 m^6, E ~ 1e-19 J).
//
// Fused step traffic per live photon-step: read r,v 24 B + write r 12 B + write v 12 B for the
// 16-byte groups that contain a scattered photon (+4 B e, +4 B id, +8 B nscat when present).
#include <stdlib.h>

#include "pcl_common.cuh"

#include "pcl_photon_body.cuh"

// ---------------------------------------------------------------------------------------------
// Fused photon step, in place, 4 photons per thread.
// ---------------------------------------------------------------------------------------------
template <bool WAVE, bool DEL, bool INJ, bool PL>
__global__ void __launch_bounds__(PCL_BLOCK, PCL_PHOTON_MINB)
pcl_k_photon_step(pcl_soa p, StepK K, int64_t *row) {
    __shared__ __align__(16) unsigned char s_tab[DEL ? 16 : PCL_TRIG_BYTES];
    __shared__ unsigned int s_acc[C_N];
    if (threadIdx.x < C_N) s_acc[threadIdx.x] = 0u;
    if (!DEL) pcl_trig_to_shared(s_tab, K.trig);
    else __syncthreads();
    const uint64_t nvec = pcl_valid_slots(p) / 4;
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    const bool has_id = p.id != nullptr;
    // whole warps enter the loop body together (warp-level tally hand-over): iterate on the warp's first group
    const uint64_t lane = threadIdx.x & 31u;
    for (uint64_t g0 = (uint64_t)blockIdx.x * PCL_BLOCK + (threadIdx.x & ~31u); g0 < nvec; g0 += stride) {
        const uint64_t g = g0 + lane;
        const bool in = g < nvec;
        const uint64_t i = (in ? g : g0) * 4;  // lanes past the end redo the warp's first group without storing
        float4 x = pcl_ld4(p.x + i), y = pcl_ld4(p.y + i), z = pcl_ld4(p.z + i);
        float4 vx = pcl_ld4(p.vx + i), vy = pcl_ld4(p.vy + i), vz = pcl_ld4(p.vz + i);
        float4 e = make_float4(1.f, 1.f, 1.f, 1.f);
        if (WAVE) e = pcl_ld4(p.e + i);
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
        if (!in) x.x = x.y = x.z = x.w = __int_as_float(0x7fc00000);  // counts nothing
        pcl_step_group4_masked<WAVE, DEL, INJ, false, PL>(p, K, s_tab, i, x, y, z, vx, vy, vz, e, id, nsc, ut4, up4, ur4, s_acc, in);
    }
    pcl_step_tail<WAVE, DEL, INJ, false, PL>(p, K, s_tab, nvec * 4, s_acc);
    __syncthreads();
    pcl_row_to_global(s_acc, row);
}

// ---------------------------------------------------------------------------------------------
// Fused photon step, in place, with the loads staged through shared memory by the TMA engine.
// One elected thread issues 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) of the next
// 1024-slot tile of every plane while the CTA computes the current one from shared memory, so the
// memory system is kept busy by the copy engine and not by however many warps happen to be waiting.
// Two stages per CTA (2 x planes x 4 KB).  Stores go straight from registers (STG.128).
// ---------------------------------------------------------------------------------------------
#define PCL_WTILE_SLOTS 128                      /* one warp-tile: 4 photons per lane */
#define PCL_WPLANE_BYTES (PCL_WTILE_SLOTS * 4)   /* 512 B of one plane */

// Every WARP owns a two-stage ring (2 x planes x 512 B) and its own pair of mbarriers: lane 0 issues
// the bulk copies of the warp's next 128-slot tile, the warp waits on its own barrier, and no warp
// ever waits for another one (the first version used CTA-wide tiles and lost ~14 % of its issue
// slots at the per-tile __syncthreads, see profiles/).
template <bool WAVE, bool DEL, bool PL>
__global__ void __launch_bounds__(PCL_BLOCK, PCL_PHOTON_MINB)
pcl_k_photon_step_tma(pcl_soa p, StepK K, int64_t *row) {
    extern __shared__ __align__(128) unsigned char s_stage_raw[];
    __shared__ __align__(8) uint64_t s_bar[PCL_WARPS][2];
    __shared__ __align__(16) unsigned char s_tab[DEL ? 16 : PCL_TRIG_BYTES];
    __shared__ unsigned int s_acc[C_N];
    if (threadIdx.x < C_N) s_acc[threadIdx.x] = 0u;
    if (!DEL) pcl_trig_to_shared(s_tab, K.trig);
    else __syncthreads();
    const bool has_id = p.id != nullptr, has_ns = p.nscat != nullptr;
    // plane order inside a stage: x y z vx vy vz [e] [id] [nscat]
    const int q_e = 6, q_id = 6 + (WAVE ? 1 : 0), q_ns = q_id + (has_id ? 1 : 0);
    const int nplanes = q_ns + (has_ns ? 1 : 0);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t valid = pcl_valid_slots(p);
    const uint64_t ntiles = valid / PCL_WTILE_SLOTS;
    uint64_t *bar = s_bar[wid];
    unsigned char *ring = s_stage_raw + (size_t)wid * 2 * nplanes * PCL_WPLANE_BYTES;
    if (lane == 0) {
        pcl_mbar_init(&bar[0], 1);
        pcl_mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](uint64_t tile, int st) {  // lane 0 only
        unsigned char *base = ring + (size_t)st * nplanes * PCL_WPLANE_BYTES;
        const uint64_t off = tile * PCL_WTILE_SLOTS;
        pcl_mbar_expect_tx(&bar[st], (uint32_t)(nplanes * PCL_WPLANE_BYTES));
        pcl_bulk_g2s(base + 0 * PCL_WPLANE_BYTES, p.x + off, PCL_WPLANE_BYTES, &bar[st]);
        pcl_bulk_g2s(base + 1 * PCL_WPLANE_BYTES, p.y + off, PCL_WPLANE_BYTES, &bar[st]);
        pcl_bulk_g2s(base + 2 * PCL_WPLANE_BYTES, p.z + off, PCL_WPLANE_BYTES, &bar[st]);
        pcl_bulk_g2s(base + 3 * PCL_WPLANE_BYTES, p.vx + off, PCL_WPLANE_BYTES, &bar[st]);
        pcl_bulk_g2s(base + 4 * PCL_WPLANE_BYTES, p.vy + off, PCL_WPLANE_BYTES, &bar[st]);
        pcl_bulk_g2s(base + 5 * PCL_WPLANE_BYTES, p.vz + off, PCL_WPLANE_BYTES, &bar[st]);
        if (WAVE) pcl_bulk_g2s(base + q_e * PCL_WPLANE_BYTES, p.e + off, PCL_WPLANE_BYTES, &bar[st]);
        if (has_id) pcl_bulk_g2s(base + q_id * PCL_WPLANE_BYTES, p.id + off, PCL_WPLANE_BYTES, &bar[st]);
        if (has_ns) pcl_bulk_g2s(base + q_ns * PCL_WPLANE_BYTES, p.nscat + off, PCL_WPLANE_BYTES, &bar[st]);
    };
    const uint64_t nwarps = (uint64_t)gridDim.x * PCL_WARPS;
    uint64_t tile = (uint64_t)blockIdx.x * PCL_WARPS + wid;
    if (lane == 0 && tile < ntiles) issue(tile, 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t it = 0; tile < ntiles; tile += nwarps, ++it) {
        const int st = (int)(it & 1u);
        const uint64_t next = tile + nwarps;
        if (lane == 0 && next < ntiles) {
            // the other stage was read (generic proxy) by this warp in the previous iteration, which
            // ended with __syncwarp; order those reads before the async-proxy writes of the new copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(next, st ^ 1);
        }
        pcl_mbar_wait(&bar[st], (it >> 1) & 1u);
        const unsigned char *base = ring + (size_t)st * nplanes * PCL_WPLANE_BYTES;
        const float4 *f4 = reinterpret_cast<const float4 *>(base);
        float4 x = f4[0 * 32 + lane], y = f4[1 * 32 + lane], z = f4[2 * 32 + lane];
        float4 vx = f4[3 * 32 + lane], vy = f4[4 * 32 + lane], vz = f4[5 * 32 + lane];
        float4 e = make_float4(1.f, 1.f, 1.f, 1.f);
        if (WAVE) e = f4[q_e * 32 + lane];
        const uint64_t i = tile * PCL_WTILE_SLOTS + (uint64_t)lane * 4;
        uint4 id = make_uint4((uint32_t)i, (uint32_t)i + 1u, (uint32_t)i + 2u, (uint32_t)i + 3u);
        if (has_id) id = reinterpret_cast<const uint4 *>(base)[q_id * 32 + lane];
        uint4 nsc = make_uint4(0u, 0u, 0u, 0u);
        if (has_ns) nsc = reinterpret_cast<const uint4 *>(base)[q_ns * 32 + lane];
        __syncwarp();  // every lane has read its operands: the stage may be refilled
        pcl_step_group4_masked<WAVE, DEL, false, false, PL>(p, K, s_tab, i, x, y, z, vx, vy, vz, e, id, nsc, zero4, zero4, zero4, s_acc, true);
    }
    // slots past the last full warp-tile: register path, block 0 only (< 128 slots), whole warps at a time
    if (blockIdx.x == 0) {
        const uint64_t nvec = valid / 4;
        for (uint64_t g0 = ntiles * 32 + (threadIdx.x & ~31u); g0 < nvec; g0 += PCL_BLOCK) {
            const uint64_t g = g0 + lane;
            const bool in = g < nvec;
            const uint64_t i = (in ? g : g0) * 4;
            float4 x = pcl_ld4(p.x + i), y = pcl_ld4(p.y + i), z = pcl_ld4(p.z + i);
            float4 vx = pcl_ld4(p.vx + i), vy = pcl_ld4(p.vy + i), vz = pcl_ld4(p.vz + i);
            float4 e = make_float4(1.f, 1.f, 1.f, 1.f);
            if (WAVE) e = pcl_ld4(p.e + i);
            uint4 id = make_uint4((uint32_t)i, (uint32_t)i + 1u, (uint32_t)i + 2u, (uint32_t)i + 3u);
            if (has_id) id = pcl_ld4u(p.id + i);
            uint4 nsc = make_uint4(0u, 0u, 0u, 0u);
            if (has_ns) nsc = pcl_ld4u(p.nscat + i);
            if (!in) x.x = x.y = x.z = x.w = __int_as_float(0x7fc00000);
            pcl_step_group4_masked<WAVE, DEL, false, false, PL>(p, K, s_tab, i, x, y, z, vx, vy, vz, e, id, nsc, zero4, zero4, zero4, s_acc, in);
        }
        pcl_step_tail<WAVE, DEL, false, false, PL>(p, K, s_tab, nvec * 4, s_acc);
    }
    __syncthreads();
    pcl_row_to_global(s_acc, row);
}

// scalar form for views that are not 16-byte aligned
template <bool WAVE, bool DEL, bool INJ, bool PL>
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_photon_step_tail(pcl_soa p, StepK K, int64_t *row, int aligned) {
    __shared__ __align__(16) unsigned char s_tab[DEL ? 16 : PCL_TRIG_BYTES];
    __shared__ unsigned int s_acc[C_N];
    if (threadIdx.x < C_N) s_acc[threadIdx.x] = 0u;
    if (!DEL) pcl_trig_to_shared(s_tab, K.trig);
    else __syncthreads();
    const uint64_t end = pcl_valid_slots(p);
    const uint64_t begin = aligned ? (end / 4) * 4 : 0;
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = begin + (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < end; i += stride)
        pcl_step_scalar<WAVE, DEL, INJ, false, PL>(p, K, s_tab, i, s_acc);
    __syncthreads();
    pcl_row_to_global(s_acc, row);
}

// ---------------------------------------------------------------------------------------------
// Several timesteps per HBM round trip.  Photons do not interact, so a thread can keep its four
// photons in registers and advance them `nsteps` timesteps (kinematics -> scatter -> escape -> tallies,
// a fresh Philox block per photon and timestep, one tally row per timestep) before anything goes
// back to HBM: the traffic per photon-step drops by nsteps and the kernel becomes bound by the
// instruction issue rate instead of by DRAM.  Results are bit-identical to nsteps single-step launches.
//
// COMPACT: retirement folded in.  The survivors of the last timestep are written densely into `d`
// (the ping-pong partner), so the next launch touches live photons only and a warp never spends
// bandwidth or issue slots on photons retired in earlier launches.  Survivors of a 1024-slot tile
// keep their order (4-bit keep masks, shuffle scan, one atomic range reservation per tile); tiles land
// in reservation order, which is why ids travel with the photons (RNG counter, identity).
// Traffic per live photon and launch: read r,v,id 28 B + write r,v,id 28 B (+ e, nscat when present).
// !COMPACT: in place; r is always written, v and nscat only by groups in which a photon scattered.
// ---------------------------------------------------------------------------------------------
#ifndef PCL_MULTI_PREFETCH_DEFAULT
#define PCL_MULTI_PREFETCH_DEFAULT 0
#endif
#ifndef PCL_MULTI_MINB_COMPACT
#define PCL_MULTI_MINB_COMPACT 4
#endif
#ifndef PCL_MULTI_MINB_INPLACE
#define PCL_MULTI_MINB_INPLACE PCL_PHOTON_MINB
#endif
template <bool WAVE, bool DEL, bool INJ, bool PL, bool COMPACT, bool SFU = false>
__global__ void __launch_bounds__(PCL_BLOCK, COMPACT ? PCL_MULTI_MINB_COMPACT : PCL_MULTI_MINB_INPLACE)
pcl_k_photon_multi(pcl_soa s, pcl_soa d, StepK K, int64_t *rows, unsigned long long *n_out, uint32_t nsteps, int prefetch) {
    constexpr int NST = COMPACT ? 9 : 1;  // staged planes: x y z vx vy vz id nscat e
    // prefetch: every thread copies ITS 16-byte pieces of the CTA's next tile into shared memory with cp.async while the
    // current tile is being stepped, and picks them up (its own slots: no barrier) at the top of the next iteration
    extern __shared__ __align__(16) float4 s_pf[];  // [plane][PCL_BLOCK], planes x y z vx vy vz [id] [e] [nscat]
    __shared__ __align__(16) float s_stage[NST][COMPACT ? PCL_BLOCK * 4 + 32 : 4];
    __shared__ uint32_t s_warp[PCL_WARPS];
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_acc[PCL_FUSE_MAX][C_N];
    __shared__ __align__(16) unsigned char s_tab[DEL ? 16 : PCL_TRIG_BYTES];
    for (uint32_t q = threadIdx.x; q < PCL_FUSE_MAX * C_N; q += PCL_BLOCK) (&s_acc[0][0])[q] = 0u;
    if (!DEL) pcl_trig_to_shared(s_tab, K.trig);
    else __syncthreads();
    const uint64_t n = pcl_valid_slots(s);
    const uint64_t ntiles = (n + PCL_BLOCK * 4 - 1) / (PCL_BLOCK * 4);
    const bool has_id = s.id != nullptr, has_ns = s.nscat != nullptr;
    // the e plane travels with the survivors whenever both sides have one, also when the law does not read it
    // (a photon keeps its energy through delete scattering and the escape sphere: light.py:34)
    const bool carry_e = COMPACT && !WAVE && s.e != nullptr && d.e != nullptr;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base_lo = (uint32_t)s.id_base;
    const bool use_pf = prefetch != 0 && !INJ;
    const int pq_id = 6, pq_e = 6 + (has_id ? 1 : 0), pq_ns = pq_e + (WAVE ? 1 : 0);
    auto pf_issue = [&](uint64_t t) {
        const uint64_t ii = (t * PCL_BLOCK + threadIdx.x) * 4;
        float4 *slot = s_pf + threadIdx.x;
        pcl_cp_async16(slot + 0 * PCL_BLOCK, s.x + ii);
        pcl_cp_async16(slot + 1 * PCL_BLOCK, s.y + ii);
        pcl_cp_async16(slot + 2 * PCL_BLOCK, s.z + ii);
        pcl_cp_async16(slot + 3 * PCL_BLOCK, s.vx + ii);
        pcl_cp_async16(slot + 4 * PCL_BLOCK, s.vy + ii);
        pcl_cp_async16(slot + 5 * PCL_BLOCK, s.vz + ii);
        if (has_id) pcl_cp_async16(slot + pq_id * PCL_BLOCK, s.id + ii);
        if (WAVE) pcl_cp_async16(slot + pq_e * PCL_BLOCK, s.e + ii);
        if (has_ns) pcl_cp_async16(slot + pq_ns * PCL_BLOCK, s.nscat + ii);
        pcl_cp_async_commit();
    };
    auto whole = [&](uint64_t t) { return (t + 1) * (uint64_t)(PCL_BLOCK * 4) <= n; };  // every slot of tile t is valid
    bool pre = false;  // the current tile sits in s_pf
    if (use_pf && blockIdx.x < ntiles && whole(blockIdx.x)) {
        pf_issue(blockIdx.x);
        pre = true;
    }
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t i = (tile * PCL_BLOCK + threadIdx.x) * 4;
        float4 x, y, z, vx, vy, vz, e = make_float4(1.f, 1.f, 1.f, 1.f);
        uint4 id = make_uint4((uint32_t)i, (uint32_t)i + 1u, (uint32_t)i + 2u, (uint32_t)i + 3u);
        uint4 nsc = make_uint4(0u, 0u, 0u, 0u);
        float4 ut4, up4, ur4;
        const float qnan = __int_as_float(0x7fc00000);
        const bool full = i + 3 < n;
        const uint64_t next_tile = tile + gridDim.x;
        const bool pre_next = use_pf && next_tile < ntiles && whole(next_tile);
        if (pre) {
            pcl_cp_async_wait<0>();
            const float4 *slot = s_pf + threadIdx.x;
            x = slot[0 * PCL_BLOCK], y = slot[1 * PCL_BLOCK], z = slot[2 * PCL_BLOCK];
            vx = slot[3 * PCL_BLOCK], vy = slot[4 * PCL_BLOCK], vz = slot[5 * PCL_BLOCK];
            if (has_id) id = *reinterpret_cast<const uint4 *>(slot + pq_id * PCL_BLOCK);
            if (WAVE) e = slot[pq_e * PCL_BLOCK];
            if (has_ns) nsc = *reinterpret_cast<const uint4 *>(slot + pq_ns * PCL_BLOCK);
        } else if (full) {
            x = pcl_ld4(s.x + i), y = pcl_ld4(s.y + i), z = pcl_ld4(s.z + i);
            vx = pcl_ld4(s.vx + i), vy = pcl_ld4(s.vy + i), vz = pcl_ld4(s.vz + i);
            if (WAVE) e = pcl_ld4(s.e + i);
            if (has_id) id = pcl_ld4u(s.id + i);
            if (s.nscat) nsc = pcl_ld4u(s.nscat + i);
            if (INJ) {
                ut4 = pcl_ld4(K.u_theta + i);
                up4 = pcl_ld4(K.u_phi + i);
                ur4 = pcl_ld4(K.u_rand + i);
            }
        } else {  // last, partial group (or past the end): guarded scalar loads, missing slots are "retired"
            x = make_float4(qnan, qnan, qnan, qnan);
            y = z = vx = vy = vz = make_float4(0.f, 0.f, 0.f, 0.f);
            ut4 = up4 = ur4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                if (i + l < n) {
                    pcl_f4(x, l) = s.x[i + l];
                    pcl_f4(y, l) = s.y[i + l];
                    pcl_f4(z, l) = s.z[i + l];
                    pcl_f4(vx, l) = s.vx[i + l];
                    pcl_f4(vy, l) = s.vy[i + l];
                    pcl_f4(vz, l) = s.vz[i + l];
                    if (WAVE) pcl_f4(e, l) = s.e[i + l];
                    if (has_id) pcl_u4(id, l) = s.id[i + l];
                    if (s.nscat) pcl_u4(nsc, l) = s.nscat[i + l];
                    if (INJ) {
                        pcl_f4(ut4, l) = K.u_theta[i + l];
                        pcl_f4(up4, l) = K.u_phi[i + l];
                        pcl_f4(ur4, l) = K.u_rand[i + l];
                    }
                }
            }
        }
        bool any_scat = false, issued = false;
        for (uint32_t st = 0; st < nsteps; ++st) {
            // all four photons of every lane of this warp retired: nothing left to do for the warp
            const bool alive = (x.x == x.x) || (x.y == x.y) || (x.z == x.z) || (x.w == x.w);
            if (!__any_sync(0xffffffffu, alive)) break;
            pcl_draw3 dr[4];
#pragma unroll
            for (int l = 0; l < 4; ++l)  // four independent Philox chains: the compiler interleaves them
                dr[l] = INJ ? pcl_draw_floats(pcl_f4(ut4, l), pcl_f4(up4, l), pcl_f4(ur4, l))
                            : pcl_draw_at_raw(K, K.step + st, base_lo + pcl_u4(id, l));
            pcl_tally4 t = {0u, 0u, 0u, 0u};
            bool hit[4];
#pragma unroll
            for (int l = 0; l < 4; l += 2)  // photons (0,1) and (2,3) as pairs: packed FP32
                pcl_photon_two<WAVE, DEL, PL, !INJ, SFU>(K, s_tab, pcl_f4(x, l), pcl_f4(x, l + 1), pcl_f4(y, l), pcl_f4(y, l + 1), pcl_f4(z, l),
                                              pcl_f4(z, l + 1), pcl_f4(vx, l), pcl_f4(vx, l + 1), pcl_f4(vy, l), pcl_f4(vy, l + 1),
                                              pcl_f4(vz, l), pcl_f4(vz, l + 1), pcl_f4(e, l), pcl_f4(e, l + 1), dr[l], dr[l + 1], t,
                                              hit[l], hit[l + 1]);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const bool sc = !DEL && hit[l];
                any_scat = any_scat || sc;
                pcl_u4(nsc, l) += sc ? 1u : 0u;
            }
            pcl_tally_to_shared<PL>(t, s_acc[st]);
            if (st == 0 && pre_next) {  // the slots' old contents are in registers and have been used
                pf_issue(next_tile);
                issued = true;
            }
        }
        if (pre_next && !issued) pf_issue(next_tile);  // a warp whose photons were all retired left the loop at once
        pre = pre_next;
        if (!COMPACT) {
            if (full) {
                pcl_st4(s.x + i, x);
                pcl_st4(s.y + i, y);
                pcl_st4(s.z + i, z);
                if (any_scat) {
                    pcl_st4(s.vx + i, vx);
                    pcl_st4(s.vy + i, vy);
                    pcl_st4(s.vz + i, vz);
                    if (s.nscat) pcl_st4u(s.nscat + i, nsc);
                }
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (i + l < n) {
                        s.x[i + l] = pcl_f4(x, l);
                        s.y[i + l] = pcl_f4(y, l);
                        s.z[i + l] = pcl_f4(z, l);
                        s.vx[i + l] = pcl_f4(vx, l);
                        s.vy[i + l] = pcl_f4(vy, l);
                        s.vz[i + l] = pcl_f4(vz, l);
                        if (s.nscat) s.nscat[i + l] = pcl_u4(nsc, l);
                    }
                }
            }
            continue;
        }
        uint32_t keep = 0u;
#pragma unroll
        for (int l = 0; l < 4; ++l) keep |= (pcl_f4(x, l) == pcl_f4(x, l)) ? (1u << l) : 0u;
        if (carry_e) {  // loaded only now: the timestep loop above does not hold it in registers
            if (full) {
                e = pcl_ld4(s.e + i);
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l)
                    if (i + l < n) pcl_f4(e, l) = s.e[i + l];
            }
        }
        // tile-local rank of this thread's first survivor: shuffle scan inside the warp, warp totals
        // through shared memory
        const uint32_t c = __popc(keep);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < PCL_WARPS; ++w) {
            uint32_t t = s_warp[w];
            wbase += (w < (int)wid) ? t : 0u;
            tot += t;
        }
        if (threadIdx.x == 0) s_base = tot ? atomicAdd(n_out, (unsigned long long)tot) : 0ull;
        __syncthreads();
        // The tile's output range [base, base + tot) starts anywhere.  Survivors are staged at their rank PLUS
        // skew = base mod 32, so that stage index q and output slot (base - skew) + q have the same alignment:
        // the copy-out then moves 16 bytes per thread and plane (LDS.128 -> STG.128), every warp store covers
        // whole 128-byte lines of the output (full lines in HBM, full-size packets when `d` is host memory),
        // and only the few threads on the two edges of the range fall back to scalar stores.
        const unsigned long long base = s_base;
        const uint32_t skew = (uint32_t)(base & 31ull);
        uint32_t rk = skew + wbase + (inc - c);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (!(keep & (1u << l))) continue;
            s_stage[0][rk] = pcl_f4(x, l);
            s_stage[1][rk] = pcl_f4(y, l);
            s_stage[2][rk] = pcl_f4(z, l);
            s_stage[3][rk] = pcl_f4(vx, l);
            s_stage[4][rk] = pcl_f4(vy, l);
            s_stage[5][rk] = pcl_f4(vz, l);
            s_stage[6][rk] = __uint_as_float(pcl_u4(id, l));
            if (has_ns) s_stage[7][rk] = __uint_as_float(pcl_u4(nsc, l));
            if (WAVE || carry_e) s_stage[NST - 1][rk] = pcl_f4(e, l);
            ++rk;
        }
        __syncthreads();
        const uint32_t end = skew + tot;
        const unsigned long long obase = base - skew;  // a multiple of 32 slots
        for (uint32_t q4 = threadIdx.x * 4; q4 < end; q4 += PCL_BLOCK * 4) {
            const unsigned long long o = obase + q4;
            if (q4 >= skew && q4 + 3 < end) {
                pcl_st4(d.x + o, *reinterpret_cast<const float4 *>(&s_stage[0][q4]));
                pcl_st4(d.y + o, *reinterpret_cast<const float4 *>(&s_stage[1][q4]));
                pcl_st4(d.z + o, *reinterpret_cast<const float4 *>(&s_stage[2][q4]));
                pcl_st4(d.vx + o, *reinterpret_cast<const float4 *>(&s_stage[3][q4]));
                pcl_st4(d.vy + o, *reinterpret_cast<const float4 *>(&s_stage[4][q4]));
                pcl_st4(d.vz + o, *reinterpret_cast<const float4 *>(&s_stage[5][q4]));
                pcl_st4(reinterpret_cast<float *>(d.id) + o, *reinterpret_cast<const float4 *>(&s_stage[6][q4]));
                if (has_ns) pcl_st4(reinterpret_cast<float *>(d.nscat) + o, *reinterpret_cast<const float4 *>(&s_stage[7][q4]));
                if (WAVE || carry_e) pcl_st4(d.e + o, *reinterpret_cast<const float4 *>(&s_stage[NST - 1][q4]));
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t q = q4 + l;
                    if (q < skew || q >= end) continue;
                    d.x[o + l] = s_stage[0][q];
                    d.y[o + l] = s_stage[1][q];
                    d.z[o + l] = s_stage[2][q];
                    d.vx[o + l] = s_stage[3][q];
                    d.vy[o + l] = s_stage[4][q];
                    d.vz[o + l] = s_stage[5][q];
                    d.id[o + l] = __float_as_uint(s_stage[6][q]);
                    if (has_ns) d.nscat[o + l] = __float_as_uint(s_stage[7][q]);
                    if (WAVE || carry_e) d.e[o + l] = s_stage[NST - 1][q];
                }
            }
        }
        __syncthreads();  // the stage and s_warp are rewritten by the next tile
    }
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < nsteps * C_N; q += PCL_BLOCK) {
        const unsigned int v = (&s_acc[0][0])[q];
        // register column -> PCL_T_* column (identical numbering by construction); row st = rows + st*COLS
        if (v) atomicAdd((unsigned long long *)&rows[q], (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone scatter: what CLProgram.run launches (physicl/__init__.py:656) plus the host
// write-back loop (light.py:325-331).  Reads the dr planes written by the kinematics step.
// ---------------------------------------------------------------------------------------------
template <bool WAVE, bool DEL, bool INJ>
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_scatter(pcl_soa p, StepK K, int32_t *flags, int64_t *row, uint64_t n) {
    __shared__ __align__(16) unsigned char s_tab[DEL ? 16 : PCL_TRIG_BYTES];
    if (!DEL) pcl_trig_to_shared(s_tab, K.trig);
    uint32_t cnt[C_PLANE0];
#pragma unroll
    for (int q = 0; q < C_PLANE0; ++q) cnt[q] = 0u;
    const float qnan = __int_as_float(0x7fc00000);
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n; i += stride) {
        float xx = p.x[i];
        int32_t flag = 0;
        if (xx == xx) {
            cnt[C_LIVEIN] += 1u;
            const pcl_draw3 d = INJ ? pcl_draw_floats(K.u_theta[i], K.u_phi[i], K.u_rand[i])
                                    : pcl_draw_at(K, K.step, (uint32_t)p.id_base + (p.id ? p.id[i] : (uint32_t)i));
            float vx = 0.f, vy = 0.f, vz = 0.f;
            float e = WAVE ? p.e[i] : 1.f;
            const bool hit = pcl_scatter_one<WAVE, DEL>(true, p.dx[i], p.dy[i], p.dz[i], e, d, K, s_tab, vx, vy, vz);
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

// Stand-alone escape sphere.
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_escape(pcl_soa p, float r2_escape, int64_t *row, uint64_t n) {
    uint32_t cnt[C_PLANE0];
#pragma unroll
    for (int q = 0; q < C_PLANE0; ++q) cnt[q] = 0u;
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n; i += stride) {
        float xx = p.x[i];
        if (xx != xx) continue;
        cnt[C_LIVEIN] += 1u;
        float yy = p.y[i], zz = p.z[i];
        float r2 = xx * xx;
        r2 = fmaf(yy, yy, r2);
        r2 = fmaf(zz, zz, r2);
        if (r2 >= r2_escape) {
            cnt[C_ESC] += 1u;
            p.x[i] = __int_as_float(0x7fc00000);
        } else {
            cnt[C_ALIVE] += 1u;
        }
    }
    if (row) pcl_flush_tally(cnt, row, 0u);
}

// Stand-alone tallies (ScatterSignMeasureStep / ScatterMeasureStep).
__global__ void __launch_bounds__(PCL_BLOCK)
pcl_k_tally(pcl_soa p, StepK K, int64_t *row, uint64_t n) {
    uint32_t cnt[C_N];
#pragma unroll
    for (int q = 0; q < C_N; ++q) cnt[q] = 0u;
    const bool has_dr = p.dx != nullptr;
    const uint64_t stride = (uint64_t)gridDim.x * PCL_BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * PCL_BLOCK + threadIdx.x; i < n; i += stride) {
        float xx = p.x[i];
        if (xx != xx) continue;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (has_dr) {
            dx = p.dx[i];
            dy = p.dy[i];
            dz = p.dz[i];
        }
        pcl_tally_one(K, true, xx, p.y[i], p.z[i], dx, dy, dz, p.vx[i], p.vy[i], p.vz[i], cnt);
    }
    pcl_flush_tally(cnt, row, K.nplanes);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 64 -> 32 bit fold of (seed, high word of the id base, stream) for the Philox2x32 key: splitmix64 finaliser
static uint32_t pcl_fold_key(uint64_t seed, uint64_t id_hi, uint64_t stream) {
    uint64_t z = seed ^ (id_hi * 0x9E3779B97F4A7C15ull) ^ (stream << 56);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

int pcl_fill_stepk(pcl_ctx *ctx, StepK &K, float dt, const pcl_scatter_params *sp, const pcl_rng *rng,
                      float escape_r2, const pcl_planes *planes, uint64_t id_base) {
    memset(&K, 0, sizeof(K));
    K.dt = dt;
    K.trig = ctx->trig;
    if (sp) {
        K.k = sp->k;
        K.c = sp->c;
        // 1/k in float64, rounded once: k = 0 -> FLT_MAX (only rand = 0 scatters, as pcoll = 0 >= 0 does in the
        // reference), k < 0 or NaN -> NaN (the test never succeeds), k = inf -> 0 (always)
        const double kd = (double)sp->k;
        if (kd > 0.0) {
            const double inv = 1.0 / kd;
            K.kinv = inv > 3.4028234663852886e38 ? 3.4028234663852886e38f : (float)inv;
        } else if (kd == 0.0) {
            K.kinv = 3.4028234663852886e38f;
        } else {
            K.kinv = nanf("");
        }
        K.kinv24 = ldexpf(K.kinv, -24);
        K.sfu = (sp->mode & PCL_SCATTER_SFU) ? 1u : 0u;
    }
    K.r2_escape = escape_r2 > 0.f ? escape_r2 : nanf("");
    if (rng) {
        uint32_t key = pcl_fold_key(rng->seed, id_base >> 32, 0);
        for (int r = 0; r < 10; ++r) K.rk[r] = key + (uint32_t)r * PCL_PHILOX2_W;
        K.step = rng->step;
        K.u_theta = rng->u_theta;
        K.u_phi = rng->u_phi;
        K.u_rand = rng->u_rand;
        if (rng->u_rand) {
            bool del = sp && (sp->mode & PCL_SCATTER_DELETE);
            PCL_REQUIRE(ctx, del || (rng->u_theta && rng->u_phi),
                        "injected uniforms need u_theta and u_phi as well as u_rand");
            if (del) {  // the delete kernel only consumes rand (light.py:235)
                if (!K.u_theta) K.u_theta = rng->u_rand;
                if (!K.u_phi) K.u_phi = rng->u_rand;
            }
        }
    }
    if (planes) {
        PCL_REQUIRE(ctx, planes->count <= PCL_MAX_PLANES, "too many planes");
        K.nplanes = planes->count;
        for (uint32_t q = 0; q < planes->count; ++q) {
            PCL_REQUIRE(ctx, planes->axis[q] < 3, "plane axis must be 0, 1 or 2");
            K.axis[q] = planes->axis[q];
            K.loc[q] = planes->loc[q];
        }
    }
    return 0;
}

// timesteps per launch for multi-step runs: PCL_PHOTON_FUSE=1 restores one launch per timestep
static uint32_t photon_fuse_max() {
    static int fuse = -1;
    if (fuse < 0) {
        const char *e = getenv("PCL_PHOTON_FUSE");
        fuse = e ? atoi(e) : PCL_FUSE_MAX;
        if (fuse < 1) fuse = 1;
        if (fuse > PCL_FUSE_MAX) fuse = PCL_FUSE_MAX;
    }
    return (uint32_t)fuse;
}

// PCL_MULTI_PREFETCH=1: the fused kernel prefetches its next tile with cp.async (tuning aid; see the kernel)
static int multi_prefetch() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCL_MULTI_PREFETCH");
        v = e ? atoi(e) : PCL_MULTI_PREFETCH_DEFAULT;
    }
    return v;
}

template <bool WAVE, bool DEL, bool INJ, bool PL>
static int launch_photon_pl(pcl_ctx *ctx, cudaStream_t st, const pcl_soa &p, const pcl_soa *dst, const StepK &K,
                            int64_t *row, uint64_t *n_out, uint32_t nsteps, bool keep_count) {
    const bool aligned = pcl_aligned16(p.x) && pcl_aligned16(p.y) && pcl_aligned16(p.z) && pcl_aligned16(p.vx) &&
                         pcl_aligned16(p.vy) && pcl_aligned16(p.vz) && pcl_aligned16(p.e) && pcl_aligned16(p.id) &&
                         pcl_aligned16(p.nscat) && pcl_aligned16(K.u_theta) && pcl_aligned16(K.u_phi) &&
                         pcl_aligned16(K.u_rand);
    if (K.sfu && !DEL) PCL_REQUIRE(ctx, aligned && !INJ, "PCL_SCATTER_SFU needs 16-byte aligned planes and Philox draws");
    if (dst) {  // retire-and-compact form: one kernel handles every slot, tail included
        PCL_REQUIRE(ctx, aligned, "the compacting step needs 16-byte aligned planes");
        if (!keep_count) PCL_CUDA(ctx, cudaMemsetAsync(n_out, 0, sizeof(uint64_t), st));
        unsigned grid = pcl_stream_grid(ctx, (p.n + 3) / 4, PCL_BLOCK, 8);
        const int pf = multi_prefetch() && !INJ;
        const size_t smem = pf ? (size_t)(6 + (p.id ? 1 : 0) + (WAVE ? 1 : 0) + (p.nscat ? 1 : 0)) * PCL_BLOCK * sizeof(float4) : 0;
        auto kern = pcl_k_photon_multi<WAVE, DEL, INJ, PL, true>;
        if constexpr (!INJ && !DEL) {
            if (K.sfu) kern = pcl_k_photon_multi<WAVE, DEL, INJ, PL, true, true>;
        }
        if (pf) PCL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PCL_BLOCK, smem, st>>>(p, *dst, K, row, (unsigned long long *)n_out, nsteps, pf);
        PCL_LAUNCHED(ctx);
        return 0;
    }
    static int single_multi = -1;
    if (single_multi < 0) {
        // Single timesteps go through the multi-step kernel too (nsteps = 1): since the arithmetic of round 2 it is the
        // fastest form at 16 Mi live photons (135 us; TMA-staged kernel 146 us, register-load kernel 143 us).
        // PCL_PHOTON_SINGLE_MULTI=0 restores the dedicated single-step kernels (PCL_PHOTON_TMA picks between them).
        const char *e = getenv("PCL_PHOTON_SINGLE_MULTI");
        single_multi = e ? atoi(e) : 1;
    }
    if (nsteps > 1 || ((single_multi || (K.sfu && !DEL)) && aligned && !INJ)) {  // several timesteps per HBM round trip, in place
        PCL_REQUIRE(ctx, aligned, "multi-step launches need 16-byte aligned planes");
        unsigned grid = pcl_stream_grid(ctx, (p.n + 3) / 4, PCL_BLOCK, 8);
        const int pf = multi_prefetch() && !INJ;
        const size_t smem = pf ? (size_t)(6 + (p.id ? 1 : 0) + (WAVE ? 1 : 0) + (p.nscat ? 1 : 0)) * PCL_BLOCK * sizeof(float4) : 0;
        auto kern = pcl_k_photon_multi<WAVE, DEL, INJ, PL, false>;
        if constexpr (!INJ && !DEL) {
            if (K.sfu) kern = pcl_k_photon_multi<WAVE, DEL, INJ, PL, false, true>;
        }
        if (pf) PCL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, PCL_BLOCK, smem, st>>>(p, p, K, row, nullptr, nsteps, pf);
        PCL_LAUNCHED(ctx);
        return 0;
    }
    if (aligned) {
        static int use_tma = -1;
        if (use_tma < 0) {
            const char *e = getenv("PCL_PHOTON_TMA");  // tuning aid: 0 = register-load kernel
            use_tma = e ? atoi(e) : 1;
        }
        if (!INJ && use_tma) {
            const int nplanes = 6 + (WAVE ? 1 : 0) + (p.id ? 1 : 0) + (p.nscat ? 1 : 0);
            const size_t smem = (size_t)2 * nplanes * PCL_WPLANE_BYTES * PCL_WARPS;
            auto kern = pcl_k_photon_step_tma<WAVE, DEL, PL>;
            PCL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const uint64_t tiles = (p.n + PCL_BLOCK * 4 - 1) / (PCL_BLOCK * 4);
            unsigned grid = (unsigned)(tiles < (uint64_t)ctx->sm_count * 4 ? (tiles ? tiles : 1) : (uint64_t)ctx->sm_count * 4);
            kern<<<grid, PCL_BLOCK, smem, st>>>(p, K, row);
        } else {
            unsigned grid = pcl_stream_grid(ctx, (p.n + 3) / 4, PCL_BLOCK, 8);
            pcl_k_photon_step<WAVE, DEL, INJ, PL><<<grid, PCL_BLOCK, 0, st>>>(p, K, row);
        }
        PCL_LAUNCHED(ctx);
    }
    if (!aligned) {  // scalar kernel for views that cannot take 128-bit accesses
        unsigned grid = pcl_stream_grid(ctx, p.n, PCL_BLOCK, 8);
        pcl_k_photon_step_tail<WAVE, DEL, INJ, PL><<<grid, PCL_BLOCK, 0, st>>>(p, K, row, 0);
        PCL_LAUNCHED(ctx);
    }
    return 0;
}

template <bool WAVE, bool DEL, bool INJ>
static int launch_photon(pcl_ctx *ctx, cudaStream_t st, const pcl_soa &p, const pcl_soa *dst, const StepK &K,
                         int64_t *row, uint64_t *n_out, uint32_t nsteps, bool keep_count) {
    return K.nplanes ? launch_photon_pl<WAVE, DEL, INJ, true>(ctx, st, p, dst, K, row, n_out, nsteps, keep_count)
                     : launch_photon_pl<WAVE, DEL, INJ, false>(ctx, st, p, dst, K, row, n_out, nsteps, keep_count);
}

static int photon_dispatch(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, const pcl_soa *dst, const StepK &K,
                           uint32_t mode, int64_t *row, uint64_t *n_out, uint32_t nsteps, bool keep_count) {
    const bool wave = mode & PCL_SCATTER_WAVELENGTH, del = mode & PCL_SCATTER_DELETE, inj = K.u_rand != nullptr;
    switch ((wave ? 4 : 0) | (del ? 2 : 0) | (inj ? 1 : 0)) {
        case 0: return launch_photon<false, false, false>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 1: return launch_photon<false, false, true>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 2: return launch_photon<false, true, false>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 3: return launch_photon<false, true, true>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 4: return launch_photon<true, false, false>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 5: return launch_photon<true, false, true>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        case 6: return launch_photon<true, true, false>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
        default: return launch_photon<true, true, true>(ctx, st, *p, dst, K, row, n_out, nsteps, keep_count);
    }
}

static int check_photon_view(pcl_ctx *ctx, const pcl_soa *p, const pcl_scatter_params *sp) {
    PCL_REQUIRE(ctx, p != nullptr && sp != nullptr, "null argument");
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->vx && p->vy && p->vz, "r and v planes are required");
    PCL_REQUIRE(ctx, p->n < (1ull << 32), "a shard holds fewer than 2^32 slots");
    // the Philox counter carries the low word of the global id and the key the high word (uniform per launch)
    PCL_REQUIRE(ctx, (p->id_base & 0xffffffffull) + p->n <= (1ull << 32), "the global ids of a shard must not cross a multiple of 2^32");
    if (sp->mode & PCL_SCATTER_WAVELENGTH) PCL_REQUIRE(ctx, p->e != nullptr, "wavelength law needs the e plane");
    return 0;
}

// nsteps timesteps in one launch (1 <= nsteps <= PCL_FUSE_MAX; tally_row: nsteps consecutive rows).
// dst: write the survivors of the last timestep densely into *dst (count in *n_out) instead of in place.
// keep_count: *n_out is not reset first, so several launches (the chunks of a host-buffer step) append
// to one output.
int pcl_photon_step_impl(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, const pcl_soa *dst, float dt,
                         const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                         int64_t *tally_row, uint64_t *n_out, uint32_t nsteps, bool keep_count) {
    int rc = check_photon_view(ctx, p, sp);
    if (rc) return rc;
    PCL_REQUIRE(ctx, nsteps >= 1 && nsteps <= PCL_FUSE_MAX, "bad number of timesteps per launch");
    PCL_REQUIRE(ctx, nsteps == 1 || !(rng && rng->u_rand), "injected uniforms are per timestep");
    PCL_REQUIRE(ctx, tally_row != nullptr, "tally_row is required");
    if (dst) {
        PCL_REQUIRE(ctx, n_out != nullptr, "n_out_dev is required");
        PCL_REQUIRE(ctx, dst->x && dst->y && dst->z && dst->vx && dst->vy && dst->vz && dst->id,
                    "dst needs r, v and id planes");
        PCL_REQUIRE(ctx, dst->x != p->x, "the compacting step writes out of place");
        if (p->e) PCL_REQUIRE(ctx, dst->e != nullptr, "dst needs the e plane (energies travel with the survivors)");
        if (p->nscat) PCL_REQUIRE(ctx, dst->nscat != nullptr, "dst needs the nscat plane");
        if (p->n == 0) {
            if (!keep_count) PCL_CUDA(ctx, cudaMemsetAsync(n_out, 0, sizeof(uint64_t), st));
            return 0;
        }
    }
    if (p->n == 0) return 0;
    StepK K;
    rc = pcl_fill_stepk(ctx, K, dt, sp, rng, escape_r2, planes, p->id_base);
    if (rc) return rc;
    return photon_dispatch(ctx, st, p, dst, K, sp->mode, tally_row, n_out, nsteps, keep_count);
}

extern "C" int pcl_photon_step(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                               const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                               const pcl_planes *planes, int64_t *tally_row) {
    PCL_ENTER(ctx);
    return pcl_photon_step_impl(ctx, (cudaStream_t)stream, p, nullptr, dt, sp, rng, escape_r2, planes, tally_row, nullptr, 1, false);
}

extern "C" int pcl_photon_step_compact(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *src, const pcl_soa *dst, float dt,
                                       const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                                       const pcl_planes *planes, int64_t *tally_row, uint64_t *n_out_dev) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, dst != nullptr, "dst is required");
    return pcl_photon_step_impl(ctx, (cudaStream_t)stream, src, dst, dt, sp, rng, escape_r2, planes, tally_row, n_out_dev, 1, false);
}

extern "C" int pcl_photon_steps_pp(pcl_ctx *ctx, uintptr_t stream, pcl_pingpong *pp, float dt,
                                   const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                                   const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps,
                                   uint32_t compact_every) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, pp != nullptr && rng != nullptr && tally_table != nullptr, "null argument");
    PCL_REQUIRE(ctx, rng->u_rand == nullptr, "multi-step runs draw from Philox; injected uniforms are per step");
    PCL_REQUIRE(ctx, pp->cur < 2 && pp->n_dev != nullptr, "bad ping-pong state");
    cudaStream_t st = (cudaStream_t)stream;
    PCL_CUDA(ctx, cudaMemsetAsync(tally_table, 0, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t), st));
    pcl_rng r = *rng;
    const uint32_t fuse = photon_fuse_max();
    for (uint32_t s = 0; s < nsteps;) {
        r.step = rng->step + s;
        int64_t *row = tally_table + (size_t)s * PCL_TALLY_COLS;
        pcl_soa src = pp->buf[pp->cur];
        src.n_dev = pp->n_dev + pp->cur;
        if (!((pp->id_valid >> pp->cur) & 1u)) src.id = nullptr;
        // this launch advances up to `fuse` timesteps in registers; it ends at the next compaction
        // boundary (every compact_every-th timestep, counted on the global step index), where the
        // survivors are written densely into the partner buffer
        uint32_t run = nsteps - s < fuse ? nsteps - s : fuse;
        bool compacting = false;
        if (compact_every) {
            const uint64_t to_boundary = compact_every - ((uint64_t)r.step % compact_every);  // 1 .. compact_every
            if (to_boundary <= run) {
                run = (uint32_t)to_boundary;
                compacting = true;
            }
        }
        int rc;
        if (compacting) {
            pcl_soa dst = pp->buf[pp->cur ^ 1];
            dst.n = src.n;
            rc = pcl_photon_step_impl(ctx, st, &src, &dst, dt, sp, &r, escape_r2, planes, row, pp->n_dev + (pp->cur ^ 1), run, false);
            if (rc == 0) {
                pp->buf[pp->cur ^ 1].n = src.n;  // upper bound; the exact count is n_dev[cur]
                pp->buf[pp->cur ^ 1].id_base = src.id_base;
                pp->cur ^= 1;
                pp->id_valid |= 1u << pp->cur;
            }
        } else {
            rc = pcl_photon_step_impl(ctx, st, &src, nullptr, dt, sp, &r, escape_r2, planes, row, nullptr, run, false);
        }
        if (rc) return rc;
        s += run;
    }
    return 0;
}

extern "C" int pcl_photon_steps(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                                const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                                const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, rng != nullptr, "rng is required");
    PCL_REQUIRE(ctx, rng->u_rand == nullptr, "multi-step runs draw from Philox; injected uniforms are per step");
    PCL_REQUIRE(ctx, tally_table != nullptr, "tally_table is required");
    cudaStream_t st = (cudaStream_t)stream;
    PCL_CUDA(ctx, cudaMemsetAsync(tally_table, 0, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t), st));
    pcl_rng r = *rng;
    const uint32_t fuse = photon_fuse_max();
    for (uint32_t s = 0; s < nsteps;) {
        const uint32_t run = nsteps - s < fuse ? nsteps - s : fuse;
        r.step = rng->step + s;
        int rc = pcl_photon_step_impl(ctx, st, p, nullptr, dt, sp, &r, escape_r2, planes,
                                      tally_table + (size_t)s * PCL_TALLY_COLS, nullptr, run, false);
        if (rc) return rc;
        s += run;
    }
    return 0;
}

extern "C" int pcl_scatter(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_scatter_params *sp,
                           const pcl_rng *rng, int32_t *flags, int64_t *tally_row) {
    PCL_ENTER(ctx);
    int rc = check_photon_view(ctx, p, sp);
    if (rc) return rc;
    PCL_REQUIRE(ctx, p->dx && p->dy && p->dz, "stand-alone scatter reads the dr planes");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    PCL_REQUIRE(ctx, !(sp->mode & PCL_SCATTER_SFU), "PCL_SCATTER_SFU applies to the fused photon steps only");
    if (p->n == 0) return 0;
    StepK K;
    rc = pcl_fill_stepk(ctx, K, 0.f, sp, rng, 0.f, nullptr, p->id_base);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = pcl_stream_grid(ctx, p->n, PCL_BLOCK, 8);
    const bool wave = sp->mode & PCL_SCATTER_WAVELENGTH, del = sp->mode & PCL_SCATTER_DELETE, inj = K.u_rand != nullptr;
#define PCL_SC(W, D, I) pcl_k_scatter<W, D, I><<<grid, PCL_BLOCK, 0, st>>>(*p, K, flags, tally_row, p->n)
    switch ((wave ? 4 : 0) | (del ? 2 : 0) | (inj ? 1 : 0)) {
        case 0: PCL_SC(false, false, false); break;
        case 1: PCL_SC(false, false, true); break;
        case 2: PCL_SC(false, true, false); break;
        case 3: PCL_SC(false, true, true); break;
        case 4: PCL_SC(true, false, false); break;
        case 5: PCL_SC(true, false, true); break;
        case 6: PCL_SC(true, true, false); break;
        default: PCL_SC(true, true, true); break;
    }
#undef PCL_SC
    PCL_LAUNCHED(ctx);
    return 0;
}

extern "C" int pcl_escape(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float r2, int64_t *tally_row) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p != nullptr && p->x && p->y && p->z, "r planes are required");
    PCL_REQUIRE(ctx, r2 > 0.f, "escape radius must be positive");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    if (p->n == 0) return 0;
    unsigned grid = pcl_stream_grid(ctx, p->n, PCL_BLOCK, 8);
    pcl_k_escape<<<grid, PCL_BLOCK, 0, (cudaStream_t)stream>>>(*p, r2, tally_row, p->n);
    PCL_LAUNCHED(ctx);
    return 0;
}

extern "C" int pcl_tally(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_planes *planes,
                         int64_t *tally_row) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, p != nullptr && tally_row != nullptr, "null argument");
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->vx && p->vy && p->vz, "r and v planes are required");
    if (planes && planes->count) PCL_REQUIRE(ctx, p->dx && p->dy && p->dz, "plane tallies read the dr planes");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    if (p->n == 0) return 0;
    StepK K;
    int rc = pcl_fill_stepk(ctx, K, 0.f, nullptr, nullptr, 0.f, planes, 0);
    if (rc) return rc;
    unsigned grid = pcl_stream_grid(ctx, p->n, PCL_BLOCK, 8);
    pcl_k_tally<<<grid, PCL_BLOCK, 0, (cudaStream_t)stream>>>(*p, K, tally_row, p->n);
    PCL_LAUNCHED(ctx);
    return 0;
}
