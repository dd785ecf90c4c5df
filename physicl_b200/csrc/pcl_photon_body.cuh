// The per-photon step body shared by every photon kernel: the hand-written ones in photon.cu and the
// run-time compiled (NVRTC) ones of pcl_jit_photon.cuh.  Device code only; the includer provides
// physicl_b200.h (pcl_soa, PCL_T_*, PCL_MAX_PLANES) and pcl_device.cuh.
#pragma once

struct StepK {
    float dt;
    float k;
    float c;
    float r2_escape;  // <= 0: no sphere
    uint32_t seed_lo, seed_hi;
    uint32_t step;
    uint32_t nplanes;
    uint32_t axis[PCL_MAX_PLANES];
    float loc[PCL_MAX_PLANES];
    const float *u_theta, *u_phi, *u_rand;
    // run-time compiled variable-density kernels only (light.py:295-299), all float64 like the reference:
    double kd;      // the kernel's scalar `A` [* (E0/(h c))^4 with the wavelength law]
    double e0;      // E = e * e0 for user expressions that read E[gid]
    double a_slot;  // what the reference binds to the kernel names `A` and `n` (light.py:287)
    double n_slot;
};

// Number density at the photon's position for ScatterIsotropicStep(variable_n=True): the user's
// OpenCL-C expression, spliced in by the run-time compiler (pcl_jit_photon.cuh defines
// PCL_USER_N_EXPR).  In the pre-compiled kernels the hook is never instantiated.
struct pcl_cl_gid {};
struct pcl_cl_arr {
    double v;
    __device__ __forceinline__ double operator[](pcl_cl_gid) const { return v; }
};
__device__ __forceinline__ double pcl_user_n(double r0_, double r1_, double r2_, double E_, double d0_, double d1_,
                                             double d2_, double norm, double A, double n) {
#ifdef PCL_USER_N_EXPR
    const pcl_cl_gid gid{};
    const pcl_cl_arr r0{r0_}, r1{r1_}, r2{r2_}, E{E_}, d0{d0_}, d1{d1_}, d2{d2_};
    (void)gid, (void)r0, (void)r1, (void)r2, (void)E, (void)d0, (void)d1, (void)d2, (void)norm, (void)A, (void)n;
    return (double)(PCL_USER_N_EXPR);
#else
    return 1.0;
#endif
}

enum { F_SCATTERED = 1, F_ABSORBED = 2, F_ESCAPED = 4 };

// tally columns held in registers per thread
enum { C_ALIVE, C_XP, C_YP, C_ZP, C_SCAT, C_ABS, C_ESC, C_LIVEIN, C_PLANE0, C_N = C_PLANE0 + PCL_MAX_PLANES };

__device__ __forceinline__ float pcl_norm3(float dx, float dy, float dz) {
    float s = dx * dx;
    s = fmaf(dy, dy, s);
    s = fmaf(dz, dz, s);
    return sqrtf(s);
}

// The scatter decision and the new direction for one photon.  dx,dy,dz is this step's dr.
// Written without branches on purpose: every lane evaluates both angles and selects, so the compiler
// can interleave the four photons a thread owns (independent chains) instead of serialising four
// divergent bodies.  At warp level nothing is lost: with pcoll ~ 0.3 some lane scatters in
// practically every warp, so the divergent form executed both sides anyway.
// VARN: the collision probability is formed in float64 from kn = kd * n(r) (see pcl_user_n).
template <bool WAVE, bool DEL, bool VARN = false>
__device__ __forceinline__ uint32_t pcl_scatter_one(bool live, float dx, float dy, float dz, float e,
                                                    float ut, float up, float ur, float k, float c,
                                                    float &vx, float &vy, float &vz, double kn = 0.0) {
    float norm = pcl_norm3(dx, dy, dz);
    bool hit;
    if (VARN) {
        double pd = kn * (double)norm;
        if (WAVE) {
            double e2 = (double)e * (double)e;
            pd = pd * (e2 * e2);
        }
        hit = live && (pd >= (double)ur);
    } else {
        float pcoll = k * norm;
        if (WAVE) {
            float e2 = e * e;
            float e4 = e2 * e2;
            pcoll = pcoll * e4;
        }
        hit = live && (pcoll >= ur);
    }
    if (DEL) return hit ? (F_SCATTERED | F_ABSORBED) : 0u;
    float st, ct, sp, cp;
    pcl_sincospi(ut + ut, st, ct);  // theta = 2*pi*u
    pcl_sincospi(up, sp, cp);       // phi   =   pi*u
    float cs = c * st;
    vx = hit ? cs * cp : vx;
    vy = hit ? cs * sp : vy;
    vz = hit ? c * ct : vz;
    return hit ? F_SCATTERED : 0u;
}

__device__ __forceinline__ void pcl_draw_at(const StepK &K, uint32_t step, uint64_t gid, float &ut, float &up, float &ur) {
    uint4 r = pcl_philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), step, 0u),
                                make_uint2(K.seed_lo, K.seed_hi));
    ut = pcl_u01(r.x);
    up = pcl_u01(r.y);
    ur = pcl_u01(r.z);
}
__device__ __forceinline__ void pcl_draw(const StepK &K, uint64_t gid, float &ut, float &up, float &ur) {
    pcl_draw_at(K, K.step, gid, ut, up, ur);
}

template <int NC>
__device__ __forceinline__ void pcl_tally_one(const StepK &K, bool on, float x, float y, float z, float dx,
                                              float dy, float dz, float vx, float vy, float vz,
                                              uint32_t (&cnt)[NC]) {
    cnt[C_ALIVE] += on ? 1u : 0u;
    cnt[C_XP] += (on && vx > 0.f) ? 1u : 0u;
    cnt[C_YP] += (on && vy > 0.f) ? 1u : 0u;
    cnt[C_ZP] += (on && vz > 0.f) ? 1u : 0u;
#pragma unroll
    for (int q = 0; q < NC - (int)C_PLANE0; ++q) {
        if ((uint32_t)q < K.nplanes) {
            uint32_t ax = K.axis[q];
            float r = ax == 0 ? x : (ax == 1 ? y : z);
            float d = ax == 0 ? dx : (ax == 1 ? dy : dz);
            float prev = r - d;  // light.py:386: obj.r[0] - obj.dr[0], evaluated after r += dr
            float loc = K.loc[q];
            bool hit = on && ((prev <= loc && loc <= r) || (prev >= loc && loc >= r));
            cnt[C_PLANE0 + q] += hit ? 1u : 0u;
        }
    }
}

template <int NC>
__device__ __forceinline__ void pcl_flush_tally(const uint32_t (&cnt)[NC], int64_t *row, uint32_t nplanes) {
    __shared__ unsigned int s_acc[C_N];
    if (threadIdx.x < C_N) s_acc[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        if (q < C_PLANE0 + (int)nplanes) {
            unsigned int w = __reduce_add_sync(0xffffffffu, cnt[q]);
            if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_acc[q], w);
        }
    }
    __syncthreads();
    if (threadIdx.x < C_N && s_acc[threadIdx.x]) {
        // register column -> PCL_T_* column (identical numbering by construction)
        atomicAdd((unsigned long long *)&row[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
    }
}

static_assert((int)C_ALIVE == (int)PCL_T_ALIVE && (int)C_XP == (int)PCL_T_XP && (int)C_ZP == (int)PCL_T_ZP &&
                  (int)C_SCAT == (int)PCL_T_SCATTERED && (int)C_ABS == (int)PCL_T_ABSORBED &&
                  (int)C_ESC == (int)PCL_T_ESCAPED && (int)C_LIVEIN == (int)PCL_T_LIVE_IN &&
                  (int)C_PLANE0 == (int)PCL_T_PLANE0 && (int)C_N == (int)PCL_TALLY_COLS,
              "register tally layout must match the ABI row layout");

// ---------------------------------------------------------------------------------------------
// One photon, one timestep: kinematics -> scatter -> escape -> tallies.  Shared by every fused kernel.
// On return x is NaN if the photon retired; v holds the new direction if it scattered.
// ---------------------------------------------------------------------------------------------
template <bool WAVE, bool DEL, bool VARN = false, int NC>
__device__ __forceinline__ uint32_t pcl_photon_one(const StepK &K, float &x, float &y, float &z, float &vx, float &vy,
                                                   float &vz, float e, float ut, float up, float ur,
                                                   uint32_t (&cnt)[NC]) {
    const bool live = x == x;  // a retired slot stays retired: NaN + dx is NaN
    cnt[C_LIVEIN] += live ? 1u : 0u;
    float dx = vx * K.dt, dy = vy * K.dt, dz = vz * K.dt;
    float xx = x + dx, yy = y + dy, zz = z + dz;
    double kn = 0.0;
    if (VARN)  // n(r) at the position reached in this timestep: the reference scatters after r += dr
        kn = K.kd * pcl_user_n((double)xx, (double)yy, (double)zz, (double)e * K.e0, (double)dx, (double)dy, (double)dz,
                               (double)pcl_norm3(dx, dy, dz), K.a_slot, K.n_slot);
    uint32_t f = pcl_scatter_one<WAVE, DEL, VARN>(live, dx, dy, dz, e, ut, up, ur, K.k, K.c, vx, vy, vz, kn);
    float r2 = xx * xx;
    r2 = fmaf(yy, yy, r2);
    r2 = fmaf(zz, zz, r2);
    const bool esc = live && !(f & F_ABSORBED) && K.r2_escape > 0.f && r2 >= K.r2_escape;
    f |= esc ? F_ESCAPED : 0u;
    cnt[C_SCAT] += (f & F_SCATTERED) ? 1u : 0u;
    cnt[C_ABS] += (f & F_ABSORBED) ? 1u : 0u;
    cnt[C_ESC] += esc ? 1u : 0u;
    const bool gone = (f & (F_ABSORBED | F_ESCAPED)) != 0u;
    pcl_tally_one(K, live && !gone, xx, yy, zz, dx, dy, dz, vx, vy, vz, cnt);
    x = (live && !gone) ? xx : __int_as_float(0x7fc00000);  // retired slots hold the canonical quiet NaN
    y = live ? yy : y;  // slots retired earlier keep their last position
    z = live ? zz : z;
    return f;
}

// number of valid slots: the view's n, or the device-resident count when the caller keeps it there
__device__ __forceinline__ uint64_t pcl_valid_slots(const pcl_soa &p) {
    if (p.n_dev) {
        uint64_t nd = *p.n_dev;
        return nd < p.n ? nd : p.n;
    }
    return p.n;
}

// Four consecutive photons held in registers: step them and write back in place (r always, v and
// nscat only when one of the four scattered).  Shared by the register-load and the TMA-staged kernels.
template <bool WAVE, bool DEL, bool INJ, bool VARN = false, int NC>
__device__ __forceinline__ void pcl_step_group4(const pcl_soa &p, const StepK &K, uint64_t i, float4 x, float4 y, float4 z,
                                                float4 vx, float4 vy, float4 vz, float4 e, uint4 id, bool has_id, uint4 nsc,
                                                float4 ut4, float4 up4, float4 ur4, uint32_t (&cnt)[NC]) {
    uint32_t any_scat = 0u;
    float ut[4], up[4], ur[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {  // four independent Philox chains: the compiler interleaves them
        if (INJ) {
            ut[l] = pcl_f4(ut4, l);
            up[l] = pcl_f4(up4, l);
            ur[l] = pcl_f4(ur4, l);
        } else {
            uint64_t gid = p.id_base + (has_id ? (uint64_t)pcl_u4(id, l) : (i + (uint64_t)l));
            pcl_draw(K, gid, ut[l], up[l], ur[l]);
        }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        uint32_t f = pcl_photon_one<WAVE, DEL, VARN>(K, pcl_f4(x, l), pcl_f4(y, l), pcl_f4(z, l), pcl_f4(vx, l),
                                                     pcl_f4(vy, l), pcl_f4(vz, l), pcl_f4(e, l), ut[l], up[l], ur[l], cnt);
        const bool sc = !DEL && (f & F_SCATTERED);
        any_scat |= sc ? 1u : 0u;
        pcl_u4(nsc, l) += sc ? 1u : 0u;
    }
    pcl_st4(p.x + i, x);
    pcl_st4(p.y + i, y);
    pcl_st4(p.z + i, z);
    if (any_scat) {
        pcl_st4(p.vx + i, vx);
        pcl_st4(p.vy + i, vy);
        pcl_st4(p.vz + i, vz);
        if (p.nscat) pcl_st4u(p.nscat + i, nsc);
    }
}

// one slot, scalar accesses
template <bool WAVE, bool DEL, bool INJ, bool VARN = false, int NC>
__device__ __forceinline__ void pcl_step_scalar(const pcl_soa &p, const StepK &K, uint64_t i, uint32_t (&cnt)[NC]) {
    float x = p.x[i];
    if (x != x) return;
    float y = p.y[i], z = p.z[i], vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
    float ut, up, ur;
    if (INJ) {
        ut = K.u_theta[i];
        up = K.u_phi[i];
        ur = K.u_rand[i];
    } else {
        uint64_t gid = p.id_base + (p.id ? (uint64_t)p.id[i] : i);
        pcl_draw(K, gid, ut, up, ur);
    }
    const float e = (WAVE || (VARN && p.e)) ? p.e[i] : 1.f;
    uint32_t f = pcl_photon_one<WAVE, DEL, VARN>(K, x, y, z, vx, vy, vz, e, ut, up, ur, cnt);
    p.x[i] = x;
    p.y[i] = y;
    p.z[i] = z;
    if (!DEL && (f & F_SCATTERED)) {
        p.vx[i] = vx;
        p.vy[i] = vy;
        p.vz[i] = vz;
        if (p.nscat) p.nscat[i] += 1u;
    }
}

// the (< 4 slot) tail after the last full group, scalar, by the first lanes of block 0
template <bool WAVE, bool DEL, bool INJ, bool VARN = false, int NC>
__device__ __forceinline__ void pcl_step_tail(const pcl_soa &p, const StepK &K, uint64_t first, uint32_t (&cnt)[NC]) {
    const uint64_t end = pcl_valid_slots(p);
    const uint64_t i = first + threadIdx.x;
    if (blockIdx.x != 0 || i >= end) return;
    pcl_step_scalar<WAVE, DEL, INJ, VARN>(p, K, i, cnt);
}
