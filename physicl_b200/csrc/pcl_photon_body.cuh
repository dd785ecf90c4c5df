// The per-photon step body shared by every photon kernel: the hand-written ones in photon.cu and the
// run-time compiled (NVRTC) ones of pcl_jit_photon.cuh.  Device code only; the includer provides
// physicl_b200.h (pcl_soa, PCL_T_*, PCL_MAX_PLANES) and pcl_device.cuh.
#pragma once

struct StepK {
    float dt;
    float k;          // A*n [* (E0/(h c))^4]: run-time compiled variable-density kernels and diagnostics
    float kinv;       // 1/k, folded in float64 on the host: the collision test is |dr|^2 [e^8] >= (rand/k)^2
    float kinv24;     // kinv * 2^-24 (exact): multiplies the raw 24-bit integer draw
    float c;
    float r2_escape;  // NaN: no sphere (r2 >= NaN is false)
    uint32_t rk[10];  // Philox2x32 round keys key + r*W (key = fold of seed, high id word, stream)
    uint32_t step;
    uint32_t sfu;     // PCL_SCATTER_SFU: directions from MUFU sin / cos (fused Philox kernels)
    uint32_t nplanes;
    uint32_t axis[PCL_MAX_PLANES];
    float loc[PCL_MAX_PLANES];
    const float *u_theta, *u_phi, *u_rand;
    const float *trig;  // sin(2 pi k / 512), k < 512, then cos(...): device memory owned by the context
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

// tally columns held in registers per thread (stand-alone kernels: one 32-bit counter per column)
enum { C_ALIVE, C_XP, C_YP, C_ZP, C_SCAT, C_ABS, C_ESC, C_LIVEIN, C_PLANE0, C_N = C_PLANE0 + PCL_MAX_PLANES };

// The fused kernels count in PACKED form: four 8-bit fields per word, in column order
//   a = ALIVE | XP << 8 | YP << 16 | ZP << 24      b = SCAT | ABS << 8 | ESC << 16 | LIVEIN << 24
//   p0 = planes 0..3                                p1 = planes 4..7
// A thread owns four photons and hands its words to pcl_tally_to_shared after every timestep, so a field never
// exceeds 4 per thread and 128 per warp: one warp reduction per word, no carries between fields.
struct pcl_tally4 {
    uint32_t a, b, p0, p1;
};

__device__ __forceinline__ float pcl_norm3(float dx, float dy, float dz) {
    float s = dx * dx;
    s = fmaf(dy, dy, s);
    s = fmaf(dz, dz, s);
    return sqrtf(s);
}

// The scatter decision and the new direction for one photon.  dx,dy,dz is this step's dr.
//   reference (light.py:305-307):  norm = sqrt(d0^2+d1^2+d2^2);  pcoll = A n norm [ (hc/E)^-4 ];  pcoll >= rand
//   here, same inequality squared (both sides are >= 0):  d0^2+d1^2+d2^2 [ e^8 ] >= (rand / k)^2
// which needs no square root (the IEEE sqrtf and its slow-path branches were ~25 instructions per photon).
// 1/k comes from the host (float64 fold): k = 0 -> FLT_MAX (only rand = 0 scatters), k < 0 -> NaN (never).
// Written without branches on purpose: every lane evaluates both angles and selects, so the compiler
// can interleave the four photons a thread owns (independent chains) instead of serialising four
// divergent bodies.  At warp level nothing is lost: with pcoll ~ 0.3 some lane scatters in
// practically every warp, so the divergent form executed both sides anyway.
// VARN: the collision probability is formed in float64 from kn = kd * n(r) (see pcl_user_n).
template <bool WAVE, bool DEL, bool VARN = false>
__device__ __forceinline__ bool pcl_scatter_one(bool live, float dx, float dy, float dz, float e, const pcl_draw3 &d,
                                                const StepK &K, const unsigned char *tab, float &vx, float &vy,
                                                float &vz, double kn = 0.0) {
    float s = dx * dx;
    s = fmaf(dy, dy, s);
    s = fmaf(dz, dz, s);
    bool hit;
    if (VARN) {
        double pd = kn * (double)sqrtf(s);
        if (WAVE) {
            double e2 = (double)e * (double)e;
            pd = pd * (e2 * e2);
        }
        hit = live && (pd >= (double)d.ur);
    } else {
        const float q = d.ur * K.kinv;
        float lhs = s;
        if (WAVE) {
            const float e2 = e * e;
            const float e4 = e2 * e2;
            lhs = s * (e4 * e4);
        }
        hit = live && (lhs >= q * q);
    }
    if (DEL) return hit;
    float st, ct, sp, cp;
    pcl_sincos_tab(tab, d.at, d.bt, st, ct);  // theta = 2 pi u
    pcl_sincos_tab(tab, d.ap, d.bp, sp, cp);  // phi   =   pi u
    const float cs = K.c * st;
    vx = hit ? cs * cp : vx;
    vy = hit ? cs * sp : vy;
    vz = hit ? K.c * ct : vz;
    return hit;
}

// this photon's draws for timestep `step`: Philox2x32-10 on (low word of the global id, step)
__device__ __forceinline__ pcl_draw3 pcl_draw_at(const StepK &K, uint32_t step, uint32_t gid_lo) {
    const uint2 w = pcl_philox2x32_10(gid_lo, step, K.rk);
    return pcl_draw_bits(w.x, w.y);
}
// the same draw with its three fields still unscaled (see pcl_draw_bits_raw): for pcl_photon_two<..., RAW = true>
__device__ __forceinline__ pcl_draw3 pcl_draw_at_raw(const StepK &K, uint32_t step, uint32_t gid_lo) {
    const uint2 w = pcl_philox2x32_10(gid_lo, step, K.rk);
    return pcl_draw_bits_raw(w.x, w.y);
}

// stand-alone tallies (ScatterSignMeasureStep / ScatterMeasureStep kernels): one counter per column
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

// Per-timestep tally hand-over of a WARP (all 32 lanes must call it).  Four counters travel as the 8-bit
// fields of one word through ONE warp reduction, and lane q then adds column q to the CTA's row with a single
// predicated shared-memory atomic.  acc: unsigned int[PCL_TALLY_COLS] in shared memory.
template <bool PL>
__device__ __forceinline__ void pcl_tally_to_shared(const pcl_tally4 &t, unsigned int *acc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sh = 8u * (lane & 3u);
    uint32_t mine = 0u;
    uint32_t w = __reduce_add_sync(0xffffffffu, t.a);
    if ((lane >> 2) == 0u) mine = w >> sh;
    w = __reduce_add_sync(0xffffffffu, t.b);
    if ((lane >> 2) == 1u) mine = w >> sh;
    if (PL) {
        w = __reduce_add_sync(0xffffffffu, t.p0);
        if ((lane >> 2) == 2u) mine = w >> sh;
        w = __reduce_add_sync(0xffffffffu, t.p1);
        if ((lane >> 2) == 3u) mine = w >> sh;
    }
    mine &= 0xffu;
    if (lane < (PL ? 16u : 8u) && mine) atomicAdd(&acc[lane], mine);
}

// the same for code that only part of a warp executes (tails, scalar kernels): plain shared atomics
__device__ __forceinline__ void pcl_tally_to_shared_divergent(const pcl_tally4 &t, unsigned int *acc) {
    const uint32_t w[4] = {t.a, t.b, t.p0, t.p1};
#pragma unroll
    for (int q = 0; q < C_N; ++q) {
        const uint32_t v = (w[q >> 2] >> (8 * (q & 3))) & 0xffu;
        if (v) atomicAdd(&acc[q], v);
    }
}

// CTA row in shared memory -> tally row in HBM (call after a CTA barrier)
__device__ __forceinline__ void pcl_row_to_global(const unsigned int *acc, int64_t *row) {
    if (threadIdx.x < C_N) {
        const unsigned int v = acc[threadIdx.x];
        if (v) atomicAdd((unsigned long long *)&row[threadIdx.x], (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------
// One photon, one timestep: kinematics -> scatter -> escape -> tallies.  Shared by every fused kernel.
// On return x is NaN if the photon retired; v holds the new direction if it scattered (return value).
// A retired slot (x = NaN) is stepped with dt = 0: it stays retired, its other planes keep their bits.
// ---------------------------------------------------------------------------------------------
template <bool WAVE, bool DEL, bool VARN, bool PL>
__device__ __forceinline__ bool pcl_photon_one(const StepK &K, const unsigned char *tab, float &x, float &y, float &z,
                                               float &vx, float &vy, float &vz, float e, const pcl_draw3 &d,
                                               pcl_tally4 &t) {
    const bool live = x == x;
    const float dte = live ? K.dt : 0.f;
    const float dx = vx * dte, dy = vy * dte, dz = vz * dte;
    const float xx = x + dx, yy = y + dy, zz = z + dz;  // NaN + dx is NaN
    double kn = 0.0;
    if (VARN)  // n(r) at the position reached in this timestep: the reference scatters after r += dr
        kn = K.kd * pcl_user_n((double)xx, (double)yy, (double)zz, (double)e * K.e0, (double)dx, (double)dy, (double)dz,
                               (double)pcl_norm3(dx, dy, dz), K.a_slot, K.n_slot);
    const bool hit = pcl_scatter_one<WAVE, DEL, VARN>(live, dx, dy, dz, e, d, K, tab, vx, vy, vz, kn);
    float r2 = xx * xx;
    r2 = fmaf(yy, yy, r2);
    r2 = fmaf(zz, zz, r2);
    const bool absorbed = DEL && hit;
    const bool esc = (r2 >= K.r2_escape) && !absorbed;  // false for retired slots (r2 is NaN) and without a sphere
    const bool gone = absorbed || esc;
    const bool on = live && !gone;
    t.b += hit ? 1u : 0u;
    if (DEL) t.b += hit ? (1u << 8) : 0u;
    t.b += esc ? (1u << 16) : 0u;
    t.b += live ? (1u << 24) : 0u;
    uint32_t m = 1u;
    m += (vx > 0.f) ? (1u << 8) : 0u;
    m += (vy > 0.f) ? (1u << 16) : 0u;
    m += (vz > 0.f) ? (1u << 24) : 0u;
    t.a += on ? m : 0u;
    if (PL) {
#pragma unroll
        for (int q = 0; q < PCL_MAX_PLANES; ++q) {
            if ((uint32_t)q < K.nplanes) {
                const uint32_t ax = K.axis[q];
                const float r = ax == 0 ? xx : (ax == 1 ? yy : zz);
                const float dd = ax == 0 ? dx : (ax == 1 ? dy : dz);
                const float prev = r - dd;  // light.py:386: obj.r[0] - obj.dr[0], evaluated after r += dr
                const float loc = K.loc[q];
                const bool cross = on && ((prev <= loc && loc <= r) || (prev >= loc && loc >= r));
                if (q < 4)
                    t.p0 += cross ? (1u << (8 * q)) : 0u;
                else
                    t.p1 += cross ? (1u << (8 * (q - 4))) : 0u;
            }
        }
    }
    x = on ? xx : __int_as_float(0x7fc00000);  // retired slots hold the quiet NaN 0x7fc00000 (NaN + 0 would be 0x7fffffff)
    y = yy;
    z = zz;
    return hit;
}

// ---------------------------------------------------------------------------------------------
// Two photons, one timestep, with the floating-point work in packed form (one FFMA2 / FMUL2 / FADD2 does the same IEEE
// operation for photon A and photon B): per pair 36 packed instructions instead of 2 x 36 scalar ones; every half is
// bit-identical to pcl_photon_one.  Decisions, selects and tallies stay per photon.  Fixed laws only (no VARN).
// ---------------------------------------------------------------------------------------------
template <bool PL>
__device__ __forceinline__ void pcl_tally_photon(const StepK &K, bool live, bool hit, bool absorbed, bool esc, bool on, float xx,
                                                 float yy, float zz, float dx, float dy, float dz, float vx, float vy, float vz,
                                                 pcl_tally4 &t) {
    t.b += hit ? 1u : 0u;
    t.b += absorbed ? (1u << 8) : 0u;
    t.b += esc ? (1u << 16) : 0u;
    t.b += live ? (1u << 24) : 0u;
    uint32_t m = 1u;
    m += (vx > 0.f) ? (1u << 8) : 0u;
    m += (vy > 0.f) ? (1u << 16) : 0u;
    m += (vz > 0.f) ? (1u << 24) : 0u;
    t.a += on ? m : 0u;
    if (PL) {
#pragma unroll
        for (int q = 0; q < PCL_MAX_PLANES; ++q) {
            if ((uint32_t)q < K.nplanes) {
                const uint32_t ax = K.axis[q];
                const float r = ax == 0 ? xx : (ax == 1 ? yy : zz);
                const float dd = ax == 0 ? dx : (ax == 1 ? dy : dz);
                const float prev = r - dd;  // light.py:386: obj.r[0] - obj.dr[0], evaluated after r += dr
                const float loc = K.loc[q];
                const bool cross = on && ((prev <= loc && loc <= r) || (prev >= loc && loc >= r));
                if (q < 4)
                    t.p0 += cross ? (1u << (8 * q)) : 0u;
                else
                    t.p1 += cross ? (1u << (8 * (q - 4))) : 0u;
            }
        }
    }
}

template <bool WAVE, bool DEL, bool PL, bool RAW, bool SFU = false>
__device__ __forceinline__ void pcl_photon_two(const StepK &K, const unsigned char *tab, float &xA, float &xB, float &yA, float &yB,
                                               float &zA, float &zB, float &vxA, float &vxB, float &vyA, float &vyB, float &vzA,
                                               float &vzB, float eA, float eB, const pcl_draw3 &dA, const pcl_draw3 &dB,
                                               pcl_tally4 &t, bool &hitA, bool &hitB) {
    const bool liveA = xA == xA, liveB = xB == xB;
    const f32x2 dte = pk(liveA ? K.dt : 0.f, liveB ? K.dt : 0.f);
    const f32x2 dx = mul2(pk(vxA, vxB), dte), dy = mul2(pk(vyA, vyB), dte), dz = mul2(pk(vzA, vzB), dte);
    const f32x2 xx = add2(pk(xA, xB), dx), yy = add2(pk(yA, yB), dy), zz = add2(pk(zA, zB), dz);  // NaN + dx is NaN
    f32x2 s = mul2(dx, dx);
    s = fma2(dy, dy, s);
    s = fma2(dz, dz, s);
    const float kq = RAW ? K.kinv24 : K.kinv;  // RAW: (U * 2^-24) * kinv == U * (kinv * 2^-24), both scalings exact
    f32x2 q = mul2(pk(dA.ur, dB.ur), pk(kq, kq));
    q = mul2(q, q);
    if (WAVE) {
        const f32x2 e = pk(eA, eB);
        const f32x2 e2 = mul2(e, e);
        const f32x2 e4 = mul2(e2, e2);
        s = mul2(s, mul2(e4, e4));
    }
    float lA, lB, qA, qB;
    upk(s, lA, lB);
    upk(q, qA, qB);
    hitA = liveA && (lA >= qA);
    hitB = liveB && (lB >= qB);
    if (!DEL) {
        f32x2 st, ct, sp, cp;
        if (SFU) {  // theta = 2 pi m / 2^24, phi = pi m / 2^16 straight from the integer fields, then MUFU
            static_assert(!SFU || RAW, "the SFU form consumes raw Philox fields");
            pcl_sincos_sfu2(mul2(pk(dA.tf, dB.tf), pk(PCL_SFU_T_SCALE, PCL_SFU_T_SCALE)), st, ct);
            pcl_sincos_sfu2(mul2(pk(dA.pf, dB.pf), pk(PCL_SFU_P_SCALE, PCL_SFU_P_SCALE)), sp, cp);
        } else {
            f32x2 bt = pk(dA.bt, dB.bt), bp = pk(dA.bp, dB.bp);
            if (RAW) {
                bt = mul2(bt, pk(PCL_BT_SCALE, PCL_BT_SCALE));
                bp = mul2(bp, pk(PCL_BP_SCALE, PCL_BP_SCALE));
            }
            pcl_sincos_tab2(tab, dA.at, dB.at, bt, st, ct);  // theta = 2 pi u
            pcl_sincos_tab2(tab, dA.ap, dB.ap, bp, sp, cp);  // phi   =   pi u
        }
        const f32x2 c2 = pk(K.c, K.c);
        const f32x2 cs = mul2(c2, st);
        float nA, nB;
        upk(mul2(cs, cp), nA, nB);
        vxA = hitA ? nA : vxA;
        vxB = hitB ? nB : vxB;
        upk(mul2(cs, sp), nA, nB);
        vyA = hitA ? nA : vyA;
        vyB = hitB ? nB : vyB;
        upk(mul2(c2, ct), nA, nB);
        vzA = hitA ? nA : vzA;
        vzB = hitB ? nB : vzB;
    }
    f32x2 r2 = mul2(xx, xx);
    r2 = fma2(yy, yy, r2);
    r2 = fma2(zz, zz, r2);
    float r2A, r2B, xxA, xxB, yyA, yyB, zzA, zzB, dxA, dxB, dyA, dyB, dzA, dzB;
    upk(r2, r2A, r2B);
    upk(xx, xxA, xxB);
    upk(yy, yyA, yyB);
    upk(zz, zzA, zzB);
    upk(dx, dxA, dxB);
    upk(dy, dyA, dyB);
    upk(dz, dzA, dzB);
    const bool absA = DEL && hitA, absB = DEL && hitB;
    const bool escA = (r2A >= K.r2_escape) && !absA, escB = (r2B >= K.r2_escape) && !absB;
    const bool onA = liveA && !(absA || escA), onB = liveB && !(absB || escB);
    pcl_tally_photon<PL>(K, liveA, hitA, absA, escA, onA, xxA, yyA, zzA, dxA, dyA, dzA, vxA, vyA, vzA, t);
    pcl_tally_photon<PL>(K, liveB, hitB, absB, escB, onB, xxB, yyB, zzB, dxB, dyB, dzB, vxB, vyB, vzB, t);
    const float qnan = __int_as_float(0x7fc00000);
    xA = onA ? xxA : qnan;
    xB = onB ? xxB : qnan;
    yA = yyA;
    yB = yyB;
    zA = zzA;
    zB = zzB;
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
// All 32 lanes of a warp must call it together (warp-level tally hand-over); a lane without a group of its own
// passes x = NaN (counts nothing) and store = false.  acc: the CTA's shared row.
template <bool WAVE, bool DEL, bool INJ, bool VARN, bool PL>
__device__ __forceinline__ void pcl_step_group4_masked(const pcl_soa &p, const StepK &K, const unsigned char *tab, uint64_t i,
                                                       float4 x, float4 y, float4 z, float4 vx, float4 vy, float4 vz, float4 e,
                                                       uint4 id, uint4 nsc, float4 ut4, float4 up4, float4 ur4,
                                                       unsigned int *acc, bool store) {
    bool any_scat = false;
    pcl_draw3 d[4];
    const uint32_t base_lo = (uint32_t)p.id_base;
#pragma unroll
    for (int l = 0; l < 4; ++l)  // four independent Philox chains: the compiler interleaves them
        d[l] = INJ ? pcl_draw_floats(pcl_f4(ut4, l), pcl_f4(up4, l), pcl_f4(ur4, l))
                   : (VARN ? pcl_draw_at(K, K.step, base_lo + pcl_u4(id, l)) : pcl_draw_at_raw(K, K.step, base_lo + pcl_u4(id, l)));
    pcl_tally4 t = {0u, 0u, 0u, 0u};
    bool hit[4];
    if (VARN) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
            hit[l] = pcl_photon_one<WAVE, DEL, VARN, PL>(K, tab, pcl_f4(x, l), pcl_f4(y, l), pcl_f4(z, l), pcl_f4(vx, l),
                                                         pcl_f4(vy, l), pcl_f4(vz, l), pcl_f4(e, l), d[l], t);
    } else {
#pragma unroll
        for (int l = 0; l < 4; l += 2)
            pcl_photon_two<WAVE, DEL, PL, !INJ>(K, tab, pcl_f4(x, l), pcl_f4(x, l + 1), pcl_f4(y, l), pcl_f4(y, l + 1), pcl_f4(z, l),
                                          pcl_f4(z, l + 1), pcl_f4(vx, l), pcl_f4(vx, l + 1), pcl_f4(vy, l), pcl_f4(vy, l + 1),
                                          pcl_f4(vz, l), pcl_f4(vz, l + 1), pcl_f4(e, l), pcl_f4(e, l + 1), d[l], d[l + 1], t, hit[l],
                                          hit[l + 1]);
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const bool sc = !DEL && hit[l];
        any_scat = any_scat || sc;
        pcl_u4(nsc, l) += sc ? 1u : 0u;
    }
    pcl_tally_to_shared<PL>(t, acc);
    if (!store) return;
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

// one slot, scalar accesses; callable from divergent code
template <bool WAVE, bool DEL, bool INJ, bool VARN, bool PL>
__device__ __forceinline__ void pcl_step_scalar(const pcl_soa &p, const StepK &K, const unsigned char *tab, uint64_t i,
                                                unsigned int *acc) {
    float x = p.x[i];  // a retired slot goes through the same arithmetic as in the vector path (dt = 0)
    float y = p.y[i], z = p.z[i], vx = p.vx[i], vy = p.vy[i], vz = p.vz[i];
    const pcl_draw3 d = INJ ? pcl_draw_floats(K.u_theta[i], K.u_phi[i], K.u_rand[i])
                            : pcl_draw_at(K, K.step, (uint32_t)p.id_base + (p.id ? p.id[i] : (uint32_t)i));
    const float e = (WAVE || (VARN && p.e)) ? p.e[i] : 1.f;
    pcl_tally4 t = {0u, 0u, 0u, 0u};
    const bool hit = pcl_photon_one<WAVE, DEL, VARN, PL>(K, tab, x, y, z, vx, vy, vz, e, d, t);
    pcl_tally_to_shared_divergent(t, acc);
    p.x[i] = x;
    p.y[i] = y;
    p.z[i] = z;
    if (!DEL && hit) {
        p.vx[i] = vx;
        p.vy[i] = vy;
        p.vz[i] = vz;
        if (p.nscat) p.nscat[i] += 1u;
    }
}

// the (< 4 slot) tail after the last full group, scalar, by the first lanes of block 0
template <bool WAVE, bool DEL, bool INJ, bool VARN, bool PL>
__device__ __forceinline__ void pcl_step_tail(const pcl_soa &p, const StepK &K, const unsigned char *tab, uint64_t first,
                                              unsigned int *acc) {
    const uint64_t end = pcl_valid_slots(p);
    const uint64_t i = first + threadIdx.x;
    if (blockIdx.x != 0 || i >= end) return;
    pcl_step_scalar<WAVE, DEL, INJ, VARN, PL>(p, K, tab, i, acc);
}
