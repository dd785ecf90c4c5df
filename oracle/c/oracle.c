/*
 * oracle.c -- CPU restatement of PhysiCL's per-particle step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; the product (physicl_b200) never does and has no CPU fallback.
 *
 * Two families live here:
 *   orc_*_f64  the reference's algorithm in the reference's own precision (double), following
 *              physicl/newton.py:14-16, physicl/light.py:146-158, :239-249, :303-315, :325-331,
 *              :374-404, :414-431, :73-104.  Pinned against golden vectors produced by running the
 *              unmodified reference (tests/golden/make_golden.py).  Also the CPU baseline that
 *              bench.py times (OpenMP over all host cores).
 *   orc_*_f32  the binary32 twin: the same law evaluated with exactly the operation sequence the
 *              CUDA kernels use (mul, add, explicit fmaf, floorf, the 512-entry direction table; no
 *              contraction: this file is compiled with -ffp-contract=off), which makes decisions and
 *              integer tallies bit-identical to the GPU's.
 * Steps that do not exist in the reference (constant acceleration, escape sphere, gravity, the
 * Philox stream) are marked NEW: their parity is unpinned by the reference; they are checked by
 * invariants in tests/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_T_ALIVE 0
#define ORC_T_XP 1
#define ORC_T_YP 2
#define ORC_T_ZP 3
#define ORC_T_SCATTERED 4
#define ORC_T_ABSORBED 5
#define ORC_T_ESCAPED 6
#define ORC_T_LIVE_IN 7
#define ORC_T_PLANE0 8
#define ORC_TALLY_COLS 16
#define ORC_MAX_PLANES 8

#define ORC_WAVELENGTH 1u
#define ORC_DELETE 2u

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  NEW: the reference draws from NumPy's     */
/* global MT19937 on the host (light.py:285).                                                    */
/* ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float u01(uint32_t r) { return (float)(r >> 8) * 0x1p-24f; }

/* Philox2x32-10 (same paper; Random123 philox2x32_R(10, ...)): one multiply per round, 64 bits per block.
 * NEW, like every in-kernel draw: the photon steps use it with counter = (low word of the global id, step). */
void orc_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
    uint32_t c0 = ctr[0], c1 = ctr[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p >> 32) ^ key ^ c1;
        c1 = (uint32_t)p;
        c0 = n0;
        key += 0x9E3779B9u;
    }
    out[0] = c0;
    out[1] = c1;
}

/* 64 -> 32 bit fold of (seed, high word of the global id, stream): splitmix64 finaliser, upper word */
uint32_t orc_fold_key(uint64_t seed, uint64_t id_hi, uint64_t stream) {
    uint64_t z = seed ^ (id_hi * 0x9E3779B97F4A7C15ull) ^ (stream << 56);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

/* One photon's draws for one timestep, in the form the step body consumes them (see pcl_device.cuh):
 * ur = 24-bit uniform; theta = 2 pi (kt + ft), phi = pi (kp + fp) with kt, kp table indices (multiples of 2 pi/256 and
 * pi/256) and bt, bp the remainders in radians. */
typedef struct {
    float ur;
    uint32_t kt, kp; /* entry numbers in the 512-entry table */
    float bt, bp;
} draw3_t;

#define ORC_TWO_PI_256 0x1.921fb6p-6f
#define ORC_PI_256 0x1.921fb6p-7f

static inline draw3_t draw_bits(uint32_t w0, uint32_t w1) {
    draw3_t d;
    d.ur = (float)(w0 >> 8) * 0x1p-24f;
    d.kt = 2u * (w1 >> 24);
    d.bt = (float)((w1 >> 8) & 0xffffu) * (ORC_TWO_PI_256 * 0x1p-16f);
    d.kp = w1 & 0xffu;
    d.bp = (float)(w0 & 0xffu) * (ORC_PI_256 * 0x1p-8f);
    return d;
}

static inline draw3_t draw_floats(float ut, float up, float ur) {
    draw3_t d;
    d.ur = ur;
    float tt = ut * 256.0f, kt = floorf(tt);
    d.kt = 2u * ((uint32_t)(int)kt & 0xffu);
    d.bt = (tt - kt) * ORC_TWO_PI_256;
    float tp = up * 256.0f, kp = floorf(tp);
    d.kp = (uint32_t)(int)kp & 0xffu;
    d.bp = (tp - kp) * ORC_PI_256;
    return d;
}

static inline draw3_t draw_at(uint64_t gid, uint64_t seed, uint32_t step) {
    uint32_t ctr[2] = {(uint32_t)gid, step}, o[2];
    orc_philox2x32_10(ctr, orc_fold_key(seed, gid >> 32, 0), o);
    return draw_bits(o[0], o[1]);
}

/* the same draws as plain uniforms: u_theta = theta / 2 pi, u_phi = phi / pi (exact in binary32), u_rand */
static inline void draw3(uint64_t gid, uint64_t seed, uint32_t step, uint32_t stream, float *a, float *b, float *c) {
    if (stream != 0u) { /* emission sampler: Philox4x32-10, counter (id_lo, id_hi, step, stream), key = seed */
        uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), step, stream};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t o[4];
        orc_philox4x32_10(ctr, key, o);
        *a = u01(o[0]);
        *b = u01(o[1]);
        *c = u01(o[2]);
        return;
    }
    uint32_t ctr[2] = {(uint32_t)gid, step}, o[2];
    orc_philox2x32_10(ctr, orc_fold_key(seed, gid >> 32, 0), o);
    *a = (float)(o[1] >> 8) * 0x1p-24f;
    *b = (float)(((o[1] & 0xffu) << 8) | (o[0] & 0xffu)) * 0x1p-16f;
    *c = (float)(o[0] >> 8) * 0x1p-24f;
}

/* fill u_theta,u_phi,u_rand[n] with the draws of photons id_base+i at `step` */
void orc_philox_uniforms(uint64_t n, uint64_t id_base, uint64_t seed, uint32_t step, uint32_t stream, float *ut,
                         float *up, float *ur) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) draw3(id_base + (uint64_t)i, seed, step, stream, &ut[i], &up[i], &ur[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* binary32 twin                                                                                */
/* ------------------------------------------------------------------------------------------ */
/* Direction table (sin, cos)(2 pi k / 512), double rounded once to binary32 (same expression as pcl_init). */
static float g_trig[512][2];
static int g_trig_ready = 0;
static void trig_init(void) {
    if (g_trig_ready) return;
#pragma omp critical(orc_trig)
    {
        if (!g_trig_ready) {
            for (int k = 0; k < 512; ++k) {
                const double a = 2.0 * 3.14159265358979323846 * (double)k / 512.0;
                g_trig[k][0] = (float)sin(a);
                g_trig[k][1] = (float)cos(a);
            }
            g_trig_ready = 1;
        }
    }
}

/* sin, cos of (table angle k) + b by the addition theorem, cos b = 1 - b^2/2, sin b = b - b^3/6; mul / fmaf only:
 * s = fma(sa, cb, ca*sb); c = fma(ca, cb, -(sa*sb)) */
void orc_sincos_tab(uint32_t k, float b, float *s, float *c) {
    trig_init();
    const float sa = g_trig[k & 511u][0], ca = g_trig[k & 511u][1];
    const float b2 = b * b;
    const float cb = fmaf(b2, -0.5f, 1.0f);
    const float sb = fmaf(b * b2, -0x1.555556p-3f, b);
    *s = fmaf(sa, cb, ca * sb);
    *c = fmaf(ca, cb, -(sa * sb));
}

/* light.py:305-311 in binary32; returns 1 if scattered.  The reference's test pcoll = k*norm [e^4] >= rand is
 * evaluated squared, norm^2 [e^8] >= (rand / k)^2 (both sides >= 0; kinv = 1/k folded in double by the caller),
 * which is the operation sequence of the CUDA kernels (no square root). */
static inline int scatter_one_f32(float dx, float dy, float dz, float e, const draw3_t *d, float kinv, float c,
                                  uint32_t mode, float *vx, float *vy, float *vz) {
    float s = dx * dx;
    s = fmaf(dy, dy, s);
    s = fmaf(dz, dz, s);
    float q = d->ur * kinv;
    float lhs = s;
    if (mode & ORC_WAVELENGTH) {
        float e2 = e * e;
        float e4 = e2 * e2;
        lhs = s * (e4 * e4);
    }
    if (!(lhs >= q * q)) return 0;
    if (mode & ORC_DELETE) return 1;
    float st, ct, sp, cp;
    orc_sincos_tab(d->kt, d->bt, &st, &ct);
    orc_sincos_tab(d->kp, d->bp, &sp, &cp);
    float cs = c * st;
    *vx = cs * cp;
    *vy = cs * sp;
    *vz = c * ct;
    return 1;
}

/* 1/k as the CUDA library folds it (pcl_fill_stepk) */
static inline float kinv_of(float k) {
    double kd = (double)k;
    if (kd > 0.0) {
        double inv = 1.0 / kd;
        return inv > 3.4028234663852886e38 ? 3.4028234663852886e38f : (float)inv;
    }
    if (kd == 0.0) return 3.4028234663852886e38f;
    return NAN;
}

static inline void tally_one_f32(float x, float y, float z, float dx, float dy, float dz, float vx, float vy,
                                 float vz, uint32_t nplanes, const uint32_t *axis, const float *loc, int64_t *row) {
    row[ORC_T_ALIVE] += 1;
    row[ORC_T_XP] += vx > 0.f;
    row[ORC_T_YP] += vy > 0.f;
    row[ORC_T_ZP] += vz > 0.f;
    for (uint32_t q = 0; q < nplanes; ++q) {
        float r = axis[q] == 0 ? x : (axis[q] == 1 ? y : z);
        float d = axis[q] == 0 ? dx : (axis[q] == 1 ? dy : dz);
        float prev = r - d; /* light.py:386 */
        float l = loc[q];
        row[ORC_T_PLANE0 + q] += (prev <= l && l <= r) || (prev >= l && l >= r);
    }
}

/* newton.py:14-16 (accel: 0 reference law; 1 a planes; 2 uniform a -- NEW) */
void orc_kinematics_f32(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz, float *dx,
                        float *dy, float *dz, const float *ax, const float *ay, const float *az, float dt, int accel,
                        const float *au) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        float u = vx[i], v = vy[i], w = vz[i];
        if (accel == 1) {
            u = u + ax[i] * dt;
            v = v + ay[i] * dt;
            w = w + az[i] * dt;
        } else if (accel == 2) {
            u = u + au[0] * dt;
            v = v + au[1] * dt;
            w = w + au[2] * dt;
        }
        float a = u * dt, b = v * dt, c = w * dt;
        x[i] = x[i] + a;
        y[i] = y[i] + b;
        z[i] = z[i] + c;
        if (accel) {
            vx[i] = u;
            vy[i] = v;
            vz[i] = w;
        }
        if (dx) {
            dx[i] = a;
            dy[i] = b;
            dz[i] = c;
        }
    }
}

/* fused step twin of pcl_photon_step: kinematics -> scatter -> escape (NEW) -> tallies */
void orc_photon_step_f32(uint64_t n, float *x, float *y, float *z, float *vx, float *vy, float *vz, const float *e,
                         const uint32_t *id, uint32_t *nscat, uint64_t id_base, float dt, float k, float c,
                         uint32_t mode, uint64_t seed, uint32_t step, const float *ut, const float *up,
                         const float *ur, float r2_escape, uint32_t nplanes, const uint32_t *axis, const float *loc,
                         int64_t *row) {
    int64_t acc[ORC_TALLY_COLS];
    memset(acc, 0, sizeof(acc));
    const float kinv = kinv_of(k);
    trig_init();
#pragma omp parallel
    {
        int64_t loc_row[ORC_TALLY_COLS];
        memset(loc_row, 0, sizeof(loc_row));
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            float xx = x[i];
            if (xx != xx) { /* retired slot: the kernels step it with dt = 0 (x stays NaN, y and z keep their value) */
                y[i] = y[i] + vy[i] * 0.f;
                z[i] = z[i] + vz[i] * 0.f;
                continue;
            }
            loc_row[ORC_T_LIVE_IN] += 1;
            float dx = vx[i] * dt, dy = vy[i] * dt, dz = vz[i] * dt;
            xx = xx + dx;
            float yy = y[i] + dy, zz = z[i] + dz;
            draw3_t d;
            if (ur) {
                d = draw_floats(ut ? ut[i] : ur[i], up ? up[i] : ur[i], ur[i]);
            } else {
                d = draw_at(id_base + (id ? (uint64_t)id[i] : (uint64_t)i), seed, step);
            }
            float nvx = vx[i], nvy = vy[i], nvz = vz[i];
            int sc = scatter_one_f32(dx, dy, dz, e ? e[i] : 1.f, &d, kinv, c, mode, &nvx, &nvy, &nvz);
            int absorbed = sc && (mode & ORC_DELETE);
            int escaped = 0;
            if (!absorbed && r2_escape > 0.f) {
                float r2 = xx * xx;
                r2 = fmaf(yy, yy, r2);
                r2 = fmaf(zz, zz, r2);
                escaped = r2 >= r2_escape;
            }
            loc_row[ORC_T_SCATTERED] += sc;
            loc_row[ORC_T_ABSORBED] += absorbed;
            loc_row[ORC_T_ESCAPED] += escaped;
            if (sc && !absorbed) {
                vx[i] = nvx;
                vy[i] = nvy;
                vz[i] = nvz;
                if (nscat) nscat[i] += 1u;
            }
            if (absorbed || escaped) {
                xx = NAN;
            } else {
                tally_one_f32(xx, yy, zz, dx, dy, dz, nvx, nvy, nvz, nplanes, axis, loc, loc_row);
            }
            x[i] = xx;
            y[i] = yy;
            z[i] = zz;
        }
#pragma omp critical
        for (int q = 0; q < ORC_TALLY_COLS; ++q) acc[q] += loc_row[q];
    }
    for (int q = 0; q < ORC_TALLY_COLS; ++q) row[q] += acc[q];
}

/* stand-alone scatter twin of pcl_scatter (reads dr planes), flags as light.py:151-155 */
void orc_scatter_f32(uint64_t n, float *x, float *vx, float *vy, float *vz, const float *dx, const float *dy,
                     const float *dz, const float *e, const uint32_t *id, uint32_t *nscat, uint64_t id_base, float k,
                     float c, uint32_t mode, uint64_t seed, uint32_t step, const float *ut, const float *up,
                     const float *ur, int32_t *flags, int64_t *row) {
    int64_t alive = 0, scat = 0, absd = 0, livein = 0;
    const float kinv = kinv_of(k);
    trig_init();
#pragma omp parallel for schedule(static) reduction(+ : alive, scat, absd, livein)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int32_t flag = 0;
        if (x[i] == x[i]) {
            livein += 1;
            draw3_t d;
            if (ur) {
                d = draw_floats(ut ? ut[i] : ur[i], up ? up[i] : ur[i], ur[i]);
            } else {
                d = draw_at(id_base + (id ? (uint64_t)id[i] : (uint64_t)i), seed, step);
            }
            float nvx = 0.f, nvy = 0.f, nvz = 0.f;
            int sc = scatter_one_f32(dx[i], dy[i], dz[i], e ? e[i] : 1.f, &d, kinv, c, mode, &nvx, &nvy, &nvz);
            if (sc) {
                flag = 1;
                scat += 1;
                if (mode & ORC_DELETE) {
                    absd += 1;
                    x[i] = NAN;
                } else {
                    vx[i] = nvx;
                    vy[i] = nvy;
                    vz[i] = nvz;
                    if (nscat) nscat[i] += 1u;
                }
            }
            if (!(sc && (mode & ORC_DELETE))) alive += 1;
        }
        if (flags) flags[i] = flag;
    }
    if (row) {
        row[ORC_T_ALIVE] += alive;
        row[ORC_T_SCATTERED] += scat;
        row[ORC_T_ABSORBED] += absd;
        row[ORC_T_LIVE_IN] += livein;
    }
}

/* light.py:414-431 and :374-404 over live slots */
void orc_tally_f32(uint64_t n, const float *x, const float *y, const float *z, const float *vx, const float *vy,
                   const float *vz, const float *dx, const float *dy, const float *dz, uint32_t nplanes,
                   const uint32_t *axis, const float *loc, int64_t *row) {
    for (uint64_t i = 0; i < n; ++i) {
        if (x[i] != x[i]) continue;
        tally_one_f32(x[i], y[i], z[i], dx ? dx[i] : 0.f, dy ? dy[i] : 0.f, dz ? dz[i] : 0.f, vx[i], vy[i], vz[i],
                      nplanes, axis, loc, row);
    }
}

/* light.py:101-104 on a prebuilt CDF; uniform from Philox4x32 stream 1, one block per four photons */
void orc_planck_sample(uint64_t n, uint64_t id_base, uint64_t seed, const double *cdf, uint32_t ncdf, float e_lo,
                       float e_step, float *e_out, int32_t *bin_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        /* photon gid takes word gid & 3 of the Philox4x32-10 block with counter (gid >> 2, step 0, stream 1) */
        const uint64_t gid = id_base + (uint64_t)i, q = gid >> 2;
        uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(q >> 32), 0u, 1u};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t o[4];
        orc_philox4x32_10(ctr, key, o);
        double u = (double)u01(o[gid & 3u]);
        int32_t bin = -1;
        /* the reference's own linear scan, verbatim in meaning: first x>=1 with cdf[x]>=u>=cdf[x-1] */
        for (uint32_t xq = 1; xq < ncdf; ++xq) {
            if (cdf[xq] >= u && u >= cdf[xq - 1]) {
                bin = (int32_t)xq;
                break;
            }
        }
        e_out[i] = bin < 0 ? NAN : fmaf((float)bin, e_step, e_lo);
        if (bin_out) bin_out[i] = bin;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* double precision: the reference's law as the reference computes it                           */
/* ------------------------------------------------------------------------------------------ */
/* newton.py:14-16 */
void orc_kinematics_f64(uint64_t n, double *x, double *y, double *z, const double *vx, const double *vy,
                        const double *vz, double *dx, double *dy, double *dz, double dt) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        dx[i] = vx[i] * dt;
        dy[i] = vy[i] * dt;
        dz[i] = vz[i] * dt;
        x[i] += dx[i];
        y[i] += dy[i];
        z[i] += dz[i];
    }
}

/* NEW (not in the reference, whose Object.a is never read): semi-implicit Euler with per-particle a,
 * v += a dt; dr = v dt; r += dr -- the law of BASELINE configs[0] / configs[4], in double, for the CPU arm */
void orc_kinematics_accel_f64(uint64_t n, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                              const double *ax, const double *ay, const double *az, double *dx, double *dy, double *dz,
                              double dt) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        vx[i] += ax[i] * dt;
        vy[i] += ay[i] * dt;
        vz[i] += az[i] * dt;
        dx[i] = vx[i] * dt;
        dy[i] = vy[i] * dt;
        dz[i] = vz[i] * dt;
        x[i] += dx[i];
        y[i] += dy[i];
        z[i] += dz[i];
    }
}

/* light.py:303-315 kernel body: res0 = NAN marks "unaffected"; hc_over applies :300-301 when E given */
void orc_scatter_sphere_f64(uint64_t n, const double *d0, const double *d1, const double *d2, const double *rtheta,
                            const double *rphi, const double *rnd, double A, double nd, const double *E, double hc,
                            double c, double *res0, double *res1, double *res2) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < (int64_t)n; ++g) {
        double norm = sqrt(pow(d0[g], 2) + pow(d1[g], 2) + pow(d2[g], 2));
        double pcoll = A * nd * norm;
        if (E) pcoll = pcoll * pow(hc / E[g], -4);
        if (pcoll >= rnd[g]) {
            res0[g] = c * sin(rtheta[g]) * cos(rphi[g]);
            res1[g] = c * sin(rtheta[g]) * sin(rphi[g]);
            res2[g] = c * cos(rtheta[g]);
        } else {
            res0[g] = NAN;
        }
    }
}

/* light.py:146-158 */
void orc_scatter_del_f64(uint64_t n, const double *d0, const double *d1, const double *d2, const double *rnd, double nd,
                         double A, int32_t *result) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < (int64_t)n; ++g) {
        double norm = sqrt(pow(d0[g], 2) + pow(d1[g], 2) + pow(d2[g], 2));
        double pcoll = A * nd * norm;
        result[g] = (pcoll >= rnd[g]) ? 1 : 0;
    }
}

/* Whole reference timestep in double for the CPU baseline: newton.py:14-16, then light.py:303-315
 * with the write-back of :325-331, then the sign tally of :414-431, plus the NEW escape sphere.
 * Uniforms come from the same Philox stream as the GPU path so both arms do the same work.
 * A retired photon has x = NaN. */
void orc_photon_step_f64(uint64_t n, double *x, double *y, double *z, double *vx, double *vy, double *vz,
                         const double *E, uint64_t id_base, double dt, double A, double nd, double hc, double c,
                         uint32_t mode, uint64_t seed, uint32_t step, double r2_escape, int64_t *row) {
    int64_t alive = 0, xp = 0, yp = 0, zp = 0, scat = 0, absd = 0, esc = 0, livein = 0;
#pragma omp parallel for schedule(static) reduction(+ : alive, xp, yp, zp, scat, absd, esc, livein)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        if (x[i] != x[i]) continue;
        livein += 1;
        double dx = vx[i] * dt, dy = vy[i] * dt, dz = vz[i] * dt;
        x[i] += dx;
        y[i] += dy;
        z[i] += dz;
        float a, b, r;
        draw3(id_base + (uint64_t)i, seed, step, 0u, &a, &b, &r);
        double rtheta = (double)a * 2 * M_PI, rphi = (double)b * M_PI;
        double norm = sqrt(pow(dx, 2) + pow(dy, 2) + pow(dz, 2));
        double pcoll = A * nd * norm;
        if (mode & ORC_WAVELENGTH) pcoll = pcoll * pow(hc / E[i], -4);
        int sc = pcoll >= (double)r;
        int dead = 0;
        if (sc) {
            scat += 1;
            if (mode & ORC_DELETE) {
                absd += 1;
                dead = 1;
            } else {
                vx[i] = c * sin(rtheta) * cos(rphi);
                vy[i] = c * sin(rtheta) * sin(rphi);
                vz[i] = c * cos(rtheta);
            }
        }
        if (!dead && r2_escape > 0.0 && x[i] * x[i] + y[i] * y[i] + z[i] * z[i] >= r2_escape) {
            esc += 1;
            dead = 1;
        }
        if (dead) {
            x[i] = NAN;
        } else {
            alive += 1;
            xp += vx[i] > 0;
            yp += vy[i] > 0;
            zp += vz[i] > 0;
        }
    }
    row[ORC_T_ALIVE] += alive;
    row[ORC_T_XP] += xp;
    row[ORC_T_YP] += yp;
    row[ORC_T_ZP] += zp;
    row[ORC_T_SCATTERED] += scat;
    row[ORC_T_ABSORBED] += absd;
    row[ORC_T_ESCAPED] += esc;
    row[ORC_T_LIVE_IN] += livein;
}

/* NEW: all-pairs softened gravity in double, from the definition */
void orc_gravity_f64(uint64_t n_local, uint64_t i_offset, uint64_t n_total, const double *px, const double *py,
                     const double *pz, const double *m, double G, double eps2, double *ax, double *ay, double *az) {
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)n_local; ++ii) {
        uint64_t i = i_offset + (uint64_t)ii;
        double sx = 0, sy = 0, sz = 0;
        for (uint64_t j = 0; j < n_total; ++j) {
            double dx = px[j] - px[i], dy = py[j] - py[i], dz = pz[j] - pz[i];
            double r2 = dx * dx + dy * dy + dz * dz + eps2;
            double inv = 1.0 / (r2 * sqrt(r2));
            sx += m[j] * dx * inv;
            sy += m[j] * dy * inv;
            sz += m[j] * dz * inv;
        }
        ax[ii] = G * sx;
        ay[ii] = G * sy;
        az[ii] = G * sz;
    }
}

/* the same definition for an arbitrary list of i-bodies (full-size spot checks) */
void orc_gravity_pick_f64(uint64_t n_pick, const uint64_t *idx, uint64_t n_total, const double *px, const double *py,
                          const double *pz, const double *m, double G, double eps2, double *ax, double *ay, double *az) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < (int64_t)n_pick; ++q) {
        const uint64_t i = idx[q];
        double sx = 0, sy = 0, sz = 0;
        for (uint64_t j = 0; j < n_total; ++j) {
            double dx = px[j] - px[i], dy = py[j] - py[i], dz = pz[j] - pz[i];
            double r2 = dx * dx + dy * dy + dz * dz + eps2;
            double inv = 1.0 / (r2 * sqrt(r2));
            sx += m[j] * dx * inv;
            sy += m[j] * dy * inv;
            sz += m[j] * dz * inv;
        }
        ax[q] = G * sx;
        ay[q] = G * sy;
        az[q] = G * sz;
    }
}

int orc_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}
