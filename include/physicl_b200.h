/*
 * physicl_b200.h -- C ABI of the B200 (sm_100a) backend for PhysiCL's per-particle step path.
 *
 * The reference (bcwarner/physicl) has no FFI: its device boundary is eight pyopencl calls made
 * from Python (physicl/__init__.py:428-429, :597, :614, :653, :656, :662).  Each entry point below
 * names the reference code it stands in for.  All pointers in pcl_soa / params are raw DEVICE
 * pointers owned by the caller (the Python side keeps torch tensors alive); host-buffer entry
 * points say so in their name (*_host).  The library allocates only small scratch in pcl_ctx.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; text via pcl_last_error()
 *   - nothing throws, nothing synchronises unless the name ends in _sync or _host
 *   - `stream` is a cudaStream_t passed as uintptr_t (0 = legacy default stream)
 *   - every call does cudaSetDevice(ctx->device) first: the reference calls steps from the
 *     Simulation thread (physicl/__init__.py:501-516), not the thread that created the context
 *   - one pcl_ctx per Simulation per GPU; concurrent calls on one ctx are not supported (the
 *     reference serialises steps under its state lock, physicl/__init__.py:513-516)
 *
 * State layout: structure-of-arrays float32 planes in HBM.  A slot whose x is NaN is a retired
 * (absorbed / escaped) photon; every kernel skips it and every tally excludes it.
 */
#ifndef PHYSICL_B200_H
#define PHYSICL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCL_ABI_VERSION 1
#define PCL_MAX_PLANES 8

/* Tally row layout (int64 per column).  Mirrors the rows the reference's measure steps append:
 * ScatterSignMeasureStep -> [t, N, xp, yp, zp] (physicl/light.py:414-431),
 * ScatterMeasureStep     -> [t, N, crossings...] (physicl/light.py:374-404). */
enum {
    PCL_T_ALIVE = 0,     /* survivors after this step (== len(sim.objects))           */
    PCL_T_XP = 1,        /* #(v_x > 0) over survivors, strict >  (light.py:424)        */
    PCL_T_YP = 2,        /* #(v_y > 0)                            (light.py:425)        */
    PCL_T_ZP = 3,        /* #(v_z > 0)                            (light.py:426)        */
    PCL_T_SCATTERED = 4, /* photons whose direction was redrawn this step              */
    PCL_T_ABSORBED = 5,  /* photons removed by delete-mode scattering (light.py:258-260) */
    PCL_T_ESCAPED = 6,   /* photons retired by the escape sphere this step             */
    PCL_T_LIVE_IN = 7,   /* live photons entering the step (particle-steps processed)  */
    PCL_T_PLANE0 = 8,    /* PCL_T_PLANE0 + k : crossings of plane k                    */
    PCL_TALLY_COLS = 16
};

typedef struct pcl_ctx pcl_ctx;

/* SoA view of one shard of particles. n = number of slots (live + retired). */
typedef struct pcl_soa {
    uint64_t n;
    float *x, *y, *z;       /* Object.r  (physicl/__init__.py:390)                      */
    float *vx, *vy, *vz;    /* Object.v  (:393)                                         */
    float *dx, *dy, *dz;    /* Object.dr (:391), nullable: fused paths keep dr in regs  */
    float *ax, *ay, *az;    /* Object.a  (:394), nullable                               */
    float *e;               /* PhotonObject.E / E0 (light.py:34), nullable              */
    uint32_t *id;           /* local id, nullable => id = slot index                    */
    uint32_t *nscat;        /* scatter count per photon, nullable (stands in for the    */
                            /* dv!=0 counting of TracePathMeasureStep, light.py:459-460)*/
    uint64_t id_base;       /* global id = id_base + local id (RNG counter, sharding)   */
    const uint64_t *n_dev;  /* nullable DEVICE pointer: when set, the number of valid   */
                            /* slots is min(n, *n_dev), read by the kernel, so stepping  */
                            /* loops that retire photons never wait for the host         */
} pcl_soa;

/* Two SoA buffers used alternately by the retire-and-compact step, plus their device-side slot
 * counts.  buf[cur] holds the photons; n_dev[k] is the valid-slot count of buf[k]. */
typedef struct pcl_pingpong {
    pcl_soa buf[2];
    uint64_t *n_dev;   /* device uint64[2] */
    uint32_t cur;      /* 0 or 1 */
    uint32_t id_valid; /* bit k: buf[k].id holds real ids; otherwise the id of a slot is its index */
} pcl_pingpong;

/* Scatter law of light_scatter_step_sphere / light_scatter_step_del
 * (physicl/light.py:303-315, :146-158, :239-249). */
enum {
    PCL_SCATTER_WAVELENGTH = 1, /* pcoll *= (h c / E)^-4   (light.py:300-301)            */
    PCL_SCATTER_DELETE = 2,     /* scattered photons are removed (light.py:146-158)      */
    PCL_SCATTER_SFU = 4         /* new directions from the SFU (MUFU.SIN / MUFU.COS, sin.approx / cos.approx) instead of
                                 * the table + addition theorem: |error| ~ 5e-7, NOT reproducible on a CPU, so results
                                 * agree with the default in law (and in the first timestep's decisions), not bit for
                                 * bit.  Opt-in; fused Philox kernels on 16-byte aligned planes only. */
};
typedef struct pcl_scatter_params {
    float k;       /* A*n, or A*n*(E0/(h c))^4 with PCL_SCATTER_WAVELENGTH (folded in f64 on host) */
    float c;       /* code-unit speed of light: what str(c) pastes into the kernel (light.py:309) */
    uint32_t mode; /* PCL_SCATTER_* bits */
    uint32_t _pad;
} pcl_scatter_params;

/* Random numbers.  Injected uniforms reproduce the reference, which draws rtheta, rphi, rand per
 * photon on the host (light.py:285, :235, :181); all three arrays hold U[0,1) float32 values
 * (the 2*pi / pi scaling of :285 is applied on the device).  When u_rand is NULL the kernel draws
 * from Philox2x32-10: counter = (low word of the global id, step), key = a 32-bit fold of (seed, high
 * word of the id base); one 64-bit block per photon and timestep gives rand (24 bits), theta (24 bits)
 * and phi (16 bits).  The global ids of one view must not cross a multiple of 2^32. */
typedef struct pcl_rng {
    uint64_t seed;
    uint32_t step;
    uint32_t _pad;
    const float *u_theta, *u_phi, *u_rand;
} pcl_rng;

/* Axis-aligned measurement planes of ScatterMeasureStep (light.py:385-399). */
typedef struct pcl_planes {
    uint32_t count;
    uint32_t axis[PCL_MAX_PLANES]; /* 0,1,2: the one coordinate that is not NaN in `loc` */
    float loc[PCL_MAX_PLANES];
} pcl_planes;

/* ---- context ------------------------------------------------------------------------------- */
/* cl.create_some_context() + cl.CommandQueue(ctx)  (physicl/__init__.py:428-429) */
int pcl_init(int device, pcl_ctx **out);
int pcl_destroy(pcl_ctx *ctx);
const char *pcl_last_error(pcl_ctx *ctx); /* ctx may be NULL: last error of this thread */
int pcl_abi_version(void);
/* Simulation.get_device_info() (physicl/__init__.py:470-499): name, SM count, bytes of HBM, L2 */
int pcl_device_info(pcl_ctx *ctx, char *name, int name_len, int *sm_count, uint64_t *hbm_bytes,
                    uint64_t *l2_bytes);
/* kernels launched through this ctx so far (bench.py "gpu_launches") */
uint64_t pcl_launch_count(pcl_ctx *ctx);
int pcl_stream_sync(pcl_ctx *ctx, uintptr_t stream);
/* Timing aid (no reference counterpart): work queued on `stream` after this call waits ON THE DEVICE
 * until *flag_host (4 bytes of page-locked, mapped host memory) is non-zero, at most timeout_ms
 * (capped at 10 s).  bench.py queues a timed region behind it, so that CUDA events bracket GPU work only. */
int pcl_stream_gate(pcl_ctx *ctx, uintptr_t stream, const uint32_t *flag_host, uint32_t timeout_ms);

/* ---- steps --------------------------------------------------------------------------------- */
/* NewtonianKinematicsStep.run (physicl/newton.py:14-16): dr = v*dt; r += dr.
 * With accel != 0: v += a*dt first (a from the ax/ay/az planes if present, else `a_uniform[3]`);
 * that integrator is not in the reference (Object.a is never read). */
int pcl_kinematics(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt, int accel,
                   const float *a_uniform);
/* nsteps timesteps of equal dt in ONE launch: particles do not interact, so each thread keeps its
 * particles in registers for all nsteps and the state crosses HBM once.  Bit-identical to nsteps
 * calls of pcl_kinematics; dr holds the last step's displacement (newton.py:15). */
int pcl_kinematics_steps(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt, int accel,
                         const float *a_uniform, uint32_t nsteps);

/* ScatterIsotropicStep.__run_cl kernel + write-back (light.py:303-315, :325-331) and
 * ScatterDeleteStep (light.py:239-249, :258-260).  Reads the dr planes the kinematics step left.
 * `flags` (nullable, int32[n]) receives the reference's int result / NaN marker: 1 = scattered. */
int pcl_scatter(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_scatter_params *sp,
                const pcl_rng *rng, int32_t *flags, int64_t *tally_row);

/* Escape sphere as a stand-alone step (not in the reference: Simulation.bounds is stored and never
 * read, physicl/__init__.py:412): photons with |r|^2 >= r2 retire.  Adds ESCAPED / ALIVE / LIVE_IN
 * into tally_row (nullable). */
int pcl_escape(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float r2, int64_t *tally_row);

/* One whole photon timestep in one HBM round trip:
 * kinematics (newton.py:14-16) -> scatter (light.py:303-315) -> escape sphere (new) ->
 * sign + plane tallies (light.py:414-431, :385-399) accumulated into tally_row[PCL_TALLY_COLS].
 * escape_r2 <= 0 disables the sphere; planes may be NULL.  tally_row must be zeroed by the caller
 * (or use pcl_photon_steps). */
int pcl_photon_step(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                    const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                    const pcl_planes *planes, int64_t *tally_row);
/* nsteps timesteps; rng->step is the first step index; tally_table is int64[nsteps][PCL_TALLY_COLS]
 * in device memory, zeroed here (one row per timestep).  Photons do not interact, so a launch keeps
 * its photons in registers for up to 8 timesteps (PCL_PHOTON_FUSE, default 8; 1 = one launch per
 * timestep) and the state crosses HBM once per launch; results are bit-identical either way. */
int pcl_photon_steps(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float dt,
                     const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                     const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps);

/* Same timestep, with Simulation.remove_obj (physicl/__init__.py:455-459) folded in: photons are read
 * from `src` and the SURVIVORS are written densely to `dst` (out of place; dst needs r, v, id and
 * whatever optional planes src has), so retired photons cost nothing from the next step on.
 * Survivors keep their order inside a 1024-slot tile; tiles land in arrival order, ids identify
 * photons.  n_out_dev: device uint64 receiving the survivor count (zeroed by the call). */
int pcl_photon_step_compact(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *src, const pcl_soa *dst,
                            float dt, const pcl_scatter_params *sp, const pcl_rng *rng,
                            float escape_r2, const pcl_planes *planes, int64_t *tally_row,
                            uint64_t *n_out_dev);
/* nsteps timesteps over a ping-pong pair: the survivors are written densely into the partner buffer
 * after every timestep with (rng->step + s + 1) % compact_every == 0 (compact_every = 0: never).  A
 * launch covers the timesteps up to the next such boundary (at most 8) with the photons in registers.
 * pp->cur and the upper bounds pp->buf[].n are updated; exact counts stay on the device. */
int pcl_photon_steps_pp(pcl_ctx *ctx, uintptr_t stream, pcl_pingpong *pp, float dt,
                        const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                        const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps,
                        uint32_t compact_every);

/* ScatterSignMeasureStep.run / ScatterMeasureStep.run as stand-alone device tallies
 * (light.py:414-431, :374-404).  Adds into tally_row[PCL_TALLY_COLS] (caller zeroes). */
int pcl_tally(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_planes *planes,
              int64_t *tally_row);

/* ScatterMeasureStep(measure_E=True) (light.py:380-402): for plane q, the (id, e = E/E0) of every live
 * photon that crossed it this timestep, appended in arrival order to out_id/out_e[q*cap ...];
 * counts_dev[q] = number of crossers (zeroed here; entries beyond cap are dropped, the count is not). */
int pcl_plane_crossers(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, const pcl_planes *planes,
                       uint32_t *out_id, float *out_e, uint64_t *counts_dev, uint64_t cap);
/* TracePathMeasureStep.run (light.py:447-460): r of every live particle written to slab[c*n_ids + id],
 * c = 0,1,2; the caller pre-fills the slab with NaN ("object does not exist", light.py:435). */
int pcl_trace_positions(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *p, float *slab, uint64_t n_ids,
                        uint32_t *nscat_by_id /* nullable: latest scatter count per id (light.py:459-460) */);

/* Simulation.remove_obj for every retired photon (physicl/__init__.py:455-459): stable stream
 * compaction of live slots of `src` into `dst` (dst planes must not alias src; dst->id required).
 * n_live_dev: device uint64 receiving the live count. */
int pcl_compact(pcl_ctx *ctx, uintptr_t stream, const pcl_soa *src, const pcl_soa *dst,
                uint64_t *n_live_dev);

/* planck_phot_distribution (light.py:73-104): inverse-CDF draw on the reference's binned law.
 * cdf: device float64[ncdf] (= bins-1 cumulative masses, light.py:88-93); grid energy of bin x
 * is e_min + x*(e_max-e_min)/(bins-1) scaled by 1/E0 by the caller (e_lo, e_step are in E0 units).
 * bin_out (nullable) gets x, or -1 where the reference falls off its loop and returns None.
 * Uniform of photon gid = id_base + i: word gid & 3 of Philox4x32-10(counter (gid >> 2, 0, 1), key seed), top 24 bits. */
int pcl_planck_sample(pcl_ctx *ctx, uintptr_t stream, uint64_t n, uint64_t id_base, uint64_t seed,
                      const double *cdf, uint32_t ncdf, float e_lo, float e_step, float *e_out,
                      int32_t *bin_out);

/* All-pairs gravity (not in the reference; behind the Step API): for local bodies i in
 * [0, n_local) a_i = G * sum_j m_j (r_j - r_i) / (|r_ij|^2 + eps2)^(3/2) over posm_all[0..n_total),
 * float4 = (x, y, z, m).  posm_local are the i-bodies. Writes ax, ay, az (adds when accumulate).
 * j-bodies with index in [j_skip_begin, j_skip_end) are left out: a sharded caller accumulates its
 * own block while the all-gather is in flight, then everything else from the gathered array. */
int pcl_gravity_accel(pcl_ctx *ctx, uintptr_t stream, const float *posm_local, uint64_t n_local,
                      const float *posm_all, uint64_t n_total, float G, float eps2, float *ax,
                      float *ay, float *az, int accumulate, uint64_t j_skip_begin,
                      uint64_t j_skip_end);
/* The same when every j-body has the same mass m (BASELINE configs[3]): pass G*m; the kernel leaves the mass out of
 * the pair sum (11 instead of 12 FP32 operations per interaction).  posm_all[].w is ignored. */
int pcl_gravity_accel_uniform(pcl_ctx *ctx, uintptr_t stream, const float *posm_local, uint64_t n_local,
                              const float *posm_all, uint64_t n_total, float G_times_m, float eps2, float *ax,
                              float *ay, float *az, int accumulate, uint64_t j_skip_begin,
                              uint64_t j_skip_end);
/* kick-drift for gravity bodies: v += a*dt; r += v*dt on the packed posm (x,y,z,m); x,y,z
 * (nullable triple) are the store's SoA position planes, refreshed in the same pass. */
int pcl_gravity_kick_drift(pcl_ctx *ctx, uintptr_t stream, uint64_t n, float *posm, float *vx,
                           float *vy, float *vz, const float *ax, const float *ay, const float *az,
                           float dt, float *x, float *y, float *z);

/* Sharded form of the kick-drift: besides updating this rank's bodies it PUBLISHES them, storing every packed body into
 * slot slot0 + i of each of the `world` arrays in peer_bufs[] (device addresses of the ranks' gathered arrays in
 * peer-mapped / symmetric memory, this rank's own included).  Replaces the per-step all-gather: the transfer rides on the
 * integration kernel's own stores over NVLink.  The caller double-buffers the gathered arrays by timestep parity and
 * puts one cross-rank, stream-ordered barrier between this call and the next pcl_gravity_accel. */
int pcl_gravity_kick_drift_p2p(pcl_ctx *ctx, uintptr_t stream, uint64_t n, float *posm, float *vx,
                               float *vy, float *vz, const float *ax, const float *ay, const float *az,
                               float dt, float *x, float *y, float *z, const uint64_t *peer_bufs,
                               uint32_t world, uint64_t slot0);

/* ---- host-buffer entry point (the reference's per-step marshalling, __init__.py:602-664) ---- */
/* One fused photon step over HOST SoA planes: chunks are copied H2D, stepped and copied back D2H
 * on rotating streams so PCIe and the kernel overlap.  Host planes should be pinned
 * (pcl_host_register).  tally_row_host: int64[PCL_TALLY_COLS] written on return (synchronous). */
int pcl_photon_step_host(pcl_ctx *ctx, const pcl_soa *host, float dt, const pcl_scatter_params *sp,
                         const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                         int64_t *tally_row_host, uint64_t chunk);
/* Same, with Simulation.remove_obj folded in: only the survivors come back, written densely from the
 * front of the host planes (host->id is required: ids identify photons afterwards); *n_out_host =
 * survivors.  PCIe bytes are proportional to live photons (28 B up + 28 B down each). */
int pcl_photon_step_host_compact(pcl_ctx *ctx, const pcl_soa *host, float dt,
                                 const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                                 const pcl_planes *planes, int64_t *tally_row_host, uint64_t chunk,
                                 uint64_t *n_out_host);
/* nsteps (1..8) timesteps per host round trip, for callers that do not look at the particles between
 * those timesteps (Simulation.run with no host step in the list, physicl/__init__.py:512-516): a chunk is
 * uploaded once, advanced nsteps timesteps by one launch and its survivors come back.
 * tally_rows_host: int64[nsteps][PCL_TALLY_COLS], one row per timestep. */
int pcl_photon_steps_host_compact(pcl_ctx *ctx, const pcl_soa *host, float dt,
                                  const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                                  const pcl_planes *planes, int64_t *tally_rows_host, uint64_t chunk,
                                  uint32_t nsteps, uint64_t *n_out_host);
/* NewtonianKinematicsStep (physicl/newton.py:14-16; with accel the v += a dt law of pcl_kinematics) over HOST planes,
 * nsteps timesteps per round trip: r, v [, a] go up, r [, v with accel] [, dr when the host view has dr planes] come
 * back; same chunked pipeline as the photon entry points.  Bit-identical to pcl_kinematics_steps on resident planes. */
int pcl_kinematics_steps_host(pcl_ctx *ctx, const pcl_soa *host, float dt, int accel, const float *a_uniform,
                              uint32_t nsteps, uint64_t chunk);
int pcl_host_register(pcl_ctx *ctx, void *ptr, uint64_t bytes);
int pcl_host_unregister(pcl_ctx *ctx, void *ptr);

/* ---- run-time compiled kernels ------------------------------------------------------------------ */
/* cl.Program(ctx, src).build() (physicl/__init__.py:597, light.py:160) for sm_100a: CUDA C++ text in,
 * loaded module out (NVRTC: --gpu-architecture=sm_100a -fmad=false).  Needed where the reference's
 * kernel only exists as run-time text: the variable_n expression (light.py:295-299) and user CLProgram
 * kernels (physicl/__init__.py:583-597).  headers: in-memory includes (name, text).  The NVRTC log is
 * the error text on failure. */
typedef struct pcl_jit_module pcl_jit_module;
typedef struct pcl_jit_kernel pcl_jit_kernel;
int pcl_jit_build(pcl_ctx *ctx, const char *source, int n_headers, const char *const *header_names,
                  const char *const *header_texts, pcl_jit_module **out);
/* prog.<kernel_name> (physicl/__init__.py:656); the handle is owned by the module */
int pcl_jit_get_kernel(pcl_ctx *ctx, pcl_jit_module *m, const char *kernel_name, pcl_jit_kernel **out);
/* prog.<kernel>(queue, (n,), None, *args) (physicl/__init__.py:656): 1-D launch over n work items
 * (256 threads per CTA, grid-stride); args[i] points at the i-th argument value. */
int pcl_jit_launch(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, uint64_t n, void **args);
int pcl_jit_free(pcl_ctx *ctx, pcl_jit_module *m);
/* compile only: needs neither a context nor a device; log receives the NVRTC log on failure */
int pcl_jit_check(const char *source, int n_headers, const char *const *header_names,
                  const char *const *header_texts, char *log, uint64_t log_cap, uint64_t *cubin_bytes);

/* ScatterIsotropicStep(variable_n=True, variable_n_fn="<expression>") (light.py:295-299): the kernel
 * computes pcoll = A * (expression) * norm [* (h c / E)^-4] in float64, where the name `A` is bound to
 * the step's n and `n` to the step's A (light.py:287).  `k` is a kernel built from
 * csrc/pcl_jit_photon.cuh with the expression spliced in: "pcl_jit_photon_step" for the fused
 * timestep, "pcl_jit_scatter" for the stand-alone scatter on the dr planes. */
typedef struct pcl_varn_params {
    double kd;     /* the kernel's scalar A, times (E0/(h c))^4 with PCL_SCATTER_WAVELENGTH            */
    double e0;     /* E[gid] = e * e0 for expressions that read the photon energy                     */
    double a_slot; /* value of the kernel name `A` (= the step's n)                                   */
    double n_slot; /* value of the kernel name `n` (= the step's A)                                   */
} pcl_varn_params;
/* nsteps fused timesteps, in place (as pcl_photon_steps); tally_table: int64[nsteps][PCL_TALLY_COLS] */
int pcl_photon_steps_jit(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, const pcl_soa *p, float dt,
                         const pcl_scatter_params *sp, const pcl_varn_params *vn, const pcl_rng *rng,
                         float escape_r2, const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps);
/* stand-alone scatter (as pcl_scatter) */
int pcl_scatter_jit(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, const pcl_soa *p,
                    const pcl_scatter_params *sp, const pcl_varn_params *vn, const pcl_rng *rng, int32_t *flags,
                    int64_t *tally_row);

/* ---- roofline denominators measured on the box --------------------------------------------- */
/* FFMA-only and copy micro-kernels; results in TFLOP/s (2 flop per FMA) and GB/s (read+write). */
int pcl_measure_fp32_peak(pcl_ctx *ctx, double *tflops);
int pcl_measure_fp32x2_peak(pcl_ctx *ctx, double *tflops); /* packed FFMA2 (fma.rn.f32x2) rate */
int pcl_measure_copy_peak(pcl_ctx *ctx, uint64_t bytes, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* PHYSICL_B200_H */
