#!/usr/bin/env python
"""Benchmark of the per-particle step hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Default workload (BASELINE.json configs[1]): isotropic photon scattering in a uniform sphere,
16 Mi photons per GPU, A = n = 1e-3, dt = 1e-3 (pcoll ~ 0.2998 per step, reference
test/test_light.py:32-34), escape sphere R = 3 mean free paths, per-step sign tallies and the
escape-time histogram.  One "step" = one timestep of the whole pipeline (kinematics + scatter +
escape + tallies) over every live photon.  metric = particle-steps/s, whole job.

  value      state resident in HBM, stepped through Simulation.run_steps (public API); CUDA events
  e2e        same step through the host-buffer C-ABI entry point (pcl_photon_step_host): particle
             planes start and end in pinned HOST memory every step (H2D + kernel + D2H timed)
  roofline   fused photon-step kernel: algorithmic bytes / launch time
  cpu_baseline / --impl reference
             the reference's law in float64 (oracle/c/oracle.c: orc_photon_step_f64, OpenMP over
             all host cores) on a bounded sample of the same workload

Other workloads (--workload): kinematics_1m (configs[0], CUDA-graph stepped), kinematics_64m,
kinematics_ref_64m, wavelength_64m (configs[2]), gravity_256k (configs[3]), sweep_1b (configs[4]);
same JSON shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_LIGHT = 299792458.0
PHOTONS_PER_GPU = 16 * 2 ** 20
A_N = 1e-3 * 1e-3  # A * n of reference test/test_light.py:34
DT = 1e-3
R_ESCAPE = 3.0 / A_N  # three mean free paths
SEED = 2024


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("PCL_NO_CLOCKS"):  # diagnosis aid: run without the sampler
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 2.0:  # nvidia-smi needs a moment to start sampling
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def init_dist(n_gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    if world != n_gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d" % (n_gpus, world, n_gpus))
    return rank, world, local


def barrier_sync(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's law in float64, OpenMP over all host cores (oracle port)
# ---------------------------------------------------------------------------------------------
def cpu_photon_sphere(n, steps, warmup):
    if "oracle" not in sys.modules:  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import oracle

    st = {k: np.zeros(n) for k in ("x", "y", "z", "vx", "vy", "vz")}
    st["vx"][:] = C_LIGHT
    for s in range(warmup):
        oracle.photon_step_f64(st, DT, 1e-3, 1e-3, 0.0, C_LIGHT, 0, SEED, s, R_ESCAPE ** 2)
    live = 0
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        row = oracle.photon_step_f64(st, DT, 1e-3, 1e-3, 0.0, C_LIGHT, 0, SEED, s, R_ESCAPE ** 2)
        live += int(row[oracle.T_LIVE_IN])
    dt = time.perf_counter() - t0
    return live / dt, dt, oracle.num_threads()


def cpu_baseline_block(steps, warmup, budget_s=12.0):
    """Bounded sample: the same step window as the GPU arm over 4 Mi photons, repeated until about
    budget_s seconds of CPU work have been timed."""
    n = 4 * 2 ** 20
    live_total, t_total, reps, cores = 0.0, 0.0, 0, 1
    while t_total < budget_s and reps < 200:
        rate, dt, cores = cpu_photon_sphere(n, steps, warmup)
        live_total += rate * dt
        t_total += dt
        reps += 1
    return {"value": live_total / t_total, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": "%d photons x steps [%d, %d) of the photon_sphere workload, %d repetitions, float64 reference law "
                      "(oracle/c/oracle.c orc_photon_step_f64, OpenMP), %.1f s timed" % (n, warmup, warmup + steps, reps, t_total)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 4 * 2 ** 20
    rate, dt, cores = cpu_photon_sphere(n, args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": "particle-steps/s", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "photon_sphere_16m", "photons_per_step_sample": n,
                   "note": "reference law on host cores; pyopencl/pocl are not installable here, so this is the C/OpenMP "
                           "port of the reference kernels + kinematics (oracle/), the closest stand-in for OpenCL-on-CPU"},
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d photons x %d steps" % (n, args.steps)},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def photon_sim(n, rank, local, wavelength=False):
    import physicl_b200 as phys
    import physicl_b200.light
    import physicl_b200.newton

    sim = phys.Simulation(cl_on=True, device=local, seed=SEED, exit=lambda s: False)
    if os.environ.get("PCL_FEEDBACK_EVERY"):  # tuning aids
        sim.feedback_every = int(os.environ["PCL_FEEDBACK_EVERY"])
    if os.environ.get("PCL_COMPACT_CADENCE"):
        sim.compact_cadence = int(os.environ["PCL_COMPACT_CADENCE"])
    r = np.zeros((3, n), np.float32)
    v = np.zeros((3, n), np.float32)
    v[0] = C_LIGHT
    sim.add_particles(r, v, E=None, id_base=rank * n)
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(DT)))
    sim.add_step(1, phys.newton.NewtonianKinematicsStep())
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
    esc = phys.light.EscapeSphereStep(R_ESCAPE)
    sim.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    sim.add_step(4, sign)
    return sim, esc, sign


def bench_photon_sphere(args, rank, world, local):
    import torch

    from physicl_b200 import _capi

    n = PHOTONS_PER_GPU
    # throw-away run of the same pipeline on 1 Mi photons: first-use costs (lazy kernel loading, the allocator, NCCL
    # set-up, page-locked buffers) are paid here and not in a timed region that only lasts a few milliseconds
    prime, _, _ = photon_sim(1 << 20, rank, local)
    prime.run_steps(24)
    del prime
    sim, esc, sign = photon_sim(n, rank, local)
    ctx = sim.cl_ctx
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.device_store()  # upload the state now: the ranks must not reach the warm-up at different times
    with ClockSampler(local) as clocks:
        # all ranks start the warm-up together, so nobody idles (and drops its clocks) at the barrier in front of a
        # timed region that only lasts a few milliseconds
        barrier_sync(world)
        sim.run_steps(args.warmup)
        store = sim.store
        row0 = store.current_row + 1
        launches0 = ctx.launches
        barrier_sync(world)
        ev0.record()
        sim.run_steps(args.steps)
        ev1.record()
        barrier_sync(world)
    ms = max_over_ranks(ev0.elapsed_time(ev1), world)
    launches = ctx.launches - launches0
    rows = np.array([store.read_row(r) for r in range(row0, store.current_row + 1)])
    fused = rows[rows[:, _capi.T_LIVE_IN] > 0]
    live = float(fused[:, _capi.T_LIVE_IN].sum())
    scat = float(fused[:, _capi.T_SCATTERED].sum())
    live_all = sum_over_ranks(live, world)
    value = live_all / (ms * 1e-3)
    alg_bytes = 36.0 * live + 12.0 * scat  # (36 + 12 f) B per photon-step
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    hist = esc.escaped

    # ---- e2e: host buffers in, host buffers out, every step ---------------------------------
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        host = {k: torch.zeros(n, dtype=torch.float32).pin_memory() for k in ("x", "y", "z", "vx", "vy", "vz")}
        host["vx"].fill_(C_LIGHT)
        host["id"] = torch.arange(n, dtype=torch.int32).pin_memory()
        soa = _capi.Soa()
        for k, t in host.items():
            setattr(soa, k, t.data_ptr())
        soa.id_base = rank * n
        sp = _capi.ScatterParams(k=A_N, c=C_LIGHT, mode=0)
        pl = _capi.make_planes([])
        row = np.zeros(_capi.TALLY_COLS, np.int64)
        n_out = C.c_uint64(0)
        k_e2e = max(3, min(args.steps, 20))
        host_chunk = int(os.environ.get("PCL_HOST_CHUNK", 1 << 20))  # photons per pipelined chunk (tuning aid)
        state = {"n": n, "up": 0, "down": 0}

        def host_step(s):
            # photons live in pinned HOST planes between steps; survivors come back densely (remove_obj)
            soa.n = state["n"]
            rg = _capi.Rng(seed=SEED, step=s)
            ctx.call("pcl_photon_step_host_compact", C.byref(soa), C.c_float(DT), C.byref(sp), C.byref(rg),
                     C.c_float(R_ESCAPE ** 2), C.byref(pl), row.ctypes.data_as(C.c_void_p), C.c_uint64(host_chunk), C.byref(n_out))
            state["up"] += 28 * state["n"]
            state["down"] += 28 * n_out.value + 8 * _capi.TALLY_COLS
            state["n"] = n_out.value
            return int(row[_capi.T_LIVE_IN])

        for s in range(args.warmup):
            host_step(s)
        state["up"] = state["down"] = 0
        barrier_sync(world)
        t0 = time.perf_counter()
        live_e = 0
        for s in range(args.warmup, args.warmup + k_e2e):
            live_e += host_step(s)
        barrier_sync(world)
        wall = max_over_ranks(time.perf_counter() - t0, world)
        e2e = {"value": sum_over_ranks(float(live_e), world) / wall, "unit": "particle-steps/s",
               "h2d_bytes_per_step": state["up"] // k_e2e, "d2h_bytes_per_step": state["down"] // k_e2e, "steps": k_e2e,
               "timer": "host wall clock around the synchronous C-ABI call, max over ranks",
               "path": "pcl_photon_step_host_compact: pinned host SoA planes (r, v, id) -> chunked H2D -> fused "
                       "retire-and-compact kernel -> D2H of the survivors, 4 streams"}
    except Exception as e:  # report, never hide
        e2e = {"value": None, "unit": "particle-steps/s", "error": repr(e)}

    cpu = cpu_baseline_block(args.steps, args.warmup) if (rank == 0 and world == 1 and not args.no_cpu) else None
    out = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "photon_sphere_16m", "photons_per_gpu": n, "A": 1e-3, "n": 1e-3, "dt": DT,
                   "escape_radius": R_ESCAPE, "seed": SEED, "rng": "philox4x32-10 in-kernel",
                   "pipeline": "kinematics+scatter+escape+sign tally fused; a launch advances m timesteps with the photons "
                               "in registers and writes the survivors densely (adaptive m <= 8)",
                   "timesteps_per_launch": args.steps / max(int(launches), 1),
                   "l2": "state 384 MiB per GPU > 126 MB L2 (inputs larger than L2, no flush needed)",
                   "live_fraction_mean": live / (n * max(len(fused), 1)), "scattered_fraction": scat / max(live, 1),
                   "escaped_in_window": int(hist[-args.steps:].sum()) if len(hist) else 0},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": 815.8e6, "traffic_note": "dram__bytes_read+write of one pcl_k_photon_multi launch advancing 16 Mi "
                     "photons by 5 timesteps (profiles/r1_ncu_full_photon_multi.csv): 9.7 B per photon-step against 39.6 B "
                     "algorithmic, because the timesteps are fused in registers; the kernel is issue-bound (75 % of peak issue rate), "
                     "not DRAM-bound",
                     "peak_source": peak_src, "kernel": "pcl_k_photon_multi<0,0,0,0,1>",
                     "algorithmic_bytes": "(36 + 12 f) B per live photon-step, f = scattered fraction (SURVEY.md 8d), summed over the "
                                          "timesteps of the timed region and divided by its CUDA-event time",
                     "per_rank": True},
        "clocks": clocks.summary(),
    }
    if cpu:
        out["cpu_baseline"] = cpu
    return out


def bench_kinematics(args, rank, world, local, n, accel, graph):
    import torch

    from physicl_b200 import _capi
    from physicl_b200.store import DeviceParticleStore

    ctx = _capi.Context(local)
    rng = np.random.default_rng(1234 + rank)
    st = DeviceParticleStore(ctx)
    r = rng.uniform(-1e3, 1e3, (3, n)).astype(np.float32)
    v = rng.normal(0, 10, (3, n)).astype(np.float32)
    a = np.zeros((3, n), np.float32)
    a[2] = -9.81
    g = st.add_group("object", r, v, a=a if accel else None, id_base=rank * n)
    g.ensure("dx", "dy", "dz")
    soa = g.soa()
    steps, chunk = args.steps, 1000
    def run(k):
        if graph:  # "fused": k timesteps per HBM round trip (pcl_kinematics_steps)
            done = 0
            while done < k:
                m = min(chunk, k - done)
                ctx.call("pcl_kinematics_steps", st.stream(), C.byref(soa), C.c_float(DT), int(accel), None, C.c_uint32(m))
                done += m
        else:
            for _ in range(k):
                ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(DT), int(accel), None)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:  # started before the warm-up: no idle gap in front of the timed region
        barrier_sync(world)
        run(args.warmup)
        l0 = ctx.launches
        barrier_sync(world)
        ev0.record()
        run(steps)
        ev1.record()
        barrier_sync(world)
    ms = max_over_ranks(ev0.elapsed_time(ev1), world)
    bpp = 72.0 if accel else 48.0
    peak, peak_src = measured_peaks()
    launches = int(ctx.launches - l0)
    # one HBM round trip of the state per LAUNCH: with the timesteps fused in registers that is bpp bytes
    # per particle per launch, not per timestep
    achieved = bpp * n * (launches if graph else steps) / (ms * 1e-3) / 1e9
    return {
        "metric": "particle-steps/s", "value": n * world * steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "particles_per_gpu": n, "law": "v+=a dt; dr=v dt; r+=dr" if accel else "dr=v dt; r+=dr",
                   "timesteps_per_launch": (steps / max(launches, 1)) if graph else 1,
                   "l2": "state %d MB per GPU %s" % (n * (48 if accel else 36) // 10 ** 6,
                                                      "fits the 126 MB L2: HBM fraction is optimistic" if n * 48 < 120e6 else "> L2")},
        "e2e": None, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read + write of one launch over 64 Mi particles with a planes, ncu --set full
                     # (profiles/r1_ncu_full_kinematics_step.csv; the fused launch moves the same 4.78 GB per launch)
                     "traffic": 4.777e9 if (accel and n == 64 * 2 ** 20) else None,
                     "peak_source": peak_src, "kernel": "pcl_k_kinematics<1,1>" if accel else "pcl_k_kinematics<0,1>",
                     "algorithmic_bytes": ("%d B per particle per LAUNCH (timesteps fused in registers: one state round trip per "
                                           "launch; %d B per particle-step when stepped one launch per timestep)" % (bpp, bpp))
                     if graph else "%d B per particle-step" % bpp,
                     "one_round_trip_per_step_equivalent_gbs": bpp * n * steps / (ms * 1e-3) / 1e9},
        "clocks": clocks.summary(),
    }


def bench_gravity(args, rank, world, local):
    import torch

    import physicl_b200 as phys
    import physicl_b200.newton

    n_total = 262144
    rng = np.random.default_rng(7)
    # Plummer sphere, G = M = 1
    m_r = rng.uniform(0, 1, n_total)
    rad = 1.0 / np.sqrt(np.maximum(m_r ** (-2.0 / 3.0) - 1.0, 1e-12))
    d = rng.normal(size=(3, n_total))
    pos = rad * d / np.linalg.norm(d, axis=0)
    vel = rng.normal(0, 0.3, (3, n_total))
    sim = phys.Simulation(cl_on=True, device=local, shard=world > 1, exit=lambda s: False)
    sim.add_particles(pos, vel, kind="object")
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
    sim.add_step(1, phys.newton.NewtonianGravityStep(G=1.0, eps2=1e-4, masses=np.full(n_total, 1.0 / n_total, np.float32)))
    ctx = sim.cl_ctx
    fp32_peak = ctx.fp32_peak_tflops()
    sim.device_store()
    steps = args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:  # started before the warm-up: no idle gap in front of the timed region
        barrier_sync(world)
        sim.run_steps(max(3, args.warmup))
        l0 = ctx.launches
        barrier_sync(world)
        ev0.record()
        sim.run_steps(steps)
        ev1.record()
        barrier_sync(world)
    ms = max_over_ranks(ev0.elapsed_time(ev1), world)
    inter = float(n_total) * n_total * steps
    tf = 20.0 * inter / world / (ms * 1e-3) / 1e12  # per GPU
    return {
        "metric": "particle-steps/s", "value": n_total * steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gravity_256k", "bodies": n_total, "eps2": 1e-4, "dt": 1e-3, "interactions_per_s": inter / (ms * 1e-3)},
        "e2e": None, "gpu_launches": int(ctx.launches - l0),
        "roofline": {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak, "traffic": None,
                     "peak_source": "FFMA-only micro-kernel measured in this run (pcl_measure_fp32_peak)",
                     "kernel": "pcl_k_gravity_x2<2,128,512>", "algorithmic_flops": "20 FLOP per pairwise interaction", "per_rank": True},
        "clocks": clocks.summary(),
    }


def bench_wavelength(args, rank, world, local):
    """configs[2]: Rayleigh (lambda^-4) scattering of photons whose energies follow the reference's binned
    "Planck" law, sampled on the device; 64 Mi photons per GPU; no retirement (in-place WAVE kernel)."""
    import torch

    import physicl_b200 as phys
    import physicl_b200.light
    import physicl_b200.newton
    from physicl_b200 import _capi

    n = 64 * 2 ** 20
    sim = phys.Simulation(cl_on=True, device=local, seed=2025, exit=lambda s: False)
    ctx = sim.cl_ctx
    dev = torch.device("cuda", local)
    E_min = float(phys.light.E_from_wavelength(2500e-9))
    E_max = float(phys.light.E_from_wavelength(200e-9))
    e0ev, e1ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0ev.record()
    phys.light.planck_sample_device(ctx, 1024, E_min, E_max, 5778.0, bins=50000, seed=1, device=dev)  # warm-up (scratch, module load)
    tm = {}
    e, E0 = phys.light.planck_sample_device(ctx, n, E_min, E_max, 5778.0, bins=50000, seed=2025, id_base=rank * n, device=dev,
                                            timing=tm)
    e1ev.record()
    torch.cuda.synchronize()
    sample_ms = tm["device_ms"]
    r = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v[0].fill_(C_LIGHT)
    e = torch.nan_to_num(e.contiguous(), nan=0.5)  # the reference's "None" draws (mass of interval 0): mid-range energy
    sim.add_particles(r, v, E=e, id_base=rank * n)
    A, nd, dt = 5.1e-31 * (532e-9) ** 4, 2.5e25, 1e-5
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(dt)))
    sim.add_step(1, phys.newton.NewtonianKinematicsStep())
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(A), n=np.double(nd), wavelength_dep_scattering=True))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    sim.add_step(3, sign)
    sim.device_store().group("photon").e0 = E0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.device_store()
    with ClockSampler(local) as clocks:
        barrier_sync(world)
        sim.run_steps(args.warmup)
        store = sim.store
        row0 = store.current_row + 1
        l0 = ctx.launches
        barrier_sync(world)
        ev0.record()
        sim.run_steps(args.steps)
        ev1.record()
        barrier_sync(world)
    ms = max_over_ranks(ev0.elapsed_time(ev1), world)
    rows = np.array([store.read_row(q) for q in range(row0, store.current_row + 1)])
    live, scat = float(rows[:, _capi.T_LIVE_IN].sum()), float(rows[:, _capi.T_SCATTERED].sum())
    peak, peak_src = measured_peaks()
    achieved = (40.0 * live + 12.0 * scat) / (ms * 1e-3) / 1e9
    return {
        "metric": "particle-steps/s", "value": sum_over_ranks(live, world) / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wavelength_64m", "photons_per_gpu": n, "T": 5778.0, "bins": 50000, "A": A, "n": nd, "dt": dt,
                   "planck_sampling_ms": sample_ms, "planck_sampling_gphotons_per_s": n / sample_ms / 1e6,
                   "scattered_fraction": scat / max(live, 1), "l2": "state 1.75 GiB per GPU > L2"},
        "e2e": None, "gpu_launches": int(ctx.launches - l0),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "pcl_k_photon_multi<1,0,0,0,0> (8 timesteps per launch, in place)",
                     "algorithmic_bytes": "(36 + 12 f + 4) B per photon-step (SURVEY.md 8d, w = 1)", "per_rank": True},
        "clocks": clocks.summary(),
    }


def bench_sweep_1b(args, rank, world, local):
    """configs[4]: 2^30 particles in total, cut into contiguous blocks over the ranks (strong scaling):
    K kinematics steps (v += a dt; dr = v dt; r += dr, 72 B/particle-step) then K photon-sphere steps."""
    import torch

    import physicl_b200 as phys
    import physicl_b200.light
    import physicl_b200.newton
    from physicl_b200 import _capi
    from physicl_b200.store import DeviceParticleStore

    n_total = 2 ** 30
    n = n_total // world
    dev = torch.device("cuda", local)
    ctx = _capi.Context(local)
    # ---- kinematics -----------------------------------------------------------------------------
    st = DeviceParticleStore(ctx)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    r = torch.empty((3, n), dtype=torch.float32, device=dev).uniform_(-1e3, 1e3, generator=gen)
    v = torch.empty((3, n), dtype=torch.float32, device=dev).normal_(0, 10, generator=gen)
    a = torch.zeros((3, n), dtype=torch.float32, device=dev)
    a[2].fill_(-9.81)
    g = st.add_group("object", r, v, a=a, id_base=rank * n)
    g.ensure("dx", "dy", "dz")
    soa = g.soa()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:  # started before the warm-up: no idle gap in front of the timed region
        barrier_sync(world)
        for _ in range(args.warmup):
            ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(DT), 1, None)
        l0 = ctx.launches
        barrier_sync(world)
        ev0.record()
        for _ in range(args.steps):
            ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(DT), 1, None)
        ev1.record()
        barrier_sync(world)
    ms_kin = max_over_ranks(ev0.elapsed_time(ev1), world)
    launches = ctx.launches - l0
    del st, g, r, v, a, soa
    torch.cuda.empty_cache()
    # ---- photons ------------------------------------------------------------------------------
    sim = phys.Simulation(cl_on=True, device=local, seed=SEED, exit=lambda s: False)
    r = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v[0].fill_(C_LIGHT)
    sim.add_particles(r, v, id_base=rank * n)
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(DT)))
    sim.add_step(1, phys.newton.NewtonianKinematicsStep())
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
    sim.add_step(3, phys.light.EscapeSphereStep(R_ESCAPE))
    sim.add_step(4, phys.light.ScatterSignMeasureStep(None, True))
    sim.device_store()
    barrier_sync(world)
    sim.run_steps(args.warmup)
    store = sim.store
    row0 = store.current_row + 1
    l0 = sim.cl_ctx.launches
    barrier_sync(world)
    ev0.record()
    sim.run_steps(args.steps)
    ev1.record()
    barrier_sync(world)
    ms_ph = max_over_ranks(ev0.elapsed_time(ev1), world)
    launches += sim.cl_ctx.launches - l0
    rows = np.array([store.read_row(q) for q in range(row0, store.current_row + 1)])
    rows = rows[rows[:, _capi.T_LIVE_IN] > 0]
    live, scat = float(rows[:, _capi.T_LIVE_IN].sum()), float(rows[:, _capi.T_SCATTERED].sum())
    live_all = sum_over_ranks(live, world)
    kin_units = float(n_total) * args.steps
    peak, peak_src = measured_peaks()
    kin_gbs = 72.0 * n * args.steps / (ms_kin * 1e-3) / 1e9
    ph_gbs = (36.0 * live + 12.0 * scat) / (ms_ph * 1e-3) / 1e9
    return {
        "metric": "particle-steps/s", "value": (kin_units + live_all) / ((ms_kin + ms_ph) * 1e-3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (ms_kin + ms_ph) / (2 * args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (generated on device)",
        "config": {"workload": "sweep_1b", "particles_total": n_total, "particles_per_gpu": n,
                   "kinematics": {"particle_steps_per_s": kin_units / (ms_kin * 1e-3), "ms_per_step": ms_kin / args.steps,
                                  "hbm_gbs_per_gpu": kin_gbs, "frac": kin_gbs / peak},
                   "photon_sphere": {"particle_steps_per_s": live_all / (ms_ph * 1e-3), "ms_per_step": ms_ph / args.steps,
                                     "hbm_gbs_per_gpu": ph_gbs, "frac": ph_gbs / peak}},
        "e2e": None, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": kin_gbs, "peak": peak, "unit": "GB/s", "frac": kin_gbs / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "pcl_k_kinematics<1,true>", "algorithmic_bytes": "72 B per particle-step",
                     "per_rank": True},
        "clocks": clocks.summary(),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="photon_sphere_16m",
                    choices=["photon_sphere_16m", "kinematics_1m", "kinematics_64m", "kinematics_64m_fused", "kinematics_ref_64m", "gravity_256k",
                             "wavelength_64m", "sweep_1b"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: physicl_b200 has no CPU path (use --impl reference for the CPU arm)")
    rank, world, local = init_dist(args.gpus)
    if args.workload == "photon_sphere_16m":
        out = bench_photon_sphere(args, rank, world, local)
    elif args.workload == "kinematics_1m":
        out = bench_kinematics(args, rank, world, local, 1_000_000, True, True)
    elif args.workload == "kinematics_64m":
        out = bench_kinematics(args, rank, world, local, 64 * 2 ** 20, True, False)
    elif args.workload == "kinematics_64m_fused":
        out = bench_kinematics(args, rank, world, local, 64 * 2 ** 20, True, True)
    elif args.workload == "kinematics_ref_64m":
        out = bench_kinematics(args, rank, world, local, 64 * 2 ** 20, False, False)
    elif args.workload == "wavelength_64m":
        out = bench_wavelength(args, rank, world, local)
    elif args.workload == "sweep_1b":
        out = bench_sweep_1b(args, rank, world, local)
    else:
        out = bench_gravity(args, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
