#!/usr/bin/env python
"""Benchmark of the per-particle step hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--no-sub]

Default workload = BASELINE.json configs[4], the configuration the north star's scaling target is quoted
on: 2^30 particles in total, cut into contiguous global-index blocks over the ranks (strong scaling), two
legs of K timesteps each:

  kinematics     v += a dt; dr = v dt; r += dr, one launch per timestep (72 B / particle-step, HBM-bound)
  photon_sphere  the configs[1] law: kinematics + isotropic scatter + escape sphere + sign tallies fused,
                 A = n = 1e-3, dt = 1e-3 (pcoll ~ 0.2998, reference test/test_light.py:32-34), R = 3/(nA)

metric = particle-steps/s over both legs, whole job.  One "step" = one timestep of a leg over every live
particle.  At N = 1 (and, for gravity, at every N) the same line carries a `sub` block with the other
BASELINE configs, each with its own roofline: photon_sphere_16m (configs[1]), kinematics_64m (the figure the
0.70-of-HBM target refers to), kinematics_1m / kinematics_1m_fused (configs[0]: 1M particles x 1000 steps, one launch per
timestep / all timesteps in registers), wavelength_64m (configs[2]), gravity_256k (configs[3]).

  value      state resident in HBM; CUDA events around the K timesteps of each leg, queued behind a
             device-side gate (pcl_stream_gate) so that no host launch latency lies between the events;
             the photon legs are driven by sim.start() / sim.join() (the reference's entry point)
  e2e        both legs through the host-buffer C-ABI entry points with the particle planes in pinned HOST
             memory between calls (H2D + kernel + D2H inside the timed region, host wall clock), weighted
             as in the resident job; each leg's own figure is in e2e.kinematics / e2e.photon_sphere
  roofline   dominant kernel (kinematics): algorithmic bytes / launch time
  tally_checksum   hash of the all-rank sums of every tally row of the timed photon leg and of an integer
             checksum of the kinematics state: identical at N = 1, 2, 4, 8 (results do not depend on sharding)
  cpu_baseline / --impl reference
             the same two legs in float64 on the host cores (oracle/c/oracle.c, C + OpenMP; kind "port") on a
             bounded sample, plus the UNMODIFIED reference end to end and its own generated kernel text on
             pre-marshalled arrays (oracle/_ref + oracle/fake_pyopencl)
"""
from __future__ import annotations

import argparse
import ctypes as C
import datetime
import hashlib
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
A_N = 1e-3 * 1e-3  # A * n of reference test/test_light.py:34
DT = 1e-3
R_ESCAPE = 3.0 / A_N  # three mean free paths
SEED = 2024
SWEEP_TOTAL = 2 ** 30
CPU_SAMPLE = 16 * 2 ** 20  # particles of the sweep the CPU arm steps
E2E_BLOCK = 64 * 2 ** 20  # photons of each rank's block the host-buffer leg steps
ISSUE_PEAK = 148 * 4 * 1.965e9  # warp instructions / s: 148 SMs x 4 schedulers x max SM clock

# static description of the default workload: identical in the GPU arm and in the reference arm
SWEEP_CONFIG = {
    "workload": "sweep_1b", "particles_total": SWEEP_TOTAL,
    "legs": ["kinematics: v += a dt; dr = v dt; r += dr, one launch per timestep",
             "photon_sphere: kinematics + isotropic scatter + escape sphere + sign tallies"],
    "dt": DT, "A": 1e-3, "n": 1e-3, "escape_radius": R_ESCAPE, "seed": SEED,
    "init": "kinematics: r ~ U(-1e3,1e3)^3, v ~ N(0,10^2), a = (0,0,-9.81), hashed from the global particle id; "
            "photons: r = 0, v = (c,0,0) (reference test/test_light.py:12-17)",
    "sharding": "contiguous global-index blocks, no data-path collective; Philox counter = global particle id",
    "l2": "inputs larger than L2 (state >= 3 GB per GPU at 8 GPUs), no flush needed",
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks: ONE nvidia-smi process for the whole job (rank 0), 100 ms period, all GPUs of the job
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "timestamp,index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices, enabled=True):
        self.indices, self.rows, self.proc, self.windows = list(indices), [], None, []
        self.enabled = enabled and not os.environ.get("PCL_NO_CLOCKS")

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # nvidia-smi needs a moment to start sampling
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(2)

    def window(self, t0, t1):
        """A span (time.time()) in which the GPUs of the job were under load (warm-up + timed region of a leg)."""
        self.windows.append((t0, t1))

    def summary(self):
        if not self.enabled:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm, mx, reasons, used = [], [], set(), 0
        for r in self.rows:
            if len(r) < 8 or not r[2].replace(".", "").isdigit():
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                inside = any(a - 0.02 <= ts <= b + 0.02 for a, b in self.windows)
            except ValueError:
                inside = True
            if not inside:
                continue
            used += 1
            sm.append(float(r[2]))
            mx.append(float(r[3]))
            reasons.update(n for n, v in zip(names, r[4:8]) if v.lower().startswith("active"))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": used,
                "sampler": "one nvidia-smi process on rank 0, 100 ms period, GPUs %s, samples inside the warm-up + timed "
                           "windows of the legs" % self.indices}


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


def _reduce(x, world, op):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def max_over_ranks(x, world):
    return _reduce(x, world, "MAX")


def sum_over_ranks(x, world):
    return _reduce(x, world, "SUM")


def sum_rows_over_ranks(rows, world):
    """Exact int64 sum of tally rows over the ranks."""
    import torch

    rows = np.ascontiguousarray(rows, np.int64)
    if world == 1:
        return rows
    import torch.distributed as dist

    t = torch.from_numpy(rows.copy()).cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


class Gate:
    """Device-side gate in front of a timed region (pcl_stream_gate): the host queues event, launches, event behind
    a one-thread kernel that spins on a word of pinned host memory, then opens it.  The two events therefore bracket
    GPU work only; host launch latency (Python, ctypes, other ranks' threads) is outside the bracket.
    Everything queued behind a closed gate must have run once before (the warm-up does that): the first launch of a
    kernel loads its module, and module loading waits for running kernels, i.e. for the gate's time-out."""

    def __init__(self, ctx):
        import torch

        self.ctx = ctx
        self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.enabled = not os.environ.get("PCL_NO_GATE")

    def close(self, stream):
        if self.enabled:
            self.flag[0] = 0
            self.ctx.call("pcl_stream_gate", stream, C.c_void_p(self.flag.data_ptr()), C.c_uint32(4000))

    def open(self):
        self.flag[0] = 1


def timed_region(world, ctx, local, run, clocks=None, warm=None):
    """[warm-up], barrier, then the timed region behind the gate.  Returns (ms max over ranks, launches, ms of this rank)."""
    import torch

    dev = torch.device("cuda", local)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate = Gate(ctx)
    barrier_sync(world)
    w0 = time.time()
    if warm is not None:
        warm()
    l0 = ctx.launches
    barrier_sync(world)
    gate.close(stream)
    ev0.record()
    run()
    ev1.record()
    gate.open()
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.window(w0, time.time())
    ms_local = ev0.elapsed_time(ev1)
    launches = ctx.launches - l0
    barrier_sync(world)
    return max_over_ranks(ms_local, world), int(launches), ms_local


def checksum_hex(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, np.int64).tobytes())
    return h.hexdigest()[:16]


# ---------------------------------------------------------------------------------------------
# CPU arm: the two legs in float64, OpenMP over all host cores (oracle port), and the unmodified reference
# ---------------------------------------------------------------------------------------------
def _oracle():
    if "oracle" not in sys.modules:  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import oracle

    return oracle


def cpu_sweep(n, steps, warmup):
    """Both legs of the sweep over n particles in double: the reference's photon law (light.py:303-315, newton.py:14-16)
    and the constant-acceleration integrator.  Returns (particle-steps, seconds, cores)."""
    oracle = _oracle()
    rng = np.random.default_rng(1234)
    r = rng.uniform(-1e3, 1e3, (3, n))
    v = rng.normal(0, 10, (3, n))
    a = np.zeros((3, n))
    a[2] = -9.81
    dr = np.zeros((3, n))
    for _ in range(warmup):
        oracle.kinematics_accel_f64(r, v, a, dr, DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.kinematics_accel_f64(r, v, a, dr, DT)
    t_kin = time.perf_counter() - t0
    del r, v, a, dr
    st = {k: np.zeros(n) for k in ("x", "y", "z", "vx", "vy", "vz")}
    st["vx"][:] = C_LIGHT
    for s in range(warmup):
        oracle.photon_step_f64(st, DT, 1e-3, 1e-3, 0.0, C_LIGHT, 0, SEED, s, R_ESCAPE ** 2)
    live = 0
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        row = oracle.photon_step_f64(st, DT, 1e-3, 1e-3, 0.0, C_LIGHT, 0, SEED, s, R_ESCAPE ** 2)
        live += int(row[oracle.T_LIVE_IN])
    t_ph = time.perf_counter() - t0
    return {"units": float(n) * steps + live, "seconds": t_kin + t_ph, "cores": oracle.num_threads(),
            "kinematics_rate": float(n) * steps / t_kin, "photon_rate": live / t_ph}


def reference_python(n=10000, steps=10, timeout=240):
    """The unmodified reference end to end (oracle/run_reference.py), in a subprocess with a clean import state."""
    try:
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--n", str(n), "--steps", str(steps)],
                             capture_output=True, text=True, timeout=timeout, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            return {"unavailable": (out.stderr or out.stdout)[-300:]}
        d = json.loads(line[-1])
        if "unavailable" in d:
            return d
        return {"value": d["particle_steps_per_s"], "unit": "particle-steps/s", "cores": 1, "kind": "reference",
                "sample": "unmodified reference (oracle/_ref/physicl: its Simulation.run, CLProgram marshalling and kernel text; "
                          "pyopencl = oracle/fake_pyopencl, kernels compiled by gcc), photon pipeline of test/test_light.py:27-37, "
                          "%d photons x %d timesteps, %.1f s" % (d["n"], d["steps"], d["wall_s"])}
    except Exception as e:  # report, never hide
        return {"unavailable": repr(e)}


def reference_kernels(n=1_000_000, reps=10, timeout=240):
    """the reference's own GENERATED kernel text (light_scatter_step_sphere) compiled by
    gcc with an OpenMP work-item loop, on pre-marshalled float64 arrays, + the kinematics law as NumPy array arithmetic."""
    try:
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--kernels", str(n), "--steps", str(reps)],
                             capture_output=True, text=True, timeout=timeout, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            return {"unavailable": (out.stderr or out.stdout)[-300:]}
        d = json.loads(line[-1])
        if "unavailable" in d:
            return d
        return {"value": d["particle_steps_per_s"], "unit": "particle-steps/s", "cores": len(os.sched_getaffinity(0)), "kind": "reference",
                "scatter_kernel_particles_per_s": d["scatter_kernel_particles_per_s"],
                "kinematics_numpy_particles_per_s": d["kinematics_numpy_particles_per_s"],
                "sample": "the kernel text the reference generates for ScatterIsotropicStep (physicl/light.py:303-315 through "
                          "CLProgram.build_kernel), compiled by gcc with an OpenMP loop over the work-items (oracle/fake_pyopencl: "
                          "pyopencl / pocl are not installable), %d pre-marshalled float64 particles x %d launches, + newton.py:14-16 "
                          "as NumPy array arithmetic; no per-particle Python" % (d["n"], d["reps"])}
    except Exception as e:  # report, never hide
        return {"unavailable": repr(e)}


def dropin_objects(n=10000, steps=10):
    """The script cpu_baseline.reference_e2e times on the unmodified reference, here with only the import changed
    (scripts/dropin_objects.py): n PhotonObjects added one by one, sim.start(); sim.join(), every object current on the
    host again at the end.  Host wall clock; the Python object bridge, not the GPU, is what it measures."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import dropin_objects as script

        script.run(256, 2)  # context, module load
        best = min((script.run(n, steps) for _ in range(3)), key=lambda d: d["wall_s"])
        return {"value": best["particle_steps_per_s"], "unit": "particle-steps/s", "n": best["n"], "steps": best["steps"],
                "wall_s": best["wall_s"], "run_s": best["run_s"], "pull_s": best["pull_s"], "last_row": best["last_row"],
                "timer": "host wall clock from sim.start() until every PhotonObject is current on the host again (best of 3)",
                "path": "sim.add_obj x n -> device store built from the objects -> chunked fused launches -> one D2H of the "
                        "planes -> per-object attribute writes; compare with cpu_baseline.reference_e2e (same script, "
                        "unmodified reference)"}
    except Exception as e:  # report, never hide
        return {"unavailable": repr(e)}


def cpu_baseline_block(steps, warmup, budget_s=12.0, with_reference=True):
    """Bounded sample of the default workload: both legs over CPU_SAMPLE particles and the same step window,
    repeated until about budget_s seconds of CPU work have been timed."""
    units, secs, reps, last = 0.0, 0.0, 0, None
    while secs < budget_s and reps < 50:
        last = cpu_sweep(CPU_SAMPLE, steps, warmup)
        units += last["units"]
        secs += last["seconds"]
        reps += 1
    out = {"value": units / secs, "unit": "particle-steps/s", "cores": last["cores"], "kind": "port",
           "sample": "%d of the 2^30 particles, both legs, timesteps [%d, %d), %d repetition(s), float64 (oracle/c/oracle.c "
                     "orc_kinematics_accel_f64 + orc_photon_step_f64, OpenMP), %.1f s timed" % (CPU_SAMPLE, warmup, warmup + steps, reps, secs),
           "kinematics_particle_steps_per_s": last["kinematics_rate"], "photon_particle_steps_per_s": last["photon_rate"]}
    if with_reference:
        out["reference_e2e"] = reference_python()
        out["reference_kernels"] = reference_kernels()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_sweep(CPU_SAMPLE, args.steps, args.warmup)
    rate = r["units"] / r["seconds"]
    out = {
        "impl": "reference", "metric": "particle-steps/s", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(2 * args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(SWEEP_CONFIG),
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": r["cores"], "kind": "port",
                         "sample": "%d of the 2^30 particles (each step a bounded sample of the workload), both legs, timesteps [%d, %d), "
                                   "float64 C + OpenMP port of the reference's kernels and kinematics (oracle/c/oracle.c); pyopencl / pocl "
                                   "are not installed and not installable here" % (CPU_SAMPLE, args.warmup, args.warmup + args.steps),
                         "kinematics_particle_steps_per_s": r["kinematics_rate"], "photon_particle_steps_per_s": r["photon_rate"],
                         "reference_e2e": reference_python(), "reference_kernels": reference_kernels()},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
# GPU arm: building blocks
# ---------------------------------------------------------------------------------------------
def _i64(c):
    return c - (1 << 64) if c >= (1 << 63) else c


def hashed_uniform(ids, salt):
    """U[0,1) float32 from int64 global ids (a 64-bit mix; elementwise, so the value of a particle does not depend on
    which rank generates it)."""
    h = ids * _i64(0x9E3779B97F4A7C15) + _i64(salt * 0xD1B54A32D192ED03 & (2 ** 64 - 1))
    h = h ^ ((h >> 32) & 0xFFFFFFFF)
    h = h * _i64(0xD6E8FEB86659FD93)
    h = h ^ ((h >> 32) & 0xFFFFFFFF)
    h = h * _i64(0xD6E8FEB86659FD93)
    return ((h >> 40) & 0xFFFFFF).to(__import__("torch").float32) * (1.0 / 16777216.0)


def kinematics_state(n, id_base, dev):
    """Config-1 initial state for global ids [id_base, id_base + n): planes generated on the device, chunk by chunk."""
    import torch

    planes = {k: torch.empty(n, dtype=torch.float32, device=dev) for k in ("x", "y", "z", "vx", "vy", "vz")}
    chunk = 1 << 24
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ids = torch.arange(id_base + lo, id_base + hi, dtype=torch.int64, device=dev)
        for q, k in enumerate(("x", "y", "z")):
            planes[k][lo:hi] = hashed_uniform(ids, 1 + q) * 2000.0 - 1000.0
        for q, k in enumerate(("vx", "vy", "vz")):
            u1 = hashed_uniform(ids, 11 + q) + (0.5 / 16777216.0)
            u2 = hashed_uniform(ids, 21 + q)
            planes[k][lo:hi] = 10.0 * torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)
    return planes


def state_checksum(planes, n):
    """Sum of the int32 bit patterns of the position planes (exact integer arithmetic: order-independent)."""
    import torch

    tot = 0
    for k in ("x", "y", "z"):
        tot += int(planes[k][:n].view(torch.int32).sum(dtype=torch.int64).item())
    return tot


def photon_sim(n, rank, local, id_base=None):
    import physicl_b200 as phys
    import physicl_b200.light
    import physicl_b200.newton
    import torch

    sim = phys.Simulation(cl_on=True, device=local, seed=SEED, exit=lambda s: False)
    if os.environ.get("PCL_FEEDBACK_EVERY"):  # tuning aids
        sim.feedback_every = int(os.environ["PCL_FEEDBACK_EVERY"])
    if os.environ.get("PCL_COMPACT_CADENCE"):
        sim.compact_cadence = int(os.environ["PCL_COMPACT_CADENCE"])
    dev = torch.device("cuda", local)
    r = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v[0].fill_(C_LIGHT)
    sim.add_particles(r, v, E=None, id_base=rank * n if id_base is None else id_base)
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(DT)))
    sim.add_step(1, phys.newton.NewtonianKinematicsStep())
    # PCL_BENCH_SFU=1 (tuning aid, not the default): directions from MUFU sin / cos instead of the reproducible table
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3), sfu_trig=os.environ.get("PCL_BENCH_SFU") == "1"))
    esc = phys.light.EscapeSphereStep(R_ESCAPE)
    sim.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    sim.add_step(4, sign)
    return sim, esc, sign


def run_threaded(sim, steps):
    """The timed timesteps go through the reference's own entry point: sim.start() ... sim.join() with an exit predicate
    on the step count (physicl/__init__.py:501-524; test/test_light.py:36-37).  The warm-up before it used run_steps
    (a Simulation thread can be started once); the Philox step counter carries on."""
    sim.exit = lambda s: len(s.ts) >= steps
    sim.start()
    sim.join()


def photon_roofline(live, scat, ms, wave, kernel, steps_per_launch):
    """The fused photon kernel is bound by instruction issue, not by DRAM (ncu: profiles/README.md).  `achieved` is
    the warp-instruction issue rate implied by the measured instructions per photon-step of the shipped build."""
    from physicl_b200 import _capi

    peak, peak_src = measured_peaks()
    loop, launch = ((_capi.PHOTON_INSTR_LOOP_WAVE, _capi.PHOTON_INSTR_LAUNCH_WAVE) if wave
                    else (_capi.PHOTON_INSTR_LOOP, _capi.PHOTON_INSTR_LAUNCH))
    ipp = loop + launch / max(steps_per_launch, 1.0)
    alg = ((40.0 if wave else 36.0) * live + 12.0 * scat) / (ms * 1e-3) / 1e9
    rate = live / (ms * 1e-3)
    issued = rate * ipp / 32.0
    return {"bound": "issue", "achieved": issued / 1e9, "peak": ISSUE_PEAK / 1e9, "unit": "G warp-instructions/s",
            "frac": issued / ISSUE_PEAK, "traffic": None, "kernel": kernel, "per_rank": True,
            "instructions_per_photon_step": ipp,
            "instructions_source": "ncu --set full of the shipped kernel (profiles/README.md): %.1f per photon-step in the timestep loop + %.1f "
                                   "per photon and launch (load, write-back / compaction) / %.2f timesteps per launch; photons that retire "
                                   "inside a launch still occupy their lanes, so the true issue rate is a little higher" % (loop, launch, steps_per_launch),
            "hbm_equivalent_gbs": alg, "hbm_equivalent_frac": alg / peak,
            "hbm_equivalent_note": "(%d + 12 f) B per live photon-step (SURVEY.md 8d: one state round trip per timestep) / time / %s; "
                                   "above 1.0 means the fused kernel beats a perfect one-round-trip-per-step streaming kernel; the state "
                                   "actually crosses HBM once per launch of several timesteps" % (40 if wave else 36, peak_src)}


def leg_photon(args, rank, world, local, n, id_base, clocks, prime=True, strong=False):
    """K timesteps of the configs[1] photon pipeline over this rank's n photons."""
    from physicl_b200 import _capi

    if prime:
        # throw-away run of the same pipeline on 1 Mi photons: first-use costs (lazy kernel loading, the allocator,
        # page-locked feedback buffers) are paid here and not in the measured one
        p, _, _ = photon_sim(1 << 20, rank, local)
        p.run_steps(24)
        del p
    sim, esc, sign = photon_sim(n, rank, local, id_base=id_base)
    ctx = sim.cl_ctx
    sim.device_store()
    store = sim.store
    state = {}

    def warm():
        sim.run_steps(args.warmup)
        state["row0"] = store.current_row + 1

    ms, launches, _ = timed_region(world, ctx, local, lambda: run_threaded(sim, args.steps), clocks, warm)
    rows = np.array([store.read_row(r) for r in range(state["row0"], store.current_row + 1)])
    assert rows.shape[0] == args.steps, rows.shape
    live = float(rows[:, _capi.T_LIVE_IN].sum())
    scat = float(rows[:, _capi.T_SCATTERED].sum())
    rows_all = sum_rows_over_ranks(rows, world)
    live_all = float(rows_all[:, _capi.T_LIVE_IN].sum())
    out = {"value": live_all / (ms * 1e-3), "ms_per_step": ms / args.steps, "gpu_launches": launches,
           "photons_per_gpu": n, "timesteps_per_launch": args.steps / max(launches, 1),
           "live_fraction_mean": live / (float(n) * args.steps), "scattered_fraction": scat / max(live, 1.0),
           "escaped_in_window": int(rows_all[:, _capi.T_ESCAPED].sum()),
           "tally_checksum": checksum_hex(rows_all),
           "roofline": photon_roofline(live, scat, ms, False, "pcl_k_photon_multi<0,0,0,0,1> (m <= 8 timesteps per launch, survivors compacted)",
                                       args.steps / max(launches, 1))}
    return out, rows_all, ms


def leg_kinematics(args, rank, world, local, n, id_base, clocks, accel=True):
    """K timesteps of the kinematics law, one launch per timestep, over this rank's n particles."""
    import torch

    from physicl_b200 import _capi
    from physicl_b200.store import DeviceParticleStore

    dev = torch.device("cuda", local)
    ctx = _capi.Context(local)
    st = DeviceParticleStore(ctx)
    g = _adopt_group(st, "object", kinematics_state(n, id_base, dev), n, id_base)  # the planes as generated: no copies
    if accel:
        g.alloc("ax", 0.0)
        g.alloc("ay", 0.0)
        g.alloc("az", -9.81)
    g.ensure("dx", "dy", "dz")
    soa = g.soa()

    def run(k):
        for _ in range(k):
            ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(DT), int(accel), None)

    ms, launches, _ = timed_region(world, ctx, local, lambda: run(args.steps), clocks, lambda: run(args.warmup))
    bpp = 72.0 if accel else 48.0
    peak, peak_src = measured_peaks()
    gbs = bpp * n * args.steps / (ms * 1e-3) / 1e9
    chk = int(sum_rows_over_ranks(np.array([state_checksum(g.planes, n)]), world)[0])
    # dram__bytes_read + write of ONE launch of this kernel by ncu, where a capture at this size exists:
    # 64 Mi particles: profiles/r1_ncu_full_kinematics_step.csv; 2^30: profiles/r2/ncu_dram_kinematics_1b.csv (38.66 GB read +
    # 38.60 GB written against 77.31 GB algorithmic: no re-reads)
    traffic = {64 * 2 ** 20: 4.777e9, 2 ** 30: 77.258e9}.get(n) if accel else None
    out = {"value": float(n) * world * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps, "gpu_launches": launches,
           "particles_per_gpu": n, "law": "v+=a dt; dr=v dt; r+=dr" if accel else "dr=v dt; r+=dr (reference newton.py:14-16)",
           "state_checksum": chk,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
                        "peak_source": peak_src, "kernel": "pcl_k_kinematics<%d,1>" % int(accel),
                        "algorithmic_bytes": "%d B per particle-step x %d particles per launch" % (bpp, n), "per_rank": True}}
    del st, g, soa
    torch.cuda.empty_cache()
    return out, chk, ms


def _adopt_group(st, kind, planes, n, id_base):
    from physicl_b200.store import Group

    g = Group(kind, st.device, n, id_base)
    for k, t in planes.items():
        g.upload(k, t)
    st.groups[kind] = g
    return g


E2E_KIN_BLOCK = 16 * 2 ** 20  # particles of each rank's block the host-buffer kinematics leg steps (12 pinned planes)


def leg_e2e_kinematics(args, ctx, world, n, m, k, host_chunk):
    """Kinematics leg through pcl_kinematics_steps_host: r, v, a, dr planes in pinned HOST memory between calls, m timesteps
    per round trip (36 B up + 36 B down per particle and round trip), the law of the resident leg (v += a dt; dr = v dt; r += dr)."""
    import torch

    from physicl_b200 import _capi

    g = torch.Generator().manual_seed(1234)
    host = {}
    for nm in ("x", "y", "z"):
        host[nm] = torch.empty(n, dtype=torch.float32, pin_memory=True).uniform_(-1e3, 1e3, generator=g)
    for nm in ("vx", "vy", "vz"):
        host[nm] = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_(0.0, 10.0, generator=g)
    for nm, val in (("ax", 0.0), ("ay", 0.0), ("az", -9.81), ("dx", 0.0), ("dy", 0.0), ("dz", 0.0)):
        host[nm] = torch.full((n,), val, dtype=torch.float32).pin_memory()
    soa = _capi.Soa()
    soa.n = n
    for nm, t in host.items():
        setattr(soa, nm, t.data_ptr())

    def call(steps):
        ctx.call("pcl_kinematics_steps_host", C.byref(soa), C.c_float(DT), 1, None, C.c_uint32(steps), C.c_uint64(host_chunk))

    s = 0
    while s < args.warmup:
        run = min(m, args.warmup - s)
        call(run)
        s += run
    trips = 0
    barrier_sync(world)
    t0 = time.perf_counter()
    while s < args.warmup + k:
        run = min(m, args.warmup + k - s)
        call(run)
        s += run
        trips += 1
    wall = time.perf_counter() - t0
    barrier_sync(world)
    wall = max_over_ranks(wall, world)
    up = down = 36 * n * trips
    return {"value": float(n) * world * k / wall, "unit": "particle-steps/s", "particles_per_gpu": n, "steps": k,
            "timesteps_per_round_trip": m, "h2d_bytes_per_step": up // k, "d2h_bytes_per_step": down // k,
            "pcie_per_gpu": {"h2d_GBps": up / wall / 1e9, "d2h_GBps": down / wall / 1e9},
            "path": "pcl_kinematics_steps_host: pinned host planes r, v, a (36 B up) -> 2 Mi-particle chunks -> one launch advancing "
                    "the timesteps of the round trip in registers -> r, v, dr back (36 B down), 4 streams"}


def leg_e2e(args, rank, world, local, n, id_base):
    """Photon leg through the host-buffer entry points: the planes (r, v, id) live in pinned HOST memory between calls.
    Primary: pcl_photon_steps_host_compact, m = 8 timesteps per host round trip (what Simulation.run needs when no host step
    sits between the device steps).  Also: one timestep per round trip (pcl_photon_step_host_compact)."""
    import torch

    from physicl_b200 import _capi

    ctx = _capi.Context(local)
    host = {k: torch.zeros(n, dtype=torch.float32, pin_memory=True) for k in ("x", "y", "z", "vx", "vy", "vz")}
    host["id"] = torch.empty(n, dtype=torch.int32, pin_memory=True)
    soa = _capi.Soa()
    for k, t in host.items():
        setattr(soa, k, t.data_ptr())
    soa.id_base = id_base
    sp = _capi.ScatterParams(k=A_N, c=C_LIGHT, mode=0)
    pl = _capi.make_planes([])
    rows = np.zeros((8, _capi.TALLY_COLS), np.int64)
    n_out = C.c_uint64(0)
    host_chunk = int(os.environ.get("PCL_HOST_CHUNK", 1 << 21))  # particles per pipelined chunk (2 Mi measured best: 1 Mi -3 %, 4 Mi -3 %)

    def reset():
        for k in ("x", "y", "z", "vy", "vz"):
            host[k].zero_()
        host["vx"].fill_(C_LIGHT)
        torch.arange(n, dtype=torch.int32, out=host["id"])
        return {"n": n, "up": 0, "down": 0, "live": 0}

    def call(state, s, m):
        soa.n = state["n"]
        rg = _capi.Rng(seed=SEED, step=s)
        ctx.call("pcl_photon_steps_host_compact", C.byref(soa), C.c_float(DT), C.byref(sp), C.byref(rg), C.c_float(R_ESCAPE ** 2),
                 C.byref(pl), rows.ctypes.data_as(C.c_void_p), C.c_uint64(host_chunk), C.c_uint32(m), C.byref(n_out))
        state["up"] += 28 * state["n"]
        state["down"] += 28 * n_out.value + 8 * _capi.TALLY_COLS * m
        state["n"] = n_out.value
        state["live"] += int(rows[:m, _capi.T_LIVE_IN].sum())

    wall_of = {}

    def measure(m, k):
        state = reset()
        s = 0
        while s < args.warmup:  # same step window as the resident leg: timesteps [warmup, warmup + k)
            run = min(m, args.warmup - s)
            call(state, s, run)
            s += run
        state.update(up=0, down=0, live=0)
        barrier_sync(world)
        t0 = time.perf_counter()
        while s < args.warmup + k:
            run = min(m, args.warmup + k - s)
            call(state, s, run)
            s += run
        wall = time.perf_counter() - t0
        barrier_sync(world)
        wall = max_over_ranks(wall, world)
        wall_of[m] = wall
        # PCIe rate of this rank's link, both directions at once (the measured bidirectional ceiling is in profiles/bench_r1/pcie_peak.txt)
        link = {"h2d_GBps": state["up"] / wall / 1e9, "d2h_GBps": state["down"] / wall / 1e9,
                # what the host's memory system sustains for all ranks' DMA together (reads + writes of pinned memory): this, not
                # the link, bounds the N-GPU figure on a box whose GPUs hang off one NUMA node
                "host_dram_GBps_all_gpus": sum_over_ranks(float(state["up"] + state["down"]), world) / wall / 1e9}
        return sum_over_ranks(float(state["live"]), world) / wall, state["up"] // k, state["down"] // k, link

    v8, up8, down8, link8 = measure(8, args.steps)
    k1 = max(3, min(args.steps, 6))
    v1, up1, down1, link1 = measure(1, k1)
    del host
    kin8 = leg_e2e_kinematics(args, ctx, world, min(n, E2E_KIN_BLOCK), 8, args.steps, host_chunk)
    kin1 = leg_e2e_kinematics(args, ctx, world, min(n, E2E_KIN_BLOCK), 1, k1, host_chunk)

    def both(r_kin, r_ph, live_frac):
        # whole job through host buffers: equal particle counts in both legs, as in the resident job; per particle and
        # timestep the kinematics leg does 1 unit, the photon leg `live_frac` units: total units / total time
        return (1.0 + live_frac) / (1.0 / r_kin + live_frac / r_ph)

    live8 = v8 * wall_of[8] / (float(n) * world * args.steps)
    live1 = v1 * wall_of[1] / (float(n) * world * k1)
    return {"value": both(kin8["value"], v8, live8), "unit": "particle-steps/s",
            "h2d_bytes_per_step": up8 + kin8["h2d_bytes_per_step"], "d2h_bytes_per_step": down8 + kin8["d2h_bytes_per_step"],
            "steps": args.steps, "legs": "kinematics + photon sphere, weighted as in the resident job (equal particle counts; the "
                                         "photon leg counts live photon-steps): (1 + f) / (1 / r_kin + f / r_photon), f = %.4f" % live8,
            "photon_sphere": {"value": v8, "h2d_bytes_per_step": up8, "d2h_bytes_per_step": down8},
            "pcie_per_gpu": link8, "kinematics": kin8,
            "timesteps_per_round_trip": 8, "photons_per_gpu": n,
            "timer": "host wall clock around the synchronous C-ABI calls, max over ranks",
            "path": "pcl_photon_steps_host_compact: pinned host SoA planes (r, v, id) -> 2 Mi-photon chunks H2D -> ONE fused launch "
                    "advancing 8 timesteps and compacting -> D2H of the survivors, 4 streams; the particles are back in host memory "
                    "after every call (every 8th timestep), which is all Simulation.run needs when no host step sits between the "
                    "device steps",
            "sample": "the first %d photons of each rank's block (the path streams 2 Mi-photon chunks: its throughput does not depend "
                      "on the block size)" % n,
            "one_timestep_per_round_trip": {"value": both(kin1["value"], v1, live1), "photon_sphere": {"value": v1},
                                            "kinematics": kin1,
                                            "h2d_bytes_per_step": up1 + kin1["h2d_bytes_per_step"],
                                            "d2h_bytes_per_step": down1 + kin1["d2h_bytes_per_step"], "steps": k1,
                                            "pcie_per_gpu": link1,
                                            "path": "pcl_photon_step_host_compact: the same, planes back in host memory after EVERY timestep "
                                                    "(28 B up + 28 B down per photon-step over PCIe)"}}


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def bench_sweep_1b(args, rank, world, local, clocks):
    n = SWEEP_TOTAL // world
    id_base = rank * n
    kin, kin_chk, ms_kin = leg_kinematics(args, rank, world, local, n, id_base, clocks)
    ph, rows_all, ms_ph = leg_photon(args, rank, world, local, n, id_base, clocks)
    kin_units = float(SWEEP_TOTAL) * args.steps
    from physicl_b200 import _capi

    live_all = float(rows_all[:, _capi.T_LIVE_IN].sum())
    e2e = None
    if not args.no_e2e:
        try:
            e2e = leg_e2e(args, rank, world, local, min(E2E_BLOCK, n), id_base)
        except Exception as e:  # report, never hide
            e2e = {"value": None, "unit": "particle-steps/s", "error": repr(e)}
    out = {
        "metric": "particle-steps/s", "value": (kin_units + live_all) / ((ms_kin + ms_ph) * 1e-3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (ms_kin + ms_ph) / (2 * args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (generated on the device)",
        "config": dict(SWEEP_CONFIG),
        "detail": {"particles_per_gpu": n, "kinematics": {k: v for k, v in kin.items() if k != "roofline"},
                   "photon_sphere": {k: v for k, v in ph.items() if k != "roofline"},
                   "timed_region": "CUDA events on the launching stream, queued behind a device-side gate (pcl_stream_gate); max over ranks"},
        "e2e": e2e, "gpu_launches": kin["gpu_launches"] + ph["gpu_launches"],
        "roofline": kin["roofline"], "roofline_photon": ph["roofline"],
        "tally_checksum": checksum_hex(rows_all, [kin_chk]),
    }
    return out


def bench_photon_sphere(args, rank, world, local, clocks):
    n = 16 * 2 ** 20
    ph, rows_all, ms = leg_photon(args, rank, world, local, n, rank * n, clocks)
    return {"metric": "particle-steps/s", "value": ph["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ph["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "photon_sphere_16m", "photons_per_gpu": n, "A": 1e-3, "n": 1e-3, "dt": DT, "escape_radius": R_ESCAPE,
                       "seed": SEED, "l2": "state 384 MiB per GPU > 126 MB L2"},
            "detail": {k: v for k, v in ph.items() if k != "roofline"}, "e2e": None, "gpu_launches": ph["gpu_launches"],
            "roofline": ph["roofline"], "tally_checksum": ph["tally_checksum"]}


def bench_kinematics(args, rank, world, local, clocks, n, accel, fused):
    import torch

    from physicl_b200 import _capi
    from physicl_b200.store import DeviceParticleStore

    if not fused:
        kin, _, ms = leg_kinematics(args, rank, world, local, n, rank * n, clocks, accel=accel)
        if n * 48 <= 120e6:  # the state never leaves the L2: the figure below is an L2 rate, not an HBM one
            kin["roofline"]["l2_resident"] = True
        return {"metric": "particle-steps/s", "value": kin["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": kin["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "particles_per_gpu": n, "law": kin["law"], "timesteps_per_launch": 1,
                           "l2": ("state %d MB per GPU > L2" % (n * (48 if accel else 36) // 10 ** 6)) if n * 48 > 120e6 else
                                 ("state %d MB per GPU fits the 126 MB L2: the HBM fraction is optimistic (SURVEY 7, hard part 4); "
                                  "take the HBM figure from kinematics_64m" % (n * (48 if accel else 36) // 10 ** 6))},
                "e2e": None, "gpu_launches": kin["gpu_launches"], "roofline": kin["roofline"]}
    # timesteps fused in registers (pcl_kinematics_steps): FP32-pipe bound, configs[0]
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
        done = 0
        while done < k:
            m = min(chunk, k - done)
            ctx.call("pcl_kinematics_steps", st.stream(), C.byref(soa), C.c_float(DT), int(accel), None, C.c_uint32(m))
            done += m

    ms, launches, _ = timed_region(world, ctx, local, lambda: run(steps), clocks, lambda: run(args.warmup))
    bpp = 72.0 if accel else 48.0
    peak, peak_src = measured_peaks()
    achieved = bpp * n * launches / (ms * 1e-3) / 1e9
    return {
        "metric": "particle-steps/s", "value": n * world * steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "particles_per_gpu": n, "law": "v+=a dt; dr=v dt; r+=dr" if accel else "dr=v dt; r+=dr",
                   "timesteps_per_launch": steps / max(launches, 1),
                   "l2": "state %d MB per GPU %s" % (n * (48 if accel else 36) // 10 ** 6,
                                                      "fits the 126 MB L2: HBM fraction is optimistic" if n * 48 < 120e6 else "> L2")},
        "e2e": None, "gpu_launches": launches,
        "roofline": {"bound": "fp32 pipe (timesteps fused in registers; 9 dependent FADD/FMUL per particle-step)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "pcl_k_kinematics<%d,1>" % int(accel),
                     "algorithmic_bytes": "%d B per particle per LAUNCH (one state round trip per launch)" % bpp,
                     "one_round_trip_per_step_equivalent_gbs": bpp * n * steps / (ms * 1e-3) / 1e9},
    }


def bench_gravity(args, rank, world, local, clocks, steps=None):
    import physicl_b200 as phys
    import physicl_b200.newton

    n_total = 262144
    steps = args.steps if steps is None else steps
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
    ms, launches, _ = timed_region(world, ctx, local, lambda: sim.run_steps(steps), clocks, lambda: sim.run_steps(max(3, args.warmup)))
    inter = float(n_total) * n_total * steps
    tf = 20.0 * inter / world / (ms * 1e-3) / 1e12  # per GPU
    # cross-N check value: float sums depend on the split, so this is a tolerance figure, not a bit pattern
    snap = sim.store.snapshot("object", live_only=False)
    com = [sum_over_ranks(float(np.sum(snap[k].astype(np.float64))), world) / n_total for k in ("x", "y", "z")]
    return {
        "metric": "particle-steps/s", "value": n_total * steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gravity_256k", "bodies": n_total, "eps2": 1e-4, "dt": 1e-3, "interactions_per_s": inter / (ms * 1e-3),
                   "exchange": _gravity_exchange_label(sim, world),
                   "centre_of_mass_after_run": com},
        "e2e": None, "gpu_launches": launches,
        "roofline": {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak, "traffic": None,
                     "peak_source": "FFMA-only micro-kernel measured in this run (pcl_measure_fp32_peak); nominal 148 x 128 x 2 x 1.965 GHz = 74.4",
                     "kernel": _gravity_kernel_name(), "algorithmic_flops": "20 FLOP per pairwise interaction", "per_rank": True},
    }


def _gravity_exchange_label(sim, world):
    if world == 1:
        return "none (1 GPU)"
    x = type((sim.steps[1]._state or {}).get("xchg")).__name__
    if x == "GravityExchangeP2P":
        return ("GravityExchangeP2P: the kick-drift kernel stores every updated body into all ranks' gathered arrays (peer-mapped "
                "symmetric memory, NVLink stores), one signal-pad barrier per step, one acceleration launch over local HBM")
    return "GravityExchange: NCCL all-gather of the packed (x,y,z,m) blocks per step, overlapped with the local-block tile loop"


def _gravity_kernel_name():
    from physicl_b200 import _capi

    return getattr(_capi, "GRAVITY_KERNEL", "pcl_k_gravity_x2<2,128,512>")


def bench_wavelength(args, rank, world, local, clocks):
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
    phys.light.planck_sample_device(ctx, 1024, E_min, E_max, 5778.0, bins=50000, seed=1, device=dev)  # warm-up (scratch, module load)
    tm = {}
    e, E0 = phys.light.planck_sample_device(ctx, n, E_min, E_max, 5778.0, bins=50000, seed=2025, id_base=rank * n, device=dev,
                                            timing=tm)
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
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(A), n=np.double(nd), wavelength_dep_scattering=True,
                                                    sfu_trig=os.environ.get("PCL_BENCH_SFU") == "1"))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    sim.add_step(3, sign)
    sim.device_store().group("photon").e0 = E0
    store = sim.store
    state = {}

    def warm():
        sim.run_steps(args.warmup)
        state["row0"] = store.current_row + 1

    ms, launches, _ = timed_region(world, ctx, local, lambda: run_threaded(sim, args.steps), clocks, warm)
    rows = np.array([store.read_row(q) for q in range(state["row0"], store.current_row + 1)])
    live, scat = float(rows[:, _capi.T_LIVE_IN].sum()), float(rows[:, _capi.T_SCATTERED].sum())
    rows_all = sum_rows_over_ranks(rows, world)
    return {
        "metric": "particle-steps/s", "value": float(rows_all[:, _capi.T_LIVE_IN].sum()) / (ms * 1e-3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wavelength_64m", "photons_per_gpu": n, "T": 5778.0, "bins": 50000, "A": A, "n": nd, "dt": dt,
                   "planck_sampling_ms": sample_ms, "planck_sampling_gphotons_per_s": n / sample_ms / 1e6,
                   "scattered_fraction": scat / max(live, 1), "l2": "state 1.75 GiB per GPU > L2"},
        "e2e": None, "gpu_launches": launches,
        "roofline": photon_roofline(live, scat, ms, True, "pcl_k_photon_multi<1,0,0,0,0> (8 timesteps per launch, in place)",
                                    args.steps / max(launches, 1)),
        "tally_checksum": checksum_hex(rows_all),
    }


def compact_sub(d):
    """The part of a workload's line that goes into the default line's `sub` block."""
    keep = ("value", "unit", "ms_per_step", "steps", "scaling", "gpu_launches", "roofline", "tally_checksum")
    out = {k: d[k] for k in keep if k in d}
    out["config"] = d["config"]
    return out


WORKLOADS = ["sweep_1b", "photon_sphere_16m", "kinematics_1m", "kinematics_1m_fused", "kinematics_64m", "kinematics_64m_fused",
             "kinematics_ref_64m", "gravity_256k", "wavelength_64m"]


def run_workload(name, args, rank, world, local, clocks):
    if name == "sweep_1b":
        return bench_sweep_1b(args, rank, world, local, clocks)
    if name == "photon_sphere_16m":
        return bench_photon_sphere(args, rank, world, local, clocks)
    if name == "kinematics_1m":  # configs[0], one launch per timestep (the 72 MB working set lives in L2)
        return bench_kinematics(args, rank, world, local, clocks, 1_000_000, True, False)
    if name == "kinematics_1m_fused":  # configs[0], the 1000 timesteps applied in registers by one launch
        return bench_kinematics(args, rank, world, local, clocks, 1_000_000, True, True)
    if name == "kinematics_64m":
        return bench_kinematics(args, rank, world, local, clocks, 64 * 2 ** 20, True, False)
    if name == "kinematics_64m_fused":
        return bench_kinematics(args, rank, world, local, clocks, 64 * 2 ** 20, True, True)
    if name == "kinematics_ref_64m":
        return bench_kinematics(args, rank, world, local, clocks, 64 * 2 ** 20, False, False)
    if name == "wavelength_64m":
        return bench_wavelength(args, rank, world, local, clocks)
    if name == "gravity_256k":
        return bench_gravity(args, rank, world, local, clocks)
    raise SystemExit("unknown workload " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep_1b", choices=WORKLOADS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-results of the default line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: physicl_b200 has no CPU path (use --impl reference for the CPU arm)")
    rank, world, local = init_dist(args.gpus)
    import gc

    with ClockSampler(range(world), enabled=rank == 0) as clocks:
        out = run_workload(args.workload, args, rank, world, local, clocks)
        if args.workload == "sweep_1b" and not args.no_sub:
            # the other BASELINE configs, each with its own roofline (weak-scaling ones at 1 GPU only; gravity, the one
            # workload with a data-path collective, at every N)
            subs = (["photon_sphere_16m", "kinematics_64m", "kinematics_1m", "kinematics_1m_fused", "wavelength_64m", "gravity_256k"]
                    if world == 1 else ["gravity_256k"])
            out["sub"] = {}
            for name in subs:
                gc.collect()
                torch.cuda.empty_cache()
                sub_args = argparse.Namespace(**vars(args))
                sub_args.workload = name
                if name == "gravity_256k":
                    sub_args.steps = min(args.steps, 10)
                if name.startswith("kinematics_1m"):
                    sub_args.steps = 1000  # BASELINE configs[0]: 1M particles x 1000 steps
                if world > 1:  # a rank that swallowed an error would leave the others waiting in a collective
                    out["sub"][name] = compact_sub(run_workload(name, sub_args, rank, world, local, clocks))
                    continue
                try:  # a sub-result that fails (e.g. no memory left on a shared box) must not cost the headline line
                    out["sub"][name] = compact_sub(run_workload(name, sub_args, rank, world, local, clocks))
                except Exception as e:  # reported in the line, never hidden
                    out["sub"][name] = {"error": repr(e)}
                    torch.cuda.synchronize()
    out["clocks"] = clocks.summary() if rank == 0 else None
    if rank == 0 and world == 1 and not args.no_cpu and args.workload == "sweep_1b":
        out["cpu_baseline"] = cpu_baseline_block(args.steps, args.warmup)
        out["dropin_objects"] = dropin_objects()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
