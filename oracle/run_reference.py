#!/usr/bin/env python
"""Time the UNMODIFIED reference end to end (its ``Simulation.run``, its Python marshalling, its own
OpenCL-C kernel text) on the host CPU.  TEST INFRASTRUCTURE: only ``bench.py``'s CPU legs run this.

The reference package comes from ``oracle/_ref`` (see ``oracle/make_ref.py``) or, in the build container,
from ``/root/reference``; ``pyopencl`` is ``oracle/fake_pyopencl`` (the reference's kernel strings compiled
with gcc; pyopencl/pocl are not installed in this image).  Pipeline = reference test/test_light.py:27-37:
UpdateTimeStep(dt=1e-3) + NewtonianKinematicsStep + ScatterIsotropicStep(A=n=1e-3) + ScatterSignMeasureStep
over N photons emitted at the origin along +x, run for ``--steps`` timesteps.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cl-off", action="store_true", help="the reference's pure-Python path (cl_on=False)")
    ap.add_argument("--kernels", type=int, default=0, metavar="N",
                    help="kernel-level baseline instead: the reference's own generated scatter kernel on N pre-marshalled particles")
    args = ap.parse_args()
    import numpy as np

    warnings.filterwarnings("ignore")
    np.int = np.int32  # the reference writes dtype=np.int (physicl/__init__.py:653); removed in NumPy 1.24
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "physicl")):
        ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "physicl")):
        print(json.dumps({"unavailable": "no copy of the reference (oracle/_ref missing: run oracle/make_ref.py in the build container)"}))
        return
    sys.path.insert(0, os.path.join(HERE, "fake_pyopencl"))
    sys.path.insert(0, ref)
    import physicl
    import physicl.light
    import physicl.newton

    if args.kernels:
        kernel_level(physicl, np, args.kernels, args.steps)
        return
    steps = args.steps
    sim = physicl.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=not args.cl_off, exit=lambda c: len(c.ts) >= steps)
    for _ in range(args.n):
        sim.add_obj(physicl.light.PhotonObject(s=np.array([0] * 3, dtype=np.double),
                                               v=np.array([physicl.light.c, 0, 0], dtype=np.double), E=np.double(1)))
    sim.add_step(0, physicl.UpdateTimeStep(lambda s: np.double(0.001)))
    sim.add_step(1, physicl.newton.NewtonianKinematicsStep())
    sim.add_step(2, physicl.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    sign = physicl.light.ScatterSignMeasureStep(None, True)
    sim.add_step(3, sign)
    np.random.seed(2024)
    t0 = time.perf_counter()
    sim.start()
    sim.join()
    wall = time.perf_counter() - t0
    rows = len(sign.data)
    print(json.dumps({"particle_steps_per_s": args.n * rows / wall, "wall_s": wall, "n": args.n, "steps": rows,
                      "cl_on": not args.cl_off, "package": os.path.dirname(physicl.__file__),
                      "last_row": [float(v) for v in sign.data[-1]]}))


def kernel_level(physicl, np, n, reps):
    """the kernel text the reference GENERATES for ScatterIsotropicStep
    (physicl/light.py:303-315 wrapped by CLProgram.build_kernel, physicl/__init__.py:583-597), compiled by gcc with an
    OpenMP loop over the work-items (oracle/fake_pyopencl), launched on n pre-marshalled float64 particles, plus the
    kinematics law of newton.py:14-16 as NumPy array arithmetic over the same n particles (the reference has no kernel
    for it).  No Python per-particle loops: this is the stand-in for "the reference's OpenCL kernels on the host cores"."""
    import pyopencl  # the shim

    sim = physicl.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 1)
    for _ in range(4):
        sim.add_obj(physicl.light.PhotonObject(v=np.array([physicl.light.c, 0, 0], dtype=np.double), E=np.double(1)))
    sim.add_step(0, physicl.UpdateTimeStep(lambda s: np.double(0.001)))
    sim.add_step(1, physicl.newton.NewtonianKinematicsStep())
    step = physicl.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001))
    sim.add_step(2, step)
    sim.start()
    sim.join()  # one timestep: the step has generated and built its kernel
    prog = step.prog
    kern = getattr(prog.prog, prog.prog_name)
    rng = np.random.default_rng(1)
    c, dt = float(physicl.light.c), 1e-3
    d = rng.normal(size=(3, n))
    v = c * d / np.linalg.norm(d, axis=0)
    r = np.zeros((3, n))
    args = []
    for ctype, is_ptr, name in kern.argspec:
        if is_ptr:
            if name in ("d0", "d1", "d2"):
                args.append(np.ascontiguousarray(v[int(name[1])] * dt))
            elif name == "rtheta":
                args.append(rng.random(n) * 2 * np.pi)
            elif name == "rphi":
                args.append(rng.random(n) * np.pi)
            elif name == "rand":
                args.append(rng.random(n))
            else:
                args.append(np.empty(n, np.double))  # res0..res2
        else:
            args.append(np.double(0.001))  # A, n
    kern(None, (n,), None, *args)  # warm
    t0 = time.perf_counter()
    for _ in range(reps):
        kern(None, (n,), None, *args)
    t_scatter = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        dr = v * dt
        r += dr
    t_kin = (time.perf_counter() - t0) / reps
    hit = int(np.sum(~np.isnan(args[-3])))
    print(json.dumps({"particle_steps_per_s": n / (t_scatter + t_kin), "scatter_kernel_particles_per_s": n / t_scatter,
                      "kinematics_numpy_particles_per_s": n / t_kin, "n": n, "reps": reps, "kernel": prog.prog_name,
                      "scattered_fraction": hit / n, "threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
