#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

How: ``/root/reference`` is put on sys.path together with ``oracle/fake_pyopencl`` (a stand-in
``pyopencl`` that compiles the reference's own OpenCL-C kernel strings with gcc).  The reference's
``Simulation`` then runs its own host code (``CLProgram.run`` marshalling, ``NewtonianKinematicsStep``,
measure steps) and its own kernel text.  ``np.random.random`` / ``np.random.rand`` are wrapped so every
uniform the reference draws is logged in draw order; kernel launches are logged by the shim.

One compatibility alias is needed on NumPy >= 1.24: ``np.int = np.int32`` (the reference writes
``dtype=np.int``, physicl/__init__.py:653, light.py:198; its kernels declare ``int``).

Each file stores inputs (uniforms, initial state, constants) and the reference's outputs (kernel
results, per-step object state, measure-step rows).  ``min_margin`` is the smallest relative distance
``|pcoll - rand| / pcoll`` over the run: the float32 device path can only flip a scatter decision if
this is below ~1e-6, so the generator insists on > 1e-4 and the parity tests may demand identical flags.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "fake_pyopencl"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
np.int = np.int32  # see module docstring

import pyopencl  # noqa: E402  (the shim)
import physicl  # noqa: E402
import physicl.light  # noqa: E402
import physicl.newton  # noqa: E402

assert physicl.__file__.startswith("/root/reference/"), physicl.__file__


class DrawLog:
    """Wraps np.random.random / np.random.rand and records every scalar draw in order."""

    def __init__(self):
        self.draws = []
        self._random, self._rand = np.random.random, np.random.rand

    def __enter__(self):
        def random(*a, **k):
            v = self._random(*a, **k)
            if not a and not k:
                self.draws.append(float(v))
            return v

        def rand(*a, **k):
            v = self._rand(*a, **k)
            if not a and not k:
                self.draws.append(float(v))
            return v

        np.random.random, np.random.rand = random, rand
        return self

    def __exit__(self, *exc):
        np.random.random, np.random.rand = self._random, self._rand

    def take(self):
        d, self.draws = self.draws, []
        return np.array(d, np.float64)


class Snapshot(physicl.Step):
    """Extra step appended to the reference simulation: copies every object's state each timestep."""

    def __init__(self, log, keep_ids=False):
        self.log, self.rows, self.keep_ids = log, [], keep_ids

    def run(self, sim):
        objs = sim.objects
        st = {
            "r": np.array([np.asarray(o.r, float) for o in objs]).reshape(-1, 3).T.copy(),
            "v": np.array([np.asarray(o.v, float) for o in objs]).reshape(-1, 3).T.copy(),
            "dr": np.array([np.asarray(o.dr, float) for o in objs]).reshape(-1, 3).T.copy(),
            "dv": np.array([np.asarray(o.dv, float) for o in objs]).reshape(-1, 3).T.copy(),
            "t": float(sim.t), "dt": float(sim.dt),
            "u": self.log.take(),
        }
        if self.keep_ids:
            st["gid"] = np.array([o.gid for o in objs], np.int64)
        self.rows.append(st)


def photons(n, E=None):
    c = physicl.light.c
    out = []
    for i in range(n):
        e = np.double(1) if E is None else E[i]
        p = physicl.light.PhotonObject(s=np.zeros(3), v=np.array([c, 0, 0], dtype=np.double), E=e)  # test/test_light.py:12-17
        p.gid = i
        out.append(p)
    return out


def run_sim(objs, steps, nsteps, dt):
    pyopencl.LAUNCH_LOG.clear()
    pyopencl.RECORD = True
    dts = dt if callable(dt) else (lambda s: np.double(dt))
    state = {"k": 0}

    def exit_fn(s):
        return state["k"] >= nsteps

    class Count(physicl.Step):
        def run(self, sim):
            state["k"] += 1

    sim = physicl.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=True, exit=exit_fn)
    sim.add_objs(objs)
    sim.add_step(0, physicl.UpdateTimeStep(dts))
    for i, s in enumerate(steps):
        sim.add_step(i + 1, s)
    sim.add_step(99, Count())
    sim.start()
    sim.join()
    launches = list(pyopencl.LAUNCH_LOG)
    pyopencl.RECORD = False
    return sim, launches


def margin(launch, with_E=False, hc=None):
    b = launch["before"]
    norm = np.sqrt(b["d0"] ** 2 + b["d1"] ** 2 + b["d2"] ** 2)
    p = launch["scalars"]["A"] * launch["scalars"]["n"] * norm
    if with_E:
        p = p * (hc / b["E"]) ** -4
    return float(np.min(np.abs(p - b["rand"]) / np.maximum(p, 1e-300)))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def pack_steps(rows, launches, prefix_keys):
    out = {}
    for s, (st, la) in enumerate(zip(rows, launches)):
        for k in ("r", "v", "dr", "dv", "u"):
            out["s%d_%s" % (s, k)] = st[k]
        out["s%d_t" % s] = st["t"]
        out["s%d_dt" % s] = st["dt"]
        for k in prefix_keys:
            src = la["after"] if k.startswith("res") else la["before"]
            out["s%d_%s" % (s, k)] = src[k]
    return out


def gen_iso(seed=11, N=512, nsteps=6):
    np.random.seed(seed)
    planes = [[4.0e5, np.nan, np.nan], [np.nan, 0.0, np.nan], [np.nan, np.nan, -1.0e5]]
    with DrawLog() as log:
        snap = Snapshot(log)
        sign = physicl.light.ScatterSignMeasureStep(None, True)
        plane = physicl.light.ScatterMeasureStep(None, True, [np.array(p, dtype=np.double) for p in planes])
        A, n = np.double(0.001), np.double(0.001)  # test/test_light.py:34
        sim, launches = run_sim(photons(N), [physicl.newton.NewtonianKinematicsStep(),
                                             physicl.light.ScatterIsotropicStep(A=A, n=n), sign, plane, snap], nsteps, 0.001)
    assert len(launches) == nsteps
    mm = min(margin(l) for l in launches)
    assert mm > 1e-4, mm
    out = pack_steps(snap.rows, launches, ["d0", "d1", "d2", "rtheta", "rphi", "rand", "res0", "res1", "res2"])
    save("iso", N=N, nsteps=nsteps, seed=seed, A=float(A), n=float(n), c=float(physicl.light.c), dt=0.001,
         planes=np.array(planes), sign_rows=np.array(sign.data), plane_rows=np.array(plane.data), min_margin=mm,
         kernel_A=launches[0]["scalars"]["A"], kernel_n=launches[0]["scalars"]["n"], **out)


def gen_wave(seed=12, N=512, nsteps=4):
    np.random.seed(seed)
    T = 5778.0
    E_min = float(physicl.light.E_from_wavelength(2500e-9))
    E_max = float(physicl.light.E_from_wavelength(200e-9))
    E = []
    while len(E) < N:  # examples/presentation_example.ipynb cell 1 idiom; the sampler may return None
        e = physicl.light.planck_phot_distribution(E_min, E_max, T, bins=200)
        if e is not None:
            E.append(np.double(e))
    E = np.array(E)
    A = np.double(5.1e-31 * (532e-9) ** 4)  # examples/presentation_example.ipynb cell 3
    n = np.double(2.5e25)  # examples/presentation_example_2.ipynb cell 0
    dt = 1e-5
    with DrawLog() as log:
        snap = Snapshot(log)
        sign = physicl.light.ScatterSignMeasureStep(None, True)
        sim, launches = run_sim(photons(N, E), [physicl.newton.NewtonianKinematicsStep(),
                                                physicl.light.ScatterIsotropicStep(A=A, n=n, wavelength_dep_scattering=True),
                                                sign, snap], nsteps, dt)
    hc = float(physicl.light.h) * float(physicl.light.c)
    mm = min(margin(l, True, hc) for l in launches)
    assert mm > 1e-4, mm
    out = pack_steps(snap.rows, launches, ["d0", "d1", "d2", "rtheta", "rphi", "rand", "E", "res0", "res1", "res2"])
    save("wave", N=N, nsteps=nsteps, seed=seed, A=float(A), n=float(n), c=float(physicl.light.c), h=float(physicl.light.h),
         dt=dt, E=E, sign_rows=np.array(sign.data), min_margin=mm, kernel_src=np.array(sim.steps[2].prog.kernel_code), **out)


def gen_delete(seed=13, N=1024, nsteps=6, reference_twin=False):
    np.random.seed(seed)
    n, A = np.double(0.001), np.double(0.001)  # test/test_light.py:54-56
    planes = [[1 / (float(n) * float(A)), np.nan, np.nan]]
    with DrawLog() as log:
        snap = Snapshot(log, keep_ids=True)
        plane = physicl.light.ScatterMeasureStep(None, True, [np.array(planes[0], dtype=np.double)])
        step = physicl.light.ScatterDeleteStepReference(n, A) if reference_twin else physicl.light.ScatterDeleteStep(n, A)
        sim, launches = run_sim(photons(N), [physicl.newton.NewtonianKinematicsStep(), step, plane, snap], nsteps, 0.001)
    mm = min(margin({"before": {"d0": l["before"].get("d0", l["before"].get("dx")),
                                "d1": l["before"].get("d1", l["before"].get("dy")),
                                "d2": l["before"].get("d2", l["before"].get("dz")), "rand": l["before"]["rand"]},
                     "scalars": l["scalars"]}) for l in launches)
    assert mm > 1e-4, mm
    out = {}
    for s, (st, la) in enumerate(zip(snap.rows, launches)):
        out["s%d_gid" % s] = st["gid"]  # survivors after the step, in list order
        out["s%d_r" % s] = st["r"]
        out["s%d_dr" % s] = st["dr"]
        out["s%d_u" % s] = st["u"]  # one draw per photon alive before the step
        out["s%d_flags" % s] = la["after"]["result" if reference_twin else "res"]
    save("delete_ref" if reference_twin else "delete", N=N, nsteps=nsteps, seed=seed, A=float(A), n=float(n),
         c=float(physicl.light.c), dt=0.001, planes=np.array(planes), plane_rows=np.array(plane.data), min_margin=mm, **out)


def gen_planck(seed=14, bins=200, ndraw=4000):
    np.random.seed(seed)
    T = 5778.0
    E_min = float(physicl.light.E_from_wavelength(2500e-9))
    E_max = float(physicl.light.E_from_wavelength(200e-9))
    physicl.light.last_planck_params = None
    with DrawLog() as log:
        vals = [physicl.light.planck_phot_distribution(E_min, E_max, T, bins=bins) for _ in range(ndraw)]
        u = log.take()
    grid = np.linspace(E_min, E_max, bins)
    picked = np.array([-1 if v is None else int(np.argmin(np.abs(grid - float(v)))) for v in vals], np.int64)
    for v, b in zip(vals, picked):
        assert v is None or float(v) == grid[b]
    save("planck", bins=bins, T=T, E_min=E_min, E_max=E_max, seed=seed, u=u, bin=picked,
         cdf=np.array(physicl.light.last_planck_cdf), gamma_norm=np.array(physicl.light.last_planck_gamma_norm),
         E=np.array([np.nan if v is None else float(v) for v in vals]), kB=float(physicl.light.kB))


def gen_kin(seed=15, N=64, nsteps=5):
    rng = np.random.RandomState(seed)
    objs = []
    for i in range(N):
        o = physicl.Object()
        o.r = physicl.Measurement(list(rng.uniform(-1e3, 1e3, 3)), "m**1")
        o.v = physicl.Measurement(list(rng.normal(0, 10, 3)), "m**1 s**-1")
        objs.append(o)
    r0 = np.array([np.asarray(o.r, float) for o in objs]).T.copy()
    v0 = np.array([np.asarray(o.v, float) for o in objs]).T.copy()
    dts = [1e-3, 2.5e-3, 1e-4, 0.5, 1e-3]
    k = {"i": 0}

    def dt_fn(s):
        k["i"] += 1
        return np.double(dts[k["i"] - 1])

    with DrawLog() as log:
        snap = Snapshot(log)
        run_sim(objs, [physicl.newton.NewtonianKinematicsStep(), snap], nsteps, dt_fn)
    out = {}
    for s, st in enumerate(snap.rows):
        out["s%d_r" % s], out["s%d_dr" % s], out["s%d_t" % s] = st["r"], st["dr"], st["t"]
    save("kin", N=N, nsteps=nsteps, seed=seed, r0=r0, v0=v0, dts=np.array(dts), **out)


class _NumpyRagged:
    """numpy >= 1.24 refuses ragged ``np.array([...])``; the reference's measure_E rows are ragged
    (physicl/light.py:404).  Inside physicl.light only, fall back to an object array, which is what
    older numpy produced."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def array(self, obj, *a, **k):
        try:
            return self._real.array(obj, *a, **k)
        except ValueError:
            out = self._real.empty(len(obj), dtype=object)
            out[:] = obj
            return out


def gen_measure_E(seed=17, N=256, nsteps=5):
    np.random.seed(seed)
    T = 5778.0
    E_min = float(physicl.light.E_from_wavelength(2500e-9))
    E_max = float(physicl.light.E_from_wavelength(200e-9))
    E = []
    while len(E) < N:
        e = physicl.light.planck_phot_distribution(E_min, E_max, T, bins=200)
        if e is not None:
            E.append(np.double(e))
    E = np.array(E)
    A, n, dt = np.double(5.1e-31 * (532e-9) ** 4), np.double(2.5e25), 1e-5
    c = float(physicl.light.c)
    planes = [[2.5 * c * dt, np.nan, np.nan], [np.nan, 0.2 * c * dt, np.nan]]
    real_np = physicl.light.np
    physicl.light.np = _NumpyRagged(real_np)
    try:
        with DrawLog() as log:
            snap = Snapshot(log)
            m = physicl.light.ScatterMeasureStep(None, True, [np.array(p, dtype=np.double) for p in planes], measure_E=True)
            sim, launches = run_sim(photons(N, E), [physicl.newton.NewtonianKinematicsStep(),
                                                    physicl.light.ScatterIsotropicStep(A=A, n=n, wavelength_dep_scattering=True),
                                                    m, snap], nsteps, dt)
    finally:
        physicl.light.np = real_np
    out = {}
    for s, (st, row) in enumerate(zip(snap.rows, m.data)):
        out["s%d_u" % s] = st["u"]
        out["s%d_counts" % s] = np.array([row[1], row[2], row[4]], np.int64)
        out["s%d_E0" % s] = np.array([float(x) for x in row[3]], np.float64)
        out["s%d_E1" % s] = np.array([float(x) for x in row[5]], np.float64)
    assert sum(len(out["s%d_E0" % s]) + len(out["s%d_E1" % s]) for s in range(nsteps)) > 0
    save("measure_E", N=N, nsteps=nsteps, seed=seed, A=float(A), n=float(n), c=c, h=float(physicl.light.h), dt=dt, E=E,
         planes=np.array(planes), **out)


def gen_trace(seed=18, N=48, nsteps=6):
    np.random.seed(seed)
    A, n = np.double(0.002), np.double(0.001)
    with DrawLog() as log:
        snap = Snapshot(log)
        tr = physicl.light.TracePathMeasureStep(None, trace_dv=True)
        sim, launches = run_sim(photons(N), [physicl.newton.NewtonianKinematicsStep(),
                                             physicl.light.ScatterIsotropicStep(A=A, n=n), tr, snap], nsteps, 0.001)
    rows = tr.data
    assert rows[0][0] == "t" and len(rows) == N + 1
    freq = np.array([r[1] for r in rows[1:]], np.int64)
    pos = np.array([[np.asarray(p, float) for p in r[2:2 + nsteps]] for r in rows[1:]])  # (N, steps, 3)
    out = {"s%d_u" % s: st["u"] for s, st in enumerate(snap.rows)}
    save("trace", N=N, nsteps=nsteps, seed=seed, A=float(A), n=float(n), c=float(physicl.light.c), dt=0.001,
         ts=np.array([float(t) for t in rows[0][1:]]), freq=freq, pos=pos, id_info=np.array([str(r[0]) for r in rows[1:]]), **out)


def eval_cl_expr(expr, **arrays):
    """The user's OpenCL-C density expression evaluated with NumPy in float64 (gid = every photon)."""
    ns = {"pow": np.power, "exp": np.exp, "sqrt": np.sqrt, "log": np.log, "sin": np.sin, "cos": np.cos, "fabs": np.abs,
          "gid": slice(None)}
    ns.update(arrays)
    return eval(expr, {"__builtins__": {}}, ns)


def gen_varn(name, expr, wave, n_kwarg, A_kwarg, dt, seed, N=512, nsteps=5):
    """ScatterIsotropicStep(variable_n=True) (light.py:295-299): the reference splices `expr` into its
    kernel; the generated kernel text and scalar bindings are stored next to the outputs."""
    np.random.seed(seed)
    T = 5778.0
    E_min = float(physicl.light.E_from_wavelength(2500e-9))
    E_max = float(physicl.light.E_from_wavelength(200e-9))
    E = []
    while len(E) < N:
        e = physicl.light.planck_phot_distribution(E_min, E_max, T, bins=200)
        if e is not None:
            E.append(np.double(e))
    E = np.array(E)
    with DrawLog() as log:
        snap = Snapshot(log)
        sign = physicl.light.ScatterSignMeasureStep(None, True)
        step = physicl.light.ScatterIsotropicStep(A=A_kwarg, n=n_kwarg, wavelength_dep_scattering=wave, variable_n=True,
                                                  variable_n_fn=expr)
        sim, launches = run_sim(photons(N, E), [physicl.newton.NewtonianKinematicsStep(), step, sign, snap], nsteps, dt)
    hc = float(physicl.light.h) * float(physicl.light.c)
    mm, fracs = 1.0, []
    for la in launches:
        b = la["before"]
        norm = np.sqrt(b["d0"] ** 2 + b["d1"] ** 2 + b["d2"] ** 2)
        dens = eval_cl_expr(expr, r0=b["r0"], r1=b["r1"], r2=b["r2"])
        pc = la["scalars"]["A"] * dens * norm
        if wave:
            pc = pc * (hc / b["E"]) ** -4
        hit = pc >= b["rand"]
        assert np.array_equal(hit, ~np.isnan(la["after"]["res0"]))  # the restated law is the reference's
        mm = min(mm, float(np.min(np.abs(pc - b["rand"]) / np.maximum(pc, 1e-300))))
        fracs.append(hit.mean())
    assert mm > 1e-4, mm
    assert 0.02 < np.mean(fracs) < 0.9, fracs
    keys = ["d0", "d1", "d2", "r0", "r1", "r2", "rtheta", "rphi", "rand", "res0", "res1", "res2"] + (["E"] if wave else [])
    out = pack_steps(snap.rows, launches, keys)
    save(name, N=N, nsteps=nsteps, seed=seed, A=float(A_kwarg), n=float(n_kwarg), c=float(physicl.light.c),
         h=float(physicl.light.h), dt=dt, E=E, wave=bool(wave), expr=np.array(expr), sign_rows=np.array(sign.data), min_margin=mm,
         kernel_A=launches[0]["scalars"]["A"], kernel_n=launches[0]["scalars"]["n"], scattered_fraction=np.array(fracs),
         kernel_src=np.array(step.prog.kernel_code), **out)


def gen_clprogram(seed=21, N=300):
    """A user-written kernel through the reference's CLInput/CLOutput/CLProgram (physicl/__init__.py:543-664):
    obj inputs, an obj_def input drawn from np.random, an obj_action type filter, an obj_track list, constants,
    a double and an int output, an early return.  Half of the objects are photons, which the filter skips."""
    np.random.seed(seed)
    rng = np.random.RandomState(seed)
    objs = []
    c = physicl.light.c
    for i in range(N):
        if i % 2:
            o = physicl.light.PhotonObject(s=np.zeros(3), v=np.array([c, 0, 0], dtype=np.double), E=np.double(1))
        else:
            o = physicl.Object()
            o.r = physicl.Measurement(list(rng.uniform(-1e3, 1e3, 3)), "m**1")
            o.v = physicl.Measurement(list(rng.normal(0, 10, 3)), "m**1 s**-1")
        o.gid = i
        objs.append(o)
    body = """
        int gid = get_global_id(0);
        double speed = sqrt(pow(v0[gid], 2) + pow(v1[gid], 2) + pow(v2[gid], 2));
        if (r2[gid] < zcut) { flag[gid] = 0; ke[gid] = NAN; return; }
        ke[gid] = 0.5 * m * speed * speed + g * r2[gid] + jitter[gid] * exp(-speed / 10.0);
        flag[gid] = 1;
    """
    sim = physicl.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=True, exit=lambda s: True)
    sim.add_objs(objs)
    prog = physicl.CLProgram(sim, "user_energy", body)
    skip = physicl.CLInput(name="skip", type="obj_action", code="if type(obj) == physicl.light.PhotonObject:\n \t\t continue")
    v = [physicl.CLInput(name="v%d" % i, type="obj", obj_attr="v[%d]" % i) for i in range(3)]
    r2 = physicl.CLInput(name="r2", type="obj", obj_attr="r[2]")
    jit = physicl.CLInput(name="jitter", type="obj_def", obj_def="np.random.random()")
    who = physicl.CLInput(name="who", type="obj_track", obj_track="obj")
    consts = [physicl.CLInput(name="m", type="const", const_value="2.5"), physicl.CLInput(name="g", type="const", const_value=str(9.81)),
              physicl.CLInput(name="zcut", type="const", const_value="-250.0")]
    prog.prep_metadata = [skip] + v + [r2, jit, who] + consts
    prog.output_metadata = [physicl.CLOutput(name="ke"), physicl.CLOutput(name="flag", ctype="int")]
    pyopencl.LAUNCH_LOG.clear()
    pyopencl.RECORD = True
    with DrawLog() as log:
        prog.build_kernel()
        out = prog.run()
        u = log.take()
    pyopencl.RECORD = False
    la = pyopencl.LAUNCH_LOG[-1]
    assert len(out["ke"]) == N // 2 and 0 < out["flag"].sum() < N // 2
    r = np.array([np.asarray(o.r, float) for o in objs]).T.copy()
    vv = np.array([np.asarray(o.v, float) for o in objs]).T.copy()
    save("clprogram", N=N, seed=seed, r=r, v=vv, is_photon=np.array([i % 2 for i in range(N)], np.int32), body=np.array(body),
         jitter=u, ke=out["ke"], flag=np.asarray(out["flag"], np.int32), tracked_gid=np.array([o.gid for o in prog.who], np.int64),
         m=2.5, g=9.81, zcut=-250.0, c=float(c), **{"in_" + k: a for k, a in la["before"].items()})


if __name__ == "__main__":
    gen_iso()
    gen_wave()
    gen_delete()
    gen_delete(seed=16, N=512, nsteps=3, reference_twin=True)
    gen_planck()
    gen_kin()
    gen_measure_E()
    gen_trace()
    # examples/presentation_example.ipynb cell 0 (radial atmosphere) and presentation_example_2.ipynb cell 0 (plane
    # atmosphere) forms, scaled so that some but not all photons scatter per step
    gen_varn("varn", "{} * exp(-1 * ({} - {})/({}))".format(6.0e26, "sqrt(pow(r0[gid], 2) + pow(r1[gid], 2) + pow(r2[gid], 2))",
                                                            1000.0, 8000.0),
             True, np.double(5.1e-31 * (532e-9) ** 4), np.double(123.0), 1e-5, seed=19)
    gen_varn("varn_z", "{} * exp(r2[gid] / {})".format(1.0e-3, 2.0e6), False, np.double(1.0e-3), np.double(7.0), 1e-3, seed=20)
    gen_clprogram()
