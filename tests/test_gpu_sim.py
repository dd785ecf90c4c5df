"""GPU tests through the public API (``Simulation`` / ``Step``), written like the reference's own
tests (test/test_light.py) plus drop-in checks against its golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

import physicl_b200 as phys  # noqa: E402
import physicl_b200.light  # noqa: E402
import physicl_b200.newton  # noqa: E402
from physicl_b200 import _capi  # noqa: E402


def rand_ray():  # reference test/test_light.py:12-17
    return {"s": np.array([0] * 3, dtype=np.double), "v": np.array([phys.light.c, 0, 0], dtype=np.double), "E": np.double(1)}


def sim(n=10000, **kw):  # reference test/test_light.py:19-24
    s = phys.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=True, exit=lambda cond: cond.t >= 0.100, **kw)
    for _ in range(n):
        s.add_obj(phys.light.PhotonObject(**rand_ray()))
    return s


def test_scatter_spherical():
    """reference test/test_light.py:27-43, unchanged apart from the import."""
    x = sim()
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    step = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(3, step)
    x.start()
    x.join()
    assert len(step.data) == 100 and step.data[0][1] == 10000
    error = (np.double(step.data[0][1] * 0.5) - (sum([y[2] for y in step.data]) / len(step.data))) / np.double(step.data[0][1] * 0.5)
    assert np.isclose(error, 0, 0, 0.10)
    # sharper than the reference's 10 %: late rows have forgotten the +x start
    late = np.array([y[2] for y in step.data[50:]])
    assert abs(late.mean() / 10000 - 0.5) < 0.02


def test_scatter_delete():
    """reference test/test_light.py:45-66 (including its row-2 quirk), then the
    Beer-Lambert statement it was after."""
    x = sim()
    x.exit = lambda x: len(x.objects) == 0
    N_i = len(x.objects)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    n, A = 0.001, 0.001
    x.add_step(2, phys.light.ScatterDeleteStep(np.double(n), np.double(A)))
    step = phys.light.ScatterMeasureStep(None, True, [[1 / (n * A), np.nan, np.nan]])
    x.add_step(3, step)
    x.start()
    x.join()
    N_x = sum(step.data[2])
    error = (np.e ** -1 - (N_x / N_i)) / (np.e ** -1)
    assert np.isclose(error, 0, 0, 0.10)
    alive = np.array([r[1] for r in step.data])
    assert np.all(np.diff(alive) <= 0) and alive[-1] == 0
    p = n * A * float(phys.light.c) * 0.001
    for s in range(6):
        assert abs(alive[s] / N_i - (1 - p) ** (s + 1)) < 0.02
    # the plane x = 1/(nA) is crossed on step 4 by everything still alive
    assert step.data[3][2] == alive[3] and sum(r[2] for r in step.data) == alive[3]


def test_numpy_rng_mode_reproduces_the_reference_run(golden):
    """Drop-in check: same seed, same host RNG stream, same pipeline as the reference run that made
    tests/golden/iso.npz -> identical measure-step rows."""
    g = golden("iso")
    N = int(g["N"])
    np.random.seed(int(g["seed"]))
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= int(g["nsteps"]))
    for _ in range(N):
        x.add_obj(phys.light.PhotonObject(**rand_ray()))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(g["A"]), n=np.double(g["n"]), rng="numpy"))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    plane = phys.light.ScatterMeasureStep(None, True, [np.array(p) for p in g["planes"]])
    x.add_step(3, sign)
    x.add_step(4, plane)
    x.start()
    x.join()
    got_s, got_p = np.array(sign.data), np.array(plane.data)
    assert np.array_equal(got_s[:, 1:], g["sign_rows"][:, 1:])
    assert np.array_equal(got_p[:, 1:], g["plane_rows"][:, 1:])
    np.testing.assert_allclose(got_s[:, 0], g["sign_rows"][:, 0], rtol=1e-12)
    # final particle state against the reference's objects (pulls the store back into sim.objects)
    last = int(g["nsteps"]) - 1
    v = np.array([np.asarray(o.v, float) for o in x.objects]).T
    r = np.array([np.asarray(o.r, float) for o in x.objects]).T
    c = float(g["c"])
    assert np.abs(v - g["s%d_v" % last]).max() <= 1e-5 * c
    assert np.abs(r - g["s%d_r" % last]).max() <= 1e-5 * c * 0.001 * (last + 1)


def test_numpy_rng_delete_reproduces_the_reference_run(golden):
    g = golden("delete")
    np.random.seed(int(g["seed"]))
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= int(g["nsteps"]))
    for _ in range(int(g["N"])):
        x.add_obj(phys.light.PhotonObject(**rand_ray()))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterDeleteStep(np.double(g["n"]), np.double(g["A"]), rng="numpy"))
    plane = phys.light.ScatterMeasureStep(None, True, [np.array(g["planes"][0])])
    x.add_step(3, plane)
    x.start()
    x.join()
    assert np.array_equal(np.array(plane.data)[:, 1:], g["plane_rows"][:, 1:])
    assert len(x.objects) == int(g["plane_rows"][-1][1])


def _pipeline(fuse, n=20000, steps=30, seed=5):
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= steps, fuse=fuse, seed=seed)
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x.add_particles(r, v, E=np.ones(n))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    esc = phys.light.EscapeSphereStep(1.0e6)
    x.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    plane = phys.light.ScatterMeasureStep(None, True, [[2.0e5, np.nan, np.nan], [np.nan, np.nan, 0.0]])
    x.add_step(4, sign)
    x.add_step(5, plane)
    x.start()
    x.join()
    return x, esc, sign, plane


def test_fused_and_unfused_pipelines_agree_exactly():
    a, esc_a, sign_a, plane_a = _pipeline(True)
    b, esc_b, sign_b, plane_b = _pipeline(False)
    assert np.array_equal(np.array(sign_a.data), np.array(sign_b.data))
    assert np.array_equal(np.array(plane_a.data), np.array(plane_b.data))
    assert np.array_equal(esc_a.escaped, esc_b.escaped)
    assert esc_a.escaped.sum() + sign_a.data[-1][1] == 20000 and esc_a.escaped.sum() > 0
    assert len(a.objects) == sign_a.data[-1][1]
    sa, sb = a.store.snapshot("photon"), b.store.snapshot("photon")
    assert np.array_equal(sa["id"], sb["id"])
    for nm in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(sa[nm].view(np.uint32), sb[nm].view(np.uint32)), nm
    assert a.cl_ctx.launches < b.cl_ctx.launches  # one launch per timestep instead of five


def test_mixed_population_and_host_step_interop():
    """Kinematics moves every object, scattering touches photons only (light.py:283), the sign tally
    counts all objects (light.py:423-426), and a user's host step sees current state."""
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 3)
    for _ in range(50):
        x.add_obj(phys.light.PhotonObject(**rand_ray()))
    rocks = [phys.Object(v=phys.Measurement([0.0, -2.0, 3.0], "m**1 s**-1")) for _ in range(7)]
    x.add_objs(rocks)
    seen = []

    class Peek(phys.Step):
        def run(self, sim):
            seen.append([np.asarray(o.r, float).copy() for o in sim.objects if type(o) is phys.Object])

    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.5)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-8), n=np.double(1.0)))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(3, sign)
    x.add_step(4, Peek())
    x.start()
    x.join()
    assert len(seen) == 3 and len(seen[0]) == 7
    np.testing.assert_allclose(seen[2][0], [0.0, -3.0, 4.5], rtol=1e-6)
    for row in sign.data:
        assert row[1] == 57 and row[4] >= 7  # 7 rocks have v_z > 0, none has v_y > 0
    assert all(type(o.v) is phys.Measurement for o in x.objects)


def test_wavelength_law_with_the_references_rayleigh_constants():
    """A = 5.1e-31 m^2 * (532 nm)^4 underflows binary32 and (h c / E)^-4 ~ 1e25 overflows the
    product order of the reference kernel; the folded constant keeps the law exact."""
    n = 200_000
    rng = np.random.default_rng(3)
    lam = rng.uniform(300e-9, 900e-9, n)
    E = 6.62607015e-34 * 299792458.0 / lam
    A, nd, dt = 5.1e-31 * (532e-9) ** 4, 2.5e25, 1e-5
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 1, seed=9)
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x.add_particles(r, v, E=E, track_nscat=True)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(dt)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(A), n=np.double(nd), wavelength_dep_scattering=True))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(3, sign)
    x.start()
    x.join()
    pcoll = np.minimum(1.0, A * nd * 299792458.0 * dt * lam ** -4.0)
    snap = x.store.snapshot("photon")
    hits = snap["nscat"].astype(np.float64)
    assert abs(hits.sum() - pcoll.sum()) < 5 * np.sqrt((pcoll * (1 - pcoll)).sum())
    blue, red = lam < 450e-9, lam > 750e-9
    assert hits[blue].mean() > 5 * hits[red].mean()  # lambda^-4
    np.testing.assert_allclose(snap["E"], E, rtol=1e-6)


def test_device_info_and_launch_accounting():
    info = phys.Simulation.get_device_info()
    dev = [v for k, v in info["physicl_b200"].items() if isinstance(v, dict)][0]
    assert dev["MAX_COMPUTE_UNITS"] >= 100 and dev["GLOBAL_MEM_SIZE"] > 1e11
    ctx = _capi.Context(0)
    assert ctx.launches == 0
    assert ctx.fp32_peak_tflops() > 20 and ctx.copy_peak_gbs(1 << 28) > 2000
    assert ctx.launches > 0
    ctx.close()


def test_bulk_run_steps_equals_threaded_run():
    """Simulation.run_steps (k timesteps per C-ABI call) and the threaded per-step loop give the
    same rows and the same particles."""
    a, esc_a, sign_a, plane_a = _pipeline(True, n=30000, steps=40, seed=12)
    x = phys.Simulation(cl_on=True, fuse=True, seed=12)
    n = 30000
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x.add_particles(r, v, E=np.ones(n))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    esc = phys.light.EscapeSphereStep(1.0e6)
    x.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    plane = phys.light.ScatterMeasureStep(None, True, [[2.0e5, np.nan, np.nan], [np.nan, np.nan, 0.0]])
    x.add_step(4, sign)
    x.add_step(5, plane)
    x.run_steps(25)
    x.run_steps(15)
    assert len(x.ts) == 40 and x.step_index == 40
    sa, sb = np.array(sign_a.data), np.array(sign.data)
    assert np.array_equal(sa[:, 1:], sb[:, 1:]) and np.allclose(sa[:, 0], sb[:, 0], rtol=1e-12)
    assert np.array_equal(np.array(plane_a.data)[:, 1:], np.array(plane.data)[:, 1:])
    assert np.array_equal(esc_a.escaped, esc.escaped)
    pa, pb = a.store.snapshot("photon"), x.store.snapshot("photon")
    assert np.array_equal(pa["id"], pb["id"])
    for nm in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(pa[nm].view(np.uint32), pb[nm].view(np.uint32)), nm
    # the threaded run is chunked too (Simulation._run_chunked): both need about one launch per 4-8 timesteps
    assert x.cl_ctx.launches <= 14 and a.cl_ctx.launches <= 14


def _sphere_run(n, steps, seed, A, nd, R, dt=0.001):
    x = phys.Simulation(cl_on=True, seed=seed)
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x.add_particles(r, v, track_nscat=True)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(dt)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    sc = phys.light.ScatterIsotropicStep(A=A, n=nd)
    x.add_step(2, sc)
    esc = None
    if R:
        esc = phys.light.EscapeSphereStep(R)
        x.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(4, sign)
    x.run_steps(steps)
    return x, sc, esc, sign


def test_escape_histogram_matches_float64_reference_law():
    """configs[1] in small: the escape-time histogram and the per-step tallies of the device run
    against the reference's law in float64 (oracle.photon_step_f64) driven by the same Philox
    uniforms.  Decisions can only differ where |pcoll - rand| or |r| - R is within float32 rounding,
    so the integer rows agree up to a handful of photons out of 300 000."""
    import oracle

    n, steps, R = 300_000, 40, 1.5e6
    x, sc, esc, sign = _sphere_run(n, steps, seed=31, A=np.double(1e-3), nd=np.double(1e-3), R=R)
    seed = sc._seed(x)
    st = {k: np.zeros(n) for k in ("x", "y", "z", "vx", "vy", "vz")}
    st["vx"][:] = float(phys.light.c)
    rows = np.array([oracle.photon_step_f64(st, 1e-3, 1e-3, 1e-3, 0.0, float(phys.light.c), 0, seed, s, R * R) for s in range(steps)])
    got_esc = esc.escaped
    got = np.array(sign.data)
    assert got_esc.sum() > 0.5 * n
    assert np.abs(got_esc - rows[:, oracle.T_ESCAPED]).max() <= 3
    assert np.abs(got[:, 1] - rows[:, oracle.T_ALIVE]).max() <= 5
    for col, ocol in ((2, oracle.T_XP), (3, oracle.T_YP), (4, oracle.T_ZP)):
        assert np.abs(got[:, col] - rows[:, ocol]).max() <= 5


def test_scatter_count_per_photon_is_binomial():
    """With a constant collision probability p the number of scatterings of a photon after s steps
    is Binomial(s, p): chi-square of the device's nscat histogram at n = 500 000 (config 3's
    scatter-count check), plus the mean against s*p at 5 sigma."""
    from scipy import stats

    n, steps = 500_000, 12
    x, sc, _, _ = _sphere_run(n, steps, seed=8, A=np.double(1e-3), nd=np.double(1e-3), R=None)
    ns = x.store.snapshot("photon")["nscat"].astype(np.int64)
    p = float(np.float32(1e-6) * np.float32(np.float32(float(phys.light.c)) * np.float32(1e-3)))
    assert abs(ns.mean() - steps * p) < 5 * np.sqrt(steps * p * (1 - p) / n)
    obs = np.bincount(ns, minlength=steps + 1)[: steps + 1].astype(float)
    exp = stats.binom.pmf(np.arange(steps + 1), steps, p) * n
    keep = exp >= 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum() + (obs[~keep].sum() - exp[~keep].sum()) ** 2 / max(exp[~keep].sum(), 1e-9)
    assert stats.chi2.sf(chi2, keep.sum()) > 1e-4, chi2


def test_code_scaled_units_give_the_same_physics():
    """set_code_scale changes every number the kernels see (c, A, n, dt-products; reference light.py:14,
    :309) but not the dimensionless collision probability: tallies agree between metres and kilometres."""
    import importlib
    import subprocess
    import sys
    import json
    import os

    prog = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
import physicl_b200 as phys
scale = float(sys.argv[1])
phys.Measurement.set_code_scale("m", scale)
import physicl_b200.light, physicl_b200.newton
x = phys.Simulation(cl_on=True, seed=4)
n = 50000
c = float(phys.light.c)
r = np.zeros((3, n)); v = np.zeros((3, n)); v[0] = c
x.add_particles(r, v)
x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
x.add_step(1, phys.newton.NewtonianKinematicsStep())
x.add_step(2, phys.light.ScatterIsotropicStep(A=phys.Measurement(1e-3, "m**2"), n=phys.Measurement(1e-3, "m**-3")))
m = phys.light.ScatterSignMeasureStep(None, True)
x.add_step(3, m)
x.run_steps(10)
print(json.dumps({"c": c, "rows": np.array(m.data)[:, 1:].tolist()}))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = []
    for scale in ("1.0", "0.001"):
        r = subprocess.run([sys.executable, "-c", prog, scale], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert out[0]["c"] == 299792458.0 and abs(out[1]["c"] - 299792.458) < 1e-6
    a, b = np.array(out[0]["rows"]), np.array(out[1]["rows"])
    assert a.shape == b.shape and np.abs(a - b).max() <= 3  # float32 rounding may flip a decision or two


# ---- list-producing measure steps, pinned to reference runs --------------
def test_measure_E_lists_match_the_reference_run(golden):
    """ScatterMeasureStep(measure_E=True) (light.py:380-402): same seed and host RNG stream as the
    reference run behind tests/golden/measure_E.npz -> same counts and the same energy lists, in order."""
    g = golden("measure_E")
    N, c, dt = int(g["N"]), float(g["c"]), float(g["dt"])
    np.random.seed(int(g["seed"]))
    phys.light.last_planck_params = None
    E_min, E_max = float(phys.light.E_from_wavelength(2500e-9)), float(phys.light.E_from_wavelength(200e-9))
    E = []
    while len(E) < N:  # the generator's loop: one np.random.rand() per call, None results skipped
        e = phys.light.planck_phot_distribution(E_min, E_max, 5778.0, bins=200)
        if e is not None:
            E.append(np.double(e))
    assert np.array_equal(np.array(E), g["E"])
    x = phys.Simulation(cl_on=True, exit=lambda s: len(s.ts) >= int(g["nsteps"]))
    for e in E:
        x.add_obj(phys.light.PhotonObject(s=np.zeros(3), v=np.array([phys.light.c, 0, 0], dtype=np.double), E=e))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(dt)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(g["A"]), n=np.double(g["n"]), wavelength_dep_scattering=True, rng="numpy"))
    m = phys.light.ScatterMeasureStep(None, True, [np.array(p) for p in g["planes"]], measure_E=True)
    x.add_step(3, m)
    x.start()
    x.join()
    assert len(m.data) == int(g["nsteps"])
    seen = 0
    for s, row in enumerate(m.data):
        assert row.dtype == object and len(row) == 6
        assert [int(row[1]), int(row[2]), int(row[4])] == [int(q) for q in g["s%d_counts" % s]]
        for got, want in ((row[3], g["s%d_E0" % s]), (row[5], g["s%d_E1" % s])):
            assert len(got) == len(want)
            if len(want):
                np.testing.assert_allclose(np.array(got, float), want, rtol=2e-7)
                seen += len(want)
    assert seen > 0


def test_trace_path_matches_the_reference_run(golden):
    """TracePathMeasureStep(trace_dv=True) (light.py:433-483): positions of every photon at every
    timestep and its scatter count, against the reference run behind tests/golden/trace.npz."""
    g = golden("trace")
    N, steps, c = int(g["N"]), int(g["nsteps"]), float(g["c"])
    np.random.seed(int(g["seed"]))
    x = phys.Simulation(cl_on=True, exit=lambda s: len(s.ts) >= steps)
    for _ in range(N):
        x.add_obj(phys.light.PhotonObject(**rand_ray()))
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(g["A"]), n=np.double(g["n"]), rng="numpy"))
    tr = phys.light.TracePathMeasureStep(None, trace_dv=True)
    x.add_step(3, tr)
    x.start()
    x.join()
    rows = tr.data
    assert rows[0][0] == "t" and np.allclose([float(t) for t in rows[0][1:]], g["ts"], rtol=1e-12)
    assert len(rows) == N + 1
    for i, row in enumerate(rows[1:]):
        assert row[0].replace("physicl_b200", "physicl") == str(g["id_info"][i])
        assert int(row[1]) == int(g["freq"][i])
        pos = np.array(row[2:2 + steps], float)
        assert np.abs(pos - g["pos"][i]).max() <= 1e-5 * c * 0.001 * steps
    assert g["freq"].sum() > 0


def test_trace_path_marks_retired_photons_and_keeps_their_counts():
    n, steps = 2000, 10
    x = phys.Simulation(cl_on=True, seed=3)
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x.add_particles(r, v)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
    x.add_step(3, phys.light.EscapeSphereStep(6.0e5))
    tr = phys.light.TracePathMeasureStep(None, trace_dv=True)
    x.add_step(4, tr)
    x.run_steps(steps)
    tr.terminate(x)
    rows = tr.data[1:]
    assert len(rows) == n
    lengths = np.array([sum(1 for q in row[2:] if isinstance(q, np.ndarray)) for row in rows])
    assert lengths.min() < steps and lengths.max() == steps  # some escaped, some survived
    for row, k in zip(rows[:200], lengths[:200]):
        assert len(row) == 2 + k + 3 * (steps - k)
        if k:
            assert np.linalg.norm(row[2 + k - 1]) < 6.0e5 + 3.0e5
    freq = np.array([row[1] for row in rows])
    assert freq.sum() > 0 and freq.max() <= steps


def test_bulk_kinematics_equals_stepwise():
    """Simulation.run_steps on [UpdateTimeStep, NewtonianKinematicsStep]: the bulk form keeps the
    particles in registers for a whole chunk of timesteps; same bits and same clock as step by step."""
    n = 100_003
    rng = np.random.default_rng(9)
    r = rng.uniform(-1e3, 1e3, (3, n)).astype(np.float32)
    v = rng.normal(0, 10, (3, n)).astype(np.float32)

    def make():
        s = phys.Simulation(cl_on=True, exit=lambda x: False)
        s.add_particles(r, v, kind="object")
        s.add_step(0, phys.UpdateTimeStep(lambda x: np.double(1e-3)))
        s.add_step(1, phys.newton.NewtonianKinematicsStep(accel=True, a_uniform=[0, 0, -9.81]))
        return s

    a, b = make(), make()
    a.run_steps(300)
    for _ in range(300):
        b.run_steps(1)
    assert len(a.ts) == len(b.ts) == 300 and float(a.t) == float(b.t)
    ga, gb = a.store.group("object"), b.store.group("object")
    for nm in ("x", "y", "z", "vx", "vy", "vz", "dx", "dy", "dz"):
        assert np.array_equal(ga.download(nm).view(np.uint32), gb.download(nm).view(np.uint32)), nm
    # free fall: z = z0 + vz0 t - g t (t + dt) / 2 for this semi-implicit Euler scheme
    t, dt = 0.3, 1e-3
    want = r[2].astype(np.float64) + v[2].astype(np.float64) * t - 9.81 * t * (t + dt) / 2
    assert np.abs(ga.download("z") - want).max() < 5e-3


def test_presentation_example_2_runs_with_only_the_import_changed():
    """examples/presentation_example_2.ipynb (cells 0, 1, 3, 4 without the plotting): plane atmosphere
    n(z) as an OpenCL-C expression, Rayleigh law, Planck energies, TracePathMeasureStep, threaded run
    polled through get_state().  One photon gets E = None, which planck_phot_distribution may return
    (light.py:102-104): the reference uploads it as NaN and it never scatters."""
    import time

    light, newton = phys.light, phys.newton
    n_0 = phys.Measurement(2.5e25, "m**-3")
    z_0 = phys.Measurement(8.6e3, "m**1")
    cl_n2 = "{} * exp(r2[gid] / {})".format(n_0, z_0)
    np.random.seed(4)
    T = 5778
    E = [light.planck_phot_distribution(light.E_from_wavelength(200e-9), light.E_from_wavelength(2500e-9), T, bins=50000)
         for x in range(300)]
    E[5] = None
    phot = light.generate_photons_from_E(E)
    for p in phot:
        p.r = phys.Measurement([-150e3, np.random.uniform(-50e3, 50e3), np.random.uniform(0, 100e3)], "m**1")
    start = np.array([np.asarray(p.r, float) for p in phot])
    runtime = ((50 * 3 + 50) * 1e3 / light.c) * 0.1  # the notebook runs 5 crossing times (333 steps); 0.1 is enough here
    A_targ = phys.Measurement(5.1e-31, "m**2") * (phys.Measurement(532e-9, "m**1") ** 4)
    sim = phys.Simulation(cl_on=True, exit=lambda cond: cond.t >= runtime)
    sim.add_step(2, phys.UpdateTimeStep(lambda t: phys.Measurement(0.00001, "s**1")))
    sim.add_step(1, newton.NewtonianKinematicsStep())
    sim.add_step(3, light.ScatterIsotropicStep(A=A_targ, variable_n=True, variable_n_fn=cl_n2, wavelength_dep_scattering=True))
    tp = light.TracePathMeasureStep(None)
    sim.add_step(0, tp)
    sim.add_objs(phot)
    sim.start()
    states = []
    while sim.running:
        time.sleep(0.05)
        states.append(sim.get_state())
    sim.join()
    assert states and states[-1]["objects"] == 300
    nsteps = len(sim.ts)
    assert nsteps == int(np.ceil(float(runtime) / 1e-5)) and tp.data[0][0] == "t" and len(tp.data) == 301
    c, dt = float(light.c), 1e-5
    path = np.array([[np.asarray(q, float) for q in row[1:1 + nsteps]] for row in tp.data[1:]])  # (300, steps, 3)
    # steps run in INSERTION order (physicl/__init__.py:514), so the tracer (idx 0, added last) runs last: column k
    # holds the position after k + 1 timesteps
    assert np.allclose(path[:, 0], start + np.array([c * dt, 0.0, 0.0]), rtol=1e-6)
    hops = np.linalg.norm(np.diff(path, axis=1), axis=2)
    assert np.allclose(hops, c * dt, rtol=2e-5)  # every photon moves c dt per step, whatever its direction
    # the E = None photon never scatters: it keeps flying along +x
    assert np.allclose(path[5, :, 1:], start[5, 1:], rtol=1e-6) and np.allclose(np.diff(path[5, :, 0]), c * dt, rtol=1e-5)
    # with variable_n the reference's kernel scalar `A` is the step's n (= 1, light.py:273, :287), so the collision
    # probability is huge at these densities and every other photon is redirected in every step: after the first
    # hop nobody is still on the +x axis direction
    dirs = np.diff(path, axis=1) / (c * dt)
    others = np.delete(np.arange(300), 5)
    assert np.all(np.abs(dirs[others, 0, 0] - 1.0) > 1e-6)


# ---- regressions for the round-1 advisor findings ---------------------------------------------------
def _energy_photons(n):
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    E = 1.0 + np.arange(n, dtype=np.float64) / n  # distinct, exactly representable after the /E0 scaling or not: compare by id
    return r, v, E


@pytest.mark.parametrize("law", ["delete", "escape"])
def test_energies_survive_device_compaction(law):
    """A photon keeps its E through delete scattering / the escape sphere (light.py:34): the compacting
    kernel must move the e plane with the survivors even though the law does not read it."""
    n, steps = 50000, 9
    r, v, E = _energy_photons(n)
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= steps, seed=11)
    x.add_particles(r, v, E=E)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    if law == "delete":
        x.add_step(2, phys.light.ScatterDeleteStep(np.double(2e-4), np.double(1e-3)))
    else:
        x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
        x.add_step(3, phys.light.EscapeSphereStep(1.2e6))
    x.add_step(4, phys.light.ScatterSignMeasureStep(None, True))
    x.run_steps(steps)
    assert x.store.compactions > 0
    snap = x.store.snapshot("photon")
    assert 0 < snap["id"].size < n
    e0 = x.store.group("photon").e0
    want = (E / e0).astype(np.float32).astype(np.float64) * e0
    assert np.array_equal(snap["E"], want[snap["id"]])
    # ... and in the objects pulled back to the host
    objs = list(x.objects)
    assert len(objs) == snap["id"].size
    assert np.array_equal(np.array([float(o.E) for o in objs]), want[snap["id"]])


def test_host_compact_step_carries_energies():
    """pcl_photon_step_host_compact moves host->e with the survivors when the law is not wavelength-dependent."""
    import ctypes as C

    ctx = _capi.Context(0)
    n = 40000
    host = {k: torch.zeros(n, dtype=torch.float32).pin_memory() for k in ("x", "y", "z", "vx", "vy", "vz")}
    host["vx"].fill_(float(phys.light.c))
    host["e"] = (1.0 + torch.arange(n, dtype=torch.float32) / n).pin_memory()
    e_by_id = host["e"].clone().numpy()
    host["id"] = torch.arange(n, dtype=torch.int32).pin_memory()
    soa = _capi.Soa()
    for k, t in host.items():
        setattr(soa, k, t.data_ptr())
    soa.n = n
    sp = _capi.ScatterParams(k=1e-6, c=float(phys.light.c), mode=_capi.SCATTER_DELETE)
    pl = _capi.make_planes([])
    row = np.zeros(_capi.TALLY_COLS, np.int64)
    n_out = C.c_uint64(0)
    for s in range(3):
        rg = _capi.Rng(seed=3, step=s)
        ctx.call("pcl_photon_step_host_compact", C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                 C.byref(pl), row.ctypes.data_as(C.c_void_p), C.c_uint64(8192), C.byref(n_out))
        soa.n = n_out.value
    m = n_out.value
    assert 0 < m < n
    ids = host["id"].numpy()[:m].view(np.uint32)
    assert np.array_equal(host["e"].numpy()[:m], e_by_id[ids])


def test_bulk_particles_are_ingested_once():
    """add_particles data must not come back a second time when the store is rebuilt (host measure step, add_obj)."""
    n = 3000
    r, v, E = _energy_photons(n)

    class CountStep(phys.MeasureStep):  # a host step: walks sim.objects like the reference's measure steps do
        def run(self, sim):
            self.data.append(sum(1 for _ in sim.objects))

    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 6, seed=2)
    x.add_particles(r, v, E=E)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterDeleteStep(np.double(1e-3), np.double(1e-3)))
    cnt = CountStep()
    x.add_step(3, cnt)
    x.start()
    x.join()
    assert cnt.data[0] < n and all(b <= a for a, b in zip(cnt.data, cnt.data[1:]))
    before = len(x.objects)
    x.add_obj(phys.light.PhotonObject(E=np.double(1), v=np.array([phys.light.c, 0, 0], dtype=np.double)))
    x.device_store()
    assert len(x.objects) == before + 1
    # every photon dies, then the store is rebuilt: the original particles must stay gone
    y = phys.Simulation(cl_on=True, exit=lambda c: len(c.objects) == 0, seed=2)
    y.add_particles(r, v, E=E)
    y.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    y.add_step(1, phys.newton.NewtonianKinematicsStep())
    y.add_step(2, phys.light.ScatterDeleteStep(np.double(1.0), np.double(1.0)))
    y.start()
    y.join()
    assert len(list(y.objects)) == 0
    y._host_dirty = True
    assert y.device_store().n_slots == 0


def test_unfused_steps_after_a_device_compaction():
    """Stand-alone device steps (tally before kinematics, escape not adjacent to scatter) keep working once the
    fused retiring step has compacted on the device (the slot count then lives in n_dev)."""
    n = 40000
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 10, seed=9)
    x.add_particles(r, v)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    first = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(1, first)  # measure step BEFORE the kinematics step: runs unfused
    x.add_step(2, phys.newton.NewtonianKinematicsStep())
    x.add_step(3, phys.light.ScatterDeleteStep(np.double(2e-4), np.double(1e-3)))
    after = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(4, after)
    x.add_step(5, phys.light.EscapeSphereStep(2.0e6))  # not adjacent to the scatter step: unfused
    x.start()
    x.join()
    a, b = np.array(first.data), np.array(after.data)
    assert a.shape[0] == b.shape[0] == 10
    assert a[0, 1] == n
    assert np.array_equal(a[1:7, 1], b[:6, 1])  # nothing reaches R = 2e6 before step 7: counts carry over
    assert b[-1, 1] < n


def test_gravity_step_sees_a_rebuilt_store():
    n = 512
    rng = np.random.default_rng(3)
    pos, vel = rng.normal(size=(3, n)), np.zeros((3, n))
    x = phys.Simulation(cl_on=True, exit=lambda c: len(c.ts) >= 2)
    x.add_particles(pos, vel, kind="object")
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
    grav = phys.newton.NewtonianGravityStep(G=1.0, eps2=1e-2, masses=np.full(n, 1.0 / n, np.float32))
    x.add_step(1, grav)
    x.run_steps(2)
    objs = list(x.objects)  # pull to host, edit, rebuild
    objs[0].r = phys.Measurement([50.0, 0.0, 0.0], "m**1")
    x._host_dirty = True
    x.run_steps(1)
    snap = x.store.snapshot("object")
    assert abs(snap["x"][0] - 50.0) < 1.0  # the edit was not overwritten by a stale packed array


# ---- Simulation.start(): the chunked main loop stops where the per-timestep loop stops ------------------------------
def _threaded(exit_fn, n=30000, feedback_every=64, delete=False, seed=17, dt_fn=None):
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    x = phys.Simulation(cl_on=True, exit=exit_fn, seed=seed)
    x.feedback_every = feedback_every
    x.add_particles(r, v)
    x.add_step(0, phys.UpdateTimeStep(dt_fn or (lambda s: np.double(0.001))))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    if delete:
        x.add_step(2, phys.light.ScatterDeleteStep(np.double(1e-3), np.double(1e-3)))
    else:
        x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
        x.add_step(3, phys.light.EscapeSphereStep(1.5e6))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    x.add_step(4, sign)
    x.start()
    x.join()
    return x, np.array(sign.data)


@pytest.mark.parametrize("case", ["time", "steps", "empty", "half", "time_and_count"])
def test_chunked_run_stops_exactly_where_the_stepwise_run_stops(case):
    """reference physicl/__init__.py:512-516 evaluates exit(sim) before EVERY timestep.  The chunked loop (64 timesteps per
    C-ABI call) must produce the same rows, the same t / ts and the same particles as the same run advanced one timestep
    per chunk, for predicates on time, on the step count and on the particle count (the latter fire in the middle of a
    chunk and force a roll-back)."""
    n = 30000
    exit_fn = {
        "time": lambda s: s.t >= 0.0375,
        "steps": lambda s: len(s.ts) >= 71,
        "empty": lambda s: len(s.objects) == 0,
        "half": lambda s: len(s.objects) <= n // 8,
        "time_and_count": lambda s: s.t >= 0.0105 and len(s.objects) <= (9 * n) // 10,
    }[case]
    delete = case in ("empty", "half")
    a, rows_a = _threaded(exit_fn, n, feedback_every=64, delete=delete)
    b, rows_b = _threaded(exit_fn, n, feedback_every=1, delete=delete)
    assert rows_a.shape == rows_b.shape and rows_a.shape[0] >= 3
    assert np.array_equal(rows_a[:, 1:], rows_b[:, 1:])
    np.testing.assert_allclose(rows_a[:, 0].astype(float), rows_b[:, 0].astype(float), rtol=0, atol=0)
    assert len(a.ts) == len(b.ts) == rows_a.shape[0] and float(a.t) == float(b.t)
    assert a.step_index == b.step_index == rows_a.shape[0]
    assert a.exit(a) and b.exit(b)
    if case == "half":  # the stepwise loop stops at the FIRST timestep with <= n/8 photons: the row before is above
        assert rows_a[-1, 1] <= n // 8 < rows_a[-2, 1]
    sa, sb = a.store.snapshot("photon"), b.store.snapshot("photon")
    assert np.array_equal(sa["id"], sb["id"])
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(sa[k].view(np.uint32), sb[k].view(np.uint32)), k


def test_chunked_run_with_a_time_step_that_changes():
    """UpdateTimeStep's fn may return a different dt every timestep (physicl/__init__.py:337-343): chunks break there."""
    dt_fn = lambda s: np.double(0.001 if len(s.ts) % 7 else 0.002)  # noqa: E731
    a, rows_a = _threaded(lambda s: len(s.ts) >= 30, 20000, feedback_every=64, dt_fn=dt_fn)
    b, rows_b = _threaded(lambda s: len(s.ts) >= 30, 20000, feedback_every=1, dt_fn=dt_fn)
    assert rows_a.shape[0] == 30 and np.array_equal(rows_a, rows_b)
    assert [float(t) for t in a.ts] == [float(t) for t in b.ts]


def test_start_after_run_steps_continues_the_random_stream():
    """run() resets t / dt / ts like the reference (physicl/__init__.py:508-510) but never the Philox step counter."""
    n = 20000
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)

    def build():
        x = phys.Simulation(cl_on=True, exit=lambda s: len(s.ts) >= 6, seed=3)
        x.add_particles(r, v)
        x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
        x.add_step(1, phys.newton.NewtonianKinematicsStep())
        x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
        sign = phys.light.ScatterSignMeasureStep(None, True)
        x.add_step(3, sign)
        return x, sign

    a, sa = build()
    a.run_steps(5)
    a.start()
    a.join()
    b, sb = build()
    b.run_steps(11)
    assert a.step_index == b.step_index == 11
    assert np.array_equal(np.array(sa.data)[:, 1:], np.array(sb.data)[:, 1:])


def test_trace_path_with_objects_added_while_running():
    """reference light.py:450-456, :477-481: an object first seen at column b gets 3*b NaNs in front of its positions
    (and, by the reference's own arithmetic, [nan]*3*(columns - positions) behind them).  Objects are added by a host
    step in the middle of the run, which rebuilds the device store: trace ids must survive that."""

    class Adder(phys.Step):
        def run(self, sim):
            if len(sim.ts) == 3:
                for q in range(2):
                    o = phys.Object()
                    o.r = phys.Measurement([100.0 + q, 0.0, 0.0], "m**1")
                    o.v = phys.Measurement([0.0, 1.0 + q, 0.0], "m**1 s**-1")
                    sim.add_obj(o)

    x = phys.Simulation(cl_on=True, exit=lambda s: len(s.ts) >= 6)
    for q in range(3):
        o = phys.Object()
        o.r = phys.Measurement([float(q), 0.0, 0.0], "m**1")
        o.v = phys.Measurement([1.0, 2.0 * q, -1.0], "m**1 s**-1")
        x.add_obj(o)
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.5)))
    x.add_step(1, Adder())
    x.add_step(2, phys.newton.NewtonianKinematicsStep())
    tr = phys.light.TracePathMeasureStep(None, id_info_fn=lambda o: "obj")
    x.add_step(3, tr)
    x.start()
    x.join()
    data = tr.data
    assert data[0][0] == "t" and len(data[0]) == 7 and len(data) == 1 + 5
    for q in range(3):  # present from the start: 6 positions, nothing else
        row = data[1 + q]
        assert row[0] == "obj" and len(row) == 1 + 6
        for k in range(6):
            np.testing.assert_allclose(row[1 + k], [q + 0.5 * (k + 1), 2.0 * q * 0.5 * (k + 1), -0.5 * (k + 1)], rtol=1e-6)
    for q in range(2):  # added before the third timestep's kinematics: first seen at column 2
        row = data[4 + q]
        b, npos = 2, 4
        assert len(row) == 1 + 3 * b + npos + 3 * (6 - npos)
        assert all(np.isnan(v) for v in row[1:1 + 3 * b])
        for k in range(npos):
            np.testing.assert_allclose(row[1 + 3 * b + k], [100.0 + q, (1.0 + q) * 0.5 * (k + 1), 0.0], rtol=1e-6)
        assert all(np.isnan(v) for v in row[1 + 3 * b + npos:])


def test_pulled_objects_carry_dv():
    """Object.dv (reference light.py:325-331): v_new - v_old for a photon that scattered in the timestep, zero otherwise.
    A host step that walks sim.objects every timestep sees exactly that."""
    seen = []

    class Watch(phys.Step):
        def run(self, sim):
            seen.append([(np.asarray(o.v, float).copy(), np.asarray(o.dv, float).copy()) for o in sim.objects])

    x = sim(400, seed=4)
    x.exit = lambda s: len(s.ts) >= 4
    x.add_step(0, phys.UpdateTimeStep(lambda s: np.double(0.001)))
    x.add_step(1, phys.newton.NewtonianKinematicsStep())
    x.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    x.add_step(3, Watch())
    x.start()
    x.join()
    assert len(seen) == 4 and all(len(s) == 400 for s in seen)
    c32 = float(np.float32(float(phys.light.c)))
    prev = [np.array([c32, 0.0, 0.0])] * 400
    scattered = 0
    for step in seen:
        for j, (v, dv) in enumerate(step):
            np.testing.assert_allclose(dv, v - prev[j], rtol=0, atol=1e-3)
            if np.any(v != prev[j]):
                scattered += 1
            else:
                assert not np.any(dv)
        prev = [v for v, _ in step]
    assert 0.2 * 1600 < scattered < 0.4 * 1600


def test_sfu_trig_option_through_the_simulation_api():
    """ScatterIsotropicStep(sfu_trig=True): same pipeline, directions from the SFU.  The first row (no photon has a new
    direction yet when it is decided) equals the default run's in its decision columns, the sign tallies stay balanced
    the same way, and an unfused pipeline refuses the option loudly."""
    n, steps, C_LIGHT = 200_000, 12, 299792458.0

    def run(sfu, fuse=True):
        sim = phys.Simulation(cl_on=True, seed=99, fuse=fuse, exit=lambda s: len(s.ts) >= steps)
        r = np.zeros((3, n), np.float32)
        v = np.zeros((3, n), np.float32)
        v[0] = C_LIGHT
        sim.add_particles(r, v)
        sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
        sim.add_step(1, phys.newton.NewtonianKinematicsStep())
        sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3), sfu_trig=sfu))
        sign = phys.light.ScatterSignMeasureStep(None, True)
        sim.add_step(3, sign)
        sim.start()
        sim.join()
        return np.array([[float(x) for x in row] for row in sign.data]), sim.device_store().snapshot("photon")

    rows_t, snap_t = run(False)
    rows_s, snap_s = run(True)
    assert rows_t.shape == rows_s.shape == (steps, 5)
    assert np.array_equal(rows_t[:, :2], rows_s[:, :2])  # t, N
    assert not np.array_equal(snap_t["vx"], snap_s["vx"])  # the option reached the kernel
    speed = np.sqrt(snap_s["vx"].astype(np.float64) ** 2 + snap_s["vy"].astype(np.float64) ** 2 + snap_s["vz"].astype(np.float64) ** 2)
    assert np.abs(speed / C_LIGHT - 1.0).max() <= 1e-5
    for col in (2, 3, 4):  # sign counts agree in law: differences within 5 sigma of two binomial counts
        assert np.all(np.abs(rows_t[:, col] - rows_s[:, col]) < 5 * np.sqrt(rows_t[:, col] + rows_s[:, col] + 1.0))
    with pytest.raises(Exception, match="SFU"):
        run(True, fuse=False)


def test_read_only_host_step_skips_the_reupload():
    """A Python measurement step that only reads sim.objects (modifies_objects = False): the objects are current when it
    runs, the device store is NOT rebuilt from them afterwards, and the physics is the run without that step, row for row.
    The same step without the declaration forces a rebuild every timestep (the safe default) and gives the same rows."""
    n, steps = 3000, 6

    class MeanX(phys.MeasureStep):
        def __init__(self, read_only):
            super().__init__(None)
            if read_only:
                self.modifies_objects = False
            self.stores = []

        def run(self, sim):
            self.data.append(float(np.mean([float(o.r[0]) for o in sim.objects])))
            self.stores.append(sim.store)  # kept alive, so that identities cannot be recycled

    def run(host_step):
        sim = phys.Simulation(cl_on=True, seed=5, exit=lambda s: len(s.ts) >= steps)
        for _ in range(n):
            sim.add_obj(phys.light.PhotonObject(v=np.array([phys.light.c, 0, 0], dtype=np.double), E=np.double(1)))
        sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
        sim.add_step(1, phys.newton.NewtonianKinematicsStep())
        sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
        sign = phys.light.ScatterSignMeasureStep(None, True)
        sim.add_step(3, sign)
        if host_step is not None:
            sim.add_step(4, host_step)
        sim.start()
        sim.join()
        return np.array([[float(x) for x in row] for row in sign.data])

    base = run(None)
    ro, rw = MeanX(True), MeanX(False)
    rows_ro, rows_rw = run(ro), run(rw)
    assert np.array_equal(rows_ro, base) and np.array_equal(rows_rw, base)
    assert len(ro.data) == steps and np.allclose(ro.data, rw.data, rtol=0, atol=0)
    assert ro.data[0] > 0 and ro.data[-1] != ro.data[0]          # the objects were current every time it looked
    assert len({id(x) for x in ro.stores}) == 1                    # one store for the whole run
    assert len({id(x) for x in rw.stores}) == steps                # rebuilt after every host step
