"""CPU-only tests: host-side mirror of the reference API, the C-ABI surface, sharding logic."""
import ctypes
import os
import re

import numpy as np
import pytest

import physicl_b200 as phys
import physicl_b200.light
import physicl_b200.newton
from physicl_b200 import _capi, dist, fused

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- code units (behaviour of reference test/test_units.py) -----------------------------------
def dict_equiv(a, b):
    return all(b.get(k, 0) in (0, v) or v == 0 for k, v in a.items()) and all(a.get(k, 0) in (0, v) or v == 0 for k, v in b.items())


def test_units_derived_equals_base():
    x = phys.Measurement(5, "kg**1 m**1 s**-2")
    y = phys.Measurement(5, "N**1")
    assert x == y and x.units == y.units == {"M": 1, "L": 1, "T": -2}


def test_units_scaling_au():
    x = phys.Measurement(1, "au**1")
    y = phys.Measurement(149597870700, "m**1")
    assert x + y == phys.Measurement(2, "au**1")
    assert y + x == phys.Measurement(149597870700 * 2, "m**1")


def test_units_photon_object():
    p = phys.light.PhotonObject(E=phys.Measurement(5, "J**1"), v=phys.Measurement([phys.light.c, 0, 0], "m**1 s**-1"))
    assert p.E.units == {"L": 2, "T": -2, "M": 1}
    assert p.v.units == {"L": 1, "T": -1}
    assert np.linalg.norm(p.v) == phys.light.c
    with pytest.raises(Exception, match="valid speed"):
        phys.light.PhotonObject(E=1, v=phys.Measurement([1.0, 0, 0], "m**1 s**-1"))
    with pytest.raises(Exception, match="valid energy"):
        phys.light.PhotonObject(v=phys.Measurement([phys.light.c, 0, 0], "m**1 s**-1"))


def test_units_wavelength_energy_roundtrip():
    E = phys.light.E_from_wavelength(phys.Measurement(633e-9, "m**1"))
    assert E == (299792458 * 6.62607015e-34) / (633e-9)
    assert E.units == {"L": 2, "T": -2, "M": 1}
    wv = phys.light.wavelength_from_E(E)
    assert wv == 633e-9
    assert dict_equiv(wv.units, {"L": 1})


def test_units_ev_conversion():
    E_g = phys.Measurement(0, "J**1") + phys.Measurement(13.6, "eV**1")
    f = E_g / phys.light.h
    lam = phys.light.c / f
    assert E_g == 1.602176634e-19 * 13.6
    assert f == (1.602176634e-19 * 13.6) / 6.62607015e-34 and dict_equiv(f.units, {"T": -1})
    assert lam == 299792458 / ((1.602176634e-19 * 13.6) / 6.62607015e-34) and dict_equiv(lam.units, {"L": 1})


def test_units_ufuncs():
    a = phys.Measurement(5, "kg**1 m**1 s**-2")
    l = phys.Measurement(5, "au**1")
    t = phys.Measurement(10, "min**2")
    assert a * t == 50
    assert phys.Measurement(0, "kg**1 m**1") + (a * t) == (60 ** 2) * 10 * 5
    assert a * l == 25
    assert (a / l).flat[0] == 5 / (5 * 149597870700)
    assert a ** 2 == 25
    # sqrt carries half powers (the reference's unit parser drops them: its test_units_6 fails upstream)
    assert float(np.sqrt(l)) == pytest.approx(np.sqrt(5 * 149597870700), rel=1e-15)
    assert np.sqrt(l).units == {"L": 0.5}
    assert float(phys.Measurement(0, "m**1") + np.sqrt(l)) == pytest.approx(np.sqrt(149597870700 * 5), rel=1e-15)
    assert str(phys.light.c) == "299792458.0" and str(phys.light.h) == "6.62607015E-34"  # what kernels get spliced


def test_code_scale_changes_constants():
    import importlib

    phys.Measurement.set_code_scale("m", 1e-3)
    try:
        assert float(phys.Measurement(1.0, "m**1")) == 1e-3
        mod = importlib.reload(phys.light)
        assert float(mod.c) == pytest.approx(299792.458)
    finally:
        phys.Measurement.reset_code_scale("m")
        importlib.reload(phys.light)
    assert float(phys.light.c) == 299792458.0


# ---- Simulation runtime (physicl/__init__.py:400-541) -----------------------------------------
class Recorder(phys.Step):
    touches_objects = False

    def __init__(self, log, tag):
        self.log, self.tag, self.done = log, tag, False

    def run(self, sim):
        self.log.append(self.tag)

    def terminate(self, sim):
        self.done = True


def test_simulation_runs_steps_in_insertion_order_until_exit():
    log = []
    s = phys.Simulation(cl_on=False, exit=lambda c: c.t >= 0.003)
    s.add_step(3, phys.UpdateTimeStep(lambda c: np.double(0.001)))
    a, b = Recorder(log, "a"), Recorder(log, "b")
    s.add_step(1, b)
    s.add_step(0, a)
    s.start()
    s.join()
    assert log == ["b", "a"] * 3  # insertion order, not idx order
    assert len(s.ts) == 3 and s.ts[-1] == pytest.approx(0.003) and not s.running and a.done and b.done
    assert s.run_time >= 0


def test_add_step_duplicate_and_remove_while_running():
    s = phys.Simulation(cl_on=False)
    s.add_step(0, phys.Step())
    with pytest.raises(NameError):
        s.add_step(0, phys.Step())
    s.running = True
    with pytest.raises(RuntimeError):
        s.remove_step(0)
    s.running = False
    s.remove_step(0)
    assert s.steps == {}


def test_default_exit_and_object_list():
    s = phys.Simulation(cl_on=False)
    o = phys.Object()
    s.add_obj(o)
    s.add_objs([phys.Object(), phys.Object()])
    assert len(s.objects) == 3 and not s.exit(s)
    s.remove_obj(o)
    assert len(s.objects) == 2 and s.get_state()["objects"] == 2

    class Drop(phys.Step):
        def run(self, sim):
            sim.remove_obj(sim.objects[0])

    s.add_step(0, Drop())
    s.start()
    s.join()
    assert len(s.objects) == 0


def test_step_errors_surface_on_join():
    class Boom(phys.Step):
        def run(self, sim):
            raise ValueError("boom")

    s = phys.Simulation(cl_on=False, exit=lambda c: False)
    s.add_step(0, Boom())
    s.start()
    with pytest.raises(ValueError, match="boom"):
        s.join()
    assert not s.running


def test_measure_step_writes_rows(tmp_path):
    fn = tmp_path / "out.csv"
    m = phys.MeasureStep(str(fn))
    m.data.append(np.array([0.001, 10, 4]))
    m.data.append(np.array([0.002, 9, 5]))
    m.terminate(None)
    lines = fn.read_text().strip().split("\n")
    assert lines[0] == "0.001, 10.0, 4.0" and len(lines) == 2


def test_device_steps_refuse_to_run_without_device():
    s = phys.Simulation(cl_on=False, exit=lambda c: c.t >= 0.001)
    s.add_obj(phys.Object())
    s.add_step(0, phys.UpdateTimeStep(lambda c: np.double(0.001)))
    s.add_step(1, phys.newton.NewtonianKinematicsStep())
    s.start()
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.join()


def test_variable_n_and_measure_E_are_in_scope():
    st = phys.light.ScatterIsotropicStep(variable_n=True, variable_n_fn="1.0", check_expression=False)
    assert st.variable_n and "PCL_USER_N_EXPR 1.0" in st._jit_source
    plan = fused.fuse_plan([phys.newton.NewtonianKinematicsStep(), st])
    assert isinstance(plan[0], fused.FusedPhotonStep) and plan[0].varn
    assert phys.light.ScatterMeasureStep(None, True, [], measure_E=True).needs_dr


def test_fuse_plan_groups_the_canonical_pipeline():
    upd = phys.UpdateTimeStep(lambda c: 1e-3)
    kin = phys.newton.NewtonianKinematicsStep()
    sc = phys.light.ScatterIsotropicStep(A=1e-3, n=1e-3)
    esc = phys.light.EscapeSphereStep(3e6)
    sign = phys.light.ScatterSignMeasureStep(None)
    pl = phys.light.ScatterMeasureStep(None, True, [[1e6, np.nan, np.nan]])
    user = Recorder([], "u")
    plan = fused.fuse_plan([upd, kin, sc, esc, sign, pl, user])
    assert plan[0] is upd and isinstance(plan[1], fused.FusedPhotonStep) and plan[2] is user and len(plan) == 3
    assert plan[1].members == [kin, sc, esc, sign, pl] and plan[1]._planes.count == 1
    # any other order stays unfused
    plan2 = fused.fuse_plan([upd, sc, kin, sign])
    assert plan2 == [upd, sc, kin, sign]
    plan3 = fused.fuse_plan([kin, user, sc])
    assert plan3 == [kin, user, sc]
    acc = phys.newton.NewtonianKinematicsStep(accel=True, a_uniform=[0, 0, -9.81])
    assert fused.fuse_plan([acc, sc]) == [acc, sc]


def test_scatter_params_fold_constants_in_float64():
    class G:
        e0 = 9.93e-19

    st = phys.light.ScatterIsotropicStep(A=np.double(5.1e-31 * (532e-9) ** 4), n=np.double(2.5e25), wavelength_dep_scattering=True)
    sp = st.scatter_params(G())
    hc = 6.62607015e-34 * 299792458.0
    want = 5.1e-31 * (532e-9) ** 4 * 2.5e25 * (G.e0 / hc) ** 4
    assert sp.mode == _capi.SCATTER_WAVELENGTH and sp.k == pytest.approx(want, rel=1e-6)
    assert 1e-12 < sp.k < 1e-2  # representable in binary32 although A ~ 4e-56 is not
    assert np.float32(5.1e-31 * (532e-9) ** 4) == 0.0


# ---- emission host path -------------------------------------------------------------------------
def test_planck_host_sampler_reproduces_reference_sequence(golden):
    g = golden("planck")
    np.random.seed(int(g["seed"]))
    phys.light.last_planck_params = None
    got = [phys.light.planck_phot_distribution(float(g["E_min"]), float(g["E_max"]), float(g["T"]), bins=int(g["bins"]))
           for _ in range(len(g["E"]))]
    none = np.array([v is None for v in got])
    assert np.array_equal(none, np.isnan(g["E"])) and none.sum() > 0
    vals = np.array([np.nan if v is None else float(v) for v in got])
    assert np.array_equal(vals[~none], g["E"][~none])
    np.testing.assert_allclose(phys.light.last_planck_cdf, g["cdf"], rtol=1e-11)
    pr = phys.light.planck_probability(float(g["E_min"]), float(g["E_max"]), float(g["T"]))
    assert 0 < pr[0] < 1


def test_generate_photons_shapes():
    ph = phys.light.generate_photons(5, min=1e-19, max=2e-19)
    assert len(ph) == 5 and all(type(p) is phys.light.PhotonObject for p in ph)
    assert all(1e-19 <= float(p.E) <= 2e-19 for p in ph)
    assert np.array_equal(np.asarray(ph[0].v), [float(phys.light.c), 0, 0])
    ph2 = phys.light.generate_photons_from_E([1.0, 2.0])
    assert [float(p.E) for p in ph2] == [1.0, 2.0]


# ---- C ABI surface ---------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "physicl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_capi.EXPORTS)
    assert _capi.load().pcl_abi_version() == 1


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_capi.Soa) == 8 + 15 * 8 + 8 + 8
    assert ctypes.sizeof(_capi.Pingpong) == 2 * ctypes.sizeof(_capi.Soa) + 8 + 4 + 4
    assert ctypes.sizeof(_capi.ScatterParams) == 16
    assert ctypes.sizeof(_capi.Rng) == 8 + 4 + 4 + 3 * 8
    assert ctypes.sizeof(_capi.Planes) == 4 + 4 * 8 + 4 * 8
    assert _capi.make_planes([(0, 1.0), (2, -3.5)]).count == 2
    with pytest.raises(ValueError):
        _capi.make_planes([(0, 0.0)] * 9)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    with pytest.raises(_capi.PclError, match="no CUDA device|failed"):
        _capi.Context(0)
    with pytest.raises(_capi.PclError):
        phys.Simulation(cl_on=True)


def test_object_attribute_gather_fast_and_fallback_paths():
    """_gather3 (the store is built from sim.objects with it): Measurements, plain arrays, lists and (3, 1) columns all
    end up as one (3, N) float64 array, whichever conversion path is taken."""
    from physicl_b200 import _gather3

    class O:
        pass

    vals = [phys.Measurement([1, 2, 3], "m**1"), np.array([4.0, 5.0, 6.0]), [7, 8, 9], np.array([[10.0], [11.0], [12.0]])]
    objs = []
    for v in vals:
        o = O()
        o.r = v
        objs.append(o)
    want = np.array([[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]], np.float64)
    assert np.array_equal(_gather3(objs[:2], "r"), want[:, :2])  # fast path: every entry is a (3,) array
    assert np.array_equal(_gather3(objs, "r"), want)             # ragged shapes: per-object conversion
    assert _gather3([], "r").shape == (3, 0)


def test_api_surface_covers_the_reference():
    """Every module-level name of the reference's physicl/__init__.py, light.py and newton.py exists here, every
    method of every class too, with the reference's positional parameters first and in the same order (so a call
    written for the reference binds the same way).  tests/golden/api_surface.json is extracted from the reference's
    source by tests/golden/make_api_surface.py."""
    import inspect
    import json
    import os

    import physicl_b200.light
    import physicl_b200.newton

    surface = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "api_surface.json")))
    mods = {"__init__": phys, "light": phys.light, "newton": phys.newton}
    # module-level names that are the reference's own imports / pyopencl plumbing, not API
    problems = []

    def check_sig(where, fn, want):
        try:
            have = list(inspect.signature(fn).parameters)
        except (TypeError, ValueError):
            return
        renamed = {("Measurement", "__scale__"), ("Measurement", "__array_finalize__")}  # positional-only in practice
        if have[:len(want["args"])] != want["args"] and where not in renamed:
            problems.append(f"{where}: parameters {have} do not start with {want['args']}")

    for mod, names in surface.items():
        m = mods[mod]
        for name, info in names.items():
            if not hasattr(m, name):
                problems.append(f"{mod}.{name} is missing")
                continue
            obj = getattr(m, name)
            if info["kind"] == "function":
                check_sig((mod, name), obj, info)
            elif info["kind"] == "class":
                for meth, want in info["methods"].items():
                    mangled = meth if not (meth.startswith("__") and not meth.endswith("__")) else f"_{name}{meth}"
                    if name == "Measurement" and meth.startswith("__") and not meth.endswith("__"):
                        continue  # private helpers of the units parser (rewritten here)
                    if not hasattr(obj, mangled) and not any(hasattr(b, f"_{b.__name__}{meth}") for b in obj.__mro__):
                        problems.append(f"{mod}.{name}.{meth} is missing")
                        continue
                    if hasattr(obj, mangled):
                        check_sig((name, meth), getattr(obj, mangled), want)
    assert not problems, "\n".join(problems)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "physicl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


# ---- sharding ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p", [(0, 4), (7, 8), (1000, 3), (16 * 2 ** 20, 8), (10 ** 9, 8)])
def test_shard_range_partitions_exactly(n, p):
    spans = [dist.shard_range(n, r, p) for r in range(p)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_sharded_oracle_steps_equal_unsharded():
    """Why sharding needs no data-path collective: the Philox counter is the GLOBAL particle id, so
    P independent shards reproduce the single-shard run bit for bit and tallies add up."""
    import oracle

    n, c, dt, k = 10_001, 299792458.0, 1e-3, 1.1e-6
    rng = np.random.default_rng(0)
    d = rng.normal(size=(3, n))
    v = (c * d / np.linalg.norm(d, axis=0)).astype(np.float32)
    base = {nm: a.copy() for nm, a in zip(("x", "y", "z", "vx", "vy", "vz"), list(np.zeros((3, n), np.float32)) + list(v))}
    whole = {nm: a.copy() for nm, a in base.items()}
    rows = [oracle.photon_step_f32(whole, dt, k, c, seed=3, step=s, r2_escape=np.float32(6e5 ** 2)) for s in range(4)]
    for p in (2, 3, 8):
        parts = []
        tot = np.zeros((4, 16), np.int64)
        for r in range(p):
            lo, hi = dist.shard_range(n, r, p)
            sh = {nm: a[lo:hi].copy() for nm, a in base.items()}
            for s in range(4):
                tot[s] += oracle.photon_step_f32(sh, dt, k, c, seed=3, step=s, r2_escape=np.float32(6e5 ** 2), id_base=lo)
            parts.append(sh)
        assert np.array_equal(tot, np.array(rows))
        for nm in base:
            assert np.array_equal(np.concatenate([q[nm] for q in parts]).view(np.uint32), whole[nm].view(np.uint32))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as td

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = dist.shard_range(1001)
        rows = np.zeros((3, 16), np.int64)
        rows[:, 0] = hi - lo
        rows[:, 1] = rank + 1
        total = dist.all_reduce_rows(rows)
        q.put((rank, lo, hi, total.tolist(), dist.all_reduce_int(hi - lo)))
    finally:
        td.destroy_process_group()


def test_gloo_world2_tally_reduction():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert [o[1:3] for o in out] == [(0, 501), (501, 1001)]
    for o in out:
        assert o[3][0][0] == 1001 and o[3][2][1] == 3 and o[4] == 1001


def test_fused_plan_is_cached_and_chunks_end_on_compaction_boundaries():
    """The fused step keeps its state (cadence, pinned feedback buffers) across run_steps calls, and a chunk
    handed to the C ABI ends where a launch would compact anyway."""
    s = phys.Simulation(cl_on=False)
    s.cl_on = True  # plan construction only; nothing is launched
    upd = phys.UpdateTimeStep(lambda c: 1e-3)
    kin = phys.newton.NewtonianKinematicsStep()
    sc = phys.light.ScatterIsotropicStep(A=1e-3, n=1e-3)
    esc = phys.light.EscapeSphereStep(3e6)
    s.add_step(0, upd)
    s.add_step(1, kin)
    s.add_step(2, sc)
    s.add_step(3, esc)
    p1, p2 = s._plan(), s._plan()
    assert p1 is p2 and isinstance(p1[1], fused.FusedPhotonStep) and p1[1].retires
    s.add_step(4, phys.light.ScatterSignMeasureStep(None))
    p3 = s._plan()
    assert p3 is not p1 and len(p3) == 2 and p3[1].measures
    f = p3[1]
    assert s.feedback_every == 64 and f.cadence == 5
    for idx, want in ((0, 65), (5, 65), (7, 63), (8, 62)):  # (5 - idx % 5) + 5 * (round(64 / 5) - 1)
        s.step_index = idx
        assert f.chunk_steps(s) == want and (idx + want) % 5 == 0
    s.compact_cadence = 1
    assert f.chunk_steps(s) == 64
    s.compact_cadence, s.feedback_every = None, 8
    f.cadence = 8
    s.step_index = 3
    assert f.chunk_steps(s) == 8  # m >= feedback_every: the chunk is feedback_every long
    # pipelines that never retire photons take whole chunks
    t = phys.Simulation(cl_on=False)
    t.cl_on = True
    t.add_step(0, upd)
    t.add_step(1, phys.newton.NewtonianKinematicsStep())
    t.add_step(2, phys.light.ScatterIsotropicStep(A=1e-3, n=1e-3))
    g = t._plan()[1]
    assert not g.retires and g.chunk_steps(t) == 64
    assert kin.chunk_steps(s) == 256 and kin.can_run_many(s)


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/physicl_b200.h is a C header (no C++ constructs): a C99 translation unit that takes the address of
    every declared entry point compiles with gcc and links against the shared library."""
    import subprocess

    hdr = open(os.path.join(REPO, "include", "physicl_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", hdr)) - {"pcl_ctx"})
    assert len(names) >= 30
    src = tmp_path / "abi.c"
    body = "\n".join("    p[%d] = (void (*)(void))%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "physicl_b200.h"\n#include <stdio.h>\nint main(void) {\n    void (*p[%d])(void);\n%s\n'
                   '    printf("%%d %%d\\n", %d, pcl_abi_version());\n    return p[0] == 0;\n}\n' % (len(names), body, len(names)))
    exe = tmp_path / "abi"
    lib = os.path.join(REPO, "physicl_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(REPO, "include"), str(src),
                        "-o", str(exe), "-L", lib, "-l:libphysicl_b200.so", "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == [str(len(names)), "1"]


# ---- the chunked main loop (Simulation._run_chunked) against a scripted device -----------------------------------------
class _ScriptedStore:
    """Stands in for DeviceParticleStore: tally rows of a population that loses a fixed share per timestep."""

    def __init__(self, n):
        self.n0 = n
        self.alive = n
        self.rows = []
        self.restored = 0

    @property
    def current_row(self):
        return len(self.rows) - 1

    def read_row(self, r):
        return self.rows[r]

    def checkpoint(self, kind):
        return (self.alive, len(self.rows))

    def restore(self, kind, ck):
        self.alive = ck[0]  # rows of the discarded timesteps stay in the table, as on the device
        self.restored += 1


class _ScriptedFused(phys.Step):
    uses_device = True
    tallies_every_timestep = True

    def __init__(self, store, chunk):
        self.store, self.chunk, self.calls = store, chunk, []

    def can_run_many(self, sim):
        return True

    def chunk_steps(self, sim):
        return self.chunk

    def checkpoint(self):
        return None

    def rollback(self, ck):
        pass

    def run_many(self, sim, k, dt, ts):
        self.calls.append(k)
        for _ in range(k):
            self.store.alive = (self.store.alive * 9) // 10
            row = np.zeros(16, np.int64)
            row[0] = self.store.alive
            self.store.rows.append(row)
        sim._mark_device_dirty()


def _scripted_sim(exit_fn, chunk, n=1000, dt_fn=None):
    sim = phys.Simulation(cl_on=False, exit=exit_fn)
    st = _ScriptedStore(n)
    fused = _ScriptedFused(st, chunk)
    upd = phys.UpdateTimeStep(dt_fn or (lambda s: np.double(0.5)))
    sim.cl_on = True  # only so that _bulk_plan accepts the pair; no context is ever touched
    sim._plan = lambda: [upd, fused]
    sim.device_store = lambda: st
    sim.store = st
    sim._device_live_count = lambda: st.alive
    sim._device_dirty = True  # the scripted device is the authority for len(sim.objects)
    sim.run()
    if sim.error:
        raise sim.error
    return sim, st, fused


@pytest.mark.parametrize("chunk", [1, 4, 64])
def test_chunked_loop_time_predicate_looks_ahead(chunk):
    """exit on t: evaluated on the host after every UpdateTimeStep of the chunk; exactly 7 timesteps run, no roll-back."""
    sim, st, fused = _scripted_sim(lambda s: s.t >= 3.5, chunk)
    assert len(sim.ts) == 7 and float(sim.t) == 3.5 and sim.step_index == 7 and len(st.rows) == 7
    assert st.restored == 0 and sum(fused.calls) == 7 and max(fused.calls) <= chunk


@pytest.mark.parametrize("chunk", [1, 4, 64])
def test_chunked_loop_particle_predicate_rolls_back(chunk):
    """exit on len(objects): checked against the tally row of every timestep of the chunk; the loop ends after the FIRST
    timestep at which it holds (1000 * 0.9^k <= 500 first at k = 7), whatever the chunk size."""
    sim, st, fused = _scripted_sim(lambda s: len(s.objects) <= 500, chunk)
    assert len(sim.ts) == 7 and sim.step_index == 7 and st.alive == 477
    assert float(sim.t) == 3.5
    assert st.restored == (0 if chunk == 1 else 1)  # chunk 4: the second chunk (timesteps 5-8) overshoots by one; 64: by 57


def test_chunked_loop_mixed_predicate_and_changing_dt():
    """A predicate that starts looking at the particles only once t is large enough, and a dt that changes."""
    dt_fn = lambda s: np.double(0.25 if len(s.ts) < 3 else 0.5)  # noqa: E731
    sim, st, fused = _scripted_sim(lambda s: s.t >= 1.0 and len(s.objects) <= 600, 16, dt_fn=dt_fn)
    ref, st1, _ = _scripted_sim(lambda s: s.t >= 1.0 and len(s.objects) <= 600, 1, dt_fn=dt_fn)
    assert [float(t) for t in sim.ts] == [float(t) for t in ref.ts] and st.alive == st1.alive and sim.step_index == ref.step_index
    assert st.alive <= 600 and len(sim.ts) == 5
