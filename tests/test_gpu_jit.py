"""GPU tests of the run-time compiled (NVRTC, sm_100a) variable-density photon kernels
(ScatterIsotropicStep(variable_n=True), physicl/light.py:295-299).

Bar: scatter decisions and sign tallies identical to runs of the unmodified reference
(tests/golden/varn*.npz, injected uniforms); velocities within 1e-5*c and positions within
1e-5*c*dt per step of its float64 state; fused and stand-alone forms bit-identical to each other."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import reference_law as law

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from physicl_b200 import _capi

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = _capi.Context(0)
    yield c
    c.close()


def _u():
    import gpu_util

    return gpu_util


def _split_u(uu):
    uu = uu.reshape(-1, 3)
    return [np.ascontiguousarray(uu[:, i], np.float32) for i in range(3)]


def _varn(gd, g):
    from physicl_b200 import _capi

    c, wave = float(gd["c"]), bool(gd["wave"])
    kd = float(gd["kernel_A"])
    if wave:
        kd *= (g.e0 / (float(gd["h"]) * c)) ** 4
    vn = _capi.VarnParams(kd=kd, e0=g.e0, a_slot=float(gd["kernel_A"]), n_slot=float(gd["kernel_n"]))
    sp = _capi.ScatterParams(k=0.0, c=c, mode=_capi.SCATTER_WAVELENGTH if wave else 0)
    return vn, sp


def _jit_step(ctx, mod, st, g, dt, sp, vn, seed=0, step=0, uniforms=None, r2_escape=0.0, planes=None, nsteps=1):
    from physicl_b200 import _capi

    rg = _capi.Rng(seed=seed, step=step)
    keep = None
    if uniforms is not None:
        keep = [torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(st.device) for x in uniforms]
        rg.u_theta, rg.u_phi, rg.u_rand = (t.data_ptr() for t in keep)
    pl = _capi.make_planes(planes)
    first = st.new_rows(nsteps)
    soa = g.soa()
    soa.dx = soa.dy = soa.dz = None
    ctx.call("pcl_photon_steps_jit", st.stream(), mod.kernel("pcl_jit_photon_step"), C.byref(soa), C.c_float(dt), C.byref(sp),
             C.byref(vn), C.byref(rg), C.c_float(r2_escape), C.byref(pl), st.row_ptr(first), C.c_uint32(nsteps))
    st.synchronize()
    return [st.read_row(first + i) for i in range(nsteps)]


@pytest.mark.parametrize("name", ["varn", "varn_z"])
def test_variable_n_fused_step_tracks_reference_golden(ctx, golden, name):
    from physicl_b200 import _capi, jit

    u = _u()
    gd = golden(name)
    N, c, dt = int(gd["N"]), float(gd["c"]), float(gd["dt"])
    r, v = u.beam_photons(N, c)
    st, g = u.make_store(ctx, r, v, E=gd["E"])
    mod = jit.Module(ctx, jit.photon_source(str(gd["expr"]), bool(gd["wave"])))
    vn, sp = _varn(gd, g)
    for s in range(int(gd["nsteps"])):
        row = _jit_step(ctx, mod, st, g, dt, sp, vn, uniforms=_split_u(gd["s%d_u" % s]))[0]
        ref_hit = ~np.isnan(gd["s%d_res0" % s])
        assert int(row[_capi.T_SCATTERED]) == int(ref_hit.sum())  # decisions identical
        vv = np.stack([g.download(nm) for nm in ("vx", "vy", "vz")]).astype(np.float64)
        rr = np.stack([g.download(nm) for nm in ("x", "y", "z")]).astype(np.float64)
        assert np.abs(vv - gd["s%d_v" % s]).max() <= 1e-5 * c
        assert np.abs(rr - gd["s%d_r" % s]).max() <= 1e-5 * c * dt * (s + 1)
        srow = gd["sign_rows"][s]
        assert [int(row[q]) for q in (_capi.T_ALIVE, _capi.T_XP, _capi.T_YP, _capi.T_ZP)] == [int(q) for q in srow[1:5]]
    mod.close()


def test_variable_n_philox_decisions_match_float64_law(ctx, golden):
    """In-kernel Philox draws, 300k photons at spread-out positions (so n(r) spans decades), with the
    escape sphere and a plane tally: decisions against the float64 restatement of the reference's law on
    the same state; ties closer than 1e-6 relative are excluded (there are none to speak of)."""
    from physicl_b200 import _capi, jit

    u = _u()
    gd = golden("varn")
    n, c, dt, seed = 300_007, float(gd["c"]), 1e-5, 99
    rng = np.random.default_rng(5)
    r, v = u.random_photons(n, seed=21, spread=2.0e4)
    E = rng.uniform(float(gd["E"].min()), float(gd["E"].max()), n)
    st, g = u.make_store(ctx, r, v, E=E)
    mod = jit.Module(ctx, jit.photon_source(str(gd["expr"]), True))
    vn, sp = _varn(gd, g)
    hc = float(gd["h"]) * c
    r2 = 3.0e4 ** 2
    for step in range(3):
        host = u.host_state(g)
        live = ~np.isnan(host["x"])
        row = _jit_step(ctx, mod, st, g, dt, sp, vn, seed=seed, step=step, r2_escape=r2, planes=[(0, 1.0e3)])[0]
        tw = {k2: a.copy() for k2, a in host.items()}
        oracle.kinematics_f32(tw, dt)  # binary32 twin of the kernel's r += v*dt
        rr = np.stack([tw[q] for q in ("x", "y", "z")]).astype(np.float64)
        dr = np.stack([host[q].astype(np.float32) * np.float32(dt) for q in ("vx", "vy", "vz")]).astype(np.float64)
        Eh = host["e"].astype(np.float64) * g.e0
        p = law.pcoll_variable_n(dr, rr, str(gd["expr"]), float(gd["kernel_A"]), float(gd["kernel_n"]), Eh, hc)
        ut, up, ur = oracle.philox_uniforms(n, 0, seed, step)
        hit = live & (p >= ur)
        sure = np.abs(p - ur) > 1e-6 * np.maximum(p, ur)
        vx_new = g.download("vx")
        changed = u.bits(vx_new) != u.bits(host["vx"])
        got_hit = changed  # a scattered photon gets a fresh direction (same vx bits: probability ~1e-7)
        assert np.array_equal(got_hit[sure & live], hit[sure & live])
        assert abs(int(row[_capi.T_SCATTERED]) - int(hit.sum())) <= int((~sure).sum())
        assert 0.02 * live.sum() < hit.sum() < 0.98 * live.sum()
        esc = live & ((rr ** 2).sum(0) >= r2)
        assert abs(int(row[_capi.T_ESCAPED]) - int(esc.sum())) <= 2  # float32 r^2 vs float64 at the boundary
        assert int(row[_capi.T_LIVE_IN]) == int(live.sum())
    mod.close()


def test_variable_n_scatter_alone_equals_fused(ctx, golden):
    """kinematics, then the stand-alone run-time scatter on the dr planes == the fused run-time step."""
    from physicl_b200 import _capi, jit

    u = _u()
    gd = golden("varn_z")
    n, c, dt = 100_003, float(gd["c"]), 1e-3
    r, v = u.random_photons(n, seed=4, spread=3.0e6)
    stA, gA = u.make_store(ctx, r, v, nscat=True)
    stB, gB = u.make_store(ctx, r, v, nscat=True)
    gB.ensure("dx", "dy", "dz")
    mod = jit.Module(ctx, jit.photon_source(str(gd["expr"]), False))
    vn, sp = _varn(gd, gA)
    for step in range(4):
        rowA = _jit_step(ctx, mod, stA, gA, dt, sp, vn, seed=11, step=step)[0]
        soa = gB.soa()
        ctx.call("pcl_kinematics", stB.stream(), C.byref(soa), C.c_float(dt), 0, None)
        rg = _capi.Rng(seed=11, step=step)
        flags = torch.empty(n, dtype=torch.int32, device=stB.device)
        r1 = stB.new_row()
        ctx.call("pcl_scatter_jit", stB.stream(), mod.kernel("pcl_jit_scatter"), C.byref(soa), C.byref(sp), C.byref(vn),
                 C.byref(rg), C.c_void_p(flags.data_ptr()), stB.row_ptr())
        rowB = stB.read_row(r1)
        assert rowB[_capi.T_SCATTERED] == rowA[_capi.T_SCATTERED] == int(flags.sum().item())
        assert 0 < rowA[_capi.T_SCATTERED] < n
    for nm in u.PLANE_NAMES + ("nscat",):
        assert u.same_bits(gA.download(nm), gB.download(nm)), nm
    assert gA.download("nscat").sum() > 0
    mod.close()


def test_variable_n_many_steps_equal_single_steps_and_unaligned_view(ctx, golden):
    from physicl_b200 import jit

    u = _u()
    gd = golden("varn_z")
    n, c, dt = 65_539, float(gd["c"]), 1e-3
    r, v = u.random_photons(n, seed=6, spread=3.0e6)
    mod = jit.Module(ctx, jit.photon_source(str(gd["expr"]), False))
    stA, gA = u.make_store(ctx, r, v)
    stB, gB = u.make_store(ctx, r, v)
    vn, sp = _varn(gd, gA)
    rowsA = _jit_step(ctx, mod, stA, gA, dt, sp, vn, seed=3, step=0, nsteps=5)
    rowsB = [_jit_step(ctx, mod, stB, gB, dt, sp, vn, seed=3, step=s)[0] for s in range(5)]
    assert np.array_equal(np.array(rowsA), np.array(rowsB))
    for nm in u.PLANE_NAMES:
        assert u.same_bits(gA.download(nm), gB.download(nm)), nm
    # a view that starts one slot in: 4-byte aligned only -> scalar path, same bits as slots 1.. of a fresh run
    stC, gC = u.make_store(ctx, r, v)
    stD, gD = u.make_store(ctx, r[:, 1:], v[:, 1:], id_base=1)
    soa = gC.soa()
    for nm in ("x", "y", "z", "vx", "vy", "vz"):
        setattr(soa, nm, getattr(soa, nm) + 4)
    soa.n, soa.id_base = n - 1, 1
    soa.dx = soa.dy = soa.dz = None
    from physicl_b200 import _capi

    rg = _capi.Rng(seed=3, step=0)
    pl = _capi.make_planes(None)
    row = stC.new_row()
    ctx.call("pcl_photon_steps_jit", stC.stream(), mod.kernel("pcl_jit_photon_step"), C.byref(soa), C.c_float(dt), C.byref(sp),
             C.byref(vn), C.byref(rg), C.c_float(0.0), C.byref(pl), stC.row_ptr(), C.c_uint32(1))
    stC.synchronize()
    rowD = _jit_step(ctx, mod, stD, gD, dt, sp, vn, seed=3, step=0)[0]
    assert np.array_equal(stC.read_row(row), rowD)
    for nm in u.PLANE_NAMES:
        assert u.same_bits(gC.download(nm)[1:], gD.download(nm)), nm
    mod.close()


def test_build_errors_carry_the_compiler_log(ctx):
    from physicl_b200 import _capi, jit

    with pytest.raises(_capi.PclError, match="nosuchfn"):
        jit.Module(ctx, jit.photon_source("nosuchfn(r0[gid])", False))
    mod = jit.Module(ctx, jit.photon_source("1.0", False))
    with pytest.raises(_capi.PclError, match="no_such_kernel"):
        mod.kernel("no_such_kernel")
    mod.close()


# ---- through the public API --------------------------------------------------------------------------
def _sim(n, expr, wave, split, seed=5):
    import physicl_b200 as phys
    import physicl_b200.light
    import physicl_b200.newton

    s = phys.Simulation(cl_on=True, seed=seed, exit=lambda x: False)
    c = float(phys.light.c)
    rng = np.random.default_rng(2)
    d = rng.normal(size=(3, n))
    v = (c * d / np.linalg.norm(d, axis=0)).astype(np.float32)
    r = rng.uniform(-5e3, 5e3, (3, n)).astype(np.float32)
    E = rng.uniform(1e-19, 9e-19, n)
    s.add_particles(r, v, E=E)
    s.add_step(0, phys.UpdateTimeStep(lambda x: np.double(1e-5)))
    s.add_step(1, phys.newton.NewtonianKinematicsStep())
    if split:  # a host step between kinematics and scatter: the plan cannot fuse them

        class Nop(phys.Step):
            uses_device = True

            def run(self, sim):
                pass

        s.add_step(5, Nop())
    s.add_step(2, phys.light.ScatterIsotropicStep(n=np.double(5.1e-31 * (532e-9) ** 4), A=np.double(1.0), variable_n=True,
                                                  variable_n_fn=expr, wavelength_dep_scattering=wave, seed=77))
    sign = phys.light.ScatterSignMeasureStep(None, True)
    s.add_step(3, sign)
    return s, sign


def test_simulation_variable_n_fused_equals_unfused():
    """The example pipeline of examples/presentation_example.ipynb (radial atmosphere, Rayleigh law):
    fused and step-by-step plans give the same rows and the same state, bit for bit."""
    expr = "{} * exp(-1 * ({} - {})/({}))".format(6.0e26, "sqrt(pow(r0[gid], 2) + pow(r1[gid], 2) + pow(r2[gid], 2))", 1000.0, 8000.0)
    a, sa = _sim(50_001, expr, True, split=False)
    b, sb = _sim(50_001, expr, True, split=True)
    a.run_steps(6)
    b.run_steps(6)
    ra, rb = np.array(sa.data), np.array(sb.data)
    assert np.array_equal(ra, rb) and ra.shape == (6, 5)
    ga, gb = a.store.group("photon"), b.store.group("photon")
    for nm in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(ga.download(nm).view(np.uint32), gb.download(nm).view(np.uint32)), nm
    vx0 = float(ra[0][2]) / ra[0][1]
    assert 0.3 < vx0 < 0.7 and len({tuple(x[2:]) for x in ra}) > 1  # something happens from step to step


# ---- user kernels through CLProgram ----------------------------------------------------------------
def test_user_program_matches_the_reference_run(golden):
    """The reference's own CLProgram ran this user kernel (tests/golden/make_golden.py:gen_clprogram);
    here the same declarations compile for sm_100a and give the same arrays: the int flags exactly, the
    float64 energies to 1e-14 (device exp/sqrt/pow against the host libm the reference kernel used)."""
    import physicl_b200 as physicl
    from test_jit_host import build_user_program, user_objects

    gd = golden("clprogram")
    sim = physicl.Simulation(cl_on=True, exit=lambda s: True)
    sim.add_objs(user_objects(gd))
    prog = build_user_program(sim, gd)
    prog.build_kernel()
    np.random.seed(int(gd["seed"]))
    out = prog.run()
    assert out["flag"].dtype == np.int32 and np.array_equal(out["flag"], gd["flag"])
    cut = gd["flag"] == 0
    assert np.all(np.isnan(out["ke"][cut]))
    assert np.allclose(out["ke"][~cut], gd["ke"][~cut], rtol=1e-14, atol=0)
    assert [o.gid for o in prog.who] == list(gd["tracked_gid"])
    out2 = prog.run()  # second run: new jitter draws, same structure, module reused
    assert np.array_equal(out2["flag"], gd["flag"]) and not np.array_equal(out2["ke"][~cut], out["ke"][~cut])


def test_user_program_inside_a_step_next_to_device_steps():
    """A user Step that runs its own kernel on the objects each timestep (the reference's extension
    story, README.md:8), mixed with the device kinematics step: the objects it sees are the ones the
    device moved."""
    import physicl_b200 as physicl
    import physicl_b200.newton

    class Height(physicl.Step):
        def __init__(self):
            self.rows, self.prog = [], None

        def run(self, sim):
            if self.prog is None:
                self.prog = physicl.CLProgram(sim, "height", "int gid = get_global_id(0); h[gid] = z[gid] - floor_z; up[gid] = vz[gid] > 0 ? 1 : 0;")
                self.prog.prep_metadata = [physicl.CLInput(name="z", type="obj", obj_attr="r[2]"), physicl.CLInput(name="vz", type="obj", obj_attr="v[2]"),
                                           physicl.CLInput(name="floor_z", type="const", const_value="-5.0")]
                self.prog.output_metadata = [physicl.CLOutput(name="h"), physicl.CLOutput(name="up", ctype="int")]
                self.prog.build_kernel()
            self.rows.append(self.prog.run())

    sim = physicl.Simulation(cl_on=True, exit=lambda s: s.t >= 0.0029)
    rng = np.random.default_rng(1)
    z0, vz0 = rng.uniform(0, 10, 64), rng.normal(0, 5, 64)
    for i in range(64):
        o = physicl.Object()
        o.r = physicl.Measurement([0.0, 0.0, float(z0[i])], "m**1")
        o.v = physicl.Measurement([1.0, 0.0, float(vz0[i])], "m**1 s**-1")
        sim.add_obj(o)
    h = Height()
    sim.add_step(0, physicl.UpdateTimeStep(lambda s: np.double(0.001)))
    sim.add_step(1, physicl.newton.NewtonianKinematicsStep())
    sim.add_step(2, h)
    sim.start()
    sim.join()
    assert len(h.rows) == 3
    for k, row in enumerate(h.rows):
        want = np.float32(z0) + (k + 1) * np.float32(vz0) * np.float32(0.001) + 5.0
        assert np.allclose(row["h"], want, rtol=0, atol=1e-5)
        assert np.array_equal(row["up"], (np.float32(vz0) > 0).astype(np.int32))


def test_a_step_written_like_the_references_own_delete_step(golden):
    """A user step that declares its kernel through CLInput / CLOutput / CLProgram the way the reference's
    ScatterDeleteStep does (physicl/light.py:231-260: type filter, dr inputs, one np.random draw per photon,
    swapped constants, int flags, sim.remove_obj per flagged photon), between the device kinematics step and a
    device measure step.  With the reference's seed the run reproduces the reference's run: same flags, same
    survivors in the same order, same plane-crossing rows (tests/golden/delete.npz)."""
    import physicl_b200 as physicl
    import physicl_b200.light
    import physicl_b200.newton

    gd = golden("delete")

    class UserDeleteStep(physicl.Step):
        def __init__(self, n, A):
            self.n, self.A, self.built, self.flags = n, A, False, []

        def run(self, sim):
            if not self.built:
                skip = physicl.CLInput(name="photon_check", type="obj_action",
                                       code="if type(obj) != physicl.light.PhotonObject:\n \t\t continue")
                d0, d1, d2 = [physicl.CLInput(name="d" + str(x), type="obj", obj_attr="dr[" + str(x) + "]") for x in range(3)]
                rand = physicl.CLInput(name="rand", type="obj_def", obj_def="np.random.random()")
                A_ = physicl.CLInput(name="A", type="const", const_value=str(self.n))
                n_ = physicl.CLInput(name="n", type="const", const_value=str(self.A))
                pht = physicl.CLInput(name="pht", type="obj_track", obj_track="obj")
                res = physicl.CLOutput(name="res", ctype="int")
                kernel = """
                    int gid = get_global_id(0);
                    double norm = sqrt(pow(d0[gid], 2) + pow(d1[gid], 2) + pow(d2[gid], 2));
                    double pcoll = A * n * norm;
                    if (pcoll >= rand[gid]){
                        res[gid] = 1;
                    } else {
                        res[gid] = 0;
                    }
                """
                self.prog = physicl.CLProgram(sim, "test", kernel)
                self.prog.prep_metadata = [skip, d0, d1, d2, rand, pht, A_, n_]
                self.prog.output_metadata = [res]
                self.prog.build_kernel()
                self.built = True
            out = self.prog.run()
            self.flags.append(out["res"].copy())
            for idx, x in enumerate(out["res"]):
                if x == 1:
                    sim.remove_obj(self.prog.pht[idx])

    np.random.seed(int(gd["seed"]))
    nsteps = int(gd["nsteps"])
    sim = physicl.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=True, exit=lambda c: len(c.ts) >= nsteps)
    for i in range(int(gd["N"])):
        p = physicl.light.PhotonObject(s=np.zeros(3), v=np.array([physicl.light.c, 0, 0], dtype=np.double), E=np.double(1))
        p.gid = i
        sim.add_obj(p)
    step = UserDeleteStep(np.double(gd["n"]), np.double(gd["A"]))
    plane = physicl.light.ScatterMeasureStep(None, True, [np.array(gd["planes"][0])])
    survivors = []

    class Snapshot(physicl.Step):
        def run(self, sim):
            survivors.append([o.gid for o in sim.objects])

    sim.add_step(0, physicl.UpdateTimeStep(lambda s: np.double(float(gd["dt"]))))
    sim.add_step(1, physicl.newton.NewtonianKinematicsStep())
    sim.add_step(2, step)
    sim.add_step(3, plane)
    sim.add_step(4, Snapshot())
    sim.start()
    sim.join()
    assert len(step.flags) == nsteps
    for s in range(nsteps):
        assert np.array_equal(step.flags[s], gd["s%d_flags" % s]), s
        assert survivors[s] == list(gd["s%d_gid" % s]), s
    assert np.array_equal(np.array(plane.data)[:, 1:], gd["plane_rows"][:, 1:])
    assert 0 < len(sim.objects) < int(gd["N"])
