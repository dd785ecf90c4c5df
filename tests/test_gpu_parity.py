"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's
golden vectors.  Bar: integer / decision / tally work bit-exact; float state bit-exact against the
binary32 twin and within 1e-5 (relative to |v| = c, resp. the path length) of the float64 reference.
"""
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


# ---------------------------------------------------------------------------------------------
# kinematics (newton.py:14-16)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 3, 4, 1023, 4096 + 5, 1_000_003])
@pytest.mark.parametrize("accel", [0, 1, 2])
def test_kinematics_bit_exact_vs_twin(ctx, n, accel):
    u = _u()
    rng = np.random.default_rng(n + accel)
    r = rng.uniform(-1e3, 1e3, (3, n))
    v = rng.normal(0, 10, (3, n))
    a = rng.normal(0, 9.81, (3, n)) if accel == 1 else None
    st, g = u.make_store(ctx, r, v, a=a, kind="object")
    g.ensure("dx", "dy", "dz")
    host = u.host_state(g)
    au = np.array([0.0, 0.0, -9.81], np.float32)
    for s, dt in enumerate([1e-3, 2.5e-3, 0.5]):
        soa = g.soa()
        ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(dt), int(accel != 0),
                 au.ctypes.data_as(C.POINTER(C.c_float)) if accel == 2 else None)
        oracle.kinematics_f32(host, dt, accel, au if accel == 2 else None)
    for nm in host:
        assert u.same_bits(g.download(nm), host[nm]), nm


def test_kinematics_unaligned_view_takes_scalar_path(ctx):
    u = _u()
    n = 10_001
    rng = np.random.default_rng(1)
    st, g = u.make_store(ctx, rng.uniform(-1e3, 1e3, (3, n)), rng.normal(0, 10, (3, n)), kind="object")
    g.ensure("dx", "dy", "dz")
    host = {k: v[1:].copy() for k, v in u.host_state(g).items()}
    soa = g.soa(offset=1)  # every plane pointer is now 4 bytes past a 16-byte boundary
    ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(1e-3), 0, None)
    oracle.kinematics_f32(host, 1e-3)
    for nm in host:
        assert u.same_bits(g.download(nm)[1:], host[nm]), nm


@pytest.mark.parametrize("accel,with_dr,pinned", [(0, True, True), (1, True, True), (1, False, False), (2, True, False), (0, False, True)])
def test_kinematics_host_buffer_steps_equal_the_twin(ctx, accel, with_dr, pinned):
    """pcl_kinematics_steps_host (planes in HOST memory, chunked H2D -> k timesteps in registers -> D2H) against the
    binary32 twin, bit for bit; planes that are inputs only (a; v without accel) keep their bits; several chunks, a
    ragged last one, and guard bands around every host plane."""
    import torch

    n, pad, chunk = 300_007, 16, 65_536
    rng = np.random.default_rng(11 + accel)
    names = ["x", "y", "z", "vx", "vy", "vz"] + (["ax", "ay", "az"] if accel == 1 else []) + (["dx", "dy", "dz"] if with_dr else [])
    full = {}
    for nm in names:
        t = torch.full((n + 2 * pad,), 12345.0, dtype=torch.float32)
        if pinned:
            t = t.pin_memory()
        scale = 1e3 if nm in "xyz" else (10.0 if nm.startswith("v") else (3.0 if nm.startswith("a") else 0.0))
        t[pad:pad + n] = torch.from_numpy((rng.normal(0, 1, n) * scale).astype(np.float32))
        full[nm] = t
    host = {nm: t[pad:pad + n].numpy().copy() for nm, t in full.items()}
    from physicl_b200 import _capi

    soa = _capi.Soa()
    soa.n = n
    for nm, t in full.items():
        setattr(soa, nm, t.data_ptr() + 4 * pad)
    au = np.array([0.5, 0.0, -9.81], np.float32)
    pau = au.ctypes.data_as(C.POINTER(C.c_float)) if accel == 2 else None
    done = 0
    for k in (3, 1, 8):
        ctx.call("pcl_kinematics_steps_host", C.byref(soa), C.c_float(1e-3), int(accel != 0), pau, C.c_uint32(k), C.c_uint64(chunk))
        for _ in range(k):
            oracle.kinematics_f32(host, 1e-3, accel, au if accel == 2 else None)
        done += k
    for nm, t in full.items():
        got = t.numpy()
        assert (got[:pad] == 12345.0).all() and (got[pad + n:] == 12345.0).all(), nm
        assert _u().same_bits(got[pad:pad + n], host[nm]), nm


@pytest.mark.parametrize("accel", [0, 1, 2])
def test_kinematics_timesteps_in_registers_equal_single_steps(ctx, accel):
    """pcl_kinematics_steps (k timesteps per HBM round trip) == k launches of pcl_kinematics == the twin."""
    u = _u()
    n = 200_003
    rng = np.random.default_rng(5)
    r, v = rng.uniform(-1e3, 1e3, (3, n)), rng.normal(0, 10, (3, n))
    a = rng.normal(0, 3, (3, n)) if accel == 1 else None
    st1, g1 = u.make_store(ctx, r, v, a=a, kind="object")
    st2, g2 = u.make_store(ctx, r, v, a=a, kind="object")
    for g in (g1, g2):
        g.ensure("dx", "dy", "dz")
    host = u.host_state(g2)
    au = np.array([0.0, 0.0, -9.81], np.float32)
    pau = au.ctypes.data_as(C.POINTER(C.c_float)) if accel == 2 else None
    for k in (25, 1, 24):
        s1 = g1.soa()
        ctx.call("pcl_kinematics_steps", st1.stream(), C.byref(s1), C.c_float(1e-3), accel, pau, C.c_uint32(k))
    for _ in range(50):
        s2 = g2.soa()
        ctx.call("pcl_kinematics", st2.stream(), C.byref(s2), C.c_float(1e-3), accel, pau)
        oracle.kinematics_f32(host, 1e-3, accel, au if accel == 2 else None)
    for nm in ("x", "y", "z", "vx", "vy", "vz", "dx", "dy", "dz"):
        assert u.same_bits(g1.download(nm), g2.download(nm)), nm
        assert u.same_bits(g1.download(nm), host[nm]), nm


def test_kinematics_matches_reference_golden(ctx, golden):
    u = _u()
    gd = golden("kin")
    st, g = u.make_store(ctx, gd["r0"], gd["v0"], kind="object")
    g.ensure("dx", "dy", "dz")
    for s in range(int(gd["nsteps"])):
        soa = g.soa()
        ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(float(gd["dts"][s])), 0, None)
        r = np.stack([g.download(nm) for nm in ("x", "y", "z")]).astype(np.float64)
        dr = np.stack([g.download(nm) for nm in ("dx", "dy", "dz")]).astype(np.float64)
        # tolerance: 1e-5 relative per step (north star); float32 gives ~1e-7
        np.testing.assert_allclose(r, gd["s%d_r" % s], rtol=1e-5, atol=1e-5 * np.abs(gd["r0"]).max())
        np.testing.assert_allclose(dr, gd["s%d_dr" % s], rtol=1e-5, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# fused photon step against the binary32 twin (Philox in-kernel)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, oracle.WAVELENGTH, oracle.DELETE, oracle.WAVELENGTH | oracle.DELETE])
@pytest.mark.parametrize("n", [1, 5, 4096, 250_007])
def test_fused_step_bit_exact_vs_twin(ctx, mode, n):
    u = _u()
    r, v = u.random_photons(n, seed=n + mode)
    rng = np.random.default_rng(99)
    E = rng.uniform(0.2, 1.0, n) * 4e-19 if mode & oracle.WAVELENGTH else None
    st, g = u.make_store(ctx, r, v, E=E, id_base=10_000_000_000, nscat=True)
    host = u.host_state(g)
    dt, k, c = 1e-3, 1.2e-6, u.C_LIGHT
    planes = [(0, 1.0e5), (1, -2.0e5), (2, 0.0)]
    r2 = 5.0e5 ** 2
    for step in range(6):
        got = u.photon_step(ctx, st, g, dt, k, c, mode, seed=2024, step=step, r2_escape=r2, planes=planes)
        want = oracle.photon_step_f32(host, dt, k, c, mode, seed=2024, step=step, r2_escape=r2, planes=planes,
                                      id_base=10_000_000_000)
        assert np.array_equal(got, want), (step, got, want)
    for nm in host:
        assert u.same_bits(g.download(nm), host[nm]), nm
    if n >= 4096:
        assert want[oracle.T_ALIVE] < n  # something retired, something scattered
        assert host["nscat"].sum() > 0 or mode & oracle.DELETE


def test_fused_equals_unfused_sequence(ctx):
    """kinematics -> scatter -> escape -> tally as four launches gives the fused launch's bits."""
    from physicl_b200 import _capi

    u = _u()
    n = 100_003
    r, v = u.random_photons(n, seed=3)
    stA, gA = u.make_store(ctx, r, v)
    stB, gB = u.make_store(ctx, r, v)
    gB.ensure("dx", "dy", "dz")
    dt, k, c, r2 = 1e-3, 1.5e-6, u.C_LIGHT, 4.5e5 ** 2
    planes = [(0, 5e4)]
    for step in range(5):
        rowA = u.photon_step(ctx, stA, gA, dt, k, c, 0, seed=7, step=step, r2_escape=r2, planes=planes)
        sp = _capi.ScatterParams(k=k, c=c, mode=0)
        rg = _capi.Rng(seed=7, step=step)
        soa = gB.soa()
        ctx.call("pcl_kinematics", stB.stream(), C.byref(soa), C.c_float(dt), 0, None)
        r1 = stB.new_row()
        ctx.call("pcl_scatter", stB.stream(), C.byref(soa), C.byref(sp), C.byref(rg), None, stB.row_ptr())
        r2row = stB.new_row()
        ctx.call("pcl_escape", stB.stream(), C.byref(soa), C.c_float(r2), stB.row_ptr())
        r3 = stB.new_row()
        pl = _capi.make_planes(planes)
        ctx.call("pcl_tally", stB.stream(), C.byref(soa), C.byref(pl), stB.row_ptr())
        sc, es, ta = stB.read_row(r1), stB.read_row(r2row), stB.read_row(r3)
        assert sc[_capi.T_SCATTERED] == rowA[_capi.T_SCATTERED]
        assert es[_capi.T_ESCAPED] == rowA[_capi.T_ESCAPED]
        for col in (_capi.T_ALIVE, _capi.T_XP, _capi.T_YP, _capi.T_ZP, _capi.T_PLANE0):
            assert ta[col] == rowA[col], col
    live = ~np.isnan(gA.download("x"))
    assert np.array_equal(live, ~np.isnan(gB.download("x"))) and 0 < live.sum() < n
    for nm in u.PLANE_NAMES:  # retired slots hold unspecified values; live ones must agree bit for bit
        assert u.same_bits(gA.download(nm)[live], gB.download(nm)[live]), nm


def test_scatter_flags_match_twin(ctx):
    from physicl_b200 import _capi

    u = _u()
    n = 50_001
    r, v = u.random_photons(n, seed=8)
    st, g = u.make_store(ctx, r, v)
    g.ensure("dx", "dy", "dz")
    soa = g.soa()
    ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(1e-3), 0, None)
    host = u.host_state(g)
    flags_d = torch.empty(n, dtype=torch.int32, device=st.device)
    sp = _capi.ScatterParams(k=1e-6, c=u.C_LIGHT, mode=_capi.SCATTER_DELETE)
    rg = _capi.Rng(seed=5, step=3)
    row = st.new_row()
    ctx.call("pcl_scatter", st.stream(), C.byref(soa), C.byref(sp), C.byref(rg), C.c_void_p(flags_d.data_ptr()), st.row_ptr())
    flags, want_row = oracle.scatter_f32(host, 1e-6, u.C_LIGHT, oracle.DELETE, seed=5, step=3)
    assert np.array_equal(flags_d.cpu().numpy(), flags)
    assert 0 < flags.sum() < n
    got = st.read_row(row)
    for col in (_capi.T_ALIVE, _capi.T_SCATTERED, _capi.T_ABSORBED, _capi.T_LIVE_IN):
        assert got[col] == want_row[col]
    assert u.same_bits(g.download("x"), host["x"])


@pytest.mark.parametrize("k,dt", [(0.0, 1e-3), (1e30, 1e-3), (1.2e-6, 0.0), (3e-39, 1e-3), (float("inf"), 1e-3)])
def test_collision_test_at_the_ends_of_the_range(ctx, k, dt):
    """The squared form of pcoll >= rand (|dr|^2 >= (rand/k)^2 with 1/k from the host) at k = 0, k huge, k denormal, k = inf
    and dt = 0: same rows and bits as the twin, and the law of the reference (light.py:305-307): k = 0 or dt = 0 scatter only
    on rand = 0 exactly, k huge always."""
    u = _u()
    n = 200_003
    r, v = u.random_photons(n, seed=3)
    st, g = u.make_store(ctx, r, v, nscat=True)
    host = u.host_state(g)
    for step in range(3):
        got = u.photon_step(ctx, st, g, dt, k, u.C_LIGHT, 0, seed=8, step=step)
        want = oracle.photon_step_f32(host, dt, k, u.C_LIGHT, 0, seed=8, step=step)
        assert np.array_equal(got, want), (step, got, want)
        scat = int(got[oracle.T_SCATTERED])
        if k >= 1e30:
            assert scat == n
        else:
            ut, up, ur = oracle.philox_uniforms(n, 0, 8, step)
            assert scat == int((ur == 0).sum())  # a handful at most
    for nm in host:
        assert u.same_bits(g.download(nm), host[nm]), nm


# ---------------------------------------------------------------------------------------------
# against the reference's own outputs (golden vectors, injected uniforms)
# ---------------------------------------------------------------------------------------------
def _split_u(uu):
    uu = uu.reshape(-1, 3)
    return [np.ascontiguousarray(uu[:, i], np.float32) for i in range(3)]


@pytest.mark.parametrize("name", ["iso", "wave"])
def test_fused_step_tracks_reference_golden(ctx, golden, name):
    from physicl_b200 import _capi

    u = _u()
    gd = golden(name)
    N, c, dt = int(gd["N"]), float(gd["c"]), float(gd["dt"])
    r, v = u.beam_photons(N, c)
    E = gd["E"] if name == "wave" else None
    st, g = u.make_store(ctx, r, v, E=E)
    k, mode = float(gd["A"]) * float(gd["n"]), 0
    if name == "wave":
        k *= (g.e0 / (float(gd["h"]) * c)) ** 4
        mode = _capi.SCATTER_WAVELENGTH
    planes = [(0, 4.0e5), (1, 0.0), (2, -1.0e5)] if name == "iso" else None
    for s in range(int(gd["nsteps"])):
        row = u.photon_step(ctx, st, g, dt, k, c, mode, uniforms=_split_u(gd["s%d_u" % s]), planes=planes)
        ref_hit = ~np.isnan(gd["s%d_res0" % s])
        assert int(row[_capi.T_SCATTERED]) == int(ref_hit.sum())  # decisions identical
        vv = np.stack([g.download(nm) for nm in ("vx", "vy", "vz")]).astype(np.float64)
        rr = np.stack([g.download(nm) for nm in ("x", "y", "z")]).astype(np.float64)
        assert np.abs(vv - gd["s%d_v" % s]).max() <= 1e-5 * c
        assert np.abs(rr - gd["s%d_r" % s]).max() <= 1e-5 * c * dt * (s + 1)
        srow = gd["sign_rows"][s]
        assert [int(row[q]) for q in (_capi.T_ALIVE, _capi.T_XP, _capi.T_YP, _capi.T_ZP)] == [int(q) for q in srow[1:5]]
        if planes:
            assert [int(row[_capi.T_PLANE0 + q]) for q in range(3)] == [int(q) for q in gd["plane_rows"][s][2:5]]


@pytest.mark.parametrize("name", ["delete", "delete_ref"])
def test_delete_matches_reference_golden(ctx, golden, name):
    from physicl_b200 import _capi

    u = _u()
    gd = golden(name)
    N, c, dt = int(gd["N"]), float(gd["c"]), float(gd["dt"])
    r, v = u.beam_photons(N, c)
    st, g = u.make_store(ctx, r, v)
    k = float(gd["A"]) * float(gd["n"])
    alive = np.arange(N)
    for s in range(int(gd["nsteps"])):
        ur = np.zeros(N, np.float32)
        ur[alive] = gd["s%d_u" % s].astype(np.float32)
        row = u.photon_step(ctx, st, g, dt, k, c, _capi.SCATTER_DELETE, uniforms=(None, None, ur),
                            planes=[(0, float(gd["planes"][0][0]))])
        alive = np.nonzero(~np.isnan(g.download("x")))[0]
        assert np.array_equal(alive, gd["s%d_gid" % s])
        prow = gd["plane_rows"][s]
        assert int(row[_capi.T_ALIVE]) == int(prow[1]) and int(row[_capi.T_PLANE0]) == int(prow[2])
        assert int(row[_capi.T_ABSORBED]) == int((gd["s%d_flags" % s] == 1).sum())


# ---------------------------------------------------------------------------------------------
# compaction (Simulation.remove_obj, physicl/__init__.py:455-459)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 1024, 1025, 300_001])
def test_compaction_is_stable_and_exact(ctx, n):
    u = _u()
    r, v = u.random_photons(n, seed=n)
    E = np.linspace(1.0, 2.0, n)
    st, g = u.make_store(ctx, r, v, E=E, nscat=True)
    rng = np.random.default_rng(n)
    dead = rng.random(n) < 0.4
    x = g.download("x").copy()
    x[dead] = np.nan
    g.upload("x", x)
    before = u.host_state(g)
    n_live = st.compact("photon")
    keep = ~dead
    assert n_live == int(keep.sum()) == g.n
    assert np.array_equal(g.download("id"), np.nonzero(keep)[0].astype(np.uint32))
    for nm in before:
        assert u.same_bits(g.download(nm), before[nm][keep]), nm
    # a second compaction with nothing retired is the identity
    again = u.host_state(g)
    assert st.compact("photon") == n_live
    for nm in again:
        assert u.same_bits(g.download(nm), again[nm]), nm


@pytest.mark.parametrize("n", [5, 1024, 4097, 300_001])
def test_compaction_carries_acceleration_and_displacement_planes(ctx, n):
    """Groups with a and dr planes (kinematics with per-particle acceleration, unfused pipelines) take the kernel form with
    all fifteen planes: every plane of every survivor arrives, in order, bit for bit."""
    u = _u()
    rng = np.random.default_rng(n + 1)
    r, v = u.random_photons(n, seed=n + 1)
    a = rng.normal(0, 3, (3, n))
    st, g = u.make_store(ctx, r, v, E=np.linspace(0.5, 1.5, n), a=a, nscat=True)
    g.ensure("dx", "dy", "dz")
    for q, nm in enumerate(("dx", "dy", "dz")):
        g.upload(nm, rng.normal(0, 1, n).astype(np.float32) + q)
    dead = rng.random(n) < 0.6
    x = g.download("x").copy()
    x[dead] = np.nan
    g.upload("x", x)
    before = u.host_state(g)
    assert {"ax", "ay", "az", "dx", "dy", "dz", "e", "nscat"} <= set(before)
    n_live = st.compact("photon")
    keep = ~dead
    assert n_live == int(keep.sum()) == g.n
    assert np.array_equal(g.download("id"), np.nonzero(keep)[0].astype(np.uint32))
    for nm in before:
        assert u.same_bits(g.download(nm), before[nm][keep]), nm


def test_tallies_do_not_depend_on_compaction(ctx):
    u = _u()
    n = 120_000
    r, v = u.beam_photons(n)
    stA, gA = u.make_store(ctx, r, v)
    stB, gB = u.make_store(ctx, r, v)
    dt, k, c, r2 = 1e-3, 1e-6, u.C_LIGHT, 1.0e6 ** 2
    rowsA, rowsB = [], []
    for step in range(28):
        rowsA.append(u.photon_step(ctx, stA, gA, dt, k, c, 0, seed=11, step=step, r2_escape=r2).copy())
        rowsB.append(u.photon_step(ctx, stB, gB, dt, k, c, 0, seed=11, step=step, r2_escape=r2).copy())
        if step % 5 == 4:
            stB.compact("photon")
    a, b = np.array(rowsA), np.array(rowsB)
    assert np.array_equal(a, b)
    assert b[-1, 0] < n and stB.compactions >= 4 and gB.n < gA.n
    # survivors carry identical state (matched by id)
    sa, sb = stA.snapshot("photon"), stB.snapshot("photon")
    assert np.array_equal(sa["id"], sb["id"])
    for nm in u.PLANE_NAMES:
        assert u.same_bits(sa[nm], sb[nm]), nm


# ---------------------------------------------------------------------------------------------
# emission (light.py:73-104)
# ---------------------------------------------------------------------------------------------
def test_planck_bins_bit_exact_and_chi_square(ctx, golden):
    import physicl_b200.light as light
    from scipy import stats

    gd = golden("planck")
    n = 400_000
    e, E0, b = light.planck_sample_device(ctx, n, float(gd["E_min"]), float(gd["E_max"]), float(gd["T"]),
                                          bins=int(gd["bins"]), seed=2025, id_base=123, want_bins=True)
    E, norm, cdf = light.planck_table(float(gd["E_min"]), float(gd["E_max"]), float(gd["T"]), int(gd["bins"]))
    np.testing.assert_allclose(cdf, gd["cdf"], rtol=1e-11)  # same table as the reference builds with quad
    step = (E[-1] - E[0]) / (len(E) - 1)
    e_or, b_or = oracle.planck_sample(n, 123, 2025, cdf, np.float32(E[0] / E0), np.float32(step / E0))
    b = b.cpu().numpy()
    e = e.cpu().numpy()
    assert np.array_equal(b, b_or)  # integer result: bit-exact
    assert np.array_equal(np.isnan(e), b < 0) and _u().same_bits(e[b >= 0], e_or[b >= 0])
    ok = b >= 0
    np.testing.assert_allclose(e[ok].astype(np.float64) * E0, E[b[ok]], rtol=1e-6)  # grid energies
    # chi-square of the bin histogram against the reference's bin masses (bins coarsened to >= 5 expected)
    ncdf = cdf.size
    counts = np.bincount(b[ok], minlength=ncdf)  # counts[x]: photons that got grid energy E[x], x = 1..ncdf-1
    expected = norm * n  # mass of interval x; interval 0 is the reference's "None" branch
    groups_c, groups_e, cc, ee = [], [], 0.0, 0.0
    for ci, ei in zip(counts[1:ncdf], expected[1:ncdf]):
        cc, ee = cc + ci, ee + ei
        if ee >= 5:
            groups_c.append(cc), groups_e.append(ee)
            cc, ee = 0.0, 0.0
    chi2 = float(((np.array(groups_c) - np.array(groups_e)) ** 2 / np.array(groups_e)).sum())
    p = stats.chi2.sf(chi2, len(groups_c) - 1)
    assert p > 1e-4, (chi2, len(groups_c), p)
    assert abs((b < 0).mean() - norm[0]) < 5 * np.sqrt(norm[0] / n) + 1e-6  # None rate = mass of interval 0


def _planck_case_cdf(kind, ncdf):
    rng = np.random.default_rng(ncdf)
    if kind == "uniform":  # every bucket of the guide table holds about ncdf / 65536 thresholds
        w = rng.uniform(0.5, 1.5, ncdf)
    elif kind == "tails":  # almost all mass in a few entries: thousands of thresholds share one bucket
        w = np.full(ncdf, 1e-9)
        w[rng.integers(0, ncdf, 7)] = 1.0
    elif kind == "over_one":  # cumulative sums that reach and pass 1 before the table ends
        w = rng.uniform(0.5, 1.5, ncdf)
        w[-ncdf // 50:] = 0.0
    else:  # entries that are exact multiples of 2^-24 (u == cdf[x] must be accepted: closed interval)
        w = rng.integers(1, 64, ncdf).astype(np.float64)
        return np.cumsum(w / 2.0 ** 24 * np.floor(2.0 ** 24 / w.sum()))
    cdf = np.cumsum(w / w.sum())
    if kind == "over_one":
        cdf = cdf * (1.0 + 3e-16) + 1e-16
        assert (cdf >= 1.0).sum() > 10
    return cdf


@pytest.mark.parametrize("kind,ncdf,n,id_base", [
    ("uniform", 49_999, 300_000, 0),        # shared-memory tables, 16-byte stores
    ("uniform", 49_999, 300_001, 2 ** 33 + 5),  # ... blocks cut by both ends of the shard, scalar stores, high id word
    ("uniform", 65_535, 270_000, 4),        # largest table that fits the shared-memory form
    ("uniform", 65_536, 270_000, 4),        # one more: guide in global memory, float64 probes
    ("uniform", 199, 1_000_003, 0),         # bins=200 of the reference's examples: most buckets empty
    ("tails", 30_000, 400_000, 8),          # crowded buckets: the byte search
    ("over_one", 20_000, 400_000, 0),
    ("grid", 5_000, 2_000_000, 0),          # equality with table entries
    ("uniform", 1, 300_000, 0),             # a one-entry table never yields a bin (the reference needs x >= 1)
    ("uniform", 49_999, 1, 3), ("uniform", 49_999, 5, 3), ("uniform", 49_999, 4097, 1),  # small draws
])
def test_planck_sampler_forms_bit_exact_vs_linear_scan(ctx, kind, ncdf, n, id_base):
    """Both kernel forms of pcl_planck_sample (tables in shared memory / guide in global memory) against the oracle's
    linear scan of the reference's rule (light.py:101-104) on synthetic tables that stress each branch."""
    cdf = _planck_case_cdf(kind, ncdf)
    dev = torch.device("cuda", ctx.device)
    cdf_d = torch.from_numpy(cdf).to(dev)
    pad = 64
    e = torch.full((n + 2 * pad,), -7.0, dtype=torch.float32, device=dev)
    b = torch.full((n + 2 * pad,), -7, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ctx.call("pcl_planck_sample", stream, C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(99), C.c_void_p(cdf_d.data_ptr()),
             C.c_uint32(ncdf), C.c_float(0.25), C.c_float(1e-5), C.c_void_p(e[pad:].data_ptr()), C.c_void_p(b[pad:].data_ptr()))
    torch.cuda.synchronize()
    e, b = e.cpu().numpy(), b.cpu().numpy()
    assert (b[:pad] == -7).all() and (b[pad + n:] == -7).all() and (e[:pad] == -7).all() and (e[pad + n:] == -7).all()
    e, b = e[pad:pad + n], b[pad:pad + n]
    e_or, b_or = oracle.planck_sample(n, id_base, 99, cdf, np.float32(0.25), np.float32(1e-5))
    assert np.array_equal(b, b_or)
    assert np.array_equal(np.isnan(e), b < 0) and _u().same_bits(e[b >= 0], e_or[b >= 0])
    # the same draw without the bin output (the form bulk emission uses): identical energies
    e2 = torch.full((n + 2 * pad,), -7.0, dtype=torch.float32, device=dev)
    ctx.call("pcl_planck_sample", stream, C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(99), C.c_void_p(cdf_d.data_ptr()),
             C.c_uint32(ncdf), C.c_float(0.25), C.c_float(1e-5), C.c_void_p(e2[pad:].data_ptr()), None)
    torch.cuda.synchronize()
    e2 = e2.cpu().numpy()
    assert (e2[:pad] == -7).all() and (e2[pad + n:] == -7).all() and _u().same_bits(e2[pad:pad + n], e)
    if kind == "grid" :
        # the closed interval of the reference: some uniforms equal a table entry and must take the LOWER bin
        m = np.round(cdf * 2.0 ** 24).astype(np.int64)
        assert np.array_equal(m / 2.0 ** 24, cdf)


# ---------------------------------------------------------------------------------------------
# stochastic parity: the reference's distributions (KS / chi-square at fixed sample size)
# ---------------------------------------------------------------------------------------------
def test_scatter_direction_distribution_matches_reference_law(ctx):
    """The reference draws theta ~ U[0,2pi) as polar and phi ~ U[0,pi) as azimuth (light.py:285,
    :309-311): v_z/c = cos(theta) is arcsine distributed, v_y >= 0 always scatters to sin(theta)
    sign.  KS against those laws at n = 10^6; scattered fraction against pcoll (binomial)."""
    from scipy import stats

    u = _u()
    n = 1_000_000
    r, v = u.beam_photons(n)
    st, g = u.make_store(ctx, r, v)
    dt, c = 1e-3, u.C_LIGHT
    k = 0.5 / (c * dt)  # pcoll = 0.5
    row = u.photon_step(ctx, st, g, dt, k, c, 0, seed=77, step=0)
    raw = [g.download(nm) for nm in ("vx", "vy", "vz")]
    hit = ~((raw[0] == np.float32(c)) & (raw[1] == 0.0) & (raw[2] == 0.0))
    vx, vy, vz = (a.astype(np.float64) / c for a in raw)
    nh = int(hit.sum())
    assert nh == int(row[4])
    assert abs(nh - 0.5 * n) < 5 * np.sqrt(n * 0.25)
    ks_z = stats.kstest(vz[hit], lambda t: 0.5 + np.arcsin(np.clip(t, -1, 1)) / np.pi)
    assert ks_z.pvalue > 1e-3, ks_z
    # phi = atan2(vy, vx) folded: uniform on [0, pi) after undoing the sign of sin(theta)
    phi = np.arctan2(np.abs(vy[hit]), vx[hit] * np.sign(vy[hit] + (vy[hit] == 0)))
    ks_p = stats.kstest(phi / np.pi, "uniform")
    assert ks_p.pvalue > 1e-3, ks_p
    assert np.abs(np.sqrt(vx[hit] ** 2 + vy[hit] ** 2 + vz[hit] ** 2) - 1).max() < 1e-6
    # same statistics from the float64 reference law driven by the same Philox uniforms
    ut, up, ur = oracle.philox_uniforms(n, 0, 77, 0)
    rt, rp = law.scale_uniforms(ut.astype(np.float64), up.astype(np.float64))
    ref_vz = np.cos(rt)[ur.astype(np.float64) <= 0.5 * (1 + 1e-7)]
    assert stats.ks_2samp(vz[hit], ref_vz).pvalue > 1e-3


def test_beer_lambert_survival(ctx):
    """Delete scattering: survivors after s steps follow (1 - pcoll)^s (the physics behind the
    reference's test_scatter_delete, test/test_light.py:45-66), checked per step at 5 sigma."""
    from physicl_b200 import _capi

    u = _u()
    n = 1_000_000
    r, v = u.beam_photons(n)
    st, g = u.make_store(ctx, r, v)
    dt, c = 1e-3, u.C_LIGHT
    p = 1e-3 * 1e-3 * c * dt  # A n c dt of the reference test
    alive = n
    for step in range(12):
        row = u.photon_step(ctx, st, g, dt, np.float32(1e-6), c, _capi.SCATTER_DELETE, seed=5, step=step)
        assert row[_capi.T_LIVE_IN] == alive
        expect = alive * (1 - p)
        assert abs(row[_capi.T_ALIVE] - expect) < 5 * np.sqrt(alive * p * (1 - p)) + 1
        assert row[_capi.T_ALIVE] + row[_capi.T_ABSORBED] == alive
        alive = int(row[_capi.T_ALIVE])


# ---------------------------------------------------------------------------------------------
# gravity (new step; oracle from the definition, float64)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 513, 3000])
def test_gravity_matches_float64_definition(ctx, n):
    rng = np.random.default_rng(n)
    pos = rng.normal(size=(3, n))
    m = rng.uniform(0.5, 1.5, n)
    G, eps2 = 1.0, 1e-4
    posm = torch.from_numpy(np.ascontiguousarray(np.vstack([pos, m[None]]).T, np.float32)).cuda()
    acc = torch.zeros((3, n), dtype=torch.float32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), C.c_float(G), C.c_float(eps2),
             p(acc[0]), p(acc[1]), p(acc[2]), 0, C.c_uint64(0), C.c_uint64(0))
    torch.cuda.synchronize()
    pos32 = posm[:, :3].T.cpu().numpy().astype(np.float64)
    m32 = posm[:, 3].cpu().numpy().astype(np.float64)
    want = oracle.gravity_f64(np.ascontiguousarray(pos32), np.ascontiguousarray(m32), G, eps2)
    got = acc.cpu().numpy().astype(np.float64)
    scale = np.abs(want).max() if n > 1 else 1.0
    assert np.abs(got - want).max() <= 1e-5 * scale + 1e-12
    if n > 1:
        # Newton's third law: total force vanishes to rounding
        f = (got * m32[None]).sum(1)
        assert np.abs(f).max() <= 1e-4 * np.abs(got * m32[None]).sum(1).max()


@pytest.mark.parametrize("n", [1, 3, 513, 3001])
def test_gravity_uniform_mass_matches_float64_definition(ctx, n):
    """pcl_gravity_accel_uniform (equal masses, G*m passed, posm.w ignored), including a ragged skip range and ragged
    tile ends (padded j-slots must contribute exactly nothing)."""
    rng = np.random.default_rng(100 + n)
    pos = rng.normal(size=(3, n))
    m0, G, eps2 = 0.37, 2.0, 1e-4
    posm = torch.from_numpy(np.ascontiguousarray(np.vstack([pos, rng.uniform(5, 9, (1, n))]).T, np.float32)).cuda()  # w: junk on purpose
    acc = torch.zeros((3, n), dtype=torch.float32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    lo, hi = (n // 3, (2 * n) // 3)
    own = posm[lo:hi].contiguous()
    if hi > lo:  # own block first, then everything else with the block skipped (the sharded calling sequence)
        ctx.call("pcl_gravity_accel_uniform", None, p(posm), C.c_uint64(n), p(own), C.c_uint64(hi - lo), C.c_float(G * m0),
                 C.c_float(eps2), p(acc[0]), p(acc[1]), p(acc[2]), 0, C.c_uint64(0), C.c_uint64(0))
    ctx.call("pcl_gravity_accel_uniform", None, p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), C.c_float(G * m0), C.c_float(eps2),
             p(acc[0]), p(acc[1]), p(acc[2]), int(hi > lo), C.c_uint64(lo), C.c_uint64(hi))
    torch.cuda.synchronize()
    pos32 = np.ascontiguousarray(posm[:, :3].T.cpu().numpy().astype(np.float64))
    want = oracle.gravity_f64(pos32, np.full(n, m0), G, eps2)
    got = acc.cpu().numpy().astype(np.float64)
    scale = np.abs(want).max() if n > 1 else 1.0
    assert np.abs(got - want).max() <= 1e-5 * scale + 1e-12


def test_gravity_split_blocks_equal_whole(ctx):
    """Sharded accumulation (local block, then remote blocks with accumulate=1) equals one pass."""
    n = 2048
    rng = np.random.default_rng(1)
    posm = torch.from_numpy(np.ascontiguousarray(np.vstack([rng.normal(size=(3, n)), np.ones((1, n))]).T, np.float32)).cuda()
    p = lambda t: C.c_void_p(t.data_ptr())
    whole = torch.zeros((3, n), dtype=torch.float32, device="cuda")
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), C.c_float(1.0), C.c_float(1e-4),
             p(whole[0]), p(whole[1]), p(whole[2]), 0, C.c_uint64(0), C.c_uint64(0))
    parts = torch.zeros((3, n), dtype=torch.float32, device="cuda")
    for b in range(4):
        blk = posm[b * 512:(b + 1) * 512]
        ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(n), p(blk), C.c_uint64(512), C.c_float(1.0), C.c_float(1e-4),
                 p(parts[0]), p(parts[1]), p(parts[2]), int(b > 0), C.c_uint64(0), C.c_uint64(0))
    # the sharded form: own block first, then the whole array with that block skipped (ragged skip range)
    two = torch.zeros((3, n), dtype=torch.float32, device="cuda")
    lo, hi = 700, 1801
    own = posm[lo:hi].contiguous()
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(n), p(own), C.c_uint64(hi - lo), C.c_float(1.0), C.c_float(1e-4),
             p(two[0]), p(two[1]), p(two[2]), 0, C.c_uint64(0), C.c_uint64(0))
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), C.c_float(1.0), C.c_float(1e-4),
             p(two[0]), p(two[1]), p(two[2]), 1, C.c_uint64(lo), C.c_uint64(hi))
    torch.cuda.synchronize()
    w, q, t2 = whole.cpu().numpy(), parts.cpu().numpy(), two.cpu().numpy()
    assert np.abs(w - q).max() <= 2e-5 * np.abs(w).max()
    assert np.abs(w - t2).max() <= 2e-5 * np.abs(w).max()


# ---------------------------------------------------------------------------------------------
# host-buffer entry point
# ---------------------------------------------------------------------------------------------
def test_host_buffer_step_equals_device_step(ctx):
    from physicl_b200 import _capi

    u = _u()
    n = 300_007
    r, v = u.random_photons(n, seed=21)
    st, g = u.make_store(ctx, r, v)
    host = {nm: torch.from_numpy(g.download(nm).copy()).pin_memory() for nm in u.PLANE_NAMES}
    dt, k, c, r2 = 1e-3, 1.5e-6, u.C_LIGHT, 4.5e5 ** 2
    for step in range(3):
        want = u.photon_step(ctx, st, g, dt, k, c, 0, seed=9, step=step, r2_escape=r2, planes=[(1, 0.0)])
        soa = _capi.Soa()
        soa.n = n
        for nm in u.PLANE_NAMES:
            setattr(soa, nm, host[nm].data_ptr())
        sp = _capi.ScatterParams(k=k, c=c, mode=0)
        rg = _capi.Rng(seed=9, step=step)
        pl = _capi.make_planes([(1, 0.0)])
        row = np.zeros(_capi.TALLY_COLS, np.int64)
        ctx.call("pcl_photon_step_host", C.byref(soa), C.c_float(dt), C.byref(sp), C.byref(rg), C.c_float(r2), C.byref(pl),
                 row.ctypes.data_as(C.c_void_p), C.c_uint64(65_536))
        assert np.array_equal(row, want)
    for nm in u.PLANE_NAMES:
        assert u.same_bits(host[nm].numpy(), g.download(nm)), nm


def test_errors_are_reported_not_swallowed(ctx):
    from physicl_b200 import _capi

    soa = _capi.Soa()
    soa.n = 16
    with pytest.raises(_capi.PclError, match="r and v planes"):
        ctx.call("pcl_kinematics", None, C.byref(soa), C.c_float(1.0), 0, None)
    with pytest.raises(_capi.PclError):
        _capi.Context(4096)


def test_size_limits_are_checked_before_any_launch(ctx):
    """A view holds fewer than 2^32 slots and its global ids must not cross a multiple of 2^32 (the Philox counter carries
    the low id word, the key the high one); ids far above 2^32 are fine.  Violations are errors, not wrong numbers."""
    from physicl_b200 import _capi

    u = _u()
    r, v = u.beam_photons(64)
    st, g = u.make_store(ctx, r, v)
    sp = _capi.ScatterParams(k=1e-6, c=u.C_LIGHT, mode=0)
    rg = _capi.Rng(seed=1, step=0)
    pl = _capi.make_planes([])
    launches = ctx.launches

    def step(soa):
        soa.dx = soa.dy = soa.dz = None
        ctx.call("pcl_photon_step", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0), C.byref(pl),
                 st.row_ptr(st.new_row()))

    soa = g.soa()
    soa.n = 1 << 32
    with pytest.raises(_capi.PclError, match=r"2\^32 slots"):
        step(soa)
    soa = g.soa()
    soa.id_base = (1 << 32) - 10  # ids [2^32 - 10, 2^32 + 54): crosses
    with pytest.raises(_capi.PclError, match=r"multiple of 2\^32"):
        step(soa)
    assert ctx.launches == launches
    soa = g.soa()
    soa.id_base = (7 << 32) + 5  # high word 7 goes into the key
    step(soa)
    st.synchronize()
    assert ctx.launches == launches + 1
    host = {nm: np.zeros(64, np.float32) for nm in u.PLANE_NAMES}
    host["vx"][:] = np.float32(u.C_LIGHT)
    want = oracle.photon_step_f32(host, 1e-3, 1e-6, u.C_LIGHT, 0, seed=1, step=0, id_base=(7 << 32) + 5)
    assert np.array_equal(st.read_row(st.current_row), want)
    for nm in u.PLANE_NAMES:
        assert u.same_bits(g.download(nm), host[nm]), nm


# ---------------------------------------------------------------------------------------------
# retire-and-compact step and the ping-pong multi-step loop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, oracle.DELETE, oracle.WAVELENGTH])
@pytest.mark.parametrize("n", [1, 6, 1023, 200_001])
def test_compacting_step_matches_twin_by_id(ctx, mode, n):
    """pcl_photon_step_compact = the in-place step followed by dropping retired slots; survivors are
    identified by id (tiles land in arrival order)."""
    from physicl_b200 import _capi

    u = _u()
    r, v = u.random_photons(n, seed=n + 3 * mode)
    E = np.random.default_rng(4).uniform(0.3, 1.0, n) * 3e-19 if mode & oracle.WAVELENGTH else None
    st, g = u.make_store(ctx, r, v, E=E, id_base=5_000_000_000, nscat=True)
    host = u.host_state(g)
    host["id"] = np.arange(n, dtype=np.uint32)
    st.reserve_spare("photon")
    dt, k, c, r2 = 1e-3, 1.3e-6, u.C_LIGHT, 4.0e5 ** 2
    planes = [(0, 1.0e5), (2, -1.0e5)]
    n_dev = torch.zeros(2, dtype=torch.int64, device=st.device)
    for step in range(5):
        sp = _capi.ScatterParams(k=k, c=c, mode=mode)
        rg = _capi.Rng(seed=31, step=step)
        pl = _capi.make_planes(planes)
        row = st.new_row()
        src = g.soa()
        src.dx = src.dy = src.dz = None
        dst = g.soa(planes=g.spare)
        dst.dx = dst.dy = dst.dz = None
        ctx.call("pcl_photon_step_compact", st.stream(), C.byref(src), C.byref(dst), C.c_float(dt), C.byref(sp), C.byref(rg),
                 C.c_float(r2), C.byref(pl), st.row_ptr(), C.c_void_p(n_dev.data_ptr()))
        n_live = int(n_dev[0].item())
        g.cur ^= 1
        g.id_valid[g.cur] = True
        g.n = n_live
        want = oracle.photon_step_f32(host, dt, k, c, mode, seed=31, step=step, r2_escape=r2, planes=planes,
                                      id_base=5_000_000_000)
        got = st.read_row(row)
        assert np.array_equal(got, want), (step, got, want)
        keep = ~np.isnan(host["x"])
        host = {nm: a[keep] for nm, a in host.items()}
        assert n_live == int(keep.sum()) == int(want[oracle.T_ALIVE])
        snap = st.snapshot("photon", live_only=False)
        assert np.array_equal(snap["id"], host["id"])
        for nm in host:
            assert u.same_bits(snap[nm], host[nm]), (step, nm)


@pytest.mark.parametrize("wave", [0, 1])
def test_sfu_directions_same_first_step_decisions_and_same_law(ctx, wave):
    """PCL_SCATTER_SFU (opt-in): new directions from MUFU.SIN / MUFU.COS instead of the reproducible table.  The first
    timestep's decisions do not depend on the direction arithmetic, so tally rows, scattered sets and positions equal
    the default form's; the new velocities agree to 2e-6 c (both approximate the same angles) and |v| = c to 1e-5;
    over more timesteps the two forms agree in law: scattered counts within 5 sigma, v_z / c arcsine-distributed (KS)."""
    from scipy import stats

    from physicl_b200 import _capi

    u = _u()
    n = 1_000_003
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = u.C_LIGHT
    E = np.random.default_rng(5).uniform(0.5, 1.0, n) if wave else None
    k = 1.0e-6 * (4.0 if wave else 1.0)  # ~30 % scatter per step (wave: e^4 ~ 0.06 .. 1)
    runs = {}
    for name, mode in (("tab", wave), ("sfu", wave | _capi.SCATTER_SFU)):
        st, g = u.make_store(ctx, r, v, E=E, nscat=True)
        sp = _capi.ScatterParams(k=k, c=u.C_LIGHT, mode=mode)
        pl = _capi.make_planes([])
        steps = 9
        first = st.new_rows(steps)
        soa = g.soa()
        soa.dx = soa.dy = soa.dz = None
        rg = _capi.Rng(seed=77, step=0)
        ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                 C.byref(pl), st.row_ptr(first), C.c_uint32(1))
        one = {nm: g.download(nm).copy() for nm in u.PLANE_NAMES + ("nscat",)}
        rg = _capi.Rng(seed=77, step=1)
        ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                 C.byref(pl), st.row_ptr(first + 1), C.c_uint32(8))
        rows = np.array([st.read_row(first + i) for i in range(steps)])
        runs[name] = (one, rows, {nm: g.download(nm).copy() for nm in u.PLANE_NAMES + ("nscat",)})
    (one_t, rows_t, end_t), (one_s, rows_s, end_s) = runs["tab"], runs["sfu"]
    # first timestep: identical decisions, positions and scatter counts; velocities of scattered photons agree closely
    assert np.array_equal(rows_t[0][[_capi.T_ALIVE, _capi.T_SCATTERED, _capi.T_LIVE_IN]], rows_s[0][[_capi.T_ALIVE, _capi.T_SCATTERED, _capi.T_LIVE_IN]])
    assert np.array_equal(one_t["nscat"], one_s["nscat"]) and 0.15 * n < one_t["nscat"].sum() < 0.6 * n
    for nm in ("x", "y", "z"):
        assert u.same_bits(one_t[nm], one_s[nm]), nm
    hit = one_t["nscat"] == 1
    vt = np.stack([one_t[q][hit] for q in ("vx", "vy", "vz")]).astype(np.float64)
    vs = np.stack([one_s[q][hit] for q in ("vx", "vy", "vz")]).astype(np.float64)
    assert np.abs(vs - vt).max() <= 2e-6 * u.C_LIGHT
    assert np.abs(np.linalg.norm(vs, axis=0) / u.C_LIGHT - 1.0).max() <= 1e-5
    assert not np.array_equal(vs, vt)  # the SFU form really ran
    for q in ("vx", "vy", "vz"):  # untouched photons keep their bits
        assert u.same_bits(one_s[q][~hit], one_t[q][~hit])
    # nine timesteps: same law
    sc_t, sc_s = rows_t[:, _capi.T_SCATTERED].sum(), rows_s[:, _capi.T_SCATTERED].sum()
    assert abs(int(sc_t) - int(sc_s)) < 5 * np.sqrt(2.0 * sc_t)
    moved = end_s["nscat"] > 0
    vz = end_s["vz"][moved].astype(np.float64) / u.C_LIGHT
    ks = stats.kstest(vz[:200_000], lambda x: 0.5 + np.arcsin(np.clip(x, -1, 1)) / np.pi)  # cos(theta), theta ~ U[0, 2 pi)
    assert ks.pvalue > 1e-4, ks
    # sign tallies of the last row: v_x balanced, v_y > 0 for every scattered photon with sin(theta) > 0 ...: compare forms
    for col in (_capi.T_XP, _capi.T_YP, _capi.T_ZP):
        a, b = int(rows_t[-1][col]), int(rows_s[-1][col])
        assert abs(a - b) < 5 * np.sqrt(a + b + 1.0)


def test_sfu_flag_is_rejected_where_it_cannot_apply(ctx):
    """The stand-alone scatter kernel and injected uniforms keep the reproducible table: asking for the SFU there fails loudly."""
    from physicl_b200 import _capi

    u = _u()
    n = 4096
    r, v = u.random_photons(n, seed=3)
    st, g = u.make_store(ctx, r, v)
    g.ensure("dx", "dy", "dz")
    sp = _capi.ScatterParams(k=1e-6, c=u.C_LIGHT, mode=_capi.SCATTER_SFU)
    rg = _capi.Rng(seed=1, step=0)
    soa = g.soa()
    with pytest.raises(_capi.PclError, match="SFU"):
        ctx.call("pcl_scatter", st.stream(), C.byref(soa), C.byref(sp), C.byref(rg), None, st.row_ptr(st.new_row()))
    un = [torch.rand(n, device=st.device) for _ in range(3)]
    rg.u_theta, rg.u_phi, rg.u_rand = (t.data_ptr() for t in un)
    soa.dx = soa.dy = soa.dz = None
    with pytest.raises(_capi.PclError, match="SFU"):
        ctx.call("pcl_photon_step", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                 C.byref(_capi.make_planes([])), st.row_ptr(st.new_row()))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [3, 1024, 70_001])
def test_timesteps_fused_in_registers_equal_single_steps(ctx, mode, n):
    """pcl_photon_steps advances up to 8 timesteps per launch with the photons held in registers: same
    tally rows and the same bits as one launch per timestep, and as the binary32 twin stepped on the CPU."""
    from physicl_b200 import _capi

    u = _u()
    r, v = u.random_photons(n, seed=40 + n % 7, spread=2e5)
    E = np.random.default_rng(3).uniform(0.2, 1.0, n) if mode & 1 else None
    k = 2.0e-7 if mode & 2 else 2.5e-6  # delete mode: ~6 % absorbed per step; else ~75 % scatter
    planes = [(0, 1.0e5), (2, -5.0e4)]
    r2 = 1.5e6 ** 2
    steps = 13  # 8 + 5: a full and a partial group
    stA, gA = u.make_store(ctx, r, v, E=E, nscat=True)
    stB, gB = u.make_store(ctx, r, v, E=E, nscat=True)
    host = u.host_state(gB)
    sp = _capi.ScatterParams(k=k, c=u.C_LIGHT, mode=mode)
    pl = _capi.make_planes(planes)
    first = stA.new_rows(steps)
    soa = gA.soa()
    soa.dx = soa.dy = soa.dz = None
    rg = _capi.Rng(seed=23, step=4)
    ctx.call("pcl_photon_steps", stA.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(r2),
             C.byref(pl), stA.row_ptr(first), C.c_uint32(steps))
    rowsA = np.array([stA.read_row(first + i) for i in range(steps)])
    rowsB, rowsT = [], []
    for s in range(steps):
        rowsB.append(u.photon_step(ctx, stB, gB, 1e-3, k, u.C_LIGHT, mode, seed=23, step=4 + s, r2_escape=r2, planes=planes).copy())
        rowsT.append(oracle.photon_step_f32(host, 1e-3, k, u.C_LIGHT, mode, seed=23, step=4 + s, r2_escape=np.float32(r2),
                                            planes=planes))
    assert np.array_equal(rowsA, np.array(rowsB))
    assert np.array_equal(rowsA, np.array(rowsT))
    live = ~np.isnan(gA.download("x"))
    assert np.array_equal(live, ~np.isnan(host["x"]))
    if n > 1000:
        assert live.sum() < n and rowsA[:, _capi.T_ESCAPED].sum() > 0 and rowsA[:, _capi.T_SCATTERED].sum() > 0
        assert live.sum() > 0 or mode & 2  # delete mode: nobody lasts 13 steps inside this sphere
    for nm in u.PLANE_NAMES + ("nscat",):
        assert u.same_bits(gA.download(nm)[live], gB.download(nm)[live]), nm
        assert u.same_bits(gA.download(nm)[live], host[nm][live]), nm


def test_pingpong_loop_equals_in_place_loop(ctx):
    """pcl_photon_steps_pp (compaction every m steps, slot count kept on the device) gives the same
    tallies and the same surviving photons as the plain in-place loop."""
    from physicl_b200 import _capi

    u = _u()
    n = 150_003
    r, v = u.beam_photons(n)
    stA, gA = u.make_store(ctx, r, v)
    sp = _capi.ScatterParams(k=1e-6, c=u.C_LIGHT, mode=0)
    pl = _capi.make_planes([(0, 5.0e5)])
    r2 = 1.2e6 ** 2
    steps = 24
    firstA = stA.new_rows(steps)
    soa = gA.soa()
    rg = _capi.Rng(seed=17, step=0)
    ctx.call("pcl_photon_steps", stA.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(r2),
             C.byref(pl), stA.row_ptr(firstA), C.c_uint32(steps))
    rowsA = np.array([stA.read_row(firstA + i) for i in range(steps)])
    for m in (1, 3, 7):
        stB, gB = u.make_store(ctx, r, v)
        rowsB = []
        done = 0
        for chunk in (5, 11, 8):  # several calls: state must carry over (cur, id_valid, n_dev)
            pp = stB.pingpong("photon")
            first = stB.new_rows(chunk)
            rg = _capi.Rng(seed=17, step=done)
            ctx.call("pcl_photon_steps_pp", stB.stream(), C.byref(pp), C.c_float(1e-3), C.byref(sp), C.byref(rg),
                     C.c_float(r2), C.byref(pl), stB.row_ptr(first), C.c_uint32(chunk), C.c_uint32(m))
            stB.adopt_pingpong("photon", pp, (done + chunk) // m - done // m)
            rowsB += [stB.read_row(first + i).copy() for i in range(chunk)]
            done += chunk
        assert np.array_equal(np.array(rowsB), rowsA), m
        sa, sb = stA.snapshot("photon"), stB.snapshot("photon")
        assert gB.n < n and gB.n >= rowsA[-1, 0]
        assert np.array_equal(sa["id"], sb["id"])
        for nm in u.PLANE_NAMES:
            assert u.same_bits(sa[nm], sb[nm]), (m, nm)


@pytest.mark.parametrize("pinned", [True, False, "zero-copy"])
def test_compacting_host_step_equals_resident_steps(ctx, pinned, monkeypatch):
    """pcl_photon_step_host_compact: photons in host planes (pinned or pageable), survivors returned
    densely; same tallies and the same survivors (by id) as the device-resident path."""
    from physicl_b200 import _capi

    u = _u()
    n = 400_003
    r, v = u.beam_photons(n)
    st, g = u.make_store(ctx, r, v, id_base=7_000_000)
    # "zero-copy": kernels store the survivors straight into the pinned host planes (opt-in form)
    monkeypatch.setenv("PCL_HOST_ZEROCOPY", "1" if pinned == "zero-copy" else "0")
    pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    host = {nm: pin(torch.from_numpy(g.download(nm).copy())) for nm in u.PLANE_NAMES}
    host["id"] = pin(torch.arange(n, dtype=torch.int32))
    dt, k, c, r2 = 1e-3, 1e-6, u.C_LIGHT, 1.0e6 ** 2
    n_live = n
    for step in range(8):
        want = u.photon_step(ctx, st, g, dt, k, c, 0, seed=9, step=step, r2_escape=r2, planes=[(0, 7.0e5)])
        soa = _capi.Soa()
        soa.n = n_live
        soa.id_base = 7_000_000
        for nm, t in host.items():
            setattr(soa, nm, t.data_ptr())
        sp = _capi.ScatterParams(k=k, c=c, mode=0)
        rg = _capi.Rng(seed=9, step=step)
        pl = _capi.make_planes([(0, 7.0e5)])
        row = np.zeros(_capi.TALLY_COLS, np.int64)
        n_out = C.c_uint64(0)
        ctx.call("pcl_photon_step_host_compact", C.byref(soa), C.c_float(dt), C.byref(sp), C.byref(rg), C.c_float(r2),
                 C.byref(pl), row.ctypes.data_as(C.c_void_p), C.c_uint64(65_536), C.byref(n_out))
        assert np.array_equal(row, want), (step, row, want)
        n_live = int(n_out.value)
        assert n_live == int(want[_capi.T_ALIVE])
    assert 0 < n_live < n
    snap = st.snapshot("photon")
    ids = host["id"].numpy()[:n_live].view(np.uint32)
    order = np.argsort(ids)
    assert np.array_equal(ids[order], snap["id"])
    for nm in u.PLANE_NAMES:
        assert u.same_bits(host[nm].numpy()[:n_live][order], snap[nm]), nm


@pytest.mark.parametrize("m", [1, 3, 8])
def test_multi_timestep_host_round_trips_equal_resident_steps(ctx, m):
    """pcl_photon_steps_host_compact: m timesteps per host round trip (one fused launch per chunk).  Every tally row
    and the surviving photons (by id) equal the device-resident path advanced one timestep at a time; n = 0 is a no-op."""
    from physicl_b200 import _capi

    u = _u()
    n, steps = 300_011, 9
    r, v = u.beam_photons(n)
    st, g = u.make_store(ctx, r, v, id_base=5_000_000)
    host = {nm: torch.from_numpy(g.download(nm).copy()).pin_memory() for nm in u.PLANE_NAMES}
    host["id"] = torch.arange(n, dtype=torch.int32).pin_memory()
    dt, k, c, r2 = 1e-3, 1e-6, u.C_LIGHT, 1.0e6 ** 2
    want = [u.photon_step(ctx, st, g, dt, k, c, 0, seed=4, step=s, r2_escape=r2, planes=[(0, 7.0e5)]) for s in range(steps)]
    sp = _capi.ScatterParams(k=k, c=c, mode=0)
    pl = _capi.make_planes([(0, 7.0e5)])
    rows = np.zeros((8, _capi.TALLY_COLS), np.int64)
    n_out = C.c_uint64(0)
    n_live, s = n, 0
    while s < steps:
        run = min(m, steps - s)
        soa = _capi.Soa()
        soa.n, soa.id_base = n_live, 5_000_000
        for nm, t in host.items():
            setattr(soa, nm, t.data_ptr())
        rg = _capi.Rng(seed=4, step=s)
        ctx.call("pcl_photon_steps_host_compact", C.byref(soa), C.c_float(dt), C.byref(sp), C.byref(rg), C.c_float(r2),
                 C.byref(pl), rows.ctypes.data_as(C.c_void_p), C.c_uint64(32_768), C.c_uint32(run), C.byref(n_out))
        for q in range(run):
            assert np.array_equal(rows[q], want[s + q]), (s + q, rows[q], want[s + q])
        n_live = int(n_out.value)
        assert n_live == int(want[s + run - 1][_capi.T_ALIVE])
        s += run
    assert 0 < n_live < n
    snap = st.snapshot("photon")
    ids = host["id"].numpy()[:n_live].view(np.uint32)
    order = np.argsort(ids)
    assert np.array_equal(ids[order], snap["id"])
    for nm in u.PLANE_NAMES:
        assert u.same_bits(host[nm].numpy()[:n_live][order], snap[nm]), nm
    soa.n = 0
    ctx.call("pcl_photon_steps_host_compact", C.byref(soa), C.c_float(dt), C.byref(sp), C.byref(rg), C.c_float(r2),
             C.byref(pl), rows.ctypes.data_as(C.c_void_p), C.c_uint64(32_768), C.c_uint32(m), C.byref(n_out))
    assert n_out.value == 0 and not rows[:m].any()


def test_empty_and_tiny_inputs_through_every_bulk_entry_point(ctx):
    """n = 0 is a no-op everywhere (the reference's loops simply do not execute); rows stay zero."""
    from physicl_b200 import _capi, jit

    u = _u()
    for n in (0, 1, 2):
        r, v = u.beam_photons(n)
        st, g = u.make_store(ctx, r, v, E=np.ones(n))
        sp = _capi.ScatterParams(k=1e-6, c=u.C_LIGHT, mode=0)
        pl = _capi.make_planes([(0, 1.0e5)])
        rg = _capi.Rng(seed=1, step=0)
        first = st.new_rows(9)
        soa = g.soa()
        soa.dx = soa.dy = soa.dz = None
        ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                 C.byref(pl), st.row_ptr(first), C.c_uint32(9))
        rows = np.array([st.read_row(first + i) for i in range(9)])
        assert np.all(rows[:, _capi.T_LIVE_IN] == n) and np.all(rows[:, _capi.T_ALIVE] == n)
        ctx.call("pcl_kinematics_steps", st.stream(), C.byref(soa), C.c_float(1e-3), 0, None, C.c_uint32(17))
        if n:
            assert np.allclose(g.download("x"), (9 + 17) * np.float32(u.C_LIGHT) * np.float32(1e-3), rtol=1e-6) or rows[:, _capi.T_SCATTERED].sum() > 0
        mod = jit.Module(ctx, jit.photon_source("1e-6", False))
        vn = _capi.VarnParams(kd=1.0, e0=1.0, a_slot=1.0, n_slot=1.0)
        first = st.new_rows(3)
        ctx.call("pcl_photon_steps_jit", st.stream(), mod.kernel("pcl_jit_photon_step"), C.byref(soa), C.c_float(1e-3), C.byref(sp),
                 C.byref(vn), C.byref(rg), C.c_float(0.0), C.byref(pl), st.row_ptr(first), C.c_uint32(3))
        rows = np.array([st.read_row(first + i) for i in range(3)])
        assert np.all(rows[:, _capi.T_LIVE_IN] == n)
        mod.close()


def test_stream_gate_holds_the_stream_until_opened(ctx):
    """pcl_stream_gate: work queued behind the gate does not start before the host writes the flag (and does start after)."""
    import time

    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    x = torch.zeros(1 << 20, device="cuda")
    x.add_(1.0)  # first use loads the kernel's module, which waits for running kernels: not behind a closed gate
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ctx.call("pcl_stream_gate", stream, C.c_void_p(flag.data_ptr()), C.c_uint32(200))  # first use of the gate kernel itself
    torch.cuda.synchronize()
    t0 = time.time()
    ctx.call("pcl_stream_gate", stream, C.c_void_p(flag.data_ptr()), C.c_uint32(5000))
    x.add_(1.0)
    ev = torch.cuda.Event()
    ev.record()
    time.sleep(0.05)
    assert not ev.query()  # still behind the gate
    flag[0] = 1
    ev.synchronize()
    assert time.time() - t0 < 2.0  # opened by the flag, not by the 5 s timeout
    assert float(x[0].item()) == 2.0
    # a gate nobody opens gives up after its timeout instead of hanging the GPU
    flag[0] = 0
    t0 = time.time()
    ctx.call("pcl_stream_gate", stream, C.c_void_p(flag.data_ptr()), C.c_uint32(200))
    torch.cuda.synchronize()
    assert 0.15 < time.time() - t0 < 2.0
