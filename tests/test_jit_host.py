"""CPU tests of the run-time compiled kernels: the variable-density law
restated in the oracle is pinned to runs of the unmodified reference, and the generated CUDA
translation units compile for sm_100a (NVRTC needs no GPU)."""
import numpy as np
import pytest

from oracle import reference_law as law

VARN = ["varn", "varn_z"]


@pytest.mark.parametrize("name", VARN)
def test_variable_n_law_matches_reference_runs(golden, name):
    """light.py:295-315 with variable_n: decisions and new velocities of the reference's own kernel."""
    gd = golden(name)
    c, hc, wave = float(gd["c"]), float(gd["h"]) * float(gd["c"]), bool(gd["wave"])
    expr = str(gd["expr"])
    assert "(" + expr + ")" in str(gd["kernel_src"])  # light.py:299: spliced in parentheses
    assert float(gd["kernel_A"]) == float(gd["n"]) and float(gd["kernel_n"]) == float(gd["A"])  # light.py:287 swap
    for s in range(int(gd["nsteps"])):
        dr = np.stack([gd["s%d_d%d" % (s, q)] for q in range(3)])
        r = np.stack([gd["s%d_r%d" % (s, q)] for q in range(3)])
        E = gd["s%d_E" % s] if wave else None
        p = law.pcoll_variable_n(dr, r, expr, float(gd["kernel_A"]), float(gd["kernel_n"]), E, hc)
        hit = p >= gd["s%d_rand" % s]
        ref_hit = ~np.isnan(gd["s%d_res0" % s])
        assert np.array_equal(hit, ref_hit)
        assert 0 < hit.sum() < hit.size
        want = c * np.sin(gd["s%d_rtheta" % s]) * np.cos(gd["s%d_rphi" % s])
        assert np.allclose(gd["s%d_res0" % s][hit], want[hit], rtol=1e-14)
        # the positions the kernel saw are the ones after this timestep's kinematics
        assert np.array_equal(r, gd["s%d_r" % s])


@pytest.mark.parametrize("name", VARN)
def test_generated_kernels_compile_for_sm100a(golden, name):
    from physicl_b200 import jit

    gd = golden(name)
    src = jit.photon_source(str(gd["expr"]), bool(gd["wave"]))
    assert jit.check(src) > 10_000  # a real cubin came back
    assert "PCL_USER_N_EXPR" in src and "pcl_jit_photon.cuh" in src


def test_bad_expression_fails_when_the_step_is_created():
    import physicl_b200 as phys
    import physicl_b200.light
    from physicl_b200 import _capi

    with pytest.raises(_capi.PclError, match="undefined|error"):
        phys.light.ScatterIsotropicStep(n=1.0, variable_n=True, variable_n_fn="2.0 * nosuchfn(r0[gid])")
    with pytest.raises(ValueError):
        phys.light.ScatterIsotropicStep(n=1.0, variable_n=True, variable_n_fn=None)
    with pytest.raises(ValueError):
        phys.light.ScatterIsotropicStep(n=1.0, variable_n=True, variable_n_fn="1.0\n#include <x>")
    ok = phys.light.ScatterIsotropicStep(n=1.0, variable_n=True, variable_n_fn="1e-3 * exp(r2[gid] / 2.0e6) * pown(1.0, 2)")
    assert ok.variable_n and ok.mode == 0


def test_variable_n_constants_follow_the_reference_binding():
    """light.py:287: the kernel scalar `A` is the step's n; the step's A only enters on request."""
    import physicl_b200 as phys
    import physicl_b200.light

    class G:
        e0 = 9.9e-19

    st = phys.light.ScatterIsotropicStep(n=np.double(4e-56), A=np.double(123.0), variable_n=True, variable_n_fn="1.0",
                                         wavelength_dep_scattering=True, check_expression=False)
    vn = st.varn_params(G)
    hc = float(phys.light.h) * float(phys.light.c)
    assert vn.a_slot == 4e-56 and vn.n_slot == 123.0 and vn.e0 == G.e0
    assert np.isclose(vn.kd, 4e-56 * (G.e0 / hc) ** 4, rtol=1e-15)
    st2 = phys.light.ScatterIsotropicStep(n=np.double(2.0), A=np.double(3.0), variable_n=True, variable_n_fn="1.0",
                                          variable_n_apply_A=True, check_expression=False)
    assert st2.varn_params(G).kd == 6.0


# ---- user kernels: CLInput / CLOutput / CLProgram (physicl/__init__.py:543-664) ----
def build_user_program(sim, gd):
    """The program of tests/golden/make_golden.py:gen_clprogram, written exactly as against the reference."""
    import physicl_b200 as physicl

    prog = physicl.CLProgram(sim, "user_energy", str(gd["body"]))
    skip = physicl.CLInput(name="skip", type="obj_action", code="if type(obj) == physicl.light.PhotonObject:\n \t\t continue")
    v = [physicl.CLInput(name="v%d" % i, type="obj", obj_attr="v[%d]" % i) for i in range(3)]
    r2 = physicl.CLInput(name="r2", type="obj", obj_attr="r[2]")
    jit = physicl.CLInput(name="jitter", type="obj_def", obj_def="np.random.random()")
    who = physicl.CLInput(name="who", type="obj_track", obj_track="obj")
    consts = [physicl.CLInput(name="m", type="const", const_value="2.5"), physicl.CLInput(name="g", type="const", const_value=str(9.81)),
              physicl.CLInput(name="zcut", type="const", const_value="-250.0")]
    prog.prep_metadata = [skip] + v + [r2, jit, who] + consts
    prog.output_metadata = [physicl.CLOutput(name="ke"), physicl.CLOutput(name="flag", ctype="int")]
    return prog


def user_objects(gd):
    import physicl_b200 as physicl
    import physicl_b200.light

    objs = []
    for i in range(int(gd["N"])):
        if gd["is_photon"][i]:
            o = physicl.light.PhotonObject(s=np.zeros(3), v=np.array([physicl.light.c, 0, 0], dtype=np.double), E=np.double(1))
        else:
            o = physicl.Object()
            o.r = physicl.Measurement(list(gd["r"][:, i]), "m**1")
            o.v = physicl.Measurement(list(gd["v"][:, i]), "m**1 s**-1")
        o.gid = i
        objs.append(o)
    return objs


def test_user_program_text_and_gather_follow_the_reference(golden):
    import physicl_b200 as physicl

    gd = golden("clprogram")
    sim = physicl.Simulation(cl_on=False)
    sim.add_objs(user_objects(gd))
    prog = build_user_program(sim, gd)
    prog.build_kernel()  # NVRTC compile for sm_100a happens here and needs no GPU
    sig = prog.source[prog.source.index('extern "C"'):].split("{")[0]
    # argument order of physicl/__init__.py:586-592: inputs and constants in prep_metadata order, then outputs
    names = [a.split()[-1].lstrip("*") for a in sig[sig.index("(") + 1:sig.rindex(")")].split(",")]
    assert names == ["pcl_n", "v0", "v1", "v2", "r2", "jitter", "m", "g", "zcut", "ke", "flag"]
    assert "int *flag" in sig and "double *ke" in sig
    np.random.seed(int(gd["seed"]))
    prog._gather()
    assert np.array_equal(prog.v0_np, gd["in_v0"]) and np.array_equal(prog.r2_np, gd["in_r2"])
    assert np.array_equal(prog.jitter_np, gd["in_jitter"])  # same np.random stream, one draw per KEPT object
    assert [o.gid for o in prog.who] == list(gd["tracked_gid"])  # the obj_action filter skipped the photons
    # float64 restatement of the kernel body against the reference's outputs
    speed = np.sqrt(gd["in_v0"] ** 2 + gd["in_v1"] ** 2 + gd["in_v2"] ** 2)
    ke = 0.5 * float(gd["m"]) * speed * speed + float(gd["g"]) * gd["in_r2"] + gd["in_jitter"] * np.exp(-speed / 10.0)
    cut = gd["in_r2"] < float(gd["zcut"])
    assert np.array_equal(gd["flag"], (~cut).astype(np.int32))
    assert np.all(np.isnan(gd["ke"][cut])) and np.allclose(gd["ke"][~cut], ke[~cut], rtol=1e-15)


def test_user_program_errors():
    import physicl_b200 as physicl
    from physicl_b200 import _capi

    sim = physicl.Simulation(cl_on=False)
    bad = physicl.CLProgram(sim, "k", "int gid = get_global_id(0); out[gid] = nosuch(x[gid]);")
    bad.prep_metadata = [physicl.CLInput(name="x", type="obj", obj_attr="r[0]")]
    bad.output_metadata = [physicl.CLOutput(name="out")]
    with pytest.raises(_capi.PclError, match="nosuch"):
        bad.build_kernel()
    odd = physicl.CLProgram(sim, "k", "")
    odd.output_metadata = [physicl.CLOutput(name="out", ctype="quaternion")]
    with pytest.raises(ValueError, match="quaternion"):
        odd.build_kernel()
    ok = physicl.CLProgram(sim, "k", "int gid = get_global_id(0); out[gid] = x[gid];")
    ok.prep_metadata = [physicl.CLInput(name="x", type="obj", obj_attr="r[0]")]
    ok.output_metadata = [physicl.CLOutput(name="out")]
    ok.build_kernel()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ok.run()
