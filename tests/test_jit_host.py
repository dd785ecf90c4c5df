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
