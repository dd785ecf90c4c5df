"""CPU oracle for the physicl_b200 hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``physicl_b200`` never does.

Contents
--------
``oracle.c`` (``liboracle.so``)
    C restatement: ``orc_*_f64`` follows the reference in double precision
    (physicl/newton.py:14-16, physicl/light.py:146-158, :303-315, :325-331, :374-404, :414-431,
    :73-104); ``orc_*_f32`` is the binary32 twin with the CUDA kernels' exact operation order.
``reference_law``
    NumPy float64 restatement of the same functions, written against the reference's Python
    (readable spec; checked against the golden vectors and against ``oracle.c``).
``fake_pyopencl``
    A stand-in ``pyopencl`` that compiles the reference's OWN OpenCL-C kernel text with gcc, so the
    unmodified reference runs here (pyopencl/pocl are not installed and cannot be).  Used only by
    ``tests/golden/make_golden.py`` to produce the committed golden vectors.

Pinning: the golden vectors in ``tests/golden/*.npz`` are outputs of the reference itself (its
Python host code + its kernel strings) run in the build container; ``tests/test_oracle_golden.py``
checks both restatements against them.  The steps the reference does not have (constant
acceleration, escape sphere, gravity, Philox stream) are "parity unpinned" and are covered by
invariant tests instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

TALLY_COLS = 16
T_ALIVE, T_XP, T_YP, T_ZP, T_SCATTERED, T_ABSORBED, T_ESCAPED, T_LIVE_IN, T_PLANE0 = range(9)
WAVELENGTH, DELETE = 1, 2


def build(force: bool = False) -> str:
    """Compile oracle/c/oracle.c with the committed Makefile (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "c", "oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (type(a), getattr(a, "dtype", None))
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def philox2x32_10(ctr, key):
    c = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in ctr])
    o = (C.c_uint32 * 2)()
    lib().orc_philox2x32_10(c, C.c_uint32(int(key) & 0xFFFFFFFF), o)
    return [int(v) for v in o]


def fold_key(seed, id_hi=0, stream=0):
    f = lib().orc_fold_key
    f.restype = C.c_uint32
    return int(f(C.c_uint64(seed), C.c_uint64(id_hi), C.c_uint64(stream)))


def philox_uniforms(n, id_base, seed, step, stream=0):
    """The photon kernels' draws as plain uniforms (u_theta = theta / 2 pi, u_phi = phi / pi, u_rand); stream 1: the
    emission sampler's Philox4x32 stream."""
    ut = np.empty(n, np.float32)
    up = np.empty(n, np.float32)
    ur = np.empty(n, np.float32)
    lib().orc_philox_uniforms(C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(seed), C.c_uint32(step),
                              C.c_uint32(stream), _p(ut, np.float32), _p(up, np.float32), _p(ur, np.float32))
    return ut, up, ur


def sincos_tab(k, b):
    """(sin, cos) of table angle 2 pi k / 512 plus remainder b (radians), as the kernels evaluate it."""
    fs, fc = C.c_float(), C.c_float()
    lib().orc_sincos_tab(C.c_uint32(int(k)), C.c_float(float(b)), C.byref(fs), C.byref(fc))
    return fs.value, fc.value


def _planes(planes):
    """planes: list of (axis, loc) -> ctypes arrays"""
    planes = planes or []
    axis = np.array([a for a, _ in planes], np.uint32)
    loc = np.array([l for _, l in planes], np.float32)
    return len(planes), axis, loc


def kinematics_f32(st, dt, accel=0, a_uniform=None):
    """st: dict of float32 planes x,y,z,vx,vy,vz,[dx,dy,dz],[ax,ay,az]; updated in place."""
    au = np.ascontiguousarray(a_uniform if a_uniform is not None else [0, 0, 0], np.float32)
    lib().orc_kinematics_f32(C.c_uint64(st["x"].size), *[_p(st[k], np.float32) for k in ("x", "y", "z", "vx", "vy", "vz")],
                             *[_p(st.get(k), np.float32) for k in ("dx", "dy", "dz", "ax", "ay", "az")],
                             C.c_float(dt), C.c_int(accel), _p(au, np.float32))


def photon_step_f32(st, dt, k, c, mode=0, seed=0, step=0, uniforms=None, r2_escape=0.0, planes=None, id_base=0):
    """Fused-step twin.  st planes updated in place; returns the int64 tally row."""
    row = np.zeros(TALLY_COLS, np.int64)
    npl, axis, loc = _planes(planes)
    ut, up, ur = uniforms if uniforms is not None else (None, None, None)
    lib().orc_photon_step_f32(
        C.c_uint64(st["x"].size), *[_p(st[q], np.float32) for q in ("x", "y", "z", "vx", "vy", "vz")],
        _p(st.get("e"), np.float32), _p(st.get("id"), np.uint32), _p(st.get("nscat"), np.uint32), C.c_uint64(id_base),
        C.c_float(dt), C.c_float(k), C.c_float(c), C.c_uint32(mode), C.c_uint64(seed), C.c_uint32(step),
        _p(ut, np.float32), _p(up, np.float32), _p(ur, np.float32), C.c_float(r2_escape), C.c_uint32(npl),
        _p(axis, np.uint32), _p(loc, np.float32), _p(row, np.int64))
    return row


def scatter_f32(st, k, c, mode=0, seed=0, step=0, uniforms=None, id_base=0, want_flags=True):
    row = np.zeros(TALLY_COLS, np.int64)
    flags = np.zeros(st["x"].size, np.int32) if want_flags else None
    ut, up, ur = uniforms if uniforms is not None else (None, None, None)
    lib().orc_scatter_f32(
        C.c_uint64(st["x"].size), *[_p(st[q], np.float32) for q in ("x", "vx", "vy", "vz", "dx", "dy", "dz")],
        _p(st.get("e"), np.float32), _p(st.get("id"), np.uint32), _p(st.get("nscat"), np.uint32), C.c_uint64(id_base),
        C.c_float(k), C.c_float(c), C.c_uint32(mode), C.c_uint64(seed), C.c_uint32(step), _p(ut, np.float32),
        _p(up, np.float32), _p(ur, np.float32), _p(flags, np.int32), _p(row, np.int64))
    return flags, row


def tally_f32(st, planes=None):
    row = np.zeros(TALLY_COLS, np.int64)
    npl, axis, loc = _planes(planes)
    lib().orc_tally_f32(C.c_uint64(st["x"].size), *[_p(st[q], np.float32) for q in ("x", "y", "z", "vx", "vy", "vz")],
                        *[_p(st.get(q), np.float32) for q in ("dx", "dy", "dz")], C.c_uint32(npl), _p(axis, np.uint32),
                        _p(loc, np.float32), _p(row, np.int64))
    return row


def planck_sample(n, id_base, seed, cdf, e_lo, e_step):
    cdf = np.ascontiguousarray(cdf, np.float64)
    e = np.empty(n, np.float32)
    b = np.empty(n, np.int32)
    lib().orc_planck_sample(C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(seed), _p(cdf, np.float64),
                            C.c_uint32(cdf.size), C.c_float(e_lo), C.c_float(e_step), _p(e, np.float32), _p(b, np.int32))
    return e, b


# ---- double precision: the reference's law ---------------------------------------------------
def kinematics_f64(r, v, dt):
    """r, v: (3, N) float64 C-contiguous rows; returns dr (3, N); r updated in place."""
    dr = np.empty_like(r)
    lib().orc_kinematics_f64(C.c_uint64(r.shape[1]), *[_p(r[i], np.float64) for i in range(3)],
                             *[_p(v[i], np.float64) for i in range(3)], *[_p(dr[i], np.float64) for i in range(3)],
                             C.c_double(dt))
    return dr


def kinematics_accel_f64(r, v, a, dr, dt):
    """NEW law (v += a dt; dr = v dt; r += dr) in double; r, v, a, dr: (3, N) float64, updated in place."""
    lib().orc_kinematics_accel_f64(C.c_uint64(r.shape[1]), *[_p(r[i], np.float64) for i in range(3)],
                                   *[_p(v[i], np.float64) for i in range(3)], *[_p(a[i], np.float64) for i in range(3)],
                                   *[_p(dr[i], np.float64) for i in range(3)], C.c_double(dt))


def scatter_sphere_f64(dr, rtheta, rphi, rnd, A, n, c, E=None, hc=0.0):
    N = rnd.size
    res = np.full((3, N), np.nan)
    lib().orc_scatter_sphere_f64(C.c_uint64(N), *[_p(dr[i], np.float64) for i in range(3)], _p(rtheta, np.float64),
                                 _p(rphi, np.float64), _p(rnd, np.float64), C.c_double(A), C.c_double(n),
                                 _p(E, np.float64), C.c_double(hc), C.c_double(c), *[_p(res[i], np.float64) for i in range(3)])
    return res


def scatter_del_f64(dr, rnd, n, A):
    out = np.empty(rnd.size, np.int32)
    lib().orc_scatter_del_f64(C.c_uint64(rnd.size), *[_p(dr[i], np.float64) for i in range(3)], _p(rnd, np.float64),
                              C.c_double(n), C.c_double(A), _p(out, np.int32))
    return out


def photon_step_f64(st, dt, A, n, hc, c, mode=0, seed=0, step=0, r2_escape=0.0, id_base=0):
    """Whole reference timestep in double (CPU baseline). st: float64 planes x,y,z,vx,vy,vz,[E]."""
    row = np.zeros(TALLY_COLS, np.int64)
    lib().orc_photon_step_f64(C.c_uint64(st["x"].size), *[_p(st[q], np.float64) for q in ("x", "y", "z", "vx", "vy", "vz")],
                              _p(st.get("E"), np.float64), C.c_uint64(id_base), C.c_double(dt), C.c_double(A),
                              C.c_double(n), C.c_double(hc), C.c_double(c), C.c_uint32(mode), C.c_uint64(seed),
                              C.c_uint32(step), C.c_double(r2_escape), _p(row, np.int64))
    return row


def gravity_f64(pos, m, G, eps2, i_offset=0, n_local=None):
    """pos (3, N) float64, m (N,) -> acc (3, n_local). NEW step: oracle from the definition."""
    N = m.size
    n_local = N if n_local is None else n_local
    acc = np.empty((3, n_local))
    lib().orc_gravity_f64(C.c_uint64(n_local), C.c_uint64(i_offset), C.c_uint64(N), *[_p(pos[i], np.float64) for i in range(3)],
                          _p(m, np.float64), C.c_double(G), C.c_double(eps2), *[_p(acc[i], np.float64) for i in range(3)])
    return acc


def gravity_pick_f64(pos, m, G, eps2, idx):
    """Accelerations of the bodies ``idx`` from ALL bodies (float64 definition); returns (3, len(idx))."""
    idx = np.ascontiguousarray(idx, np.uint64)
    acc = np.empty((3, idx.size))
    lib().orc_gravity_pick_f64(C.c_uint64(idx.size), _p(idx, np.uint64), C.c_uint64(m.size), *[_p(pos[i], np.float64) for i in range(3)],
                               _p(m, np.float64), C.c_double(G), C.c_double(eps2), *[_p(acc[i], np.float64) for i in range(3)])
    return acc
