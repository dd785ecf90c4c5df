"""Helpers for the GPU parity tests: every call goes through the C ABI (``physicl_b200._capi``)."""
import ctypes as C

import numpy as np

from physicl_b200 import _capi
from physicl_b200.store import DeviceParticleStore

C_LIGHT = 299792458.0
PLANE_NAMES = ("x", "y", "z", "vx", "vy", "vz")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def random_photons(n, seed, spread=1e5, c=C_LIGHT):
    """Random positions, unit directions scaled to c (float32-rounded), as float64 (3, n) arrays."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(-spread, spread, (3, n))
    d = rng.normal(size=(3, n))
    v = (c * d / np.linalg.norm(d, axis=0)).astype(np.float32).astype(np.float64)
    return r.astype(np.float32).astype(np.float64), v


def beam_photons(n, c=C_LIGHT):
    """All photons at the origin moving along +x (reference test/test_light.py:12-17)."""
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = c
    return r, v


def make_store(ctx, r, v, E=None, a=None, kind="photon", id_base=0, nscat=False):
    st = DeviceParticleStore(ctx)
    g = st.add_group(kind, r, v, E=E, a=a, id_base=id_base, track_nscat=nscat)
    return st, g


def host_state(g, extra=()):
    """float32 host copy of a group's planes, as the oracle twin wants them."""
    names = list(PLANE_NAMES) + [nm for nm in ("dx", "dy", "dz", "ax", "ay", "az", "e", "id", "nscat") if nm in g.planes]
    return {nm: g.download(nm).copy() for nm in names}


def photon_step(ctx, st, g, dt, k, c, mode=0, seed=0, step=0, uniforms=None, r2_escape=0.0, planes=None):
    """One fused step on the device; returns the tally row (int64[16])."""
    import torch

    sp = _capi.ScatterParams(k=k, c=c, mode=mode)
    rg = _capi.Rng(seed=seed, step=step)
    keep = None
    if uniforms is not None:
        keep = [None if u is None else torch.from_numpy(np.ascontiguousarray(u, np.float32)).to(st.device) for u in uniforms]
        rg.u_theta, rg.u_phi, rg.u_rand = (None if t is None else t.data_ptr() for t in keep)
    pl = _capi.make_planes(planes)
    row = st.new_row()
    soa = g.soa()
    soa.dx = soa.dy = soa.dz = None
    ctx.call("pcl_photon_step", st.stream(), C.byref(soa), C.c_float(dt), C.byref(sp), C.byref(rg), C.c_float(r2_escape),
             C.byref(pl), st.row_ptr())
    st.synchronize()
    return st.read_row(row)
