"""Newtonian steps.  Mirror of the reference's ``physicl/newton.py`` plus the two steps the
north star adds behind the same ``Step`` API (constant acceleration, all-pairs gravity)."""
from __future__ import annotations

import ctypes as C

import numpy as np

import physicl_b200 as physicl

from . import _capi


class NewtonianKinematicsStep(physicl.Step):
    """Moves all objects: ``dr = v * dt; r += dr`` (reference physicl/newton.py:14-16), as one
    streaming kernel over the HBM-resident planes.

    ``accel=True`` (not in the reference, where ``Object.a`` is never read) first does
    ``v += a * dt`` with ``a`` from the per-particle planes, or from ``a_uniform`` (3 floats)."""

    uses_device = True

    def __init__(self, accel=False, a_uniform=None, write_dr=True):
        self.accel = bool(accel)
        self.a_uniform = None if a_uniform is None else np.asarray(a_uniform, np.float32).reshape(3)
        self.write_dr = write_dr

    def can_run_many(self, sim):
        return True

    def chunk_steps(self, sim):
        return 256

    def run_many(self, sim, k, dt, ts):
        """k timesteps of equal dt with one C-ABI call (``pcl_kinematics_steps``): the particles stay
        in registers for all k steps, one HBM round trip, same bits as k single steps."""
        self._launch(sim, float(dt), int(k))

    def run(self, sim):
        self._launch(sim, float(sim.dt), 1)

    def _launch(self, sim, dt, k):
        st = sim.device_store()
        au = None
        if self.a_uniform is not None:
            au = self.a_uniform.ctypes.data_as(C.POINTER(C.c_float))
        for kind, g in st.groups.items():
            st.sync_n(kind)  # after a device-side compaction the exact slot count lives on the device
            if self.write_dr:
                g.ensure("dx", "dy", "dz")
            if self.accel and self.a_uniform is None and "ax" not in g.planes:
                g.ensure("ax", "ay", "az")
            soa = g.soa()
            if self.accel and self.a_uniform is not None:
                soa.ax = soa.ay = soa.az = None
            if k == 1:
                sim.cl_ctx.call("pcl_kinematics", st.stream(), C.byref(soa), C.c_float(dt), int(self.accel), au)
            else:
                sim.cl_ctx.call("pcl_kinematics_steps", st.stream(), C.byref(soa), C.c_float(dt), int(self.accel), au,
                                C.c_uint32(k))
        sim._mark_device_dirty()


class NewtonianGravityStep(physicl.Step):
    """All-pairs softened gravity followed by kick-drift (NOT in the reference; SURVEY.md section 8
    a14): ``a_i = G sum_j m_j (r_j - r_i) / (|r_ij|^2 + eps2)^(3/2)``; ``v += a dt``; ``r += v dt``.

    Bodies live in the ``object`` group with a packed ``(x, y, z, m)`` float4 array.  When the
    simulation is sharded, each rank owns a contiguous block of i-bodies and all-gathers the packed
    positions each step (NCCL over NVLink), overlapping the gather with the local-block tile loop."""

    uses_device = True

    def __init__(self, G=1.0, eps2=1e-4, masses=None):
        self.G, self.eps2 = float(G), float(eps2)
        self.masses = masses
        self._state = None

    def _setup(self, sim):
        import torch

        st = sim.device_store()
        g = st.group("object")
        if g is None:
            raise RuntimeError("NewtonianGravityStep acts on generic objects; none are present")
        n = g.n
        m = np.ones(n, np.float32) if self.masses is None else np.asarray(self.masses, np.float32).reshape(-1)
        # equal masses (BASELINE configs[3]): the common mass leaves the pair sum (pcl_gravity_accel_uniform)
        uniform = bool(m.size) and bool(np.all(m == m[0]))
        m0 = float(m[0]) if m.size else 1.0
        if m.size != n:
            lo = g.id_base
            m = m[lo:lo + n]
        posm = torch.empty((n, 4), dtype=torch.float32, device=st.device)
        posm[:, 0], posm[:, 1], posm[:, 2] = g.planes["x"][:n], g.planes["y"][:n], g.planes["z"][:n]
        posm[:, 3] = torch.from_numpy(m).to(st.device)
        acc = torch.zeros((3, n), dtype=torch.float32, device=st.device)
        self._state = dict(posm=posm, acc=acc, n=n, all=None, store=st, uniform=uniform, m0=m0)
        if sim.shard:
            from .dist import gravity_exchange

            self._state["xchg"] = gravity_exchange(posm, st.device)
        return self._state

    def run(self, sim):
        st = sim.device_store()
        s = self._state
        g = st.group("object")
        if s is None or s["store"] is not st or (g is not None and s["n"] != g.n):
            s = self._setup(sim)  # first use, or the store was rebuilt (host step, changed object list): repack
        n, posm, acc = s["n"], s["posm"], s["acc"]
        ctx, stream = sim.cl_ctx, st.stream()
        p = lambda t: C.c_void_p(t.data_ptr())
        fn = "pcl_gravity_accel_uniform" if s["uniform"] else "pcl_gravity_accel"
        g_eff = self.G * s["m0"] if s["uniform"] else self.G
        args = (C.c_float(g_eff), C.c_float(self.eps2), p(acc[0]), p(acc[1]), p(acc[2]))
        if "xchg" in s:
            s["xchg"].accelerations(ctx, st, posm, n, args, fn)
        else:
            ctx.call(fn, stream, p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), *args, 0, C.c_uint64(0), C.c_uint64(0))
        if "xchg" in s:
            s["xchg"].kick_drift(ctx, st, n, posm, g, acc, float(sim.dt))
        else:
            ctx.call("pcl_gravity_kick_drift", stream, C.c_uint64(n), p(posm), p(g.planes["vx"]), p(g.planes["vy"]),
                     p(g.planes["vz"]), p(acc[0]), p(acc[1]), p(acc[2]), C.c_float(float(sim.dt)),
                     p(g.planes["x"]), p(g.planes["y"]), p(g.planes["z"]))
        sim._mark_device_dirty()
