"""Step fusion: one kernel launch per timestep for the reference's canonical photon pipeline.

The reference runs, per timestep and per photon, NewtonianKinematicsStep -> Scatter*Step ->
measure steps (test/test_light.py:31-36, examples/*), each a separate pass over all particles.
When a simulation's step list contains that run, this module replaces it by ``FusedPhotonStep``,
which issues the fused photon kernel (kinematics + scatter + escape + sign/plane tallies, one HBM
round trip, dr kept in registers) and hands the tally row to the member measure steps.  Any other
step order runs unfused, kernel by kernel, in the user's order.

Retirement policy.  When the pipeline can retire photons (delete scattering, escape sphere) the
step runs on the store's ping-pong plane sets through ``pcl_photon_steps_pp``: every m-th timestep is
a retire-and-compact step, the others update in place.  In-place steps move ~44 B per SLOT, a
compacting step ~56 B per live photon, so with a death rate d per step the traffic per live
photon-step is about 44 + 12/m + 22 m d: minimal at m = sqrt(0.545 / d).  d is re-estimated from the
tallies every ``sim.compact_every`` timesteps (the only host<->device sync in the loop), which also
refreshes the host's upper bound on the slot count.
"""
from __future__ import annotations

import ctypes as C
import math

import physicl_b200 as physicl

from . import _capi, light, newton


class FusedPhotonStep(physicl.Step):
    uses_device = True

    def __init__(self, kin, scatter, escape, measures):
        self.kin, self.scatter, self.escape, self.measures = kin, scatter, escape, measures
        self.members = [kin, scatter] + ([escape] if escape else []) + list(measures)
        planes = []
        self._plane_slices = []
        for m in measures:
            pl = m._planes()
            self._plane_slices.append((len(planes), len(pl)))
            planes.extend(pl)
        if len(planes) > _capi.MAX_PLANES:
            raise ValueError("at most %d measurement planes per fused timestep" % _capi.MAX_PLANES)
        self._planes = _capi.make_planes(planes)
        self._multi_plane = sum(1 for _, n in self._plane_slices if n) > 1
        self.retires = bool(escape or scatter.mode & _capi.SCATTER_DELETE)
        self.cadence = 8  # m: every m-th timestep compacts; adapted at sync points
        self._last_live = None  # (step index, live count) at the previous sync point

    # ---- helpers --------------------------------------------------------------------------------
    def _fallback(self, st):
        return "object" in st.groups or st.group("photon") is None or self._multi_plane

    def can_run_many(self, sim):
        st = sim.device_store()
        return self.scatter.rng == "philox" and not self._fallback(st)

    def _note(self, sim, first, k, ts):
        for i in range(k):
            row = first + i
            if self.escape:
                self.escape._note_row(sim, row)
            for m, sl in zip(self.measures, self._plane_slices):
                m._note_row(sim, _FusedRow(row, sl), t=None if ts is None else ts[i])

    def _sync_point(self, sim, st, g, last_row, now):
        """Every sim.compact_every timesteps: read the live count (128-byte D2H), make the slot count
        exact again and re-derive the compaction cadence from the observed death rate."""
        live = int(st.peek_row(last_row)[_capi.T_ALIVE])
        g.n_live = live
        st.sync_n("photon")
        if getattr(sim, "compact_cadence", None):
            self.cadence = int(sim.compact_cadence)
        elif self._last_live is not None and self._last_live[1] > 0 and now > self._last_live[0]:
            ratio = max(live, 1) / self._last_live[1]
            d = 1.0 - ratio ** (1.0 / (now - self._last_live[0]))
            self.cadence = 64 if d <= 1e-4 else int(min(64, max(1, round(math.sqrt(0.545 / d)))))
        self._last_live = (now, live)

    # ---- k timesteps with one C-ABI call ----------------------------------------------------------
    def run_many(self, sim, k, dt, ts):
        """k fused timesteps of equal dt, launched back to back from C (``pcl_photon_steps`` /
        ``pcl_photon_steps_pp``): the host cost per timestep is a few hundred nanoseconds.  ``ts``:
        the simulation times of the k steps, for the measure rows."""
        st = sim.device_store()
        g = st.group("photon")
        if not self.retires:
            st.sync_n("photon")
        if g.n == 0:
            return
        sp = self.scatter.scatter_params(g)
        rng = _capi.Rng(seed=self.scatter._seed(sim), step=sim.step_index & 0xFFFFFFFF)
        first = st.new_rows(k)
        r2 = self.escape.R ** 2 if self.escape else 0.0
        for nm in ("dx", "dy", "dz"):  # dr stays in registers; stale planes would mislead host readers
            g.planes.pop(nm, None)
        if self.retires:
            if getattr(sim, "compact_cadence", None):
                self.cadence = int(sim.compact_cadence)
            pp = st.pingpong("photon")
            for b in pp.buf:
                b.dx = b.dy = b.dz = None
            m = self.cadence
            s0 = sim.step_index
            compacted = (s0 + k) // m - s0 // m
            sim.cl_ctx.call("pcl_photon_steps_pp", st.stream(), C.byref(pp), C.c_float(float(dt)), C.byref(sp), C.byref(rng),
                            C.c_float(r2), C.byref(self._planes), st.row_ptr(first), C.c_uint32(k), C.c_uint32(m))
            st.adopt_pingpong("photon", pp, compacted)
        else:
            soa = g.soa()
            soa.dx = soa.dy = soa.dz = None
            sim.cl_ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(float(dt)), C.byref(sp), C.byref(rng),
                            C.c_float(r2), C.byref(self._planes), st.row_ptr(first), C.c_uint32(k))
        self._note(sim, first, k, ts)
        last = first + k - 1
        sim._mark_device_dirty(live_row=last)
        if self.retires and sim.compact_every and (sim.step_index + k) % sim.compact_every == 0:
            self._sync_point(sim, st, g, last, sim.step_index + k)

    # ---- one timestep --------------------------------------------------------------------------------
    def run(self, sim):
        st = sim.device_store()
        if self._fallback(st):
            # mixed populations / several plane sets: run the member steps one by one
            for m in self.members:
                m.run(sim)
            return
        if self.scatter.rng == "philox":
            self.run_many(sim, 1, float(sim.dt), None)
            return
        # host-drawn uniforms (the reference's np.random stream): in place + stable compaction
        g = st.group("photon")
        if g.n == 0:
            return
        sp = self.scatter.scatter_params(g)
        rng, keep = self.scatter.rng_params(sim, st, g)
        row = st.new_row()
        for nm in ("dx", "dy", "dz"):
            g.planes.pop(nm, None)
        soa = g.soa()
        r2 = self.escape.R ** 2 if self.escape else 0.0
        sim.cl_ctx.call("pcl_photon_step", st.stream(), C.byref(soa), C.c_float(float(sim.dt)), C.byref(sp), C.byref(rng),
                        C.c_float(r2), C.byref(self._planes), st.row_ptr())
        st.synchronize()  # the injected uniform tensors must outlive the launch
        if self.retires:
            g.n_live = int(st.peek_row(row)[_capi.T_ALIVE])
        self._note(sim, row, 1, None)
        sim._mark_device_dirty(live_row=row)


class _FusedRow(int):
    """A tally row number that remembers which plane columns belong to the measure step."""

    def __new__(cls, row, plane_slice):
        o = int.__new__(cls, row)
        o.plane_slice = plane_slice
        return o


def fuse_plan(steps):
    out, i = [], 0
    while i < len(steps):
        s = steps[i]
        if (type(s) is newton.NewtonianKinematicsStep and not s.accel and i + 1 < len(steps)
                and isinstance(steps[i + 1], light._ScatterBase)):
            j = i + 2
            esc = None
            if j < len(steps) and type(steps[j]) is light.EscapeSphereStep:
                esc = steps[j]
                j += 1
            meas = []
            while j < len(steps) and type(steps[j]) in (light.ScatterSignMeasureStep, light.ScatterMeasureStep):
                meas.append(steps[j])
                j += 1
            out.append(FusedPhotonStep(s, steps[i + 1], esc, meas))
            i = j
        else:
            out.append(s)
            i += 1
    return out
