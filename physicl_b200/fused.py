"""Step fusion: one kernel launch per timestep for the reference's canonical photon pipeline.

The reference runs, per timestep and per photon, NewtonianKinematicsStep -> Scatter*Step ->
measure steps (test/test_light.py:31-36, examples/*), each a separate pass over all particles.
When a simulation's step list contains that run, this module replaces it by ``FusedPhotonStep``,
which issues ``pcl_photon_step`` (kinematics + scatter + escape + sign/plane tallies, one HBM round
trip, dr kept in registers) and hands the tally row to the member measure steps.  Any other step
order runs unfused, kernel by kernel, in the user's order.
"""
from __future__ import annotations

import ctypes as C

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

    def run(self, sim):
        st = sim.device_store()
        g = st.group("photon")
        if "object" in st.groups or g is None or self._multi_plane:
            # mixed populations / several plane sets: run the member steps one by one
            for m in self.members:
                m.run(sim)
            return
        if g.n == 0:
            return
        sp = self.scatter.scatter_params(g)
        rng, keep = self.scatter.rng_params(sim, st, g)
        row = st.new_row()
        soa = g.soa()
        soa.dx = soa.dy = soa.dz = None
        r2 = self.escape.R ** 2 if self.escape else 0.0
        sim.cl_ctx.call("pcl_photon_step", st.stream(), C.byref(soa), C.c_float(float(sim.dt)), C.byref(sp), C.byref(rng),
                        C.c_float(r2), C.byref(self._planes), st.row_ptr())
        if keep is not None:
            st.synchronize()
            if self.scatter.mode & _capi.SCATTER_DELETE:
                g.n_live = int(st.peek_row(row)[_capi.T_ALIVE])
        if "dx" in g.planes:  # stale once dr lives in registers only
            for nm in ("dx", "dy", "dz"):
                g.planes.pop(nm)
        if (self.escape or self.scatter.mode & _capi.SCATTER_DELETE) and sim.compact_every and \
                (sim.step_index + 1) % sim.compact_every == 0:
            # photons retire in this pipeline: every compact_every timesteps read the live count
            # (one 128-byte D2H) and squeeze the planes when enough slots are dead
            st.maybe_compact("photon", int(st.peek_row(row)[_capi.T_ALIVE]))
        if self.escape:
            self.escape._note_row(sim, row)
        for m in self.measures:
            m._note_row(sim, _FusedRow(row, self._plane_slices[self.measures.index(m)]))
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
