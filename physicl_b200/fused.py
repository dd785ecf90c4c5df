"""Step fusion: one kernel launch per timestep for the reference's canonical photon pipeline.

The reference runs, per timestep and per photon, NewtonianKinematicsStep -> Scatter*Step ->
measure steps (test/test_light.py:31-36, examples/*), each a separate pass over all particles.
When a simulation's step list contains that run, this module replaces it by ``FusedPhotonStep``,
which issues the fused photon kernel (kinematics + scatter + escape + sign/plane tallies, one HBM
round trip, dr kept in registers) and hands the tally row to the member measure steps.  Any other
step order runs unfused, kernel by kernel, in the user's order.

Several timesteps per HBM round trip.  In the bulk form (``Simulation.run_steps``) a launch keeps its
photons in registers for up to 8 timesteps (``pcl_k_photon_multi``; results identical bit for bit to
one launch per timestep), so the state crosses HBM once per launch and the kernel is bound by the
integer/FP32 pipes, not by DRAM.

Retirement policy.  When the pipeline can retire photons (delete scattering, escape sphere) the
step runs on the store's ping-pong plane sets through ``pcl_photon_steps_pp``: a launch covers m
timesteps and writes the survivors of the last one densely into the partner buffer, ALWAYS (a
compacting launch moves 56 B per slot against 48 B for an in-place one, and neither is the bound).
Measured launch time at 16 Mi live photons (scripts/time_multi.py): 96 us + 54 us per timestep, i.e. the
load + compaction phase of a launch costs F = 1.8 timesteps; photons that die inside a launch idle in their
lanes until it ends (a fraction ~ d (m-1)/2 of the slots for a death rate d per step).  Cost per live
photon-step ~ (F + m) / (m (1 - d (m-1)/2)): m = 8 wins below d ~ 5 % per step, above it the curve is flat
between m = 5 and m = 8, so m is 8 while few photons die and 5 otherwise.  d comes from the tally rows
with a lag: after every chunk (one C-ABI call, about ``sim.feedback_every`` timesteps) the last row and the
device slot counters are copied to pinned host memory asynchronously, and the host reads them one or two
chunks later (it only ever waits on work the GPU has long finished), so the stepping loop has no blocking
host<->device round trip.
"""
from __future__ import annotations

import ctypes as C

import physicl_b200 as physicl

from . import _capi, light, newton


class FusedPhotonStep(physicl.Step):
    uses_device = True
    tallies_every_timestep = True  # one tally row per timestep: Simulation._run_chunked can replay exit predicates

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
        # variable-density steps run a run-time compiled kernel (light.py:295-299), in place only
        self.varn = bool(getattr(scatter, "variable_n", False))
        self.cadence = 5  # m: timesteps per compacting launch (5 or 8, from lagged tally feedback)
        self._fb = []  # pending feedback: (event, pinned int64[18], buffer index at enqueue time)
        self._fb_pool = []

    # ---- helpers --------------------------------------------------------------------------------
    def _fallback(self, st):
        g = st.group("photon")
        # photons that carry acceleration planes: the retire-and-compact kernel does not move those, the stand-alone
        # steps + pcl_compact do
        return "object" in st.groups or g is None or self._multi_plane or (self.retires and "ax" in g.planes)

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

    def chunk_steps(self, sim):
        """Timesteps per C-ABI call: a whole number of compaction periods close to sim.feedback_every."""
        fe = int(sim.feedback_every) if sim.feedback_every else 64
        if not self.retires or self.varn:
            return max(fe, 1)
        m = int(sim.compact_cadence) if getattr(sim, "compact_cadence", None) else self.cadence
        if m >= fe:
            return max(fe, 1)
        # end on a compaction boundary (every m-th timestep of the global step count)
        return (m - sim.step_index % m) + m * (max(1, round(fe / m)) - 1)

    def _enqueue_feedback(self, st, g, last_row):
        """Async copy of a chunk's last tally row and of the device slot counters to pinned memory."""
        import torch

        if not self._fb_pool and not self._fb:  # first use: page-locking is slow, do it once (in the warm-up)
            self._fb_pool = [torch.empty(18, dtype=torch.int64).pin_memory() for _ in range(4)]
        buf = self._fb_pool.pop() if self._fb_pool else torch.empty(18, dtype=torch.int64).pin_memory()
        buf[:16].copy_(st.tally[last_row - st._row_base], non_blocking=True)
        buf[16:].copy_(g.n_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(st.device))
        self._fb.append((ev, buf, g.cur))

    def _consume_feedback(self, sim, g, max_pending=2):
        """Use whatever feedback has already arrived (never waits unless more than max_pending chunks
        are outstanding, i.e. the GPU is at least that far behind the host anyway)."""
        while self._fb and (self._fb[0][0].query() or len(self._fb) > max_pending):
            ev, buf, cur = self._fb.pop(0)
            ev.synchronize()
            row = buf[:16].numpy()
            live_in, alive = int(row[_capi.T_LIVE_IN]), int(row[_capi.T_ALIVE])
            died = int(row[_capi.T_ESCAPED]) + int(row[_capi.T_ABSORBED])
            g.n_live = alive
            slots = int(buf[16 + cur])
            if 0 < slots < g.n and not g.n_exact:
                g.n = slots  # slot counts only shrink: a stale exact count is still an upper bound
            if getattr(sim, "compact_cadence", None):
                self.cadence = int(sim.compact_cadence)
            elif live_in > 0:
                # two levels are enough (the cost curve is flat between 5 and 8 above d ~ 5 %, see the module docstring),
                # and a stale estimate then costs a few per cent at worst
                self.cadence = 8 if died / live_in < 0.05 else 5
            self._fb_pool.append(buf)

    # ---- roll-back support for Simulation._run_chunked ------------------------------------------------
    def checkpoint(self):
        return {"escape": len(self.escape._rows) if self.escape else 0, "measures": [len(m._pending) for m in self.measures],
                "cadence": self.cadence}

    def rollback(self, ck):
        if self.escape:
            del self.escape._rows[ck["escape"]:]
        for m, n in zip(self.measures, ck["measures"]):
            del m._pending[n:]
        self.cadence = ck["cadence"]
        while self._fb:  # feedback of the discarded timesteps must not shrink the restored slot count
            ev, buf, _ = self._fb.pop()
            ev.synchronize()
            self._fb_pool.append(buf)

    # ---- k timesteps with one C-ABI call ----------------------------------------------------------
    def run_many(self, sim, k, dt, ts):
        """k fused timesteps of equal dt, launched back to back from C (``pcl_photon_steps`` /
        ``pcl_photon_steps_pp``): the host cost per timestep is a few hundred nanoseconds.  ``ts``:
        the simulation times of the k steps, for the measure rows."""
        st = sim.device_store()
        g = st.group("photon")
        if not self.retires or self.varn:
            st.sync_n("photon")
        if g.n == 0:  # every photon is gone: the measure steps still get their (all-zero) rows, as in the reference
            self._note(sim, st.new_rows(k), k, ts)
            return
        sp = self.scatter.scatter_params(g)
        rng = _capi.Rng(seed=self.scatter._seed(sim), step=sim.step_index & 0xFFFFFFFF)
        first = st.new_rows(k)
        r2 = self.escape.R ** 2 if self.escape else 0.0
        for nm in ("dx", "dy", "dz"):  # dr stays in registers; stale planes would mislead host readers
            g.planes.pop(nm, None)
        if self.varn:
            soa = g.soa()
            soa.dx = soa.dy = soa.dz = None
            vn = self.scatter.varn_params(g)
            sim.cl_ctx.call("pcl_photon_steps_jit", st.stream(), self.scatter.jit_kernel(sim.cl_ctx, "pcl_jit_photon_step"),
                            C.byref(soa), C.c_float(float(dt)), C.byref(sp), C.byref(vn), C.byref(rng), C.c_float(r2),
                            C.byref(self._planes), st.row_ptr(first), C.c_uint32(k))
        elif self.retires:
            self._consume_feedback(sim, g)
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
        if self.retires and not self.varn and sim.feedback_every:
            self._enqueue_feedback(st, g, last)

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
            self._note(sim, st.new_row(), 1, None)
            return
        sp = self.scatter.scatter_params(g)
        rng, keep = self.scatter.rng_params(sim, st, g)
        row = st.new_row()
        for nm in ("dx", "dy", "dz"):
            g.planes.pop(nm, None)
        soa = g.soa()
        r2 = self.escape.R ** 2 if self.escape else 0.0
        if self.varn:
            vn = self.scatter.varn_params(g)
            sim.cl_ctx.call("pcl_photon_steps_jit", st.stream(), self.scatter.jit_kernel(sim.cl_ctx, "pcl_jit_photon_step"),
                            C.byref(soa), C.c_float(float(sim.dt)), C.byref(sp), C.byref(vn), C.byref(rng), C.c_float(r2),
                            C.byref(self._planes), st.row_ptr(), C.c_uint32(1))
        else:
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
    if any(getattr(s, "needs_dr", False) for s in steps):
        return list(steps)  # a step reads the dr planes (e.g. measure_E): kinematics must write them
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
