"""physicl_b200: a B200 (sm_100a) backend behind PhysiCL's ``Simulation`` / ``Step`` API.

Host-side mirror of ``physicl/__init__.py`` of the reference (bcwarner/physicl): same class names,
same argument meaning, same error behaviour for the per-particle step path, so a user of the
reference can ``import physicl_b200 as physicl``.  What differs is underneath: particle state is
uploaded once into structure-of-arrays float32 planes in HBM (``store.DeviceParticleStore``) and
every step is a hand-written CUDA kernel reached through a C ABI (``include/physicl_b200.h``).

There is no CPU fallback: ``cl_on=True`` (the reference's "use the device" switch,
physicl/__init__.py:413) needs the built library and a B200; device steps raise otherwise.
``cl_on=False`` builds a host-only ``Simulation`` (time stepping, user steps, measure-step files).
"""
from __future__ import annotations

import copy
import gc
import os
import threading
import time

import numpy as np

from .units import Measurement, MeasurementError  # noqa: F401

__all__ = ["Measurement", "MeasurementError", "Step", "UpdateTimeStep", "MeasureStep", "Object", "Simulation",
           "IndexException"]


class IndexException(NameError):
    """Raised by ``Simulation.add_step`` for a duplicate index.  The reference raises an undefined
    name there (physicl/__init__.py:441), i.e. a ``NameError``; this class keeps both readings."""


class Step:
    """Plugin protocol of the reference (physicl/__init__.py:293-322): ``run(sim)`` once per
    timestep, ``terminate(sim)`` once at the end.  ``uses_device`` marks steps that operate on the
    HBM-resident store; everything else is treated as a host step that reads ``sim.objects``: the objects are made
    current before it runs and re-uploaded before the next device step.  A host step that only READS the objects (a
    measurement written in Python) can set ``modifies_objects = False`` to skip that re-upload (new, optional; the
    default is the safe one)."""

    uses_device = False

    def __init__(self):
        pass

    def __compile_cl__(self, sim):
        """physicl/__init__.py:300-304: hook for a step's kernel build; a no-op in the base class, as there."""
        pass

    def __run_cl(self, sim):
        pass

    def __run_py(self, sim):
        pass

    def run(self, sim):
        pass

    def terminate(self, sim):
        pass


class UpdateTimeStep(Step):
    """physicl/__init__.py:324-343: ``dt = fn(sim); t += dt; ts.append(copy of t)``."""

    touches_objects = False

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def run(self, sim):
        sim.dt = self.fn(sim)
        sim.t += sim.dt
        sim.ts.append(copy.deepcopy(sim.t))


class MeasureStep(Step):
    """physicl/__init__.py:345-378: accumulates ``data`` and writes comma-separated rows on
    ``terminate`` when an output file name was given."""

    def __init__(self, out_fn=None):
        self.out_fn = out_fn
        self.data = []

    def run(self, sim):
        pass

    def terminate(self, sim):
        if self.out_fn is None:
            return
        rows = self.data.values() if isinstance(self.data, dict) else self.data
        with open(self.out_fn, "w") as f:
            for row in rows:
                f.write(", ".join(str(i) for i in list(row)) + "\n")


class Object:
    """physicl/__init__.py:381-396: position ``r``, last displacement ``dr``, ``dv``, velocity ``v``,
    acceleration ``a`` (3-vector Measurements) plus arbitrary keyword attributes (e.g. ``E``)."""

    def __init__(self, **kwargs):
        self.r = Measurement([0] * 3, "m**1")
        self.dr = Measurement([0] * 3, "m**1")
        self.dv = Measurement([0] * 3, "m**1 s**-2")
        self.v = Measurement([0] * 3, "m**1 s**-1")
        self.a = Measurement([0] * 3, "m**1 s**-2")
        for attr, val in kwargs.items():
            setattr(self, attr, val)


def _gather3(objs, attr):
    """(3, N) float64 array of a 3-vector attribute of every object (code units: a Measurement's stored values)."""
    try:  # the common case, all (3,) arrays: one C-level conversion of the list
        out = np.array([getattr(o, attr) for o in objs], dtype=np.float64)
        if out.shape == (len(objs), 3):
            return out.T
    except (TypeError, ValueError):
        pass
    if not objs:
        return np.zeros((3, 0))
    return np.array([np.asarray(getattr(o, attr), np.float64).reshape(3) for o in objs]).T  # lists, (3, 1) columns, ...


class _ObjectList(list):
    """``sim.objects``: a list that knows when the device store is the authority.

    Mutations mark the host side dirty (the store is rebuilt before the next device step);
    ``len()`` answers from the device's live count when device steps have retired photons, so the
    reference's default exit predicate ``len(x.objects) == 0`` (physicl/__init__.py:414) works
    without copying particles back."""

    def __init__(self, sim):
        super().__init__()
        self._sim = sim

    def _touch(self):
        self._sim._host_dirty = True

    def append(self, o):
        self._touch()
        super().append(o)

    def extend(self, it):
        self._touch()
        super().extend(it)

    def remove(self, o):
        self._sim._pull_objects()
        self._touch()
        super().remove(o)

    def insert(self, i, o):
        self._touch()
        super().insert(i, o)

    def pop(self, *a):
        self._sim._pull_objects()
        self._touch()
        return super().pop(*a)

    def clear(self):
        self._touch()
        super().clear()

    def __len__(self):
        sim = self._sim
        sim._probe_touched = True  # an exit predicate that gets here looks at the particles (Simulation._run_chunked)
        if sim._len_override is not None:  # replay of an exit predicate against the tally row of an earlier timestep
            return sim._len_override
        if sim._device_dirty and sim.store is not None:
            return sim._device_live_count()
        n = super().__len__()
        for p in (sim._pending or {}).values():  # bulk particles that have not been uploaded yet
            r = p["r"]
            n += int(r.shape[1]) if hasattr(r, "shape") and len(r.shape) == 2 else int(np.asarray(r).reshape(3, -1).shape[1])
        if sim.store is not None and not sim._host_dirty:  # bulk groups live on the device only
            n += sum(g.n_live for g in sim.store.groups.values() if g.host_objs is None)
        return n

    def __iter__(self):
        self._sim._probe_touched = True
        self._sim._pull_objects()
        return super().__iter__()

    def __getitem__(self, i):
        self._sim._probe_touched = True
        self._sim._pull_objects()
        return super().__getitem__(i)


class Simulation(threading.Thread):
    """physicl/__init__.py:400-541.  A thread that repeatedly runs its steps, in insertion order,
    until ``exit(sim)`` is true, then calls every step's ``terminate``.

    Keyword arguments become attributes exactly as in the reference (``bounds``, ``cl_on``, ``exit``,
    ``state_fn``, ``state_need_lock``).  Added, all optional: ``device`` (CUDA ordinal; default
    ``$PHYSICL_B200_DEVICE`` or ``$LOCAL_RANK`` or 0), ``seed`` (Philox key for in-kernel draws),
    ``fuse`` (merge kinematics+scatter+escape+tallies into one kernel per timestep, default True),
    ``shard`` (under ``torch.distributed``: this rank owns a contiguous slice of the particles)."""

    def __init__(self, **kwargs):
        threading.Thread.__init__(self)
        self.bounds = np.zeros(3)
        self.cl_on = True
        self.exit = lambda x: len(x.objects) == 0
        self.state_fn = lambda x: {"objects": len(x.objects), "t": x.t, "dt": x.dt, "run_time": time.time() - x.start_time}
        self.state_need_lock = False
        self.device = None
        self.seed = 0
        self.fuse = True
        self.shard = False
        # timesteps per device chunk (one C-ABI call, one asynchronous tally read-back).  Large on purpose: a chunk is
        # queued in ~0.1 ms of host time, so the GPU keeps running through host-side hiccups (other ranks' threads,
        # the scheduler); the retirement policy needs no finer feedback (physicl_b200/fused.py)
        self.feedback_every = 64
        self.compact_cadence = None  # survivors are compacted after every m-th timestep; None = chosen by the fused step
        for attr, val in kwargs.items():
            setattr(self, attr, val)
        self.dt = Measurement(np.double(0), "s**1")
        self.t = Measurement(np.double(0), "s**1")
        self.ts = []
        self.store = None
        self._host_dirty = True
        self._device_dirty = False
        self._pending = None  # bulk particles registered with add_particles()
        self._len_override = None
        self._probe_touched = False  # set by _ObjectList accessors: did the exit predicate look at the particles?
        self.objects = _ObjectList(self)
        self.steps = {}
        self._state_lock = threading.Lock()
        self.running = False
        self.start_time = 0
        self.step_index = 0
        self.error = None
        if self.cl_on:
            from . import _capi

            if self.device is None:
                self.device = int(os.environ.get("PHYSICL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            # stands in for cl.create_some_context() + cl.CommandQueue() (physicl/__init__.py:428-429)
            self.cl_ctx = _capi.Context(self.device)
            self.cl_q = self.cl_ctx
        else:
            self.cl_ctx = None
            self.cl_q = None

    # ---- registration (physicl/__init__.py:434-468) -------------------------------------------
    def add_step(self, idx, step):
        if idx in self.steps:
            raise IndexException("Cannot add a step to an existing index.")
        self.steps[idx] = step

    def add_obj(self, obj):
        self.objects.append(obj)

    def add_objs(self, objs):
        self.objects.extend(objs)

    def remove_obj(self, obj):
        self.objects.remove(obj)

    def remove_step(self, idx):
        if self.running:
            raise RuntimeError("Cannot remove a Step while the simulation is running.")
        self.steps.pop(idx)

    def add_particles(self, r, v, E=None, a=None, kind="photon", track_nscat=False, id_base=None):
        """Bulk ingest for particle counts where one Python object per particle is not an option
        (hard part 7).  ``r``, ``v``, ``a``: ``(3, N)`` arrays in code units;
        ``E``: ``(N,)``.  Additive to ``add_obj``; particles added this way have no Python objects
        until something asks for them.  ``id_base``: global id of the first particle when the caller
        has already cut its own shard (the Philox counter is the global id)."""
        if self._pending is None:
            self._pending = {}
        if kind in self._pending:
            raise ValueError("add_particles: one call per kind")
        self._pending[kind] = dict(r=r, v=v, E=E, a=a, track_nscat=track_nscat, id_base=id_base)
        self._host_dirty = True

    # ---- device store bridge ------------------------------------------------------------------
    def _shard_slice(self, n):
        if not self.shard:
            return 0, n
        from .dist import shard_range

        return shard_range(n)

    def device_store(self):
        """The HBM-resident store, (re)built from ``sim.objects`` / ``add_particles`` when the host
        side changed since the last device step."""
        if not self.cl_on:
            raise RuntimeError("physicl_b200 has no CPU path: device steps need Simulation(cl_on=True)")
        if self.store is not None and not self._host_dirty:
            return self.store
        from .store import DeviceParticleStore
        from . import light

        self._pull_objects()
        store = DeviceParticleStore(self.cl_ctx)
        objs = list.__iter__(self.objects)
        photons, others = [], []
        for o in objs:
            (photons if type(o) is light.PhotonObject else others).append(o)
        for kind, group in (("photon", photons), ("object", others)):
            if group:
                lo, hi = self._shard_slice(len(group))
                mine = group[lo:hi]
                r = _gather3(mine, "r")
                v = _gather3(mine, "v")
                E = None
                if kind == "photon":
                    try:  # every photon has a numeric E: one C-level conversion
                        E = np.array([o.E for o in mine], dtype=np.float64).reshape(len(mine))
                    except (AttributeError, TypeError, ValueError):  # E = None is allowed (planck_phot_distribution may return it)
                        E = np.array([np.nan if getattr(o, "E", None) is None else float(np.asarray(o.E)) for o in mine])
                a = _gather3(mine, "a")
                g = store.add_group(kind, r, v, E=E, a=a if np.any(a) else None, id_base=lo, host_objs=mine)
                # Object.dr (physicl/__init__.py:391) is state too: measure steps that run after a host step read the
                # displacement of this timestep (light.py:385-399), so it travels with the rebuild
                dr = _gather3(mine, "dr")
                if np.any(dr):
                    for q, nm in enumerate(("dx", "dy", "dz")):
                        g.upload(nm, dr[q])
                if any(hasattr(o, "nscat") for o in mine):
                    g.upload("nscat", np.array([int(getattr(o, "nscat", 0)) for o in mine], np.uint32))
        for kind, p in (self._pending or {}).items():
            if hasattr(p["r"], "is_cuda") and p["r"].is_cuda:  # device tensors: this rank's block as is
                store.add_group(kind, p["r"], p["v"], E=p["E"], a=p["a"], id_base=int(p["id_base"] or 0),
                                track_nscat=p["track_nscat"])
                continue
            r = np.asarray(p["r"]).reshape(3, -1)
            lo, hi = self._shard_slice(r.shape[1]) if p["id_base"] is None else (0, r.shape[1])
            sl = slice(lo, hi)
            lo = lo if p["id_base"] is None else int(p["id_base"])
            store.add_group(kind, r[:, sl], np.asarray(p["v"]).reshape(3, -1)[:, sl],
                            E=None if p["E"] is None else np.asarray(p["E"])[sl],
                            a=None if p["a"] is None else np.asarray(p["a"]).reshape(3, -1)[:, sl],
                            id_base=lo, track_nscat=p["track_nscat"])
        # the device store is the only authority for bulk particles from here on: a later rebuild pulls them back as
        # objects (they must not be ingested a second time from the caller's original arrays)
        self._pending = None
        self.store = store
        self._host_dirty = False
        self._device_dirty = False
        self._live_row = None
        return store

    def _mark_device_dirty(self, live_row=None):
        self._device_dirty = True
        if live_row is not None:
            self._live_row = live_row

    def _device_live_count(self):
        """Live particles according to the device (all ranks when sharded)."""
        st = self.store
        n = 0
        for kind, g in st.groups.items():
            if kind == "photon" and getattr(self, "_live_row", None) is not None:
                n_live = int(st.peek_row(self._live_row)[0])
                # get_state() may call this from another thread while the simulation thread is stepping
                # (physicl/__init__.py:532-541 takes the lock only on request): only the owner of the
                # stepping loop may move planes around
                if not self.running or threading.current_thread() is self:
                    st.maybe_compact("photon", n_live)
                n += n_live
            else:
                n += g.n_live
        if self.shard:
            from .dist import all_reduce_int

            n = all_reduce_int(n)
        return n

    def _pull_objects(self):
        """Make the Python objects current again (bulk D2H, then per-object attribute writes).
        Host steps and user code that walk ``sim.objects`` get here; the hot path never does."""
        if not self._device_dirty or self.store is None:
            return
        self._device_dirty = False
        st = self.store
        keep = []
        gc_was_on = gc.isenabled()
        gc.disable()  # millions of small allocations and no cycles: the collector's generation scans would double the time
        try:
            self._pull_groups(st, keep)
        finally:
            if gc_was_on:
                gc.enable()
        list.clear(self.objects)
        list.extend(self.objects, keep)

    def _pull_groups(self, st, keep):
        for kind, g in st.groups.items():
            snap = st.snapshot(kind, live_only=True)
            fresh = g.host_objs is None
            if fresh:
                from . import light

                cls = light.PhotonObject if kind == "photon" else Object
                g.host_objs = {}
                for i in snap["id"]:
                    o = cls.__new__(cls)
                    Object.__init__(o)
                    o._pcl_origin = (id(st), kind, int(i))  # which bulk particle this object stands for
                    g.host_objs[int(i)] = o
            # whole-group arrays first, then one row view per attribute and object (a Measurement built from a list
            # parses units and copies; with 10^4..10^6 objects that was most of the wall time of a pulled run)
            ids = [int(i) for i in snap["id"]]
            objs = [g.host_objs[i] for i in ids]
            n_o = len(objs)
            one = np.double(1)
            V = np.stack([snap["vx"], snap["vy"], snap["vz"]], axis=1).astype(np.float64) if n_o else np.zeros((0, 3))
            R = np.stack([snap["x"], snap["y"], snap["z"]], axis=1).astype(np.float64) if n_o else np.zeros((0, 3))
            # Object.dv (physicl/__init__.py:392; written by the scatter step, light.py:325-331: v_new - v_old for a
            # photon that scattered, zero otherwise): the change of v since the object was last current on the host.
            # With a host step in the pipeline that is once per timestep, i.e. exactly the reference's value.
            if fresh or not n_o:
                DV = np.zeros((n_o, 3))
            else:  # state is binary32 on the device
                DV = V - np.array([o.v for o in objs], dtype=np.float64).reshape(n_o, 3).astype(np.float32).astype(np.float64)
            DR = np.stack([snap["dx"], snap["dy"], snap["dz"]], axis=1).astype(np.float64) if "dx" in snap and n_o else None
            E = snap["E"] if "E" in snap and kind == "photon" else None
            NS = snap["nscat"] if "nscat" in snap else None
            M = Measurement
            for j, o in enumerate(objs):
                m = DV[j].view(M)
                m.scale, m.units, m.original_units = one, {"L": 1, "T": -1}, {"m": 1, "s": -1}
                o.dv = m
                m = R[j].view(M)
                m.scale, m.units, m.original_units = one, {"L": 1}, {"m": 1}
                o.r = m
                m = V[j].view(M)
                m.scale, m.units, m.original_units = one, {"L": 1, "T": -1}, {"m": 1, "s": -1}
                o.v = m
                if DR is not None:
                    m = DR[j].view(M)
                    m.scale, m.units, m.original_units = one, {"L": 1}, {"m": 1}
                    o.dr = m
                if E is not None:
                    o.E = np.double(E[j])
                if NS is not None:
                    o.nscat = int(NS[j])
            keep.extend(objs)

    # ---- main loop (physicl/__init__.py:501-524) ----------------------------------------------
    def _plan(self):
        """Execution plan for one timestep: the steps in insertion order, with a maximal
        kinematics -> scatter [-> escape] [-> tallies...] run replaced by one fused launch."""
        steps = list(self.steps.values())
        for ordinal, s in enumerate(steps):  # distinct Philox keys for distinct stochastic steps
            if hasattr(s, "_salt"):
                s._salt = ((ordinal + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        for s in steps:  # one-time set-up that must precede the first timestep (e.g. scatter counters)
            if hasattr(s, "prepare"):
                s.prepare(self)
        if not (self.fuse and self.cl_on):
            return steps
        # the fused plan (and with it the fused step's adaptive cadence and pinned feedback buffers) is kept as
        # long as the step list stays the same: page-locking buffers inside every run_steps call costs
        # milliseconds when several ranks do it at once
        key = tuple(id(s) for s in steps)
        cached = getattr(self, "_plan_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        from .fused import fuse_plan

        plan = fuse_plan(steps)
        self._plan_cache = (key, plan, steps)  # `steps` keeps the ids alive
        return plan

    def run(self):
        self.start_time = time.time()
        self.t = 0
        self.dt = 0
        self.ts = []
        self.running = True  # step_index is NOT reset: it is the Philox step counter and must never repeat
        try:
            if self.cl_on and self.cl_ctx is not None:  # a new host thread: its current CUDA device is 0 until told otherwise
                import torch

                torch.cuda.set_device(self.device)
            plan = self._plan()
            fast = self._bulk_plan(plan)
            if fast is not None:
                self._run_chunked(*fast)
            else:
                while not self.exit(self):
                    with self._state_lock:
                        for step in plan:
                            if not step.uses_device and getattr(step, "touches_objects", True):
                                self._pull_objects()
                                # a host step may change the objects: the store is rebuilt from them before the next device
                                # step -- unless the step says it only reads (modifies_objects = False, see Step)
                                if getattr(step, "modifies_objects", True):
                                    self._host_dirty = self._host_dirty or self.store is not None
                            step.run(self)
                        self.step_index += 1
            with self._state_lock:
                for step in self.steps.values():
                    step.terminate(self)
        except BaseException as e:  # re-raised by join(): a dead thread must not look like success
            self.error = e
        finally:
            self.run_time = time.time() - self.start_time
            self.running = False

    # ---- chunked form of the main loop ------------------------------------------------------------------
    def _bulk_plan(self, plan):
        """(UpdateTimeStep, fused step) when the whole plan is that pair and the fused step can advance several
        timesteps per C-ABI call; None otherwise."""
        if (self.cl_on and len(plan) == 2 and type(plan[0]) is UpdateTimeStep and hasattr(plan[1], "run_many")
                and plan[1].can_run_many(self)):
            return plan[0], plan[1]
        return None

    def _advance(self, upd, fused, k, probe_exit=False):
        """Up to k timesteps: UpdateTimeStep on the host (stopping early when dt changes or, with ``probe_exit``, when
        the exit predicate fires or starts looking at the particles), then ONE bulk call for all of them.
        Returns the (t, dt) of every timestep done."""
        dts, ts = [], []
        while len(dts) < k:
            upd.run(self)
            dts.append(float(self.dt))
            ts.append(self.t)
            if dts[-1] != dts[0]:
                break
            if probe_exit and len(dts) < k:
                self._probe_touched = False
                if self.exit(self) or self._probe_touched:
                    break
        if dts[-1] != dts[0]:  # dt changed inside the chunk: these timesteps go one by one
            for dt_i, t_i in zip(dts, ts):
                fused.run_many(self, 1, dt_i, [t_i])
                self.step_index += 1
        else:
            fused.run_many(self, len(dts), dts[0], ts)
            self.step_index += len(dts)
        return list(zip(ts, dts))

    def _run_chunked(self, upd, fused):
        """``while not exit(sim): one timestep`` (physicl/__init__.py:512-516) in chunks of timesteps, stopping at
        EXACTLY the timestep the per-step loop would stop at.

        A predicate that only reads ``t`` / ``dt`` / ``ts`` (``t >= 0.1``, reference test/test_light.py:20) is evaluated
        ahead of the device: after every host-side UpdateTimeStep of the chunk the predicate is asked again, and the
        timesteps counted that way go to the device in one call.  A predicate that looks at the particles
        (``len(objects) == 0``, the reference's default, physicl/__init__.py:414) is checked AFTER the chunk, against
        the tally row of every timestep in it; if it fired early the chunk is rolled back to a device-side copy taken
        before it and re-run up to the firing timestep (the draws are a pure function of particle id and step index,
        so the re-run reproduces those timesteps bit for bit)."""
        while True:
            self._probe_touched = False
            if self.exit(self):
                return
            looks_at_particles = self._probe_touched
            kmax = max(1, min(fused.chunk_steps(self), 256))
            with self._state_lock:
                if not looks_at_particles:
                    self._advance(upd, fused, kmax, probe_exit=True)
                elif not getattr(fused, "tallies_every_timestep", False):
                    self._advance(upd, fused, 1)  # no per-timestep rows to replay the predicate against
                else:
                    ck = self._checkpoint(fused)
                    done = self._advance(upd, fused, kmax)
                    fire = self._first_fire(ck, done)
                    if fire is not None:
                        self._rollback(fused, ck)
                        self._advance(upd, fused, fire)

    def _checkpoint(self, fused):
        st = self.device_store()
        return {"t": self.t, "dt": self.dt, "nts": len(self.ts), "step_index": self.step_index,
                "store": st.checkpoint("photon"), "rows": st.current_row, "members": fused.checkpoint()}

    def _rollback(self, fused, ck):
        self.t, self.dt, self.step_index = ck["t"], ck["dt"], ck["step_index"]
        del self.ts[ck["nts"]:]
        self.store.restore("photon", ck["store"])
        fused.rollback(ck["members"])

    def _first_fire(self, ck, done):
        """Number of timesteps of the chunk after which the exit predicate first holds (None: not inside the chunk),
        evaluated with t, dt, ts and len(objects) as they were after each timestep."""
        from . import _capi

        st = self.store
        first = ck["rows"] + 1
        alive = np.array([int(st.read_row(first + j)[_capi.T_ALIVE]) for j in range(len(done))], np.int64)
        if self.shard:
            from .dist import all_reduce_rows

            alive = all_reduce_rows(alive)
        full_ts, end = self.ts, (self.t, self.dt)
        try:
            for j, (t_j, dt_j) in enumerate(done[:-1]):
                self.t, self.dt = t_j, dt_j
                self.ts = full_ts[:ck["nts"] + j + 1]
                self._len_override = int(alive[j])
                if self.exit(self):
                    return j + 1
        finally:
            self._len_override = None
            self.ts = full_ts
            self.t, self.dt = end
        return None

    def join(self, timeout=None):
        super().join(timeout)
        if self.error is not None:
            err, self.error = self.error, None
            raise err

    def run_steps(self, nsteps):
        """Run exactly ``nsteps`` timesteps on the calling thread (no exit predicate, no thread):
        the bulk entry point used by bench.py."""
        plan = self._plan()
        if not self.ts:
            self.t, self.dt = 0, 0
        nsteps = int(nsteps)
        # bulk form: [UpdateTimeStep, fused step] goes to the device in chunks of about feedback_every timesteps per
        # C-ABI call (a whole number of compaction periods)
        fast = self._bulk_plan(plan)
        if fast is not None:
            upd, fused = fast
            while nsteps > 0:
                nsteps -= len(self._advance(upd, fused, min(nsteps, fused.chunk_steps(self), 256)))
            return
        for _ in range(nsteps):
            for step in plan:
                if not step.uses_device and getattr(step, "touches_objects", True):
                    self._pull_objects()
                    if getattr(step, "modifies_objects", True):
                        self._host_dirty = self._host_dirty or self.store is not None
                step.run(self)
            self.step_index += 1

    # ---- introspection --------------------------------------------------------------------------
    @staticmethod
    def get_device_info():
        """physicl/__init__.py:470-499: {platform: {property..., device: {property...}}}."""
        from . import _capi
        import torch

        out = {"NAME": "physicl_b200 (CUDA, sm_100a)"}
        for d in range(torch.cuda.device_count()):
            ctx = _capi.Context(d)
            info = ctx.device_info()
            out[info["NAME"] + " #%d" % d] = info
            ctx.close()
        return {"physicl_b200": out}

    @staticmethod
    def set_dev(id):
        """physicl/__init__.py:526-529 (a stub there): choose the default device ordinal."""
        os.environ["PHYSICL_B200_DEVICE"] = str(id)

    def get_state(self):
        if self.state_need_lock:
            with self._state_lock:
                return self.state_fn(self)
        return self.state_fn(self)


from .clprogram import CLInput, CLOutput, CLProgram  # noqa: E402,F401  (physicl/__init__.py:543-664)
