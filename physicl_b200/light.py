"""Photon objects, emission, scattering steps and their measure steps.

Mirror of the reference's ``physicl/light.py`` (same names, same arguments).  The three OpenCL
kernels there (light.py:146-158, :239-249, :303-315) and the Python loops around them become
sm_100a kernels reached through the C ABI; see ``csrc/photon.cu``.
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np
import numpy.linalg as np_lin

import physicl_b200 as physicl

from . import _capi

# SI constants, scaled to code units at import exactly like the reference (light.py:14-16):
# call Measurement.set_code_scale(...) BEFORE importing this module.
c = physicl.Measurement(np.double(299792458), "m**1 s**-1")
h = physicl.Measurement(np.double(6.62607015e-34), "J**1 s**1")
kB = physicl.Measurement(np.double(1.380649e-23), "J**1 K**-1")


class PhotonObject(physicl.Object):
    """light.py:18-35: needs ``E`` and a velocity whose norm equals ``c`` exactly."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if np_lin.norm(self.v) != np_lin.norm(c):
            raise Exception("Not a valid speed.")
        if "E" not in kwargs:
            raise Exception("Needs a valid energy.")


def E_from_wavelength(wavelength):
    """light.py:39-43"""
    return (h * c) / wavelength


def wavelength_from_E(E):
    """light.py:45-49"""
    return (h * c) / E


# ---- emission (light.py:53-128) ------------------------------------------------------------------
def _plain(x):
    return x.__unscaled__() if isinstance(x, physicl.Measurement) else x


def planck_distribution(E, T):
    """light.py:53-60: 15/(pi^4 kT) (E/kT)^3 e^(-E/kT), as a ``J**-1`` Measurement."""
    E_conv, T_conv, kB_conv = _plain(E), _plain(T), kB.__unscaled__()
    x = E_conv / (kB_conv * T_conv)
    return physicl.Measurement(15 / (np.pi ** 4 * kB_conv * T_conv) * x ** 3 / np.e ** x, "J**-1")


def _planck_primitive(E, T):
    x = E / (float(kB.__unscaled__()) * T)
    return -np.exp(-x) * (x ** 3 + 3 * x ** 2 + 6 * x + 6) * (15.0 / np.pi ** 4)


def planck_probability(E_min, E_max, T, integrator=None):
    """light.py:63-64: integral of the density over [E_min, E_max] -> ``(value, abserr)``.
    Default: the closed form (agrees with the reference's scipy.integrate.quad to 1e-13)."""
    if integrator is not None:
        return integrator(lambda x: planck_distribution(x, T), E_min, E_max)
    E_min, E_max, T = float(_plain(E_min)), float(_plain(E_max)), float(_plain(T))
    return (float(_planck_primitive(E_max, T) - _planck_primitive(E_min, T)), 0.0)


def planck_table(E_min, E_max, T, bins):
    """The reference's binned law (light.py:82-93): grid ``linspace(E_min, E_max, bins)`` and the
    cumulative, normalised masses of the ``bins-1`` intervals (float64)."""
    E_min, E_max, T, bins = float(_plain(E_min)), float(_plain(E_max)), float(_plain(T)), int(_plain(bins))
    E = np.linspace(E_min, E_max, bins)
    F = _planck_primitive(E, T)
    gamma = F[1:] - F[:-1]
    tot = 0.0
    for gm in gamma:
        tot += gm
    norm = gamma / tot
    cdf = np.empty_like(norm)
    acc = 0.0
    for i, gm in enumerate(norm):
        acc = gm if i == 0 else acc + gm
        cdf[i] = acc
    return E, norm, cdf


last_planck_params = None
last_planck_gamma_norm = None
last_planck_cdf = None


def planck_phot_distribution(E_min, E_max, T, bins=1000):
    """light.py:73-104, host form: one ``np.random.rand()`` per call, returns the GRID energy ``E[x]``
    of the first ``x >= 1`` with ``cdf[x-1] <= rand <= cdf[x]``, or ``None`` when there is none."""
    global last_planck_params, last_planck_gamma_norm, last_planck_cdf
    params = [_plain(x) for x in (E_min, E_max, T, bins)]
    E = np.linspace(params[0], params[1], params[3])
    if last_planck_params != params:
        _, norm, cdf = planck_table(*params)
        last_planck_params, last_planck_gamma_norm, last_planck_cdf = params, list(norm), list(cdf)
    cdf = np.asarray(last_planck_cdf)
    rand = np.random.rand()
    x = int(np.searchsorted(cdf, rand, side="left"))
    if x >= cdf.size:
        return None
    if x == 0:
        if cdf.size > 1 and rand == cdf[0]:
            x = 1
        else:
            return None
    return physicl.Measurement(E[x], "J**1")


def planck_sample_device(ctx, n, E_min, E_max, T, bins=1000, seed=0, id_base=0, device=None, want_bins=False, timing=None):
    """Device form of the same sampler for bulk emission: one Philox4x32 block per four photons (stream 1),
    guide-table look-up of the float64 table (held as integers in shared memory when it fits).  Returns ``(e, E0[, bin])``: ``e`` is a float32 CUDA tensor of
    ``E / E0`` with ``E0 = E_max``; ``bin`` (int32) is the grid index, -1 where the reference yields None.
    ``timing``: a dict that receives ``device_ms``, the CUDA-event time of the sampling launches."""
    import torch

    E, _, cdf = planck_table(E_min, E_max, T, bins)
    E0 = float(E[-1])
    dev = torch.device("cuda", ctx.device) if device is None else device
    cdf_d = torch.from_numpy(cdf).to(dev)
    e = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    b = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_bins else None
    step = (E[-1] - E[0]) / (len(E) - 1)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timing is not None else None
    if ev:
        ev[0].record(torch.cuda.current_stream(dev))
    ctx.call("pcl_planck_sample", stream, C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(seed), C.c_void_p(cdf_d.data_ptr()),
             C.c_uint32(cdf.size), C.c_float(E[0] / E0), C.c_float(step / E0), C.c_void_p(e.data_ptr()),
             C.c_void_p(b.data_ptr()) if b is not None else None)
    if ev:
        ev[1].record(torch.cuda.current_stream(dev))
    torch.cuda.current_stream(dev).synchronize()  # cdf_d may be freed after return
    if ev:
        timing["device_ms"] = ev[0].elapsed_time(ev[1])  # the sampling kernels alone (the table is built on the host)
    return (e[:n], E0, b[:n]) if want_bins else (e[:n], E0)


def generate_photons_from_E(E):
    """light.py:109-110"""
    return [PhotonObject(E=x, v=c * [1, 0, 0]) for x in E]


def generate_photons(n, fn=lambda: np.random.power(3), min=0, max=0, bins=-1):
    """light.py:112-128: ``E = min + (max - min) * fn()``, photons start at the origin along +x."""
    out = []
    for _ in range(int(n)):
        Eo = min + (max - min) * fn()
        out.append(PhotonObject(E=Eo, v=physicl.Measurement([c, 0, 0], "m**1 s**-1")))
    return out


# ---- scattering steps ------------------------------------------------------------------------------
class _ScatterBase(physicl.Step):
    uses_device = True
    mode = 0

    def _init_rng(self, rng, seed):
        if rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' (in-kernel) or 'numpy' (the reference's host draws)")
        self.rng = rng
        self.seed = seed
        self._salt = 0  # set from the step's position in the simulation (Simulation._plan)

    def _seed(self, sim):
        return (int(self.seed) if self.seed is not None else (int(sim.seed) ^ self._salt)) & 0xFFFFFFFFFFFFFFFF

    def scatter_params(self, group):
        k = float(self.A) * float(self.n)
        mode = self.mode
        if getattr(self, "wavelength_dep_scattering", False):
            # pcoll = A n |dr| (h c / E)^-4 (light.py:300-301) = [A n (E0 / (h c))^4] |dr| (E/E0)^4
            k = k * (group.e0 / (float(h) * float(c))) ** 4
            mode |= _capi.SCATTER_WAVELENGTH
        if getattr(self, "sfu_trig", False) and not (mode & _capi.SCATTER_DELETE):
            mode |= _capi.SCATTER_SFU
        return _capi.ScatterParams(k=k, c=float(c), mode=mode)

    def varn_params(self, group):
        """float64 constants of the variable-density kernel (light.py:287, :299-301)."""
        a_slot, n_slot = float(self.n), float(self.A)  # the reference binds name "A" to n and "n" to A
        kd = a_slot * (n_slot if getattr(self, "variable_n_apply_A", False) else 1.0)
        if getattr(self, "wavelength_dep_scattering", False):
            kd = kd * (group.e0 / (float(h) * float(c))) ** 4
        return _capi.VarnParams(kd=kd, e0=float(group.e0), a_slot=a_slot, n_slot=n_slot)

    def jit_kernel(self, ctx, name):
        from . import jit

        m = self._jit.get(id(ctx))
        if m is None or m.ctx is not ctx:
            m = self._jit[id(ctx)] = jit.Module(ctx, self._jit_source)
        return m.kernel(name)

    def rng_params(self, sim, store, group):
        """Philox: nothing to upload.  'numpy': draw on the host in the reference's order
        (light.py:285: rtheta, rphi, rand per photon; light.py:235: rand only) and inject."""
        r = _capi.Rng(seed=self._seed(sim), step=sim.step_index & 0xFFFFFFFF)
        keep = None
        if self.rng == "numpy":
            import torch

            if group.n_live != group.n:
                store.compact("photon")  # one draw per LIVE photon, in list order (light.py:235)
            n = group.n
            if self.mode & _capi.SCATTER_DELETE:
                u = np.random.random(n).astype(np.float32)
                keep = (torch.from_numpy(u).to(store.device),)
                r.u_rand = keep[0].data_ptr()
            else:
                u = np.random.random((n, 3)).astype(np.float32)
                keep = tuple(torch.from_numpy(np.ascontiguousarray(u[:, i])).to(store.device) for i in range(3))
                r.u_theta, r.u_phi, r.u_rand = (t.data_ptr() for t in keep)
        return r, keep

    def run(self, sim):
        """Unfused form: reads the dr planes the kinematics step wrote (as the reference kernel does)."""
        st = sim.device_store()
        g = st.group("photon")
        if g is None or g.n == 0:
            return
        if "dx" not in g.planes:
            raise RuntimeError("%s needs the displacement of this timestep: add NewtonianKinematicsStep before it "
                               "(reference pipelines do: test/test_light.py:33-34)" % type(self).__name__)
        st.sync_n("photon")  # stand-alone kernels take the exact slot count from the host
        sp = self.scatter_params(g)
        rng, keep = self.rng_params(sim, st, g)
        row = st.new_row()
        soa = g.soa()
        if getattr(self, "variable_n", False):
            vn = self.varn_params(g)
            sim.cl_ctx.call("pcl_scatter_jit", st.stream(), self.jit_kernel(sim.cl_ctx, "pcl_jit_scatter"), C.byref(soa),
                            C.byref(sp), C.byref(vn), C.byref(rng), None, st.row_ptr())
        else:
            sim.cl_ctx.call("pcl_scatter", st.stream(), C.byref(soa), C.byref(sp), C.byref(rng), None, st.row_ptr())
        if keep is not None:
            st.synchronize()
        sim._mark_device_dirty(live_row=row)
        if self.mode & _capi.SCATTER_DELETE and self.rng == "numpy":
            g.n_live = int(st.peek_row(row)[_capi.T_ALIVE])


class ScatterIsotropicStep(_ScatterBase):
    """light.py:262-359.  ``n`` number density, ``A`` cross-section; ``wavelength_dep_scattering``
    multiplies the collision probability by ``(h c / E)^-4``.  New keyword arguments: ``rng``
    ('philox' in-kernel draws, or 'numpy' for the reference's host-side ``np.random`` stream) and ``seed``.

    ``variable_n=True, variable_n_fn="<expression>"`` (light.py:295-299): the number density is the
    user's OpenCL-C expression over ``r0[gid]``, ``r1[gid]``, ``r2[gid]`` (also ``E[gid]``, ``norm``),
    evaluated in float64; the kernel is compiled at run time for sm_100a (``physicl_b200/jit.py``).  As in
    the reference, that kernel computes ``A * (expression) * norm`` with the NAME ``A`` bound to this
    step's ``n`` (light.py:287 swaps the two constants), so the step's ``A`` does not enter unless
    ``variable_n_apply_A=True`` is passed.

    ``sfu_trig=True`` (new, opt-in): the sines and cosines of a new direction come from the GPU's special-function unit
    (MUFU.SIN / MUFU.COS, |error| ~ 5e-7) instead of the table + addition theorem the CPU twin reproduces bit for bit: 7 %
    fewer instructions in the fused photon loop; results then agree with the default in law (same decisions in the first
    timestep, same distributions), not bit for bit.  Fused pipelines with in-kernel draws only.

    Note the direction law is the reference's: theta ~ U[0, 2 pi) polar, phi ~ U[0, pi) azimuth
    (light.py:285, :309-311), which is not uniform on the sphere."""

    def __init__(self, **kwargs):
        self.n = kwargs.get("n", 1)
        self.A = kwargs.get("A", 1)
        self.wavelength_dep_scattering = kwargs.get("wavelength_dep_scattering", False)
        self.variable_n = kwargs.get("variable_n", False)
        self.variable_n_fn = kwargs.get("variable_n_fn", None)
        self.variable_n_apply_A = kwargs.get("variable_n_apply_A", False)
        self.sfu_trig = bool(kwargs.get("sfu_trig", False))
        self._jit = {}  # ctx id -> jit.Module
        if self.variable_n:
            from . import jit

            # built when the step is created, so that a bad expression fails here and not mid-run
            self._jit_source = jit.photon_source(self.variable_n_fn, self.wavelength_dep_scattering)
            if kwargs.get("check_expression", True):
                jit.check(self._jit_source)
        self._init_rng(kwargs.get("rng", "philox"), kwargs.get("seed", None))


class ScatterDeleteStep(_ScatterBase):
    """light.py:225-260: photons whose collision test succeeds are removed from the simulation."""

    mode = _capi.SCATTER_DELETE

    def __init__(self, n, A, rng="philox", seed=None):
        self.n, self.A = n, A
        self._init_rng(rng, seed)


class ScatterDeleteStepReference(ScatterDeleteStep):
    """light.py:131-223: the hand-written twin of ``ScatterDeleteStep``; same law, same kernel here."""


class EscapeSphereStep(physicl.Step):
    """Photons with ``|r| >= R`` leave the simulation (NOT in the reference, whose ``bounds`` is never
    read).  ``escaped`` collects the per-timestep escape counts: the
    escape-time histogram."""

    uses_device = True

    def __init__(self, R):
        self.R = float(R)
        self._rows = []

    def run(self, sim):
        st = sim.device_store()
        g = st.group("photon")
        if g is None or g.n == 0:
            return
        st.sync_n("photon")
        row = st.new_row()
        soa = g.soa()
        sim.cl_ctx.call("pcl_escape", st.stream(), C.byref(soa), C.c_float(self.R * self.R), st.row_ptr())
        self._rows.append((sim.store, row))
        sim._mark_device_dirty(live_row=row)

    def _note_row(self, sim, row):
        self._rows.append((sim.store, row))

    @property
    def escaped(self):
        return np.array([int(st.read_row(r)[_capi.T_ESCAPED]) for st, r in self._rows], np.int64)

    def escaped_all_ranks(self):
        from .dist import all_reduce_rows

        return all_reduce_rows(self.escaped)


# ---- measure steps -------------------------------------------------------------------------------
class _DeviceMeasureStep(physicl.MeasureStep):
    """Rows are tallied on the device and fetched lazily: reading ``data`` costs one D2H copy for all
    pending rows instead of one blocking read per timestep."""

    uses_device = True

    def __init__(self, out_fn):
        super().__init__(out_fn)
        self._data = []
        self._pending = []  # (t, store, global row)

    @property
    def data(self):
        if self._pending:
            pend, self._pending = self._pending, []
            rows = np.stack([st.read_row(row) for _, st, row, _ in pend])
            if pend[0][3]:  # sharded: every rank holds partial counts; sum them once for all rows
                from .dist import all_reduce_rows

                rows = all_reduce_rows(rows)
            for (t, _, grow, _), r in zip(pend, rows):
                try:
                    self._data.append(self._format(t, r, grow))
                except TypeError:
                    self._data.append(self._format(t, r))
        return self._data

    @data.setter
    def data(self, v):
        self._data = v
        self._pending = []

    def _note_row(self, sim, row, t=None):
        if hasattr(row, "plane_slice") and hasattr(self, "_plane0"):
            self._plane0 = row.plane_slice[0]
        self._pending.append((sim.t if t is None else t, sim.store, int(row), bool(sim.shard)))

    def _planes(self):
        return []

    def run(self, sim):
        st = sim.device_store()
        row = st.new_row()
        pl = _capi.make_planes(self._planes())
        for kind, g in st.groups.items():
            st.sync_n(kind)
            if g.n == 0:
                continue
            if pl.count:
                g.ensure("dx", "dy", "dz")
            soa = g.soa()
            sim.cl_ctx.call("pcl_tally", st.stream(), C.byref(soa), C.byref(pl), st.row_ptr())
        self._note_row(sim, row)


class ScatterMeasureStep(_DeviceMeasureStep):
    """light.py:361-404: row ``[t, N, crossings of each plane...]``; a plane is a 3-vector with NaN in
    the two free coordinates; a crossing is ``r - dr <= loc <= r`` or ``r - dr >= loc >= r``.

    ``measure_E=True`` (light.py:380-402) adds, after each plane's count, the list of the energies of
    the photons that crossed it, in object order: ``[t, N, n_0, [E...], n_1, [E...], ...]`` (an object
    array, as ragged rows have to be on current NumPy).  That form reads the ``dr`` planes, so a
    pipeline containing it runs unfused."""

    def __init__(self, out_fn, measure_n=True, measure_locs=[], measure_E=False):
        super().__init__(out_fn)
        self.measure_locs = measure_locs
        self.measure_n = measure_n
        self.measure_E = measure_E
        self.needs_dr = bool(measure_E)
        self._plane0 = 0  # first tally column of this step's planes (non-zero only inside a fused row)
        self._elists = {}  # global row -> [list of E per plane]

    def _planes(self):
        out = []
        for loc in self.measure_locs:
            loc = np.asarray(loc, np.float64)
            ax = 0 if not np.isnan(loc[0]) else (1 if not np.isnan(loc[1]) else 2)  # light.py:385-395
            out.append((ax, float(loc[ax])))
        return out

    def run(self, sim):
        super().run(sim)
        if not self.measure_E or not self.measure_locs:
            return
        import torch

        st = sim.device_store()
        row = self._pending[-1][2]
        pl = _capi.make_planes(self._planes())
        lists = [[] for _ in self.measure_locs]
        for kind, g in st.groups.items():
            if g.n == 0 or kind != "photon":
                continue  # only photons carry E (light.py:34); the reference would raise on other objects
            st.sync_n(kind)
            g.ensure("dx", "dy", "dz")
            cap = g.n
            ids = torch.empty(pl.count * cap, dtype=torch.int32, device=st.device)
            es = torch.empty(pl.count * cap, dtype=torch.float32, device=st.device)
            cnt = torch.zeros(pl.count, dtype=torch.int64, device=st.device)
            soa = g.soa()
            sim.cl_ctx.call("pcl_plane_crossers", st.stream(), C.byref(soa), C.byref(pl), C.c_void_p(ids.data_ptr()),
                            C.c_void_p(es.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_uint64(cap))
            counts = cnt.cpu().numpy()
            for q in range(pl.count):
                k = int(counts[q])
                i = ids[q * cap:q * cap + k].cpu().numpy().view(np.uint32)
                e = es[q * cap:q * cap + k].cpu().numpy().astype(np.float64) * g.e0
                lists[q] = list(e[np.argsort(i, kind="stable")])  # object order = id order
        self._elists[row] = lists

    def _format(self, t, r, row=None):
        out = [t]
        if self.measure_n:
            out.append(int(r[_capi.T_ALIVE]))
        for k in range(len(self.measure_locs)):
            out.append(int(r[_capi.T_PLANE0 + self._plane0 + k]))
            if self.measure_E:
                out.append(self._elists.get(row, [[]] * len(self.measure_locs))[k])
        if self.measure_E:
            arr = np.empty(len(out), dtype=object)
            arr[:] = out
            return arr
        return np.array(out)


class ScatterSignMeasureStep(_DeviceMeasureStep):
    """light.py:406-431: row ``[t, N, #(v_x > 0), #(v_y > 0), #(v_z > 0)]`` over all objects."""

    def __init__(self, out_fn, measure_n=True):
        super().__init__(out_fn)
        self.measure_n = measure_n

    def _format(self, t, r):
        out = [t]
        if self.measure_n:
            out.append(int(r[_capi.T_ALIVE]))
        out.extend(int(r[q]) for q in (_capi.T_XP, _capi.T_YP, _capi.T_ZP))
        return np.array(out)


class TracePathMeasureStep(physicl.MeasureStep):
    """light.py:433-483: the position of every object at every timestep, plus (``trace_dv=True``) how
    often its velocity changed, i.e. how often it scattered (light.py:459-460).

    Device form: each timestep one kernel scatters r by particle id into that step's slab of a
    trajectory buffer (NaN where the object does not exist any more); ``terminate`` downloads the
    slabs and lays ``data`` out exactly like the reference: first row ``["t", t_0, t_1, ...]``, then per
    object ``[id_info, (freq,) nan, nan, nan, ... , r_k, r_k+1, ..., nan, nan, nan, ...]``: three NaNs for every
    timestep before the object's first appearance (light.py:477-479; objects may be added while the simulation runs),
    its positions, and ``[nan, nan, nan] * (columns - positions)`` after them (the reference's own arithmetic, light.py:478,
    :481).  The scatter count comes from the per-photon ``nscat`` plane that every scatter kernel maintains.

    Identity: like the reference (light.py:450-456) every object gets a trace id the first time it is seen; it is kept
    in the object's ``__trace_path_id`` attribute, so it survives rebuilds of the device store (a host step, add_obj in
    the middle of a run).  Bulk particles (``add_particles``) are numbered by their particle id."""

    uses_device = True

    def __init__(self, out_fn, trace_type=physicl.Object, id_info_fn=lambda x: str(type(x)), trace_dv=False):
        super().__init__(out_fn)
        self.trace_type = trace_type
        self.id_info_fn = id_info_fn
        self.trace_dv = trace_dv
        self.id_counter = 0
        self.id_dict = {}  # trace id -> id_info string
        self._start = {}  # trace id -> column of the first appearance
        self._freq_done = {}  # trace id -> scatter count, for store epochs that are over
        self._epochs = {}  # (store id, kind) -> dict(tid=int64[n_ids], n_ids, freq=device tensor or None)
        self._bulk_base = {}  # (store id, kind) -> trace id of local id 0, for groups without Python objects
        self._slabs = []  # (epoch key, device tensor [3*n_ids], column)

    def prepare(self, sim):
        """Called once before the first timestep: the scatter counters must exist before any scatter."""
        if self.trace_dv:
            st = sim.device_store()
            for g in st.groups.values():
                g.ensure("nscat", fill=0)

    def _epoch(self, sim, st, kind, g, col):
        """Trace ids of one group of one store (a store lives until the host side changes the object list)."""
        key = (id(st), kind)
        ep = self._epochs.get(key)
        if ep is not None and ep["store"] is st:
            return key, ep
        self._close_epochs(keep_store=st)
        n_ids = g.n
        g.n0 = n_ids  # ids are local indices of the group as ingested
        tid = np.empty(n_ids, np.int64)
        objs = g.host_objs if isinstance(g.host_objs, list) else None
        cls = PhotonObject if kind == "photon" else physicl.Object
        if objs is not None:
            for i, o in enumerate(objs):
                t = getattr(o, "__trace_path_id", None)
                origin = getattr(o, "_pcl_origin", None)
                if t is None and origin is not None and origin[:2] in self._bulk_base:
                    t = self._bulk_base[origin[:2]] + origin[2]  # a bulk particle that became an object on a pull
                if t is None:
                    t = self.id_counter
                    self.id_counter += 1
                    self.id_dict[t] = self.id_info_fn(o)
                    self._start[t] = col
                setattr(o, "__trace_path_id", t)
                tid[i] = t
        else:
            base = self.id_counter
            self._bulk_base[key] = base
            self.id_counter += n_ids
            info = str(cls)
            for i in range(n_ids):
                self.id_dict[base + i] = info
                self._start[base + i] = col
            tid[:] = base + np.arange(n_ids)
        ep = {"store": st, "tid": tid, "n_ids": n_ids, "freq": None}
        if self.trace_dv:
            import torch

            ep["freq"] = torch.zeros(max(n_ids, 1), dtype=torch.int32, device=st.device)
        self._epochs[key] = ep
        return key, ep

    def _close_epochs(self, keep_store=None):
        """Fold the scatter counts of finished store epochs into the per-trace-id table."""
        for key, ep in list(self._epochs.items()):
            if ep["store"] is keep_store or ep.get("closed"):
                continue
            if ep["freq"] is not None:
                f = ep["freq"].cpu().numpy()[: ep["n_ids"]]
                for i, t in enumerate(ep["tid"]):
                    self._freq_done[int(t)] = max(self._freq_done.get(int(t), 0), int(f[i]))
                ep["freq"] = None
            ep["closed"] = True

    def run(self, sim):
        import torch

        st = sim.device_store()
        col = len(sim.ts) - 1
        for kind, g in st.groups.items():
            st.sync_n(kind)
            key, ep = self._epoch(sim, st, kind, g, col)
            n_ids = ep["n_ids"]
            slab = torch.full((3 * max(n_ids, 1),), float("nan"), dtype=torch.float32, device=st.device)
            freq = C.c_void_p(ep["freq"].data_ptr()) if ep["freq"] is not None else None
            soa = g.soa()
            sim.cl_ctx.call("pcl_trace_positions", st.stream(), C.byref(soa), C.c_void_p(slab.data_ptr()), C.c_uint64(n_ids), freq)
            self._slabs.append((key, slab, col))

    def terminate(self, sim):
        ts = list(sim.ts)
        cols = len(ts)
        dat = [["t"] + copy.deepcopy(ts)]
        self._close_epochs()
        pos = {t: [] for t in self.id_dict}  # trace id -> positions in time order
        for key, slab, col in sorted(self._slabs, key=lambda e: e[2]):
            ep = self._epochs[key]
            sl = slab.cpu().numpy().reshape(3, -1)
            live = np.nonzero(~np.isnan(sl[0, : ep["n_ids"]]))[0]
            for i in live:
                pos[int(ep["tid"][i])].append(np.array(sl[:, i], np.float64))
        for t in range(self.id_counter):
            row = [self.id_dict[t]]
            if self.trace_dv:
                row.append(int(self._freq_done.get(t, 0)))
            p = pos.get(t, [])
            row.extend([np.nan, np.nan, np.nan] * self._start.get(t, 0))  # light.py:479
            row.extend(p)
            row.extend([np.nan, np.nan, np.nan] * (cols - len(p)))  # light.py:478, :481
            dat.append(row)
        self.data = dat
        super().terminate(sim)
