"""Device-resident particle state: structure-of-arrays float32 planes in HBM.

The reference keeps particles as a Python list of ``Object`` instances whose fields are 3-vector
``Measurement`` arrays (physicl/__init__.py:381-396) and re-marshals them to the device on every
step (``CLProgram.run``, :602-664).  Here the state is uploaded once and stays on the GPU; PyTorch
is used only to own the buffers (``torch.empty(..., device="cuda")``) and every kernel launch goes
through the C ABI with raw pointers.

A store holds up to two homogeneous groups, because the reference's scatter steps act on
``PhotonObject`` only (light.py:283) while kinematics and the sign tally act on every object
(newton.py:14, light.py:423): ``photon`` and ``object``.

Retirement.  A retired photon (absorbed, escaped) is a slot whose x is NaN.  Pipelines that retire
photons run on a PING-PONG pair of plane sets: a launch advances the photons m timesteps in registers
and then writes the survivors densely into the partner set (``pcl_photon_steps_pp``; one timestep at
a time: ``pcl_photon_step_compact``).  The number of valid slots then lives on the device (``n_dev``) and
kernels read it there, so the stepping loop never waits for the host; ``n`` on the host is an upper
bound until ``sync_n`` is called.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

_PLANES_F32 = ("x", "y", "z", "vx", "vy", "vz", "dx", "dy", "dz", "ax", "ay", "az", "e")
_PLANES_U32 = ("id", "nscat")


def _torch():
    import torch

    return torch


class Group:
    """One SoA block.  ``n`` counts slots (live + retired), exactly or as an upper bound."""

    def __init__(self, kind, device, n, id_base=0):
        self.kind, self.device, self.n, self.id_base = kind, device, int(n), int(id_base)
        self.bufs = [{}, {}]  # ping-pong plane sets; bufs[cur] is current
        self.cur = 0
        self.id_valid = [False, False]  # does bufs[k]["id"] hold real ids (else id = slot index)
        self.n_exact = True  # False after device-side compaction: the exact count is n_dev[cur]
        self.n_dev = None  # int64[2] device tensor, allocated with the partner set
        self.e0 = 1.0  # energy scale: the e plane holds E / e0
        self.host_objs = None  # python objects by local id, when the group came from sim.objects
        self.n_live = int(n)  # last known live count

    @property
    def planes(self):
        return self.bufs[self.cur]

    @property
    def spare(self):
        return self.bufs[self.cur ^ 1]

    def alloc(self, name, fill=None):
        torch = _torch()
        dt = torch.int32 if name in _PLANES_U32 else torch.float32
        t = torch.empty(max(self.n, 1), dtype=dt, device=self.device)
        if fill is not None:
            t.fill_(fill)
        self.planes[name] = t
        return t

    def ensure(self, *names, fill=0):
        for nm in names:
            if nm not in self.planes:
                self.alloc(nm, fill)

    def upload(self, name, host):
        torch = _torch()
        if isinstance(host, torch.Tensor) and host.is_cuda:
            # adopt a device tensor as the plane (bulk set-ups that never touch the host)
            want = torch.int32 if name in _PLANES_U32 else torch.float32
            assert host.dtype == want and host.is_contiguous() and host.numel() == self.n, (name, host.dtype, host.shape)
            assert host.data_ptr() % 16 == 0, "device planes must be 16-byte aligned"
            self.planes[name] = host
            if name == "id":
                self.id_valid[self.cur] = True
            return
        if name in _PLANES_U32:
            arr = np.ascontiguousarray(host, np.uint32).view(np.int32)
        else:
            arr = np.ascontiguousarray(host, np.float32)
        assert arr.size == self.n, (name, arr.size, self.n)
        t = torch.from_numpy(arr.copy() if arr.size else np.zeros(1, arr.dtype))
        self.planes[name] = t.to(self.device, non_blocking=False)
        if name == "id":
            self.id_valid[self.cur] = True

    def download(self, name):
        assert self.n_exact, "call store.sync_n() first"
        a = self.planes[name][: self.n].cpu().numpy()
        return a.view(np.uint32) if name in _PLANES_U32 else a

    def _fill_soa(self, s, src, with_id, offset=0, count=None):
        s.n = (self.n - offset) if count is None else count
        for nm in _PLANES_F32 + _PLANES_U32:
            t = src.get(nm)
            if nm == "id" and not with_id:
                t = None
            setattr(s, nm, (t.data_ptr() + 4 * offset) if t is not None else None)
        s.id_base = self.id_base + (offset if not with_id else 0)
        s.n_dev = None
        return s

    def soa(self, planes=None, offset=0, count=None):
        """Fill a ``pcl_soa`` with raw device pointers (optionally a sub-range of slots).  When the
        slot count is only known on the device the view carries ``n_dev``."""
        if planes is None:
            s = self._fill_soa(_capi.Soa(), self.planes, self.id_valid[self.cur], offset, count)
            if not self.n_exact:
                s.n_dev = self.n_dev.data_ptr() + 8 * self.cur
            return s
        return self._fill_soa(_capi.Soa(), planes, "id" in planes, offset, count)

    def state_bytes_per_slot(self):
        return 4 * len(self.planes)


class DeviceParticleStore:
    """All particles of one ``Simulation`` on one GPU, plus the device tally table."""

    TALLY_ROWS = 4096

    def __init__(self, ctx: _capi.Context, compact_threshold=0.125):
        torch = _torch()
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        self.groups = {}
        self.compact_threshold = compact_threshold
        self.tally = torch.zeros((self.TALLY_ROWS, _capi.TALLY_COLS), dtype=torch.int64, device=self.device)
        self._row = -1  # row currently being accumulated
        self._rows_host = {}  # flushed rows: global row number -> np.int64[16]
        self._row_base = 0  # global row number of tally[0]
        self._live_scratch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.compactions = 0

    # ---- streams ----------------------------------------------------------------------------
    def stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        _torch().cuda.current_stream(self.device).synchronize()

    # ---- construction -----------------------------------------------------------------------
    def add_group(self, kind, r, v, E=None, a=None, id_base=0, host_objs=None, track_nscat=False):
        """r, v, a: (3, N) array-likes in code units; E: (N,) or None."""
        torch = _torch()
        on_device = isinstance(r, torch.Tensor) and r.is_cuda
        if not on_device:
            r = np.asarray(r).reshape(3, -1)
            v = np.asarray(v).reshape(3, -1)
        n = r.shape[1]
        if kind in self.groups:
            raise ValueError("group '%s' already present; build the store once from all particles" % kind)
        if n >= 2 ** 32:
            raise ValueError("a shard holds fewer than 2^32 slots")
        g = Group(kind, self.device, n, id_base)
        for i, nm in enumerate(("x", "y", "z")):
            g.upload(nm, r[i])
        for i, nm in enumerate(("vx", "vy", "vz")):
            g.upload(nm, v[i])
        if a is not None:
            if not on_device:
                a = np.asarray(a).reshape(3, -1)
            for i, nm in enumerate(("ax", "ay", "az")):
                g.upload(nm, a[i])
        if E is not None and isinstance(E, torch.Tensor) and E.is_cuda:
            g.e0 = 1.0  # the caller supplies e = E / E0 and sets g.e0 itself
            g.upload("e", E)
        elif E is not None:
            E = np.asarray(E, np.float64).reshape(-1)
            finite = E[np.isfinite(E)]
            g.e0 = float(np.max(np.abs(finite))) if finite.size and np.max(np.abs(finite)) > 0 else 1.0
            g.upload("e", E / g.e0)
        if track_nscat:
            g.alloc("nscat", 0)
        g.host_objs = host_objs
        self.groups[kind] = g
        return g

    def group(self, kind):
        return self.groups.get(kind)

    @property
    def n_slots(self):
        return sum(g.n for g in self.groups.values())

    # ---- tally rows -------------------------------------------------------------------------
    def new_row(self):
        """Start a fresh tally row (one per launch) and return its global number."""
        if self._row + 1 >= self.TALLY_ROWS:
            self.flush_rows()
        self._row += 1
        return self._row_base + self._row

    def new_rows(self, k):
        """Reserve k consecutive fresh rows (a multi-step launch fills one row per timestep)."""
        if k > self.TALLY_ROWS:
            raise ValueError("at most %d tally rows per call" % self.TALLY_ROWS)
        if self._row + k >= self.TALLY_ROWS:
            self.flush_rows()
        first = self._row_base + self._row + 1
        self._row += k
        return first

    def row_ptr(self, global_row=None):
        local = self._row if global_row is None else global_row - self._row_base
        assert 0 <= local < self.TALLY_ROWS
        return C.c_void_p(self.tally.data_ptr() + local * _capi.TALLY_COLS * 8)

    @property
    def current_row(self):
        return self._row_base + self._row

    def flush_rows(self):
        """Bring every accumulated row to the host (one D2H copy) and recycle the table."""
        if self._row >= 0:
            host = self.tally[: self._row + 1].cpu().numpy()
            for i in range(host.shape[0]):
                self._rows_host[self._row_base + i] = host[i].copy()
            self.tally[: self._row + 1].zero_()
            self._row_base += self._row + 1
            self._row = -1

    def read_row(self, global_row):
        if global_row not in self._rows_host:
            self.flush_rows()
        return self._rows_host[global_row]

    def peek_row(self, global_row):
        """Read one row without recycling the table (a small blocking D2H)."""
        if global_row in self._rows_host:
            return self._rows_host[global_row]
        return self.tally[global_row - self._row_base].cpu().numpy()

    # ---- slot counts ------------------------------------------------------------------------
    def sync_n(self, kind="photon"):
        """Make ``g.n`` exact again (one 8-byte D2H) after device-side compaction."""
        g = self.groups.get(kind)
        if g is not None and not g.n_exact:
            g.n = int(g.n_dev[g.cur].item())
            g.n_exact = True
        return None if g is None else g.n

    # ---- device-side copies (Simulation._run_chunked rolls a chunk of timesteps back to one) ------------
    def checkpoint(self, kind="photon"):
        g = self.groups.get(kind)
        if g is None:
            return None
        n = max(g.n, 1)
        return {"planes": {nm: t[:n].clone() for nm, t in g.planes.items()}, "cur": g.cur, "id_valid": list(g.id_valid),
                "n": g.n, "n_exact": g.n_exact, "n_live": g.n_live, "n_dev": None if g.n_dev is None else g.n_dev.clone()}

    def restore(self, kind, ck):
        if ck is None:
            return
        g = self.groups[kind]
        g.cur, g.id_valid, g.n, g.n_exact, g.n_live = ck["cur"], list(ck["id_valid"]), ck["n"], ck["n_exact"], ck["n_live"]
        for nm, t in ck["planes"].items():
            if nm in g.planes and g.planes[nm].numel() >= t.numel():
                g.planes[nm][: t.numel()].copy_(t)
            else:
                g.planes[nm] = t.clone()
        for nm in [nm for nm in g.planes if nm not in ck["planes"]]:
            del g.planes[nm]
        if ck["n_dev"] is not None:
            g.n_dev.copy_(ck["n_dev"])

    # ---- compaction -------------------------------------------------------------------------
    def reserve_spare(self, kind="photon"):
        """Allocate the ping-pong partner planes and the device-side slot counters now (pipelines that
        retire photons call this on their first timestep), so no cudaMalloc lands inside a stepped
        region later."""
        torch = _torch()
        g = self.groups.get(kind)
        if g is None:
            return
        for nm, t in g.planes.items():
            if nm not in g.spare or g.spare[nm].numel() < t.numel():
                g.spare[nm] = torch.empty_like(t)
        for buf in g.bufs:
            if "id" not in buf:
                buf["id"] = torch.empty(max(g.n, 1), dtype=torch.int32, device=self.device)
        if g.n_dev is None:
            g.n_dev = torch.zeros(2, dtype=torch.int64, device=self.device)

    def pingpong(self, kind="photon"):
        """``pcl_pingpong`` view of a group for the multi-step retire-and-compact loop."""
        g = self.groups[kind]
        self.reserve_spare(kind)
        if g.n_exact:
            g.n_dev[g.cur] = g.n
        pp = _capi.Pingpong()
        for k in (0, 1):
            g._fill_soa(pp.buf[k], g.bufs[k], True)
            pp.buf[k].id_base = g.id_base
        pp.n_dev = g.n_dev.data_ptr()
        pp.cur = g.cur
        pp.id_valid = (1 if g.id_valid[0] else 0) | (2 if g.id_valid[1] else 0)
        return pp

    def adopt_pingpong(self, kind, pp, compacted):
        """Take the buffer state back from C after ``pcl_photon_steps_pp``."""
        g = self.groups[kind]
        g.cur = int(pp.cur)
        g.id_valid = [bool(pp.id_valid & 1), bool(pp.id_valid & 2)]
        if compacted:
            g.n_exact = False
            self.compactions += compacted

    def compact(self, kind="photon"):
        """Stable compaction of the live slots (``pcl_compact``).  Returns the live count."""
        g = self.groups[kind]
        self.sync_n(kind)
        if g.n == 0:
            return 0
        self.reserve_spare(kind)
        src = g.soa()
        dst = g.soa(planes=g.spare)
        self.ctx.call("pcl_compact", self.stream(), C.byref(src), C.byref(dst), C.c_void_p(self._live_scratch.data_ptr()))
        n_live = int(self._live_scratch.item())
        g.cur ^= 1
        g.id_valid[g.cur] = True
        g.n = n_live
        g.n_live = n_live
        self.compactions += 1
        return n_live

    def maybe_compact(self, kind, n_live):
        g = self.groups.get(kind)
        if g is None:
            return False
        g.n_live = int(n_live)
        self.sync_n(kind)
        if g.n and (g.n - n_live) > self.compact_threshold * g.n:
            self.compact(kind)
            return True
        return False

    # ---- host views -------------------------------------------------------------------------
    def snapshot(self, kind="photon", live_only=True):
        """Download one group as float32/uint32 numpy planes (plus ``id``), ordered by id."""
        g = self.groups[kind]
        self.sync_n(kind)
        out = {nm: g.download(nm) for nm in g.planes if nm != "id"}
        out["id"] = g.download("id") if g.id_valid[g.cur] else np.arange(g.n, dtype=np.uint32)
        if live_only:
            keep = ~np.isnan(out["x"])
            out = {k: v[keep] for k, v in out.items()}
        order = np.argsort(out["id"], kind="stable")
        out = {k: v[order] for k, v in out.items()}
        if "e" in out:
            out["E"] = out["e"].astype(np.float64) * g.e0
        return out
