"""Multi-GPU plumbing: one process per GPU under ``torch.distributed`` (NCCL over NVLink on the
box, gloo in CPU tests).  The reference has no distributed code at all;
the partitioning below is what the north star adds.

Independent-particle steps (kinematics, every scatter flavour, emission, tallies) shard by a
contiguous block of global particle indices with NO data-path collective: the Philox counter is the
GLOBAL particle id, so results do not depend on the number of ranks.  Integer tallies are summed
across ranks only when a measure step's ``data`` is read.  Gravity needs every body's position each
step: one all-gather of the packed (x, y, z, m) array, overlapped with the local-block tile loop.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def world():
    d = _dist()
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank``: the first ``n % P`` ranks get
    one extra item, and every block start is a multiple of 4 when ``n / P`` allows it (keeps the
    128-bit vector path aligned)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def all_reduce_rows(rows):
    """Sum an int64 numpy array over all ranks (tally rows, live counts)."""
    import torch

    d = _dist()
    if not (d.is_available() and d.is_initialized()) or d.get_world_size() == 1:
        return rows
    rows = np.ascontiguousarray(rows, np.int64)
    t = torch.from_numpy(rows.copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.cpu().numpy()


def all_reduce_int(v):
    return int(all_reduce_rows(np.array([v], np.int64))[0])


class GravityExchange:
    """Per-step all-gather of the packed positions for sharded all-pairs gravity.

    The gather runs on a side stream (NCCL over NVLink / NVSwitch) while the compute stream is
    already accumulating the local block's contribution; the remote blocks follow once the gather
    has landed.  Accelerations are summed in block order local-first, so the result for a body does
    not depend on which rank computed it beyond float summation order (tolerance, not bit-exact)."""

    def __init__(self, posm, device):
        import torch

        self.rank, self.world = world()
        self.n_local = posm.shape[0]
        counts = np.zeros(self.world, np.int64)
        counts[self.rank] = self.n_local
        counts = all_reduce_rows(counts)
        if len(set(int(c) for c in counts)) != 1:
            raise ValueError("sharded gravity needs equal block sizes per rank (got %s)" % list(counts))
        self.all = torch.empty((self.world * self.n_local, 4), dtype=torch.float32, device=device)
        self.side = torch.cuda.Stream(device=device)
        self.device = device

    def kick_drift(self, ctx, store, n, posm, g, acc, dt):
        p = lambda t: C.c_void_p(t.data_ptr())
        ctx.call("pcl_gravity_kick_drift", store.stream(), C.c_uint64(n), p(posm), p(g.planes["vx"]), p(g.planes["vy"]),
                 p(g.planes["vz"]), p(acc[0]), p(acc[1]), p(acc[2]), C.c_float(dt), p(g.planes["x"]), p(g.planes["y"]),
                 p(g.planes["z"]))

    def accelerations(self, ctx, store, posm, n, args, fn="pcl_gravity_accel"):
        import torch

        d = _dist()
        p = lambda t: C.c_void_p(t.data_ptr())
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            d.all_gather_into_tensor(self.all, posm)
        # local block first (overlaps the gather) ...
        ctx.call(fn, store.stream(), p(posm), C.c_uint64(n), p(posm), C.c_uint64(n), *args, 0,
                 C.c_uint64(0), C.c_uint64(0))
        cur.wait_stream(self.side)
        # ... then every other rank's block in ONE launch over the gathered array, own block skipped
        nl = self.n_local
        ctx.call(fn, store.stream(), p(posm), C.c_uint64(n), p(self.all), C.c_uint64(self.world * nl), *args, 1,
                 C.c_uint64(self.rank * nl), C.c_uint64((self.rank + 1) * nl))



class GravityExchangeP2P:
    """The same exchange without a collective call: every rank's gathered array lives in peer-mapped (symmetric)
    memory, and the kick-drift kernel stores each updated body straight into slot ``rank * n_local + i`` of ALL ranks'
    arrays (``pcl_gravity_kick_drift_p2p``: 16-byte NVLink stores issued by the integration kernel itself).  The arrays
    are double-buffered by timestep parity; one stream-ordered barrier on the symmetric-memory signal pads per timestep
    separates the stores from the next acceleration pass, which then reads all j-bodies from local HBM in ONE launch.
    PyTorch's symmetric-memory allocator is used for what it is: buffer ownership and the handle exchange."""

    def __init__(self, posm, device):
        import torch
        import torch.distributed._symmetric_memory as symm_mem

        d = _dist()
        self.rank, self.world = world()
        self.n_local = posm.shape[0]
        counts = np.zeros(self.world, np.int64)
        counts[self.rank] = self.n_local
        counts = all_reduce_rows(counts)
        if len(set(int(c) for c in counts)) != 1:
            raise ValueError("sharded gravity needs equal block sizes per rank (got %s)" % list(counts))
        self.device = device
        group = d.group.WORLD
        self.bufs, self.handles, self.peer_ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty((self.world * self.n_local, 4), dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, group.group_name)
            self.bufs.append(t)
            self.handles.append(h)
            self.peer_ptrs.append((C.c_uint64 * self.world)(*[int(q) for q in h.buffer_ptrs]))
        self.cur = 0
        # the starting positions: one ordinary all-gather into the first array
        d.all_gather_into_tensor(self.bufs[0], posm)
        torch.cuda.current_stream(device).synchronize()
        d.barrier()

    def accelerations(self, ctx, store, posm, n, args, fn="pcl_gravity_accel"):
        p = lambda t: C.c_void_p(t.data_ptr())
        total = self.world * self.n_local
        ctx.call(fn, store.stream(), p(posm), C.c_uint64(n), p(self.bufs[self.cur]), C.c_uint64(total), *args, 0,
                 C.c_uint64(0), C.c_uint64(0))

    def kick_drift(self, ctx, store, n, posm, g, acc, dt):
        import torch

        p = lambda t: C.c_void_p(t.data_ptr())
        nxt = self.cur ^ 1
        ctx.call("pcl_gravity_kick_drift_p2p", store.stream(), C.c_uint64(n), p(posm), p(g.planes["vx"]), p(g.planes["vy"]),
                 p(g.planes["vz"]), p(acc[0]), p(acc[1]), p(acc[2]), C.c_float(dt), p(g.planes["x"]), p(g.planes["y"]),
                 p(g.planes["z"]), self.peer_ptrs[nxt], C.c_uint32(self.world), C.c_uint64(self.rank * self.n_local))
        with torch.cuda.stream(torch.cuda.current_stream(self.device)):
            self.handles[nxt].barrier(channel=0)  # every rank's stores into array `nxt` are done and visible
        self.cur = nxt


def gravity_exchange(posm, device):
    """P2P stores from the kick-drift kernel when symmetric memory is available (``PCL_GRAVITY_EXCHANGE=nccl`` forces the
    all-gather form), otherwise the NCCL all-gather overlapped with the local-block pass."""
    import os

    d = _dist()
    mode = os.environ.get("PCL_GRAVITY_EXCHANGE", "p2p")
    if mode == "p2p" and d.get_backend() == "nccl":
        try:
            return GravityExchangeP2P(posm, device)
        except Exception as e:  # symmetric memory not available on this box: say so once, use the collective
            if world()[0] == 0:
                print("physicl_b200: symmetric memory unavailable (%r); gravity uses the NCCL all-gather" % (e,), flush=True)
    return GravityExchange(posm, device)
