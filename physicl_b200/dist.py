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
