#!/usr/bin/env python
"""Launch time of the fused photon kernel against the number of timesteps per launch (in place and compacting):
T(m) = F + c*m separates the load/store/compaction phase of a launch from its per-timestep work.  Run under gpurun."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physicl_b200 import _capi  # noqa: E402
from physicl_b200.store import DeviceParticleStore  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 * 2 ** 20
ctx = _capi.Context(0)
dev = torch.device("cuda", 0)
c = 299792458.0


def fresh():
    st = DeviceParticleStore(ctx)
    r = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v = torch.zeros((3, n), dtype=torch.float32, device=dev)
    v[0].fill_(c)
    g = st.add_group("photon", r, v)
    return st, g


sp = _capi.ScatterParams(k=1e-6, c=c, mode=0)
pl = _capi.make_planes([])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for compact in (0, 1):
    for m in (1, 2, 3, 4, 6, 8):
        st, g = fresh()
        times = []
        for rep in range(4):
            rg = _capi.Rng(seed=1, step=rep * 8)
            first = st.new_rows(m)
            if compact:
                pp = st.pingpong("photon")
                for b in pp.buf:
                    b.dx = b.dy = b.dz = None
                # step index chosen so that the launch ends exactly on a compaction boundary
                rg.step = rep * 8
                ev[0].record()
                ctx.call("pcl_photon_steps_pp", st.stream(), C.byref(pp), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                         C.byref(pl), st.row_ptr(first), C.c_uint32(m), C.c_uint32(m if (rep * 8) % m == 0 else 0))
                ev[1].record()
                st.adopt_pingpong("photon", pp, 1)
            else:
                soa = g.soa()
                soa.dx = soa.dy = soa.dz = None
                ev[0].record()
                ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
                         C.byref(pl), st.row_ptr(first), C.c_uint32(m))
                ev[1].record()
            torch.cuda.synchronize()
            times.append(ev[0].elapsed_time(ev[1]) * 1e3)
        print("compact=%d m=%d  launch %.1f us (min of %s)" % (compact, m, min(times[1:]), ["%.0f" % t for t in times]), flush=True)
        del st, g
        torch.cuda.empty_cache()
