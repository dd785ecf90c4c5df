"""Kernel-timing aid (GPU box): stable compaction (pcl_compact) of 16 Mi slots at several live fractions."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from physicl_b200 import _capi
from physicl_b200.store import DeviceParticleStore

n = 16 * 2 ** 20
ctx = _capi.Context(0)
for frac in (1.0, 0.5, 0.17):
    st = DeviceParticleStore(ctx)
    rng = np.random.default_rng(1)
    r = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    r[0, rng.uniform(size=n) > frac] = np.nan
    v = np.ones((3, n), np.float32)
    g = st.add_group("photon", r, v)
    st.reserve_spare("photon")
    best = 1e9
    for _ in range(3):
        src, dst = g.soa(), g.soa(planes=g.spare)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.call("pcl_compact", st.stream(), C.byref(src), C.byref(dst), C.c_void_p(st._live_scratch.data_ptr()))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    live = int(st._live_scratch.item())
    gb = (n * 28 + live * 28) / 1e9
    print("live fraction %.2f: %7.1f us, %d survivors, %.0f GB/s of (28 B/slot read + 28 B/survivor written)" % (frac, best * 1e3, live, gb / (best * 1e-3)))
