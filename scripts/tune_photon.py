"""Kernel-tuning aid (GPU box): per-launch time of the in-place and the compacting photon kernels
on 16 Mi live photons, timed with CUDA events inside a continuously busy stream."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from physicl_b200 import _capi
from physicl_b200.store import DeviceParticleStore

n = 16 * 2 ** 20
ctx = _capi.Context(0)
st = DeviceParticleStore(ctx)
r = np.zeros((3, n), np.float32)
v = np.zeros((3, n), np.float32)
v[0] = 299792458.0
g = st.add_group("photon", r, v)
st.reserve_spare("photon")
sp = _capi.ScatterParams(k=1e-6, c=299792458.0, mode=0)
pl = _capi.make_planes([])
tab = torch.zeros((64, 16), dtype=torch.int64, device="cuda")


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        fn()  # warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


K = 20
step = [0]


def inplace():
    soa = g.soa()
    rg = _capi.Rng(seed=1, step=step[0])
    step[0] += K
    ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
             C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(K))


def compacting():
    pp = st.pingpong("photon")
    rg = _capi.Rng(seed=1, step=step[0])
    step[0] += K
    ctx.call("pcl_photon_steps_pp", st.stream(), C.byref(pp), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(0.0),
             C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(K), C.c_uint32(1))
    st.adopt_pingpong("photon", pp, K)


t_in = timed(inplace) / K * 1e3
f = float(tab[K - 1, 4].item()) / float(tab[K - 1, 7].item())
t_cp = timed(compacting) / K * 1e3
st.sync_n("photon")
assert g.n == n
print("lib=%s in-place %.1f us/step (%.0f GB/s algorithmic at f=%.3f)  compacting %.1f us/step" % (
    _capi.LIB_PATH.split("/")[-1], t_in, (36 + 12 * f) * n / t_in / 1e3, f, t_cp))
