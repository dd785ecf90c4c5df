"""Kernel-timing aid (GPU box): per-launch time of the run-time compiled variable-density photon step
(radial-atmosphere expression, Rayleigh law) next to the pre-compiled wavelength-law step, 16 Mi photons."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from physicl_b200 import _capi, jit
from physicl_b200.store import DeviceParticleStore

n = 16 * 2 ** 20
ctx = _capi.Context(0)
st = DeviceParticleStore(ctx)
rng = np.random.default_rng(1)
r = rng.uniform(-2e4, 2e4, (3, n)).astype(np.float32)
d = rng.normal(size=(3, n))
v = (299792458.0 * d / np.linalg.norm(d, axis=0)).astype(np.float32)
E = rng.uniform(1e-19, 9e-19, n)
g = st.add_group("photon", r, v, E=E)
hc = 6.62607015e-34 * 299792458.0
pl = _capi.make_planes([])
tab = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
K = 20
step = [0]
expr = "6e26 * exp(-1 * (sqrt(pow(r0[gid], 2) + pow(r1[gid], 2) + pow(r2[gid], 2)) - 1000.0)/(8000.0))"
mod = jit.Module(ctx, jit.photon_source(expr, True))
a = 5.1e-31 * (532e-9) ** 4
vn = _capi.VarnParams(kd=a * (g.e0 / hc) ** 4, e0=g.e0, a_slot=a, n_slot=1.0)
spw = _capi.ScatterParams(k=a * 2.5e25 * (g.e0 / hc) ** 4, c=299792458.0, mode=_capi.SCATTER_WAVELENGTH)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run_jit():
    soa = g.soa()
    rg = _capi.Rng(seed=1, step=step[0])
    step[0] += K
    ctx.call("pcl_photon_steps_jit", st.stream(), mod.kernel("pcl_jit_photon_step"), C.byref(soa), C.c_float(1e-7), C.byref(spw),
             C.byref(vn), C.byref(rg), C.c_float(0.0), C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(K))


def run_fixed():
    soa = g.soa()
    rg = _capi.Rng(seed=1, step=step[0])
    step[0] += K
    ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-7), C.byref(spw), C.byref(rg), C.c_float(0.0),
             C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(K))


for name, fn in (("pcl_jit_photon_step (variable n, wavelength law)", run_jit), ("pcl_k_photon_step_tma<1,0,0>", run_fixed)):
    us = timed(fn) / K * 1e3
    row = tab[K - 1].cpu().numpy()
    f = row[_capi.T_SCATTERED] / max(row[_capi.T_LIVE_IN], 1)
    b = (40 + 12 * f) * n
    print("%-52s %8.1f us/step  f=%.3f  %6.0f GB/s algorithmic  %5.1f G photon-steps/s" % (name, us, f, b / us / 1e3, n / us / 1e3))
