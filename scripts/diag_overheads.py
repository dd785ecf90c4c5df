"""Scratch diagnostics (GPU box): host enqueue cost per step, compaction cost, kernel-only loop."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from physicl_b200 import _capi

n = bench.PHOTONS_PER_GPU
sim, esc, sign = bench.photon_sim(n, 0, 0)
sim.feedback_every = 0
sim.run_steps(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
sim.run_steps(10)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("run_steps(10): enqueue %.1f us/step, total %.1f us/step" % ((t1 - t0) * 1e5, (t2 - t0) * 1e5))
st = sim.store
g = st.group("photon")
ctx = sim.cl_ctx
sp = _capi.ScatterParams(k=bench.A_N, c=bench.C_LIGHT, mode=0)
rg = _capi.Rng(seed=1, step=100)
pl = _capi.make_planes([])
soa = g.soa()
tab = torch.zeros((64, 16), dtype=torch.int64, device="cuda")
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(bench.R_ESCAPE ** 2),
             C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(10))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("pcl_photon_steps(10): enqueue %.1f us/step, total %.1f us/step" % ((t1 - t0) * 1e5, (t2 - t0) * 1e5))
    print(tab[:10, [0, 6, 7]].cpu().numpy().T)
for rep in range(3):
    torch.cuda.synchronize()
    live_before = g.n
    t0 = time.perf_counter()
    nl = st.compact("photon")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("compact: %d -> %d slots in %.1f us" % (live_before, nl, (t2 - t0) * 1e6))
    soa = g.soa()
    ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(bench.R_ESCAPE ** 2),
             C.byref(pl), C.c_void_p(tab.data_ptr()), C.c_uint32(6))
print("fp32 peak TF", ctx.fp32_peak_tflops(), "copy GB/s", ctx.copy_peak_gbs(1 << 30))
