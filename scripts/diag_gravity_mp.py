"""Tuning aid (multi-GPU box): where a sharded gravity step spends its time.
torchrun --nproc-per-node N scripts/diag_gravity_mp.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from physicl_b200 import _capi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _capi.Context(local)
N = 262144
nl = N // world
rng = np.random.default_rng(rank)
posm = torch.from_numpy(rng.normal(size=(nl, 4)).astype(np.float32)).cuda()
allp = torch.empty((N, 4), dtype=torch.float32, device="cuda")
acc = torch.zeros((3, nl), dtype=torch.float32, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
args = (C.c_float(1.0), C.c_float(1e-4), p(acc[0]), p(acc[1]), p(acc[2]))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = np.zeros(3)
for it in range(8):
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    dist.all_gather_into_tensor(allp, posm)
    ev[1].record()
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(nl), p(posm), C.c_uint64(nl), *args, 0, C.c_uint64(0), C.c_uint64(0))
    ev[2].record()
    ctx.call("pcl_gravity_accel", None, p(posm), C.c_uint64(nl), p(allp), C.c_uint64(N), *args, 1, C.c_uint64(rank * nl), C.c_uint64((rank + 1) * nl))
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        tot += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
tot /= 5
inter = nl * N
print("rank %d/%d: gather %.3f ms, local %.3f ms, remote %.3f ms; %.2f TFLOP/s in the two kernels" % (
    rank, world, tot[0], tot[1], tot[2], 20 * inter / ((tot[1] + tot[2]) * 1e-3) / 1e12))
dist.destroy_process_group()
