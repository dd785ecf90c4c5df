"""Diagnosis aid (GPU box, 1..N ranks under torchrun): per-chunk host and device timeline of the default bench's
timed region, to find where multi-rank runs lose time."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import bench

rank, world, local = bench.init_dist(int(os.environ.get("WORLD_SIZE", "1")))
for trial in range(6):
    sim, esc, sign = bench.photon_sim(bench.PHOTONS_PER_GPU, rank, local)
    sim.device_store()
    bench.barrier_sync(world)
    sim.run_steps(5)
    plan = sim._plan()
    fused = plan[1]
    orig = fused.run_many
    marks = []

    def traced(sim_, k, dt, ts, orig=orig, marks=marks):
        e = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        orig(sim_, k, dt, ts)
        e.record()
        marks.append((k, t0, time.perf_counter(), e))

    fused.run_many = traced
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.barrier_sync(world)
    h0 = time.perf_counter()
    ev0.record()
    sim.run_steps(40)
    ev1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    line = "rank %d trial %d: gpu %.3f ms, host enqueue %.3f ms |" % (rank, trial, ms, (h1 - h0) * 1e3)
    for k, t0, t1, e in marks:
        line += " k=%d host[%.2f..%.2f] gpu_done@%.2f" % (k, (t0 - h0) * 1e3, (t1 - h0) * 1e3, ev0.elapsed_time(e))
    print(line, flush=True)
    del sim, esc, sign, fused, plan
    torch.cuda.empty_cache()
if world > 1:
    import torch.distributed as dist

    dist.destroy_process_group()
