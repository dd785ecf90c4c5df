"""Tuning aid (GPU box): where the HOST time of Simulation.run_steps goes on the default bench pipeline."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench

sim, esc, sign = bench.photon_sim(bench.PHOTONS_PER_GPU, 0, 0)
sim.run_steps(5)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
sim.run_steps(40)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue time %.3f ms, until GPU idle %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
