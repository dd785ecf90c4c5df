#!/bin/bash
# Tuning aid (GPU box): default bench under different timesteps-per-launch / compaction cadences.
# usage: scripts/sweep_fuse.sh > gpurun_out/sweep_fuse.log
for fuse in 1 8; do
  for fe in 4 8 16; do
    for cad in "" 2 3 4 8; do
      out=$(PCL_PHOTON_FUSE=$fuse PCL_FEEDBACK_EVERY=$fe PCL_COMPACT_CADENCE=$cad python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1)
      echo "fuse=$fuse feedback_every=$fe cadence=${cad:-adaptive} $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value=%.1fG ms/step=%.4f frac=%.3f launches=%d" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))')"
    done
  done
done
