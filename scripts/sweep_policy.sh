#!/bin/bash
# Tuning aid (GPU box): sensitivity of the default bench to the feedback chunk and the compaction cadence.
run() { python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 ms/step=%.4f value=%.3e frac=%.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))"; }
run "default"
for f in 2 8 16; do PCL_FEEDBACK_EVERY=$f run "feedback_every=$f"; done
for m in 1 2 3 4 6; do PCL_COMPACT_CADENCE=$m run "cadence=$m"; done
for m in 2 3; do PCL_FEEDBACK_EVERY=16 PCL_COMPACT_CADENCE=$m run "feedback_every=16 cadence=$m"; done
