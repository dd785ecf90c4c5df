#!/bin/bash
# compute-sanitizer over a slice of the parity tests (small sizes; every kernel family of the hot path is launched):
# memcheck (global/shared out-of-bounds, misaligned), racecheck (shared-memory hazards), synccheck (barrier misuse).
# Run on the GPU box:  gpurun --timeout 1500 -- 'bash scripts/sanitize_r2.sh'
# Results: gpurun_out/sanitize/<tool>.log (summary lines copied to profiles/r2/sanitize_summary.txt)
mkdir -p gpurun_out/sanitize
SEL='fused_step_bit_exact_vs_twin or compacting_step_matches_twin_by_id or timesteps_fused_in_registers_equal_single_steps or gravity_uniform_mass or gravity_split_blocks or gravity_matches or compaction_is_stable or empty_and_tiny or kinematics_bit_exact or kinematics_timesteps_in_registers or multi_timestep_host_round_trips or scatter_flags or planck_bins or pingpong_loop'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize/$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize/$tool.log | tail -3
done
