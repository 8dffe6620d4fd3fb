#!/bin/bash
# compute-sanitizer over the AMG kernels (small fixtures): memcheck, racecheck (shared memory), initcheck subset
mkdir -p gpurun_out
K='other_block_sizes or aggregation_variants or spgemm_against_scipy'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_amg.py -x -q -k "$K" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|=========.*(Invalid|hazard|Error)" gpurun_out/sanitize_$tool.log | head -12
done
