#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small cases). usage: TAG
TAG=${1:-san}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "assemble_rhs or fused_peer or integrate_k_mech or test_assemble" > gpurun_out/${TAG}_memcheck_kernels.log 2>&1; echo "memcheck kernels rc=$?"
tail -4 gpurun_out/${TAG}_memcheck_kernels.log
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_assembly.py -q -m gpu -x -k "long_rows" > gpurun_out/${TAG}_memcheck_longrows.log 2>&1; echo "memcheck long rows rc=$?"
tail -4 gpurun_out/${TAG}_memcheck_longrows.log
timeout 900 $CS --tool memcheck --error-exitcode 9 python tools/damg_check.py --edge 8 > gpurun_out/${TAG}_memcheck_damg.log 2>&1; echo "memcheck damg world1 rc=$?"
tail -4 gpurun_out/${TAG}_memcheck_damg.log | cut -c1-400
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "assemble_rhs or integrate_k_mech" > gpurun_out/${TAG}_racecheck_kernels.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/${TAG}_racecheck_kernels.log
