#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: k_integrate_elastic (all element types, warp barriers and
# shared-memory phases), k_assemble_solver (shared-memory staging), k_sell_spmm. usage: TAG
TAG=${1:-san3}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="integrate_k_mech or integrate_k_heat or solver_order or multi_vector or negative_jacobian"
timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/${TAG}_memcheck.log
timeout 1200 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "integrate_k_mech or integrate_k_heat or solver_order" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/${TAG}_racecheck.log
timeout 1200 $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "integrate_k_mech or integrate_k_heat or solver_order" > gpurun_out/${TAG}_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -4 gpurun_out/${TAG}_synccheck.log
