#!/bin/bash
# Single-GPU call of round 2: selected or full parity tests, bench (both arms).
# usage: tools/gpu_r2_single.sh TAG [pytest args...]
TAG=${1:-r2s}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
