#!/bin/bash
# Multi-GPU call of round 2: parity tests (driver-style pytest), N-GPU bench with the in-kernel wait trace.
# usage: tools/gpu_r2_multi.sh TAG NGPUS [extra pytest -k expression]
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --trace --no-amg > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/${TAG}_bench_fused.err; echo "bench fused rc=$?"
cat gpurun_out/${TAG}_bench_fused.json
tail -5 gpurun_out/${TAG}_bench_fused.err
