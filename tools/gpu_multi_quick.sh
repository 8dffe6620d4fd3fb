#!/bin/bash
# Quick multi-GPU regression: parity of the fused peer-to-peer CG on small cubes, then the weak-scaling bench (fused).
TAG=${1:-mq}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/multi_gpu_check.py --edge 24 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?" >> gpurun_out/${TAG}_check.log
timeout 300 $RUN --master-port 29512 tools/multi_gpu_check.py --edge 20 --general >> gpurun_out/${TAG}_check.log 2>&1; echo "check general rc=$?" >> gpurun_out/${TAG}_check.log
grep -E "^\{|rc=|Error|error" gpurun_out/${TAG}_check.log | tail -12 | cut -c1-400
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/${TAG}_bench_fused.err; echo "bench fused rc=$?"
cat gpurun_out/${TAG}_bench_fused.json | cut -c1-1500
tail -5 gpurun_out/${TAG}_bench_fused.err
