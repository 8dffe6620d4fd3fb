#!/bin/bash
# Final 8-GPU call of round 2. usage: tools/gpu_r2_final8.sh TAG
TAG=${1:-r2f8}; N=8
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29532 tools/damg_check.py --edge 16 > gpurun_out/${TAG}_check_damg.log 2>&1; echo "damg check rc=$?"
grep "^{" gpurun_out/${TAG}_check_damg.log | tail -1 | cut -c1-1500
timeout 600 $RUN --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --trace > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
grep "TfemError\|Error" gpurun_out/${TAG}_bench.err | tail -3
TFEM_DAMG_GATHER_MAX=100000 TFEM_AMG_TIMING=1 timeout 600 $RUN --master-port 29535 bench.py --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_gather100k.json 2> gpurun_out/${TAG}_bench_gather100k.err; echo "bench gather100k rc=$?"
cat gpurun_out/${TAG}_bench_gather100k.json
