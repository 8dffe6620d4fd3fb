#!/bin/bash
# usage: tools/gpu_r2_bench_only.sh TAG NGPUS
TAG=${1:-r2b}; N=${2:-4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
TFEM_AMG_TIMING=1 timeout 600 $RUN --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
grep "TfemError\|Error" gpurun_out/${TAG}_bench.err | tail -3
