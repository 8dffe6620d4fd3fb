#!/bin/bash
TAG=${1:-dbg}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29521 tools/damg_check.py --edge 16 > gpurun_out/${TAG}_damg.log 2>&1; echo "damg rc=$?"
grep "^{" gpurun_out/${TAG}_damg.log | tail -1 | cut -c1-2500
TFEM_AMG_TIMING=1 timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
grep "TfemError\|Error" gpurun_out/${TAG}_bench.err | tail -3
timeout 300 python tools/prof_hyper.py > gpurun_out/${TAG}_hyper_profile.txt 2>&1; head -50 gpurun_out/${TAG}_hyper_profile.txt | cut -c1-200
