#!/bin/bash
# The multi-GPU calls exactly as the driver launches them (defaults of bench.py). usage: TAG NGPUS
TAG=${1:-r2drv}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29541 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm rc=$?"
cut -c1-300 gpurun_out/${TAG}_ref.json
timeout 900 $RUN --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
grep "TfemError\|Error" gpurun_out/${TAG}_bench.err | tail -3
