#!/bin/bash
TAG=${1:-r2c2}; N=2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29535 bench.py --gpus $N --config C --steps 1 --warmup 1 --probe > gpurun_out/${TAG}_bench_configC.json 2> gpurun_out/${TAG}_bench_configC.err; echo "bench C rc=$?"
cat gpurun_out/${TAG}_bench_configC.json | cut -c1-1500
grep "TfemError\|Error" gpurun_out/${TAG}_bench_configC.err | tail -3
timeout 900 $RUN --master-port 29536 bench.py --gpus $N --config C --steps 1 --warmup 1 --probe --rtol 1e-10 --no-amg > gpurun_out/${TAG}_bench_configC_rtol1e-10.json 2> gpurun_out/${TAG}_bench_configC_rtol.err; echo "bench C parity rc=$?"
cat gpurun_out/${TAG}_bench_configC_rtol1e-10.json | cut -c1-1500
