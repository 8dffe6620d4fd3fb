#!/bin/bash
# Multi-GPU measurement call of round 2. usage: tools/gpu_r2_scale.sh TAG NGPUS [c-parity]
TAG=${1:-r2n}; N=${2:-8}; PAR=${3:-}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 240 $RUN --master-port 29531 tools/multi_gpu_check.py --edge 16 > gpurun_out/${TAG}_check_fused.log 2>&1; echo "fused check rc=$?"
grep "^{" gpurun_out/${TAG}_check_fused.log | tail -1
timeout 240 $RUN --master-port 29532 tools/damg_check.py --edge 16 > gpurun_out/${TAG}_check_damg.log 2>&1; echo "damg check rc=$?"
grep "^{" gpurun_out/${TAG}_check_damg.log | tail -1
timeout 240 $RUN --master-port 29533 tools/multi_gpu_check.py --edge 6 --hexa2 > gpurun_out/${TAG}_check_hexa2.log 2>&1; echo "hexa2 check rc=$?"
grep "^{" gpurun_out/${TAG}_check_hexa2.log | tail -1
timeout 600 $RUN --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --trace > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
grep "TfemError\|Error" gpurun_out/${TAG}_bench.err | tail -3
timeout 900 $RUN --master-port 29535 bench.py --gpus $N --config C --steps 1 --warmup 1 --probe > gpurun_out/${TAG}_bench_configC.json 2> gpurun_out/${TAG}_bench_configC.err; echo "bench C rc=$?"
cat gpurun_out/${TAG}_bench_configC.json
grep "TfemError\|Error" gpurun_out/${TAG}_bench_configC.err | tail -3
if [ -n "$PAR" ]; then
  timeout 900 $RUN --master-port 29536 bench.py --gpus $N --config C --steps 1 --warmup 1 --probe --rtol 1e-10 --no-amg > gpurun_out/${TAG}_bench_configC_rtol1e-10.json 2> gpurun_out/${TAG}_bench_configC_rtol.err; echo "bench C parity rc=$?"
  cat gpurun_out/${TAG}_bench_configC_rtol1e-10.json
fi
