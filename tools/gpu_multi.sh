#!/bin/bash
# Multi-GPU call: parity of the fused peer-to-peer CG, then the weak-scaling bench (fused and NCCL variants).
# usage: tools/gpu_multi.sh TAG NGPUS
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused_peer or config_a_solve" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 $RUN --master-port 29511 tools/multi_gpu_check.py --edge 24 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?" >> gpurun_out/${TAG}_check.log
timeout 300 $RUN --master-port 29512 tools/multi_gpu_check.py --edge 20 --general >> gpurun_out/${TAG}_check.log 2>&1; echo "check general rc=$?" >> gpurun_out/${TAG}_check.log
grep -E "^\{|rc=|Error|error" gpurun_out/${TAG}_check.log | tail -12
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_fused.json 2> gpurun_out/${TAG}_bench_fused.err; echo "bench fused rc=$?"
timeout 600 $RUN --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --dist-cg nccl > gpurun_out/${TAG}_bench_nccl.json 2> gpurun_out/${TAG}_bench_nccl.err; echo "bench nccl rc=$?"
cat gpurun_out/${TAG}_bench_fused.json gpurun_out/${TAG}_bench_nccl.json
tail -5 gpurun_out/${TAG}_bench_fused.err
