#!/bin/bash
# usage: tools/gpu_r2_damg.sh TAG NGPUS
TAG=${1:-r2d}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python tools/damg_check.py --edge 12 > gpurun_out/${TAG}_damg_w1.log 2>&1; echo "damg world1 rc=$?"
tail -5 gpurun_out/${TAG}_damg_w1.log
timeout 300 $RUN --master-port 29521 tools/damg_check.py --edge 16 > gpurun_out/${TAG}_damg.log 2>&1; echo "damg rc=$?"
tail -8 gpurun_out/${TAG}_damg.log
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --trace > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --gpus 1 --dist-path --steps 2 --warmup 2 --trace --no-amg > gpurun_out/${TAG}_bench_n1_distpath.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/${TAG}_bench_n1_distpath.json
