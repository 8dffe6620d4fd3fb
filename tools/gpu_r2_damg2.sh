#!/bin/bash
# usage: tools/gpu_r2_damg2.sh TAG NGPUS
TAG=${1:-r2d}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --trace > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
grep -v "^$" gpurun_out/${TAG}_bench.err | grep -A12 "Traceback" | tail -14
timeout 300 python bench.py --gpus 1 --dist-path --steps 2 --warmup 2 --trace --no-amg > gpurun_out/${TAG}_bench_n1_distpath.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/${TAG}_bench_n1_distpath.json
