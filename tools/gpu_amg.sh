#!/bin/bash
# AMG bring-up on the GPU box: parity tests of the AMG kernels, then timings on the benchmark cube.
TAG=${1:-amg}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/amg_check.py --edge 32 64 100 150 --jacobi > gpurun_out/${TAG}_check.jsonl 2> gpurun_out/${TAG}_check.err
echo "check rc=$?"
cat gpurun_out/${TAG}_check.jsonl
tail -20 gpurun_out/${TAG}_check.err
