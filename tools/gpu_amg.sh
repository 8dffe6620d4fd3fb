#!/bin/bash
# AMG on the GPU box: parity tests of the AMG kernels, timings on the benchmark cube, launch list of setup + solve.
TAG=${1:-amg}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/amg_check.py --edge 64 150 --jacobi --hostprof > gpurun_out/${TAG}_check.jsonl 2> gpurun_out/${TAG}_check.err
echo "check rc=$?"
cat gpurun_out/${TAG}_check.jsonl
tail -20 gpurun_out/${TAG}_check.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python tools/amg_check.py --edge 150 --rtol 1e-2 > gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_prof.log
