#!/bin/bash
# Launch list of the final bench command and full captures of the two matrix-build kernels of the step. usage: TAG
TAG=${1:-r2fin}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
   --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-api --no-amg --no-cpu-baseline > gpurun_out/${TAG}_prof_bench.log 2>&1
tail -2 gpurun_out/${TAG}_prof_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_elastic|k_assemble_solver' -c 2 \
   -o gpurun_out/${TAG}_build -f python tools/time_k1.py 150 > gpurun_out/${TAG}_ncu_build.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_build.log
