#!/bin/bash
# One GPU call: parity tests, bench, launch list, one full ncu capture of the dominant kernel.
# usage: tools/gpu_round.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python tools/prof_driver.py --edge 150 --iters 20 > gpurun_out/${TAG}_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -s 30 -c 1 \
   -o gpurun_out/${TAG}_bsell_spmv -f python tools/prof_driver.py --edge 150 --iters 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
