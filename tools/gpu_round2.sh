#!/bin/bash
# One single-GPU call of round 2: parity tests, bench (both arms), launch list + full ncu captures of the kernels of a
# step, the reference's benchmark problems through the model API, Assembly with a reference point (long-row side path).
# usage: tools/gpu_round2.sh TAG
TAG=${1:-r2x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
# launch list of the bench command itself (first 700 launches: setup, warm-up step and the head of the solve)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-api --no-amg --no-cpu-baseline > gpurun_out/${TAG}_prof_bench.log 2>&1
# launch list of a whole short step (20 CG iterations)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python tools/prof_driver.py --edge 150 --iters 20 > gpurun_out/${TAG}_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate|k_assemble' -c 3 \
   -o gpurun_out/${TAG}_integrate_assemble -f python tools/prof_driver.py --edge 150 --iters 3 > gpurun_out/${TAG}_ncu_ia.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell_spmv|k_cg_update|k_cg_direction' -s 28 -c 6 \
   -o gpurun_out/${TAG}_cg_iteration -f python tools/prof_driver.py --edge 150 --iters 3 > gpurun_out/${TAG}_ncu_cg.log 2>&1
timeout 300 python tools/run_workloads.py --method cg > gpurun_out/${TAG}_workloads_cg.jsonl 2> gpurun_out/${TAG}_workloads_cg.err
cat gpurun_out/${TAG}_workloads_cg.jsonl
TFEM_MATERIAL_GRAPH=0 timeout 300 python tools/run_workloads.py --method cg --cube 0 --topopt 0 > gpurun_out/${TAG}_workloads_cg_nograph.jsonl 2>> gpurun_out/${TAG}_workloads_cg.err
cat gpurun_out/${TAG}_workloads_cg_nograph.jsonl
timeout 300 python tools/run_workloads.py --method amgx > gpurun_out/${TAG}_workloads_amgx.jsonl 2> gpurun_out/${TAG}_workloads_amgx.err
cat gpurun_out/${TAG}_workloads_amgx.jsonl
timeout 200 python tools/assembly_check.py --nodes 61 --method cg > gpurun_out/${TAG}_assembly_check.jsonl 2> gpurun_out/${TAG}_assembly_check.err
cat gpurun_out/${TAG}_assembly_check.jsonl
tail -3 gpurun_out/${TAG}_assembly_check.err
