#!/bin/bash
# One GPU call: parity tests, bench (both arms), launch list, full ncu captures of the kernels of a step.
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
# integrate + the two assemblies of the driver
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate|k_assemble' -c 3 \
   -o gpurun_out/${TAG}_integrate_assemble -f python tools/prof_driver.py --edge 150 --iters 3 > gpurun_out/${TAG}_ncu_ia.log 2>&1
# the kernels of a CG iteration (26 k_sell_spmv launches of the SpMV timing loops are skipped)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sell_spmv|k_cg_update|k_cg_direction' -s 28 -c 6 \
   -o gpurun_out/${TAG}_cg_iteration -f python tools/prof_driver.py --edge 150 --iters 3 > gpurun_out/${TAG}_ncu_cg.log 2>&1
# AMG: launch list of one hierarchy build + 2 PCG iterations, full captures of the numeric SpGEMMs and of one V cycle
# (the 8 k_amg_spmv launches of the level-0 power iteration are skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_amg_launches.csv python tools/prof_amg.py > gpurun_out/${TAG}_amg_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spgemm_num|k_prolongator' -c 4 \
   -o gpurun_out/${TAG}_amg_setup -f python tools/prof_amg.py > gpurun_out/${TAG}_ncu_amg_setup.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_amg_spmv' -s 8 -c 6 \
   -o gpurun_out/${TAG}_amg_cycle -f python tools/prof_amg.py > gpurun_out/${TAG}_ncu_amg_cycle.log 2>&1
# Assembly at 1.35 M DOFs: tie and tie + reference point, node blocks; then the opt-in bordered PCG on the point case
timeout 120 python tools/assembly_check.py --nodes 61 --method cg > gpurun_out/${TAG}_assembly_check.jsonl 2> gpurun_out/${TAG}_assembly_check.err
TFEM_TEST_BORDERED=1 timeout 120 python -m pytest tests/test_gpu_assembly.py -q -k bordered > gpurun_out/${TAG}_pytest_bordered.log 2>&1
timeout 120 python tools/assembly_check.py --nodes 61 --method cg --cases point --long-rows 2000 >> gpurun_out/${TAG}_assembly_check.jsonl 2>> gpurun_out/${TAG}_assembly_check.err
tail -3 gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest_bordered.log
cat gpurun_out/${TAG}_assembly_check.jsonl
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
