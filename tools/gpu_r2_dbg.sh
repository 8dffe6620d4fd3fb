#!/bin/bash
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 300 python tools/run_workloads.py --method cg --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_cg_graph.jsonl 2> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_cg_graph.jsonl
TFEM_MATERIAL_GRAPH=0 timeout 300 python tools/run_workloads.py --method cg --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_cg_nograph.jsonl 2>> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_cg_nograph.jsonl
timeout 300 python tools/run_workloads.py --method amgx --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_amgx.jsonl 2>> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_amgx.jsonl
TFEM_MATERIAL_GRAPH=0 timeout 300 python tools/run_workloads.py --method amgx --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_amgx_nograph.jsonl 2>> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_amgx_nograph.jsonl
tail -3 gpurun_out/${TAG}_hyper.err
timeout 300 python tools/assembly_check.py --nodes 61 --method cg --cases point > gpurun_out/${TAG}_assembly_point.jsonl 2> gpurun_out/${TAG}_assembly_point.err; cat gpurun_out/${TAG}_assembly_point.jsonl; tail -3 gpurun_out/${TAG}_assembly_point.err
TFEM_AMG_TIMING=1 timeout 300 python tools/damg_check.py --edge 100 > gpurun_out/${TAG}_damg_w1_e100.log 2>&1; echo "damg world1 e100 rc=$?"
tail -3 gpurun_out/${TAG}_damg_w1_e100.log | cut -c1-3000
timeout 300 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_k2.sh > gpurun_out/${TAG}_k2.log 2>&1; cat gpurun_out/${TAG}_k2.log
