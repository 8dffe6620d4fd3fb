#!/bin/bash
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 300 python tools/run_workloads.py --method cg --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_cg_graph.jsonl 2> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_cg_graph.jsonl
TFEM_MATERIAL_GRAPH=0 timeout 300 python tools/run_workloads.py --method cg --cube 0 --topopt 0 > gpurun_out/${TAG}_hyper_cg_nograph.jsonl 2>> gpurun_out/${TAG}_hyper.err; cat gpurun_out/${TAG}_hyper_cg_nograph.jsonl
tail -3 gpurun_out/${TAG}_hyper.err
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_reference_suite.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
