#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "multi_vector" 2>&1 | tail -3
python tools/time_spmm.py 150
