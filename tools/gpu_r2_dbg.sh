#!/bin/bash
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 600 python tools/damg_check.py --edge 100 > gpurun_out/${TAG}_damg_w1_e100.log 2>&1; echo "damg world1 e100 rc=$?"
tail -12 gpurun_out/${TAG}_damg_w1_e100.log | cut -c1-1500
