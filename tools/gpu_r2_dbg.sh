#!/bin/bash
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 400 python tools/prof_api.py --edge 150 --method amgx --rows 40 > gpurun_out/${TAG}_api_amgx_profile.txt 2>&1; head -60 gpurun_out/${TAG}_api_amgx_profile.txt | cut -c1-190
