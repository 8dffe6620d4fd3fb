#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py -x -q 2>&1 | tail -3
timeout 600 python tools/amg_check.py --edge 150 --hostprof 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:300]); continue
    print({k: d[k] for k in ('setup_phases_ms','resetup_phases_ms','amg_resetup_ms','amg_solve_ms','amg_iterations','hostprof_total_ms')})
"
