#!/bin/bash
# AMG tuning knobs on the GPU box (per-iteration time of AMG-PCG at config B).
mkdir -p gpurun_out
for cfg in "" "TFEM_AMG_OCC6=1" "TFEM_AMG_BCSR_MIN_AVG=96" "TFEM_AMG_BCSR_MIN_AVG=24"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/amg_check.py --edge 150 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:300]); continue
    print({k: d[k] for k in ('amg_setup_ms','amg_resetup_ms','amg_solve_ms','amg_iterations','amg_ms_per_iteration','levels')})
"
done
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/tune_bench.json 2> gpurun_out/tune_bench.err
python -c "
import json; d = json.load(open('gpurun_out/tune_bench.json')); print(d['value'], d['e2e']['public_api'], d['config']['amg_pcg'])"
tail -3 gpurun_out/tune_bench.err
