#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/misc_bench.json 2> gpurun_out/misc_bench.err
python -c "
import json; d = json.load(open('gpurun_out/misc_bench.json')); print(d['value'], d['config']['phases_ms'], d['e2e']['value'], d['e2e']['public_api']['value'], d['e2e']['public_api']['with_amg'], d['config']['amg_pcg']['setup_ms'], d['config']['amg_pcg']['solve_ms'])"
tail -3 gpurun_out/misc_bench.err
for m in cg amgx; do
  timeout 600 python tools/run_workloads.py --method $m > gpurun_out/misc_workloads_$m.jsonl 2> gpurun_out/misc_workloads_$m.err
  cat gpurun_out/misc_workloads_$m.jsonl | cut -c1-600; tail -2 gpurun_out/misc_workloads_$m.err
done
