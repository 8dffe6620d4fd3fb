#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 1 0; do
  echo "== TFEM_CG_COOP=$c"
  TFEM_CG_COOP=$c timeout 600 python tools/amg_check.py --edge 16 32 64 --jacobi 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:200]); continue
    print({k: d[k] for k in ('edge','n_dofs','jacobi_solve_ms','jacobi_iterations','rel_diff_amg_vs_jacobi')})
"
  TFEM_CG_COOP=$c timeout 600 python tools/run_workloads.py --cube 40 --topopt 20 --hyper 65 --method cg 2>/dev/null | cut -c1-260
done
