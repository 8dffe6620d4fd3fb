#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for gflag in 1 0; do
  echo "== TFEM_CG_GRAPH=$gflag"
  TFEM_CG_GRAPH=$gflag timeout 600 python tools/amg_check.py --edge 16 32 64 --jacobi 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:200]); continue
    print({k: d[k] for k in ('edge','n_dofs','jacobi_solve_ms','jacobi_iterations','amg_solve_ms','amg_iterations')})
"
done
