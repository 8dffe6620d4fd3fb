#!/bin/bash
# cooperative single-kernel Krylov loops on small systems: parity tests, then solve times with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 1 0; do
  echo "== TFEM_CG_COOP=$c"
  TFEM_CG_COOP=$c python - <<'PY'
import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
torch.set_default_dtype(torch.float64)
import bench, torchfem_b200 as T
from torchfem_b200 import csr
from oracle import fem_oracle as O
dev = torch.device("cuda", 0)
for E in (16, 32, 48, 64):
    nodes, elements, con, disp = bench.build_problem(T, torch, E, dev)
    bref, w = (torch.as_tensor(t) for t in O.hexa1_tables())
    C = torch.as_tensor(O.isotropic_C3d(1000.0, 0.3, 1)).expand(len(elements), 3, 3, 3, 3).contiguous().to(dev)
    nodes, elements = nodes.to(dev), elements.to(dev)
    is_con = con.ravel().to(torch.uint8).to(dev); ubc = (disp.ravel() * con.ravel()).to(dev)
    p = csr.Pattern(elements, nodes.shape[0], 3)
    k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C)
    b = torch.empty(p.n_dofs, device=dev)
    A = p.matrix(csr.assemble(p, k, is_con, ubc=ubc, lift=b)); A.sell()
    out = {"edge": E, "n_dofs": p.n_dofs}
    for m in ("cg", "minres"):
        for rep in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            x, _, st = csr.krylov_solve(A, b, method=m, rtol=1e-8)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
        res = float((A.matvec(x) - b).norm() / b.norm())
        out[m] = {"ms": round(dt, 3), "its": st["iterations"], "us_per_it": round(1e3 * dt / st["iterations"], 1), "relres": res}
    print(json.dumps(out))
PY
done
