"""Device time of integrate_k and assemble at config B (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.set_default_dtype(torch.float64)
import torchfem_b200 as T  # noqa: E402
from torchfem_b200 import csr  # noqa: E402
from torchfem_b200.elements import Hexa1  # noqa: E402
from torchfem_b200.materials import IsotropicElasticity3D  # noqa: E402

dev = torch.device("cuda", 0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 150
nodes, elements, con, disp = bench.build_problem(T, torch, E, dev)
bref, w = bench.element_tables(Hexa1)
C = IsotropicElasticity3D(torch.full((len(elements),), 1000.0, device=dev), torch.full((len(elements),), 0.3, device=dev)).C
nodes, elements = nodes.to(dev), elements.to(dev)
is_con = con.ravel().to(torch.uint8).to(dev)
p = csr.Pattern(elements, nodes.shape[0], 3)


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False)
vals = torch.empty(p.nnz, device=dev)
sv = torch.empty(p.sell_structure.padded, device=dev)
dinv = torch.empty(p.n_dofs, device=dev)
A = p.matrix(vals)
print("integrate_k %.3f ms   assemble %.3f ms   assemble in solver order (SELL-32 + 1/diag, no CSR) %.3f ms   "
      "both orders %.3f ms   CSR + 1/diag by the solver-order kernel %.3f ms   CSR -> SELL copy %.3f ms" % (
          ev(lambda: csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False)),
          ev(lambda: csr.assemble(p, k, is_con, out=vals)),
          ev(lambda: csr.assemble(p, k, is_con, csr=False, sell_out=sv, dinv_out=dinv)),
          ev(lambda: csr.assemble(p, k, is_con, out=vals, sell_out=sv, dinv_out=dinv)),
          ev(lambda: csr.assemble(p, k, is_con, out=vals, dinv_out=dinv)),
          ev(lambda: (setattr(A, "_sell_vals", None), A._sell_mats.clear(), A.sell()))))
