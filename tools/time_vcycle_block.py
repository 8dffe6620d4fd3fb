"""Device time of the block V cycle (4 vectors per pass over the finest matrix) against single cycles at config B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.set_default_dtype(torch.float64)
import torchfem_b200 as T  # noqa: E402
from torchfem_b200 import csr  # noqa: E402
from torchfem_b200.amg import AMGPreconditioner  # noqa: E402
from torchfem_b200.elements import Hexa1  # noqa: E402
from torchfem_b200.materials import IsotropicElasticity3D  # noqa: E402

dev = torch.device("cuda", 0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 150
nodes, elements, con, disp = bench.build_problem(T, torch, E, dev)
nodes, elements = nodes.to(dev), elements.to(dev)
is_con = con.ravel().to(torch.uint8).to(dev)
p = csr.Pattern(elements, nodes.shape[0], 3)
bref, w = bench.element_tables(Hexa1)
C = IsotropicElasticity3D(torch.full((len(elements),), 1000.0, device=dev), torch.full((len(elements),), 0.3, device=dev)).C
k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False)
A = p.matrix(csr.assemble(p, k, is_con))
del k, C
amg = AMGPreconditioner(A)


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


r = torch.randn(p.n_dofs, device=dev)
t1 = ev(lambda: amg.apply(r), 10)
print(f"levels {[int(lv.n) for lv in amg.levels]}  single V cycle {t1:.3f} ms")
for m in (4, 9):
    R = torch.randn(p.n_dofs, m, device=dev)
    tm = ev(lambda: amg.apply_block(R))
    print(f"m = {m}: block cycle {tm:.3f} ms = {tm / m:.3f} ms per vector ({m * t1 / tm:.2f}x the single cycles)")
