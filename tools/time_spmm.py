"""Device time of the multi-vector SELL product against single-vector products at config B (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.set_default_dtype(torch.float64)
import torchfem_b200 as T  # noqa: E402
from torchfem_b200 import csr  # noqa: E402

dev = torch.device("cuda", 0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 150
nodes, elements, con, disp = bench.build_problem(T, torch, E, dev)
nodes, elements = nodes.to(dev), elements.to(dev)
p = csr.Pattern(elements, nodes.shape[0], 3)
A = p.matrix(torch.randn(p.nnz, device=dev))
A.sell()


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


x = torch.randn(p.n_dofs, device=dev)
y = torch.empty_like(x)
t1 = ev(lambda: A.matvec(x, out=y, fmt="sell"), 20)
print(f"single product {t1:.3f} ms")
for m in (1, 2, 4, 8, 9, 12, 16):
    X = torch.randn(p.n_dofs, m, device=dev)
    Y = torch.empty_like(X)
    tm = ev(lambda: A.matmat(X, out=Y))
    print(f"m = {m:2d}: block product {tm:.3f} ms = {tm / m:.3f} ms per vector ({m * t1 / tm:.2f}x the single products)")
