"""Short driver for ncu captures: one pass of pattern -> integrate -> assemble -> a few SpMV / CG iterations.
    python tools/prof_driver.py --edge 150 --iters 5
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=150)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--method", default="cg")
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200 import csr
    from torchfem_b200.elements import Hexa1
    from torchfem_b200.materials import IsotropicElasticity3D

    dev = torch.device("cuda", 0)
    nodes, elements, con, disp = bench.build_problem(T, torch, a.edge, dev)
    n_dofs = nodes.numel()
    bref, w = bench.element_tables(Hexa1)
    C = IsotropicElasticity3D(torch.full((len(elements),), 1000.0, device=dev), torch.full((len(elements),), 0.3, device=dev)).C
    nodes, elements = nodes.to(dev), elements.to(dev)
    is_con = con.ravel().to(torch.uint8).to(dev)
    disp = disp.ravel().to(dev)
    p = csr.Pattern(elements, nodes.shape[0], 3)
    k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C)
    vals = csr.assemble(p, k, is_con)
    vals_free = csr.assemble(p, k, None)
    del k
    A = p.matrix(vals)
    Af = p.matrix(vals_free)
    rhs = Af.matvec(disp * is_con)
    rhs.masked_fill_(is_con.bool(), 0.0)
    y = torch.empty_like(rhs)
    for fmt in ("csr", "sell-scalar", "sell"):
        for _ in range(3):
            A.matvec(rhs, out=y, fmt=fmt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            A.matvec(rhs, out=y, fmt=fmt)
        e1.record()
        torch.cuda.synchronize()
        print(f"spmv {fmt}: {e0.elapsed_time(e1) / 10:.4f} ms")
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            csr.krylov_solve(A, rhs, method=a.method, rtol=1e-8, maxiter=a.iters, check_every=a.iters)
        except RuntimeError as e:
            if rep == 0:
                print("expected:", e)
        e1.record()
        torch.cuda.synchronize()
        print(f"{a.method} per iteration ({a.iters} its incl. init): {e0.elapsed_time(e1) / a.iters:.4f} ms")
    print("done", p.nnz)


if __name__ == "__main__":
    main()
