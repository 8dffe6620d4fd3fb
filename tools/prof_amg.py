"""Short driver for ncu captures of the AMG kernels: one hierarchy build and two AMG-PCG iterations at config B.
    python tools/prof_amg.py --edge 150
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
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200 import csr
    from torchfem_b200.amg import AMGPreconditioner
    from torchfem_b200.elements import Hexa1
    from torchfem_b200.materials import IsotropicElasticity3D

    dev = torch.device("cuda", 0)
    nodes, elements, con, disp = bench.build_problem(T, torch, a.edge, dev)
    bref, w = bench.element_tables(Hexa1)
    C = IsotropicElasticity3D(torch.full((len(elements),), 1000.0, device=dev), torch.full((len(elements),), 0.3, device=dev)).C
    nodes, elements = nodes.to(dev), elements.to(dev)
    is_con = con.ravel().to(torch.uint8).to(dev)
    ubc = (disp.ravel() * con.ravel()).to(dev)
    p = csr.Pattern(elements, nodes.shape[0], 3)
    k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C)
    b = torch.empty(p.n_dofs, device=dev)
    vals = csr.assemble(p, k, is_con, ubc=ubc, lift=b)
    del k
    A = p.matrix(vals)
    amg = AMGPreconditioner(A)
    try:
        amg.solve(b, rtol=1e-8, maxiter=2)
    except RuntimeError as e:
        print("expected:", e)
    print("done", [lv.n for lv in amg.levels])


if __name__ == "__main__":
    main()
