"""Diagnostics for AMG-PCG inside a Newton loop (Neo-Hookean block): every linear system is also solved by the numpy
oracle (fresh hierarchy) and by Jacobi-CG; prints curvature / eigenvalue information."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
torch.set_default_device("cuda")
import torchfem_b200 as T  # noqa: E402
from torchfem_b200 import sparse as S  # noqa: E402
from torchfem_b200.amg import AMGPreconditioner  # noqa: E402
from torchfem_b200.materials import Hyperelastic3D  # noqa: E402
from torchfem_b200.mesh import cube_hexa  # noqa: E402
from oracle import amg_oracle as M  # noqa: E402
from oracle import fem_oracle as O  # noqa: E402

En, NU = 1000.0, 0.3
LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
MU = En / (2.0 * (1.0 + NU))


def psi(F, params):
    Cg = F.transpose(-1, -2) @ F
    logJ = 0.5 * torch.logdet(Cg)
    return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2


nodes, elements = cube_hexa(17, 9, 9, 2.0, 1.0, 1.0)
box = T.Solid(nodes, elements, Hyperelastic3D(psi, torch.tensor([MU, LBD])))
left, right = nodes[:, 0] == 0.0, nodes[:, 0] == 2.0
box.constraints[left, :] = True
box.constraints[right, 0] = True
box.displacements[right, 0] = 0.4

real = S.sparse_solve
calls = [0]


def spy(A, b, B=None, stol=1e-10, device=None, method=None, M_=None, x0=None):
    calls[0] += 1
    c = calls[0]
    Ac = S._as_csr(A)
    n = Ac.n
    Asp = O.to_csr(Ac.values_.cpu().numpy(), Ac._indices().cpu().numpy(), n)
    bn = b.detach().cpu().numpy()
    sym = abs(Asp - Asp.T).max() / abs(Asp).max()
    dense = Asp.toarray()
    ev = np.linalg.eigvalsh(0.5 * (dense + dense.T))
    print(f"[solve {c}] n={n} |b|={np.linalg.norm(bn):.3e} asym={sym:.2e} eig min={ev[0]:.3e} max={ev[-1]:.3e} "
          f"diag min={Asp.diagonal().min():.3e} reuse_M={isinstance(M_, AMGPreconditioner)}", flush=True)
    lv = M.build_hierarchy(Asp, 3)
    xo, info_o, its_o = M.amg_pcg(Asp, bn, lv, rtol=stol, maxiter=200)
    print(f"   oracle: levels {[L.n for L in lv]} info={info_o} its={its_o} rho={[getattr(L,'rho',None) for L in lv]}", flush=True)
    try:
        xj, _ = real(A, b, B, stol, device, "cg", None, x0)
        print(f"   jacobi-cg ok, |x|={float(xj.norm()):.6e}")
    except RuntimeError as e:
        print("   jacobi-cg failed:", e)
    for tag, Mx in (("fresh", None), ("given", M_)):
        try:
            x, Mo = real(A, b, B, stol, device, "amgx", Mx, x0)
            st = Mo.solve(b.detach().to(torch.float64), rtol=stol)[1]
            print(f"   amgx[{tag}] ok its={st['iterations']} levels={[l.n for l in Mo.levels]} rho={[getattr(l,'rho',None) for l in Mo.levels[:-1]]} "
                  f"|x-xo|/|xo|={np.linalg.norm(x.cpu().numpy()-xo)/max(np.linalg.norm(xo),1e-300):.2e}", flush=True)
        except RuntimeError as e:
            Mo = AMGPreconditioner(Ac) if Mx is None else Mx
            z = Mo.apply(b.detach().to(torch.float64))
            zo = M.vcycle(lv, bn)
            print(f"   amgx[{tag}] FAILED: {e}; vcycle finite={bool(torch.isfinite(z).all())} "
                  f"b.z={float((z*b).sum()):.3e} oracle b.z={bn@zo:.3e} "
                  f"dinv finite={[bool(torch.isfinite(l.dinv).all()) for l in Mo.levels]} "
                  f"inv finite={bool(torch.isfinite(Mo.levels[-1].inv).all())}", flush=True)
    return real(A, b, B, stol, device, "cg", None, x0)[0], M_


S.sparse_solve = spy
try:
    u, f, *_ = box.solve(increments=torch.tensor([0.0, 0.5, 1.0]), nlgeom=True, method="amgx", stol=1e-11)
    print("solve finished", float(u.abs().max()))
except RuntimeError as e:
    print("solve failed:", e)
