"""Differentiable geometry contractions of the residual evaluation on kernels K9/K10.

The reference evaluates, for every Gauss point q inside `integrate_material` (src/torchfem/base.py:1050-1083),
`H_inc = du_e @ B[q]^T`, the material update (torch; stays torch here — the adjoint differentiates through it,
sparse.py:689-705) and `f += w_q * compute_f(detJ[q], B[q], P)`. The two geometric lines are linear maps fixed by
the mesh; `elem_grad` and `elem_force` run each for ALL Gauss points in one launch and are each other's
transpose, which is also their backward:

    H = elem_grad(u_e)              dH/du_e^T g  = elem_force(g, weighted=False)
    f = elem_force(P, weighted)     df/dP^T  g  = elem_grad(g, weighted)

They are used when the node coordinates are not being differentiated (shape optimisation keeps the torch path).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib as L


class Geometry:
    """What the kernels need of a model: reference tables (host), nodes / connectivity (device)."""

    def __init__(self, bref: Tensor, w: Tensor, nodes: Tensor, elements: Tensor, dpn: int):
        L.require_cuda(nodes, elements)
        self.bref = np.ascontiguousarray(bref.detach().cpu().numpy(), dtype=np.float64)
        self.w = np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float64)
        self.n_int, self.dim, self.nn = (int(s) for s in bref.shape)
        self.nodes = nodes.detach().to(torch.float64).contiguous()
        self.elements = elements.to(torch.int64).contiguous()
        self.n_elem = int(elements.shape[0])
        self.dpn = int(dpn)
        self.flag = torch.zeros(1, dtype=torch.int32, device=nodes.device)
        self.validated = False

    def check(self):
        """Raise the reference's error if a kernel met det J <= 0 (base.py:311-312). The Jacobians depend on the mesh
        only, so one clean evaluation validates this geometry for good: later calls skip the device read-back (a host
        synchronisation per residual evaluation otherwise)."""
        if self.validated:
            return
        if int(self.flag.item()) != 0:
            self.flag.zero_()
            raise ValueError("Negative Jacobian. Check element numbering.")
        self.validated = True


def _grad(g: Geometry, u_e: Tensor, weighted: bool) -> Tensor:
    u_e = u_e.to(torch.float64).contiguous()
    H = torch.empty(g.n_int, g.n_elem, g.dpn, g.dim, dtype=torch.float64, device=u_e.device)
    L.check(L.lib.tfem_elem_grad(g.dim, g.nn, g.n_int, g.dpn, g.bref.ctypes.data, g.w.ctypes.data, L.ptr(g.nodes),
                                 L.ptr(g.elements), g.n_elem, L.ptr(u_e), None, 1 if weighted else 0, L.ptr(H),
                                 L.ptr(g.flag), L.stream()))
    return H


def _force(g: Geometry, P: Tensor, weighted: bool) -> Tensor:
    P = P.to(torch.float64).contiguous()
    f = torch.empty(g.n_elem, g.nn, g.dpn, dtype=torch.float64, device=P.device)
    L.check(L.lib.tfem_elem_force(g.dim, g.nn, g.n_int, g.dpn, g.bref.ctypes.data, g.w.ctypes.data, L.ptr(g.nodes),
                                  L.ptr(g.elements), g.n_elem, L.ptr(P), None, 1 if weighted else 0, L.ptr(f),
                                  L.ptr(g.flag), L.stream()))
    return f


class _ElemGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_e: Tensor, g: Geometry, weighted: bool):
        ctx.g, ctx.weighted = g, weighted
        return _grad(g, u_e, weighted)

    @staticmethod
    def backward(ctx, gH: Tensor):
        return _ElemForce.apply(gH, ctx.g, ctx.weighted), None, None


class _ElemForce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P: Tensor, g: Geometry, weighted: bool):
        ctx.g, ctx.weighted = g, weighted
        return _force(g, P, weighted)

    @staticmethod
    def backward(ctx, gf: Tensor):
        return _ElemGrad.apply(gf, ctx.g, ctx.weighted), None, None


def elem_grad(g: Geometry, u_e: Tensor) -> Tensor:
    """[n_elem, nn, dpn] nodal values per element -> field gradient [n_int, n_elem, dpn, dim] at the Gauss points."""
    return _ElemGrad.apply(u_e, g, False)


def elem_force(g: Geometry, P: Tensor) -> Tensor:
    """[n_int, n_elem, dpn, dim] flux at the Gauss points -> sum_q w_q detJ_q B_q^T P_q, [n_elem, nn, dpn]."""
    return _ElemForce.apply(P, g, True)
