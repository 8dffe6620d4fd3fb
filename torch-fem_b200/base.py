"""FEM driver of the hot path: DOF map and sparsity pattern, Gauss-point integration, global assembly and
the incremental Newton solve — the drop-in for the reference's `torchfem.base` (src/torchfem/base.py).

Same model API (`FEM`, `Mechanics`, `Heat`; attributes `forces / displacements / constraints /
ext_strain / heat_flux / temperatures`, `idx`, `k_map`, `diag_map`, `glob_idx`, `K`; methods `k0`,
`integrate_material`, `assemble_matrix`, `assemble_rhs`, `eval_shape_functions`, `compute_B`, `solve` with
every keyword of reference base.py:597-616). What runs underneath is different:

* `__init__` builds the pattern with kernel K0 (node graph, `csr.Pattern`) instead of sorting packed keys
  (base.py:78-118); `k_map` / `glob_idx` are materialised lazily, only if somebody reads them.
* `integrate_material` computes the element tangent matrices with kernel K1 (`csr.integrate_k`) from the
  material's `ddsdde`; the residual side (`H_inc`, `Material.step`, `compute_f`) stays in torch because the
  adjoint differentiates through it (reference sparse.py:689-705). Shape gradients are cached per model.
* `assemble_matrix` runs the deterministic gather kernel K2/K3 and returns a device `CSRMatrix`
  (not a COO tensor); `sparse.sparse_solve` consumes it without any format conversion.

Everything lives on the CUDA device; CPU inputs are moved there once in the constructor.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from collections.abc import Iterable
from functools import cached_property

import torch
from torch import Tensor

from . import _lib as L
from . import csr as _csr
from . import residual as _res
from .elements import Element
from .materials import Material
from .materials import small_det as _small_det
from .materials import small_matmul as _small_matmul
from .sparse import CachedSolve, describe_method, differentiable_sparse_solve, newton_solve  # noqa: F401


def _cuda_device(t: Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("torch-fem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def _step_points(material, H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent):
    """Material update at all Gauss points: the material's own batched `step_points` if it has one, else the
    reference's per-point loop over `step` (any duck-typed material works)."""
    fn = getattr(material, "step_points", None)
    if fn is not None:
        return fn(H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent=need_tangent)
    from .materials import step_points_loop

    return step_points_loop(material, H_all, F_all, stress_all, state_all, de0, cl, iter)


def _det_inv(J: Tensor) -> tuple[Tensor, Tensor]:
    """Closed-form determinant and inverse of batched 2x2 / 3x3 matrices (differentiable)."""
    d = J.shape[-1]
    if d == 1:
        det = J[..., 0, 0]
        return det, 1.0 / J
    if d == 2:
        a, b, c, e = J[..., 0, 0], J[..., 0, 1], J[..., 1, 0], J[..., 1, 1]
        det = a * e - b * c
        inv = torch.stack([torch.stack([e, -b], -1), torch.stack([-c, a], -1)], -2) / det[..., None, None]
        return det, inv
    c00 = J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1]
    c01 = J[..., 1, 2] * J[..., 2, 0] - J[..., 1, 0] * J[..., 2, 2]
    c02 = J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0]
    det = J[..., 0, 0] * c00 + J[..., 0, 1] * c01 + J[..., 0, 2] * c02
    r0 = torch.stack([c00, J[..., 0, 2] * J[..., 2, 1] - J[..., 0, 1] * J[..., 2, 2],
                      J[..., 0, 1] * J[..., 1, 2] - J[..., 0, 2] * J[..., 1, 1]], -1)
    r1 = torch.stack([c01, J[..., 0, 0] * J[..., 2, 2] - J[..., 0, 2] * J[..., 2, 0],
                      J[..., 0, 2] * J[..., 1, 0] - J[..., 0, 0] * J[..., 1, 2]], -1)
    r2 = torch.stack([c02, J[..., 0, 1] * J[..., 2, 0] - J[..., 0, 0] * J[..., 2, 1],
                      J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]], -1)
    inv = torch.stack([r0, r1, r2], -2) / det[..., None, None]
    return det, inv


_SOLVER_ORDER_MIN_DOFS = 10_000   # = sparse.DIRECT_LIMIT: below it `sparse_solve` factorises and no SELL copy is read


class _IntegrateK(torch.autograd.Function):
    """k_e = sum_q w_q detJ_q s_e B_q^T C B_q (kernel K1) with the adjoint contraction for the tangent:
    dL/dC[e,i,J,k,L] = sum_q w detJ s sum_{p,r} B_q[J,p] B_q[L,r] dL/dk[e,(p,i),(r,k)]  (mechanics, base.py:1088),
    dL/dkappa[e,i,j] = sum_q w detJ s sum_{p,r} B_q[i,p] B_q[j,r] dL/dk[e,p,r]            (heat, base.py:1274).
    The backward runs in torch: it is only reached by eigenvalue sensitivities, never by the solve path."""

    @staticmethod
    def forward(ctx, tangent: Tensor, model: "FEM"):
        ctx.model = model
        ctx.per_gp = tangent.dim() == (6 if model.KIND == L.KIND_MECH else 4)
        return model._integrate_k_raw(tangent)

    @staticmethod
    def backward(ctx, gk: Tensor):
        model = ctx.model
        _, B, detJ = model._ip_shape()
        w = model.etype.iweights.to(device=gk.device, dtype=gk.dtype)
        scale = model._k_scale
        nn, d = model.etype.nodes, model.n_dof_per_node
        out = []
        for q in range(B.shape[0]):
            f = w[q] * detJ[q] * (scale if scale is not None else 1.0)
            if model.KIND == L.KIND_MECH:
                g = gk.reshape(model.n_elem, nn, d, nn, d)
                out.append(torch.einsum("e,eJp,epirk,eLr->eiJkL", f, B[q], g, B[q]))
            else:
                out.append(torch.einsum("e,eip,epr,ejr->eij", f, B[q], gk, B[q]))
        grad = torch.stack(out) if ctx.per_gp else torch.stack(out).sum(0)
        return grad, None


class _ElementSystemSolve(torch.autograd.Function):
    """x = A^-1 b for A = c_m M(m_e) + c_k K(k_e) assembled by the kernels, with gradients to b AND to the element
    matrices:  dL/dk_e[a,b] = -c_k lam[idx[e,a]] x[idx[e,b]],  dL/dm_e likewise with c_m  (lam = A^-T dL/dx).
    This is what the reference obtains by back-propagating `Solve.backward`'s sparse gradA (sparse.py:212-216) through
    the `index_add_` of `assemble_matrix` (base.py:407-419); entries of constrained rows / columns carry no gradient
    there either (they are overwritten by constants, base.py:414-419), hence the masks."""

    @staticmethod
    def forward(ctx, b, k_e, m_e, A, idx, free, c_k, c_m, cfg):
        from .sparse import sparse_solve

        B, stol, device, method, cached, update_cache = cfg
        x0 = cached.previous_x if cached is not None else None
        x, M = sparse_solve(A, b, B, stol, device, method, None, x0)
        if update_cache and cached is not None:
            cached.update_x(x)
        ctx.save_for_backward(x)
        ctx.A, ctx.idx, ctx.free, ctx.c, ctx.cfg, ctx.M = A, idx, free, (c_k, c_m), cfg, M
        ctx.needs = (k_e is not None and k_e.requires_grad, m_e is not None and m_e.requires_grad)
        return x

    @staticmethod
    def backward(ctx, g):
        from .sparse import sparse_solve

        (x,) = ctx.saved_tensors
        B, stol, device, method, cached, update_cache = ctx.cfg
        x0 = cached.previous_grad if cached is not None else None
        lam, _ = sparse_solve(ctx.A.T, g, B, stol, device, method, ctx.M, x0)
        if update_cache and cached is not None:
            cached.update_grad(lam)
        gk = gm = None
        if any(ctx.needs):
            idx = ctx.idx.long()
            outer = -(lam * ctx.free)[idx][:, :, None] * (x * ctx.free)[idx][:, None, :]
            if ctx.needs[0]:
                gk = ctx.c[0] * outer
            if ctx.needs[1]:
                gm = ctx.c[1] * outer
        return lam, gk, gm, None, None, None, None, None, None


class _ElementMatvec(torch.autograd.Function):
    """y = M(m_e) x on the kernel-assembled CSR, differentiable w.r.t. x and the element matrices
    (dL/dm_e[a,b] = g[idx[e,a]] x[idx[e,b]] off the constrained rows / columns) — the reference's `self.M @ du` on a
    differentiable sparse tensor (base.py:1483)."""

    @staticmethod
    def forward(ctx, x, m_e, A, idx, free):
        ctx.save_for_backward(x)
        ctx.A, ctx.idx, ctx.free = A, idx, free
        ctx.need_m = m_e.requires_grad
        return A.matvec(x.detach().to(torch.float64).contiguous())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gx = ctx.A.T.matvec(g.contiguous())
        gm = None
        if ctx.need_m:
            idx = ctx.idx.long()
            gm = (g * ctx.free)[idx][:, :, None] * (x * ctx.free)[idx][:, None, :]
        return gx, gm, None, None, None


class FEM(ABC):
    """Abstract finite-element model (reference base.py:23-132)."""

    KIND = L.KIND_MECH

    def __init__(self, nodes: Tensor, elements: Tensor, material: Material | None):
        dev = _cuda_device(nodes)
        self.device = dev
        self.nodes = nodes.to(dev)
        self.elements = elements.to(dev)
        self.n_nod, self.n_dim = self.nodes.shape
        self.n_dofs = self.n_dof_per_node * self.n_nod
        self.n_elem = len(self.elements)
        self.n_int = len(self.etype.iweights)
        dt = self.nodes.dtype
        dpn = self.n_dof_per_node

        self._neumann = torch.zeros(self.n_nod, dpn, dtype=dt, device=dev)
        self._dirichlet = torch.zeros(self.n_nod, dpn, dtype=dt, device=dev)
        self._constraints = torch.zeros(self.n_nod, dpn, dtype=torch.bool, device=dev)
        self._external_gradient = torch.zeros(self.n_elem, *self.n_flux, dtype=dt, device=dev)

        # local -> global DOF map (base.py:73-76,119), int32 like the reference
        dofs = dpn * self.elements.unsqueeze(-1) + torch.arange(dpn, device=dev)
        self.idx = dofs.reshape(self.n_elem, -1).to(torch.int32)

        # sparsity pattern + assembly permutation (kernel K0; replaces base.py:78-118)
        self.pattern = _csr.Pattern(self.elements, self.n_nod, dpn)

        self.material: Material | None
        if material is None:
            self.material = None
        else:
            m = material if material.is_vectorized else material.vectorize(self.n_elem)
            self.material = m.to(dev)
        self.cached_solve = CachedSolve()
        self.K = torch.empty(0, device=dev)
        self._shape_cache = None

    # ---- reference-compatible integer structure (bit-identical, materialised on demand)
    @property
    def k_map(self) -> Tensor:
        return self.pattern.k_map

    @property
    def diag_map(self) -> Tensor:
        return self.pattern.diag_map

    @property
    def glob_idx(self) -> Tensor:
        return self.pattern.glob_idx

    @property
    def n_state(self) -> int:
        assert self.material is not None
        return self.material.n_state

    @property
    def volume_scale(self) -> Tensor:
        return torch.ones(self.n_elem, dtype=self.nodes.dtype, device=self.device)

    @property
    def _k_scale(self) -> Tensor | None:
        """Per-element factor of compute_k beyond detJ (thickness), handed to kernel K1."""
        return None

    @property
    @abstractmethod
    def n_flux(self) -> list[int]:
        ...

    @property
    @abstractmethod
    def n_dof_per_node(self) -> int:
        ...

    @property
    @abstractmethod
    def etype(self) -> type[Element]:
        ...

    @property
    @abstractmethod
    def initial_grad(self) -> Tensor:
        ...

    @abstractmethod
    def compute_k(self, detJ: Tensor, BCB: Tensor) -> Tensor:
        ...

    @abstractmethod
    def compute_f(self, detJ: Tensor, B: Tensor, S: Tensor) -> Tensor:
        ...

    def compute_m(self, detJ: Tensor, rho: Tensor) -> Tensor:
        raise NotImplementedError

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is out of scope of torch-fem_b200 (SURVEY §2, row 19)")

    # ---- boundary conditions
    @property
    def constraints(self) -> Tensor:
        return self._constraints

    @constraints.setter
    def constraints(self, value: Tensor):
        if value.shape != (self.n_nod, self.n_dof_per_node):
            raise ValueError("Constraints must have the same shape as nodes.")
        if value.dtype != torch.bool:
            raise TypeError("Constraints must be a boolean tensor.")
        self._constraints = value.to(self.device)

    def _set_field(self, name: str, label: str, value: Tensor, shape_msg: str):
        if value.shape != (self.n_nod, self.n_dof_per_node):
            raise ValueError(shape_msg)
        if not torch.is_floating_point(value):
            raise TypeError(f"{label} must be a floating-point tensor.")
        setattr(self, name, value.to(self.device))

    # ---- shape functions
    def eval_shape_functions(self, xi: Tensor) -> tuple[Tensor, Tensor, Tensor]:
        """N, B = J^-1 dN/dxi and detJ at arbitrary reference points (reference base.py:293-314).
        Differentiable w.r.t. `self.nodes`; raises the reference's ValueError on detJ <= 0."""
        X = self.nodes[self.elements]
        xi = xi.to(device=self.device, dtype=self.nodes.dtype)
        b = self.etype.B(xi)
        J = torch.einsum("...iN,ANj->...Aij", b, X)
        detJ, Jinv = _det_inv(J)
        if torch.any(detJ <= 0.0):
            raise ValueError("Negative Jacobian. Check element numbering.")
        B = torch.einsum("...Eij,...jN->...EiN", Jinv, b)
        return self.etype.N(xi), B, detJ

    def _ip_shape(self) -> tuple[Tensor, Tensor, Tensor]:
        """Shape data at the integration points, cached while `nodes` is the same non-differentiable
        tensor (the reference recomputes it on every `integrate_material` call, base.py:1048)."""
        if self.nodes.requires_grad:
            return self.eval_shape_functions(self.etype.ipoints)
        if not self._cache_valid(self._shape_cache):
            self._shape_cache = (self._cache_key(), self.eval_shape_functions(self.etype.ipoints))
        return self._shape_cache[1]

    def _cache_key(self):
        # the tensors themselves (kept alive by the entry, so their addresses cannot be recycled) + the in-place version
        return (self.nodes, self.nodes._version, self.elements, self.elements._version)

    def _cache_valid(self, entry) -> bool:
        if entry is None:
            return False
        nodes, nv, elements, ev = entry[0]
        return nodes is self.nodes and elements is self.elements and nv == nodes._version and ev == elements._version

    def _geometry(self) -> "_res.Geometry | None":
        """Kernel-side geometry of the residual contractions (K9/K10), or None when the node coordinates are
        being differentiated (shape optimisation keeps the torch path)."""
        if self.nodes.requires_grad or self.nodes.dtype != torch.float64:
            return None
        if not self._cache_valid(getattr(self, "_geom_cache", None)):
            bref, w = self._tables()
            self._geom_cache = (self._cache_key(), _res.Geometry(bref, w, self.nodes, self.elements, self.n_dof_per_node))
        return self._geom_cache[1]

    @property
    def _f_scale(self) -> Tensor | None:
        """Per-element factor of compute_f beyond w detJ (thickness; differentiable), applied in torch."""
        return None

    def _tables(self) -> tuple[Tensor, Tensor]:
        ip = self.etype.ipoints.to(torch.float64)
        return self.etype.B(ip).cpu(), self.etype.iweights.to(torch.float64).cpu()

    def _integrate_k(self, tangent: Tensor) -> Tensor:
        """Element matrices from the material tangent with kernel K1. `tangent` is one tensor per
        element or a stack over Gauss points. When the tangent carries a graph (differentiable material
        parameters) the result does too: k is linear in the tangent and `_IntegrateK.backward` applies the transposed
        contraction (needed by `solve_modes`, whose eigenvalue sensitivities flow through k; `FEM.solve` detaches K)."""
        if torch.is_grad_enabled() and tangent.requires_grad and not self.nodes.requires_grad:
            return _IntegrateK.apply(tangent, self)
        return self._integrate_k_raw(tangent)

    def _integrate_k_raw(self, tangent: Tensor) -> Tensor:
        bref, w = self._tables()
        # det J depends on the mesh only: once a kernel has seen this geometry with all Jacobians positive, K1's flag
        # need not be read back again (a host synchronisation per Newton iteration otherwise)
        ok = getattr(self, "_k1_validated", None)
        check = not self._cache_valid(ok)
        k = _csr.integrate_k(self.KIND, bref, w, self.nodes.detach(), self.elements,
                             tangent.detach().to(torch.float64), self._k_scale, check=check)
        if check and not self.nodes.requires_grad:
            self._k1_validated = (self._cache_key(),)
        return k

    def compute_B(self) -> Tensor:
        """Rigid-body modes (near null space for AMG back ends; unused by Jacobi but part of the
        `sparse_solve` signature, reference base.py:316-344)."""
        d, n = self.n_dof_per_node, self.n_nod
        kw = dict(dtype=self.nodes.dtype, device=self.device)
        if d == 3:
            B = torch.zeros(3 * n, 6, **kw)
            x, y, z = self.nodes.detach().unbind(1)
            for i in range(3):
                B[i::3, i] = 1
            B[1::3, 3], B[2::3, 3] = -z, y
            B[0::3, 4], B[2::3, 4] = z, -x
            B[0::3, 5], B[1::3, 5] = -y, x
        elif d == 2:
            B = torch.zeros(2 * n, 3, **kw)
            x, y = self.nodes.detach().unbind(1)
            B[0::2, 0] = 1
            B[1::2, 1] = 1
            B[1::2, 2], B[0::2, 2] = -x, y
        else:
            B = torch.ones(d * n, 1, **kw)
        return B

    def integrate_shape_functions(self) -> Tensor:
        N, _, detJ = self._ip_shape()
        w = self.etype.iweights.to(device=self.device, dtype=self.nodes.dtype)
        return torch.einsum("i,in,ie->en", w, N, detJ)

    def integrate_field(self, field: Tensor | None = None) -> Tensor:
        """Integral of a nodal scalar field (or of 1) over each element (reference base.py:353-374)."""
        wN = self.integrate_shape_functions()
        if field is None:
            return wN.sum(dim=1)
        return (wN * field.to(self.device)[self.elements]).sum(dim=1)

    def integrate_mass(self) -> Tensor:
        """Consistent element mass matrices (reference base.py:376-396)."""
        assert self.material is not None
        nn, dpn = self.etype.nodes, self.n_dof_per_node
        N, _, detJ = self._ip_shape()
        eye = torch.eye(dpn, dtype=self.nodes.dtype, device=self.device)
        m = torch.zeros(self.n_elem, nn * dpn, nn * dpn, dtype=self.nodes.dtype, device=self.device)
        for q, w in enumerate(self.etype.iweights):
            dens = self.compute_m(detJ[q], self.material.rho)
            blk = torch.einsum("N,M,E,ij->ENiMj", N[q], N[q], dens * torch.ones_like(detJ[q]), eye)
            m += float(w) * blk.reshape(self.n_elem, nn * dpn, nn * dpn)
        return m

    # ---- consistent nodal loads (reference base.py:446-568); host-side torch, differentiable w.r.t. the load
    def _nodal_sum(self, conn: Tensor, contrib: Tensor) -> Tensor:
        """Sum per-node contributions [n, nodes per entity, k] of elements or facets into [n_nod, k]."""
        out = torch.zeros(self.n_nod, contrib.shape[-1], dtype=contrib.dtype, device=self.device)
        return out.index_add_(0, conn.reshape(-1), contrib.reshape(-1, contrib.shape[-1]))

    def _boundary_facets(self, mask: Tensor) -> Tensor:
        """Facets whose nodes all lie in the nodal mask and that belong to one element only (an interior facet is
        listed by both neighbours), in the winding of the element that owns them (reference base.py:452-470)."""
        table = self.etype.facets.to(self.device)
        facets = self.elements[:, table].reshape(-1, table.shape[1])
        facets = facets[mask.to(self.device)[facets].all(dim=1)]
        _, inverse, count = torch.unique(facets.sort(dim=1).values, dim=0, return_inverse=True, return_counts=True)
        order = torch.argsort(inverse, stable=True)             # occurrences grouped by facet, first occurrence first
        start = torch.cumsum(count, 0) - count
        return facets[order[start[count == 1]]]

    def _integrate_facet_load(self, conn: Tensor, ftype: type[Element], load: Tensor) -> Tensor:
        kw = dict(dtype=self.nodes.dtype, device=self.device)
        xi = ftype.ipoints.to(**kw)
        w, N = ftype.iweights.to(**kw), ftype.N(xi)
        J = torch.einsum("qaN,eNj->qeaj", ftype.B(xi), self.nodes[conn])      # tangent vectors of the facet
        # outward normal scaled by the facet metric: t_1 x t_2 on a face, the tangent turned clockwise on an edge
        if J.shape[-2] == 2:
            normal = torch.linalg.cross(J[..., 0, :], J[..., 1, :], dim=-1)
        else:
            normal = torch.stack([J[..., 0, 1], -J[..., 0, 0]], dim=-1)
        metric = torch.linalg.norm(normal, dim=-1)
        if load.dim() == 0 and self.n_dof_per_node > 1:
            # a scalar on a vector field is a pressure along the outward normal
            value = load * normal / metric[..., None]
        else:
            per_facet = load if load.dim() == 2 else load.reshape(1, -1)
            value = per_facet.expand(len(conn), -1).expand(len(xi), -1, -1)
        return self._nodal_sum(conn, torch.einsum("q,qn,qe,qek->enk", w, N, metric, value))

    def integrate_body_load(self, load: float | Tensor) -> Tensor:
        """Consistent nodal loads [n_nod, k] of a load per unit volume: a vector [k] for the whole model or one per
        element [n_elem, k] (gravity, heat source); planar models include their thickness (base.py:496-513)."""
        load = torch.as_tensor(load, dtype=self.nodes.dtype, device=self.device)
        per_elem = load if load.dim() == 2 else load.reshape(1, -1)
        weights = self.integrate_shape_functions() * self.volume_scale[:, None]
        return self._nodal_sum(self.elements, weights[:, :, None] * per_elem.expand(self.n_elem, -1)[:, None, :])

    def integrate_surface_load(self, mask: Tensor, load: float | Tensor) -> Tensor:
        """Consistent nodal loads of a load per unit area on the boundary faces inside a nodal mask: a scalar is a
        pressure along the outward normal (a flux for scalar fields), a vector a traction in global coordinates, a
        [n_facets, k] tensor one value per face (base.py:515-539)."""
        if self.etype.iso_dim != 3:
            raise NotImplementedError(f"{type(self).__name__} has no surfaces to load. "
                                      "Use integrate_line_load(...) or integrate_body_load(...) instead.")
        load = torch.as_tensor(load, dtype=self.nodes.dtype, device=self.device)
        return self._integrate_facet_load(self._boundary_facets(mask), self.etype.facet_type, load)

    def integrate_line_load(self, mask: Tensor, load: float | Tensor) -> Tensor:
        """The same per unit length on the boundary edges of a planar model (base.py:541-568)."""
        if self.etype.iso_dim != 2:
            raise NotImplementedError(f"{type(self).__name__} has no edges to load. "
                                      "Use integrate_surface_load(...) or integrate_body_load(...) instead.")
        load = torch.as_tensor(load, dtype=self.nodes.dtype, device=self.device)
        if load.dim() == 0 and self.n_dim == 3:
            raise ValueError("A line in 3D has no unique normal, so a scalar load is ambiguous. "
                             "Pass a load vector instead.")
        return self._integrate_facet_load(self._boundary_facets(mask), self.etype.facet_type, load)

    # ---- element matrices / assembly
    def k0(self) -> Tensor:
        """Element matrix of the reference state, [n_elem, nd, nd] (reference base.py:228-245)."""
        kw = dict(dtype=self.nodes.dtype, device=self.device)
        u = torch.zeros(self.n_nod, self.n_dof_per_node, **kw)
        grad = torch.zeros(self.n_int, self.n_elem, *self.n_flux, **kw)
        grad[:] = self.initial_grad.to(**kw)
        flux = torch.zeros(self.n_int, self.n_elem, *self.n_flux, **kw)
        state = torch.zeros(self.n_int, self.n_elem, self.n_state, **kw)
        de0 = torch.zeros(self.n_elem, *self.n_flux, **kw)
        self.K = torch.empty(0, device=self.device)
        k, *_ = self.integrate_material(u, grad, flux, state, u.clone(), de0, 0, False)
        assert k is not None
        return k

    @abstractmethod
    def integrate_material(self, u_prev, grad_prev, flux_prev, state_prev, du, de0, iter, nlgeom,
                           compute_stiffness: bool = True):
        ...

    def assemble_matrix(self, k: Tensor, con: Tensor) -> _csr.CSRMatrix:
        """Global tangent with Dirichlet rows/columns zeroed and unit diagonal (reference
        base.py:398-426), assembled by the deterministic kernel K2/K3. Returns a device CSR matrix
        whose `_values()` / `_indices()` are the reference's COO values / `glob_idx`."""
        is_con = torch.zeros(self.n_dofs, dtype=torch.bool, device=self.device)
        is_con[con.to(self.device)] = True
        self.is_constrained = is_con
        symmetric = bool(getattr(self.material, "symmetric_tangent", True))
        mask = is_con.to(torch.uint8) if con.numel() else None
        if self.n_dofs >= _SOLVER_ORDER_MIN_DOFS and self.pattern.sell_structure.long_rows is None:
            # systems the iterative solvers take: the same pass also writes the values in their SELL-32 order
            vals, sell_vals, _ = _csr.assemble(self.pattern, k.detach().to(torch.float64), mask, sell_out=True)
            return self.pattern.matrix(vals, symmetric=symmetric, sell_vals=sell_vals)
        vals = _csr.assemble(self.pattern, k.detach().to(torch.float64), mask)
        return self.pattern.matrix(vals, symmetric=symmetric)

    def assemble_rhs(self, f: Tensor) -> Tensor:
        """Global vector from element vectors; differentiable (reference base.py:428-445). A deterministic gather
        over the pattern's incidence lists (kernel `tfem_assemble_rhs`) instead of `index_add_`'s floating-point
        atomics: two evaluations of a residual agree bit for bit, like on the reference's CPU path."""
        return _csr.assemble_rhs(self.pattern, f)

    # ---- incremental Newton solve
    def solve(self, increments: Tensor | None = None, max_iter: int = 10, rtol: float = 1e-8,
              atol: float = 1e-6, stol: float = 1e-10, cutback_factor: float = 0.5,
              growth_factor: float = 1.1, max_cutbacks: int = 10, verbose: bool = False,
              method: str | None = None, device: str | None = None, return_intermediate: bool = False,
              aggregate_integration_points: bool = True, use_cached_solve: bool = False,
              nlgeom: bool = False, alpha: float = 0.0,
              differentiable_parameters: Tensor | Iterable[Tensor] | None = None):
        """Quasi-static solve by load increments with automatic cutback (reference base.py:597-923).
        Returns (u, f, flux, grad, state) at the last increment, or stacked over increments when
        `return_intermediate`."""
        kw = dict(dtype=self.nodes.dtype, device=self.device)
        increments = torch.tensor([0.0, 1.0]) if increments is None else increments
        inc = [float(v) for v in increments]
        N = len(inc)
        dpn = self.n_dof_per_node

        m = self.integrate_mass() if alpha > 0.0 else None
        self.stabilization_energy = torch.zeros(N, **kw)

        if differentiable_parameters is None:
            params: tuple = ()
        elif isinstance(differentiable_parameters, Tensor):
            params = (differentiable_parameters,)
        else:
            params = tuple(differentiable_parameters)
        track = any(p.requires_grad for p in params)

        B = self.compute_B()
        con = torch.nonzero(self.constraints.ravel(), as_tuple=False).ravel()

        u = torch.zeros(N, self.n_nod, dpn, **kw)
        f = torch.zeros(N, self.n_nod, dpn, **kw)
        flux = torch.zeros(N, self.n_int, self.n_elem, *self.n_flux, **kw)
        grad = torch.zeros(N, self.n_int, self.n_elem, *self.n_flux, **kw)
        grad[:] = self.initial_grad.to(**kw)
        state = torch.zeros(N, self.n_int, self.n_elem, self.n_state, **kw)

        if verbose:
            print(f"torch-fem_b200 | solve | {type(self).__name__} | {self.n_elem:,} elem | {self.n_dofs:,} dof"
                  f" | {describe_method(self.n_dofs, 'cuda', method)}")

        self.K = torch.empty(0, device=self.device)
        du = torch.zeros(self.n_dofs, **kw)

        def make_eval_residual(F_ext, DU, de0, k_visc):
            # loads are bound per substep so that the adjoint replay sees the right ones (base.py:704-743)
            def eval_residual(du, i, u_prev, grad_prev, flux_prev, state_prev):
                du_bc = du.clone()
                du_bc[con] = DU[con]
                k, f_i, _, _, _ = self.integrate_material(u_prev, grad_prev, flux_prev, state_prev, du_bc,
                                                          de0, i, nlgeom)
                if k_visc is not None:
                    du_e = du_bc.view(-1, dpn)[self.elements].flatten(1)
                    f_i = f_i + torch.einsum("...ij,...j->...i", k_visc, du_e)
                    if k is not None:
                        k = k + k_visc
                if k is not None:
                    self.K = self.assemble_matrix(k, con)
                res = self.assemble_rhs(f_i) - F_ext
                res[con] = 0.0
                return res, self.K

            return eval_residual

        u_cur, f_cur = u[0].clone(), f[0].clone()
        grad_cur, flux_cur, state_cur = grad[0].clone(), flux[0].clone(), state[0].clone()
        energy = torch.zeros((), **kw)
        lam = inc[0]
        step_frac = 1.0
        k_step = 0.0

        for n in range(1, N):
            target = inc[n]
            span = target - lam
            direction = math.copysign(1.0, span)
            step_size = step_frac * abs(span)
            min_step = abs(span) * cutback_factor ** max_cutbacks
            while abs(target - lam) > 1e-12 * max(1.0, abs(target)):
                step = direction * min(step_size, abs(target - lam))
                F_ext = (lam + step) * self._neumann.ravel()
                DU = step * self._dirichlet.ravel()
                de0 = step * self._external_gradient

                k_visc = None
                if m is not None:
                    k_visc = alpha / abs(step) * m
                    if abs(step) != k_step:
                        self.K = torch.empty(0, device=self.device)
                    k_step = abs(step)

                if track:
                    prev = (u_cur.clone(), grad_cur.clone(), flux_cur.clone(), state_cur.clone())
                else:
                    prev = (u_cur.detach(), grad_cur.detach(), flux_cur.detach(), state_cur.detach())
                cached = self.cached_solve if use_cached_solve else CachedSolve()
                try:
                    du = newton_solve(make_eval_residual(F_ext, DU, de0, k_visc), du.detach(), B, max_iter,
                                      rtol, atol, stol, None, method, device, cached, use_cached_solve,
                                      *prev, *params)
                except RuntimeError as err:
                    step_size = cutback_factor * abs(step)
                    if step_size < min_step:
                        raise RuntimeError(f"Newton-Raphson did not converge in increment {n} "
                                           f"after {max_cutbacks} cutbacks.") from err
                    if verbose:
                        print(f"  increment {n}: cutback to step {step_size:.3e}")
                    continue

                # converged state (no tangent needed)
                du_eval = du.clone()
                du_eval[con] = DU[con]
                _, f_i, grad_cur, flux_cur, state_cur = self.integrate_material(
                    u_cur, grad_cur, flux_cur, state_cur, du_eval, de0, max_iter, nlgeom,
                    compute_stiffness=False)
                F_int = self.assemble_rhs(f_i)
                if k_visc is not None:
                    du_e = du_eval.view(-1, dpn)[self.elements]
                    F_v = self.assemble_rhs(torch.einsum("...ij,...j->...i", k_visc, du_e.flatten(1)))
                    F_int = F_int + F_v
                    energy = energy + torch.dot(du_eval, F_v).detach()
                f_cur = F_int.reshape(-1, dpn)
                u_cur = u_cur + du_eval.reshape(-1, dpn)
                du = du_eval
                lam += step
                step_size = min(growth_factor * step_size, abs(span))
            step_frac = min(step_size / abs(span), 1.0) if span != 0.0 else 1.0
            u[n], f[n], grad[n], flux[n], state[n] = u_cur, f_cur, grad_cur, flux_cur, state_cur
            self.stabilization_energy[n] = energy
            if verbose:
                print(f"  increment {n}/{N - 1} done (load factor {target:g})")

        out = [u, f, flux, grad, state]
        if aggregate_integration_points:
            out[2], out[3], out[4] = out[2].mean(dim=1), out[3].mean(dim=1), out[4].mean(dim=1)
        out[2] = out[2].squeeze((-2, -1))
        out[3] = out[3].squeeze((-2, -1))
        if not track:
            out = [t.detach() for t in out]
        if return_intermediate:
            return tuple(out)
        return tuple(t[-1] for t in out)


class Mechanics(FEM, ABC):
    """Vector-valued (displacement) problems (reference base.py:926-1129)."""

    KIND = L.KIND_MECH

    @property
    def n_dof_per_node(self) -> int:
        return self.nodes.shape[1]

    @property
    def initial_grad(self) -> Tensor:
        return torch.eye(self.n_flux[0])

    def solve_modes(self, n_modes: int) -> tuple[Tensor, Tensor]:
        """Natural frequencies and mode shapes, K phi = omega^2 M phi (reference base.py:1097-1129). Returns
        `(omega_sq [n_modes], differentiable w.r.t. material parameters; modes [n_modes, n_nod, n_dof_per_node],
        detached)`. Solved on the free-DOF subspace: dense `eigh` for small models, AMG-preconditioned LOBPCG on the
        device CSR otherwise (`modal.py`)."""
        from .modal import ModesFromElements

        k = self.k0()
        m = self.integrate_mass()
        omega_sq, phis = ModesFromElements.apply(k, m, self, n_modes)
        modes = phis.detach().T.reshape(n_modes, self.n_nod, self.n_dof_per_node)
        return omega_sq, modes

    @property
    def forces(self) -> Tensor:
        return self._neumann

    @forces.setter
    def forces(self, value: Tensor):
        self._set_field("_neumann", "Forces", value, "Forces must have the same shape as nodes.")

    @property
    def displacements(self) -> Tensor:
        return self._dirichlet

    @displacements.setter
    def displacements(self, value: Tensor):
        self._set_field("_dirichlet", "Displacements", value, "Displacements must have the same shape as nodes.")

    @property
    def ext_strain(self) -> Tensor:
        return self._external_gradient

    @ext_strain.setter
    def ext_strain(self, value: Tensor):
        if value.shape != (self.n_elem, self.n_dof_per_node, self.n_dim):
            raise ValueError("External strain must have the same shape as strains.")
        if not torch.is_floating_point(value):
            raise TypeError("External strain must be a floating-point tensor.")
        self._external_gradient = value.to(self.device)

    def integrate_material(self, u_prev, grad_prev, flux_prev, state_prev, du, de0, iter, nlgeom,
                           compute_stiffness: bool = True):
        """Gauss-point loop (reference base.py:982-1092): residual quantities in torch (differentiable),
        element tangent matrices on kernel K1. The tangent is needed only when none is cached or the
        material / geometry is nonlinear (same rule as base.py:1031-1033)."""
        assert self.material is not None
        d = self.n_flux[0]
        nd = self.etype.nodes * self.n_dof_per_node
        need_k = compute_stiffness and (self.K.numel() == 0 or self.n_state != 0 or nlgeom)
        geom = self._geometry() if du.dtype == torch.float64 else None
        grads, fluxes, states, tangents = [], [], [], []
        cl = self.char_lengths
        if geom is not None:
            # kernels K9/K10: gradient at all Gauss points in one launch, forces in another; the material update
            # in between is the caller's torch code (base.py:1050-1083 does ~6 small-matrix launches per point)
            H_all = _res.elem_grad(geom, du.view(-1, self.n_dof_per_node)[self.elements])
            P_all, alpha_all, tangent = _step_points(self.material, H_all, grad_prev, flux_prev, state_prev, de0,
                                                     cl, iter, need_k)
            F_all = grad_prev + H_all
            flux_all = _small_matmul(F_all, P_all) / _small_det(F_all)[..., None, None] if nlgeom else P_all
            f = _res.elem_force(geom, P_all)
            geom.check()
            if self._f_scale is not None:
                f = f * self._f_scale[:, None, None]
            k = self._integrate_k(tangent) if need_k else None
            return k, f.reshape(self.n_elem, nd), F_all, flux_all, alpha_all
        else:
            du_e = du.view(-1, self.n_dof_per_node)[self.elements].reshape(self.n_elem, -1, d).transpose(-1, -2)
            _, B, detJ = self._ip_shape()
            f = torch.zeros(self.n_elem, nd, dtype=du.dtype, device=du.device)
            for q, w in enumerate(self.etype.iweights):
                H_inc = du_e @ B[q].transpose(-1, -2)
                F_new = grad_prev[q] + H_inc
                P, alpha, ddsdde = self.material.step(H_inc, grad_prev[q], flux_prev[q], state_prev[q], de0, cl, iter)
                grads.append(F_new)
                fluxes.append(_small_matmul(F_new, P) / _small_det(F_new)[:, None, None] if nlgeom else P)
                states.append(alpha)
                f = f + float(w) * self.compute_f(detJ[q], B[q], P).reshape(-1, nd)
                if need_k:
                    tangents.append(ddsdde)
        k = None
        if need_k:
            same = all(t is tangents[0] for t in tangents)  # elastic: one tensor for every Gauss point
            k = self._integrate_k(tangents[0] if same else torch.stack(tangents))
        return k, f, torch.stack(grads), torch.stack(fluxes), torch.stack(states)


class Heat(FEM, ABC):
    """Scalar (temperature) problems (reference base.py:1132-1286)."""

    KIND = L.KIND_HEAT

    @property
    def n_dof_per_node(self) -> int:
        return 1

    @property
    def n_flux(self) -> list[int]:
        return [1, self.n_dim]

    @property
    def initial_grad(self) -> Tensor:
        return torch.zeros(1)

    @property
    def heat_flux(self) -> Tensor:
        return self._neumann

    @heat_flux.setter
    def heat_flux(self, value: Tensor):
        if value.shape != (self.n_nod, 1):
            raise ValueError("Heat flux must have the same shape as nodes.")
        self._set_field("_neumann", "Heat flux", value, "Heat flux must have the same shape as nodes.")

    @property
    def temperatures(self) -> Tensor:
        return self._dirichlet

    @temperatures.setter
    def temperatures(self, value: Tensor):
        self._set_field("_dirichlet", "Temperatures", value, "Temperatures must have the same shape as nodes.")

    def integrate_material(self, u_prev, grad_prev, flux_prev, state_prev, du, de0, iter, nlgeom,
                           compute_stiffness: bool = True):
        """Thermal Gauss-point loop (reference base.py:1177-1286); conductivity matrices on kernel K1."""
        assert self.material is not None
        nn = self.etype.nodes
        du_e = du.view(-1, 1)[self.elements].reshape(self.n_elem, -1, 1)
        need_k = compute_stiffness and (self.K.numel() == 0 or self.n_state != 0)
        geom = self._geometry() if du.dtype == torch.float64 else None
        grads, fluxes, states, tangents = [], [], [], []
        cl = self.char_lengths
        if geom is not None:  # kernels K9/K10, see Mechanics.integrate_material
            g_all = _res.elem_grad(geom, du_e)
            flux_all, state_all, tangent = _step_points(self.material, g_all, grad_prev, flux_prev, state_prev, de0,
                                                        cl, iter, need_k)
            f = _res.elem_force(geom, flux_all)
            geom.check()
            if self._f_scale is not None:
                f = f * self._f_scale[:, None, None]
            k = self._integrate_k(tangent) if need_k else None
            return k, f.reshape(self.n_elem, nn), grad_prev + g_all, flux_all, state_all
        else:
            _, B, detJ = self._ip_shape()
            f = torch.zeros(self.n_elem, nn, dtype=du.dtype, device=du.device)
            for q, w in enumerate(self.etype.iweights):
                g_inc = torch.einsum("...ij,...jk->...ki", B[q], du_e)
                grads.append(grad_prev[q] + g_inc)
                flux_q, state_q, kappa = self.material.step(g_inc, grad_prev[q], flux_prev[q], state_prev[q], de0, cl, iter)
                fluxes.append(flux_q)
                states.append(state_q)
                f = f + float(w) * self.compute_f(detJ[q], B[q], flux_q).reshape(-1, nn)
                if need_k:
                    tangents.append(kappa)
        k = None
        if need_k:
            same = all(t is tangents[0] for t in tangents)
            k = self._integrate_k(tangents[0] if same else torch.stack(tangents))
        return k, f, torch.stack(grads), torch.stack(fluxes), torch.stack(states)

    def time_integration(self, t_output: Tensor | None = None, delta_t: float = 1.0e-1, max_iter: int = 100,
                         verbose: bool = False, rtol: float = 1e-8, atol: float = 1e-6, stol: float = 1e-10,
                         device: str | None = None, method: str | None = None,
                         aggregate_integration_points: bool = True, use_cached_solve: bool = False,
                         differentiable_parameters: Tensor | Iterable[Tensor] | None = None):
        """Implicit (trapezoidal) time integration of the heat equation (reference base.py:1288-1553): an
        equilibrium state at t = 0 under the current boundary conditions, then steps of at most `delta_t` to
        every requested output time. Same arguments, return values, error messages and stepping rules as the
        reference. Per step: Newton on  M du + dt/2 (f_int_old + f_int + f_ext) = 0  with the linear solves
        `differentiable_sparse_solve(M + dt/2 K, -residual)` on the device CSR (M, K assembled once by kernel
        K2/K3; the combination is an entry-wise sum on the shared pattern). Gradients flow through the right-hand
        sides (loads, previous states) and — when a material parameter is differentiated — through the system matrix
        and `M @ du` as per-element contractions (`_ElementSystemSolve`, `_ElementMatvec`), like the reference's
        differentiable `assemble_matrix` + `Solve.backward`."""
        kw = dict(dtype=self.nodes.dtype, device=self.device)
        t_output = torch.tensor([0.0, 1.0]) if t_output is None else t_output
        t_output = t_output.detach().to("cpu", torch.float64)
        if t_output.numel() == 0:
            raise ValueError("t_output must contain at least one time.")
        if t_output.min() < 0.0:
            raise ValueError("t_output must not contain negative times.")
        if (t_output[1:] <= t_output[:-1]).any():
            raise ValueError("t_output must be strictly increasing.")

        # initial conditions: every DOF held at its prescribed value (base.py:1348-1358)
        bc_constraints = self._constraints.clone()
        self._constraints = torch.ones_like(bc_constraints)
        try:
            temp_eq, _, flux_eq, grad_eq, alpha_eq = self.solve(
                aggregate_integration_points=False, use_cached_solve=use_cached_solve,
                differentiable_parameters=differentiable_parameters)
        finally:
            self._constraints = bc_constraints

        # internal time grid: integration starts at t = 0 even if it is not an output time (base.py:1360-1387)
        t_list = [float(v) for v in t_output.tolist()]
        knots = t_list if t_list[0] <= 0.0 else [0.0] + t_list
        times, rows, row = [knots[0]], ([] if t_list[0] > 0.0 else [0]), 0
        for t0, t1 in zip(knots[:-1], knots[1:]):
            ratio = (t1 - t0) / delta_t
            n_sub = max(1, math.ceil(ratio - 1e-9 * max(1.0, ratio)))
            sub = torch.linspace(t0, t1, n_sub + 1, dtype=torch.float64, device="cpu")[1:].tolist()
            sub[-1] = t1                                   # linspace can miss the knot by an ulp
            times += sub
            row += n_sub
            rows.append(row)
        dts = [b_ - a_ for a_, b_ in zip(times[:-1], times[1:])]
        n_steps = len(times)

        B = self.compute_B()
        con = torch.nonzero(self._constraints.ravel(), as_tuple=False).ravel()
        dpn, shape4 = self.n_dof_per_node, (self.n_int, self.n_elem, self.n_dof_per_node, self.n_dim)
        u0 = temp_eq.clone()
        u0.view(-1)[con] = self._dirichlet.view(-1)[con]       # base.py:1452
        u = [u0]
        f = [torch.zeros(self.n_nod, dpn, **kw)]
        flux, grad, state = [flux_eq.reshape(shape4)], [grad_eq.reshape(shape4)], [alpha_eq]

        self.K = torch.empty(0, device=self.device)
        self.M = None
        m = self.integrate_mass()
        k_el = None                                   # element conductivity matrices behind self.K (keep their graph)
        free = (~self._constraints.ravel()).to(kw["dtype"])
        if verbose:
            print(f"torch-fem_b200 | time integration | {type(self).__name__} | {self.n_dofs:,} dof | "
                  f"{n_steps - 1} steps | dt <= {delta_t:g} | {describe_method(self.n_dofs, 'cuda', method)}")

        sys_cache = (None, None, None, None)
        for n in range(1, n_steps):
            u_guess = u[n - 1].clone()
            dt_n = dts[n - 1]
            f_old = f[n - 1]
            res_norm = res_norm0 = None
            f_int = grad_n = flux_n = state_n = None
            for it in range(max_iter):
                du = u_guess - u[n - 1]
                k, f_e, grad_n, flux_n, state_n = self.integrate_material(
                    u[n - 1], grad[n - 1], flux[n - 1], state[n - 1], du, self._external_gradient, it, False)
                f_int = self.assemble_rhs(f_e)
                if k is not None:
                    self.K = self.assemble_matrix(k, con)
                    k_el = k
                if self.M is None:
                    self.M = self.assemble_matrix(m, con)
                elem_grad = torch.is_grad_enabled() and (m.requires_grad or (k_el is not None and k_el.requires_grad))
                Mdu = (_ElementMatvec.apply(du.reshape(-1), m, self.M, self.idx, free) if elem_grad and m.requires_grad
                       else self.M @ du.reshape(-1))
                residual = Mdu + 0.5 * dt_n * (f_old.reshape(-1) + f_int + self._neumann.ravel())
                mask = torch.ones_like(residual)
                mask[con] = 0.0
                residual = residual * mask
                res_norm = torch.linalg.norm(residual.detach())
                if it == 0:
                    res_norm0 = res_norm
                if res_norm < rtol * res_norm0 or res_norm < atol:
                    break
                cached = self.cached_solve if (it == 0 and use_cached_solve) else CachedSolve()
                # the system matrix M + dt/2 K (reference base.py:1513 re-forms it in every iteration) is kept while
                # the step size and both matrices are unchanged: its SELL copy — and, with method="amgx", the
                # refreshed hierarchy — are then reused by every step
                if not (sys_cache[0] == float(dt_n) and sys_cache[1] is self.M and sys_cache[2] is self.K):
                    sys_cache = (float(dt_n), self.M, self.K, self.M + 0.5 * dt_n * self.K)
                if elem_grad:
                    step = _ElementSystemSolve.apply(-residual, k_el, m, sys_cache[3], self.idx, free, 0.5 * dt_n, 1.0,
                                                     (B, stol, device, method, cached, it == 0))
                else:
                    step = differentiable_sparse_solve(sys_cache[3], -residual, B, stol, device, method,
                                                       None, cached, it == 0)
                u_guess = u_guess + step.reshape(-1, dpn)
            if res_norm > rtol * res_norm0 and res_norm > atol:
                raise RuntimeError("Newton-Raphson iteration did not converge.")
            u.append(u_guess)
            f.append(f_int.reshape(-1, dpn))
            grad.append(grad_n)
            flux.append(flux_n)
            state.append(state_n)
            if verbose:
                print(f"  step {n:5d}  t = {times[n]:.6g}  |res| = {float(res_norm):.3e}")

        pick = lambda seq: torch.stack([seq[i] for i in rows])  # noqa: E731
        out_u, out_f, out_flux, out_grad, out_state = pick(u), pick(f), pick(flux), pick(grad), pick(state)
        if aggregate_integration_points:
            out_grad, out_flux, out_state = out_grad.mean(dim=1), out_flux.mean(dim=1), out_state.mean(dim=1)
        out_flux, out_grad = out_flux.squeeze((-2, -1)), out_grad.squeeze((-2, -1))
        return out_u, out_f, out_flux, out_grad, out_state
