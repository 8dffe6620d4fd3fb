"""Material-tangent interface of the hot path.

The contract is the reference's (src/torchfem/materials/base.py:31-91):
`Material.step(H_inc, F, stress, state, de0, cl, iter) -> (stress_new, state_new, ddsdde)` evaluated per
Gauss point on tensors batched over elements, `vectorize(n_elem)`, `rotate(R)`, attributes `n_state`,
`is_vectorized`, `rho`. The tangent `ddsdde` ([n_elem,d,d,d,d] mechanics, [n_elem,d,d] heat) is what the
CUDA integration kernel K1 consumes; the stress update itself stays in torch because the adjoint
differentiates the residual through it (reference sparse.py:689-705).

Materials on the path: isotropic elasticity (3-D, plane stress, plane strain; reference
materials/elasticity.py:11-241), hyperelasticity from a strain-energy callable (materials/hyperelasticity.py:
66-127, 309-355) and isotropic conductivity (materials/conductivity.py:11-240). Any other object with the
same `step` contract (e.g. a plasticity model) plugs into the models unchanged.
"""
from __future__ import annotations

import copy
import os
from typing import Callable

import torch
from torch import Tensor
from torch.func import jacrev, vmap
from torch.overrides import TorchFunctionMode


def small_det(a: Tensor) -> Tensor:
    """Closed-form determinant of batched 1x1 / 2x2 / 3x3 matrices (elementwise kernels only)."""
    n = a.shape[-1]
    if n == 1:
        return a[..., 0, 0]
    if n == 2:
        return a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]
    return (a[..., 0, 0] * (a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1])
            - a[..., 0, 1] * (a[..., 1, 0] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 0])
            + a[..., 0, 2] * (a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]))


def small_matmul(a: Tensor, b: Tensor) -> Tensor:
    """a @ b for batched matrices with trailing dims <= 3 as one broadcast multiply + reduction."""
    return (a.unsqueeze(-1) * b.unsqueeze(-3)).sum(-2)


_MATMULS = {torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__, torch.bmm, torch.mm}
_DETS = {torch.det, torch.linalg.det, torch.Tensor.det}
_LOGDETS = {torch.logdet, torch.Tensor.logdet}


def _is_small(t) -> bool:
    return isinstance(t, Tensor) and t.dim() >= 2 and t.shape[-1] <= 3 and t.shape[-2] <= 3


class SmallMatrixMode(TorchFunctionMode):
    """While a user's strain-energy function is differentiated on the device, products / determinants of 2x2 and 3x3
    matrices are rewritten as elementwise expressions. torch lowers `F.T @ F`, `logdet` and their derivative rules
    under vmap(jacrev(jacrev(psi))) to batched cuBLAS / cuSOLVER calls on 3x3 matrices (FP64 tensor-core GEMM tiles,
    LU + triangular solves): ~300 us per launch, 77 % of the reference's hyperelasticity benchmark (1.5 of 2.2 s,
    `tools/prof_hyper.py`). The rewritten expressions are mathematically identical (1e-15 relative on P and the
    tangent) and differentiate through elementwise kernels."""

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.tensor and len(args) == 1 and "device" not in kwargs and isinstance(args[0], (list, tuple)) \
                and args[0] and all(type(v) is int for v in args[0]):
            # jacrev builds its basis offsets with torch.tensor([numels]) and reads them back with int(): under
            # torch.set_default_device("cuda") (the reference's benchmark setting) that is a host-to-device copy plus
            # a device synchronisation per call. Integer bookkeeping lists stay on the host.
            return func(*args, device="cpu", **kwargs)
        if not kwargs:
            if func in _MATMULS and len(args) == 2 and _is_small(args[0]) and _is_small(args[1]):
                return small_matmul(args[0], args[1])
            if len(args) == 1 and _is_small(args[0]) and args[0].shape[-1] == args[0].shape[-2]:
                if func in _DETS:
                    return small_det(args[0])
                if func in _LOGDETS:
                    return torch.log(small_det(args[0]))
        return func(*args, **kwargs)


class _NoMode:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def small_matrix_mode(t: Tensor):
    """`SmallMatrixMode` for device tensors, nothing on the host (CPU bmm is fine and the host path is test-only)."""
    return SmallMatrixMode() if t.is_cuda else _NoMode()


class Material:
    """Base class: bookkeeping of batched parameters (reference materials/base.py:10-103)."""

    n_state: int = 0
    is_vectorized: bool = False
    # False for a material whose algorithmic tangent is not major-symmetric (non-associated plasticity, damage with
    # a secant / consistent mix ...): the assembled K is then tagged non-symmetric, so the adjoint solves with a real
    # transpose (reference sparse.py:671 `K.T`) and CG / AMG (SPD methods) are not auto-selected for it
    symmetric_tangent: bool = True

    def __init__(self):
        self.n_state = 0
        self.is_vectorized = False
        self.rho = torch.tensor(1.0)

    def vectorize(self, n_elem: int) -> "Material":
        """Copy whose tensor attributes carry one entry per element (shared if already batched)."""
        if self.is_vectorized:
            return self
        out = copy.copy(self)
        for name, val in list(vars(out).items()):
            if isinstance(val, Tensor):
                setattr(out, name, val.repeat(n_elem, *([1] * val.dim())))
        out.is_vectorized = True
        return out

    def to(self, device) -> "Material":
        """The material with every tensor attribute on `device` (the models keep their material on the GPU): the
        object itself when nothing has to move — like the reference, which keeps the user's object when it is already
        vectorized (materials/base.py:43-47), so re-assigning `material.C = ...` still reaches the model — else a
        shallow copy."""
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None and torch.cuda.is_available():
            dev = torch.device("cuda", torch.cuda.current_device())
        if all(v.device == dev for v in vars(self).values() if isinstance(v, Tensor)):
            return self
        out = copy.copy(self)
        for name, val in list(vars(out).items()):
            if isinstance(val, Tensor):
                setattr(out, name, val.to(device))
        return out

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        raise NotImplementedError

    def step_points(self, H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent: bool = True):
        """`step` at all Gauss points: inputs stacked over the leading integration-point axis. Returns
        (stress [n_int, ...], state [n_int, ...], tangent) where tangent is ONE per-element tensor if `step`
        returned the same object at every point (elastic) and a stack otherwise. The default is the reference's
        loop (base.py:1050-1058); materials whose update is launch-bound override it with one batched call."""
        return step_points_loop(self, H_all, F_all, stress_all, state_all, de0, cl, iter)

    def rotate(self, R: Tensor) -> "Material":
        return self


def step_points_loop(material, H_all, F_all, stress_all, state_all, de0, cl, iter):
    """One `material.step` per Gauss point, exactly as the reference's Gauss loop calls it; works for any object
    with the `step` contract (reference materials/base.py:50-91)."""
    Ps, sts, tans = [], [], []
    for q in range(H_all.shape[0]):
        P, a, t = material.step(H_all[q], F_all[q], stress_all[q], state_all[q], de0, cl, iter)
        Ps.append(P)
        sts.append(a)
        tans.append(t)
    same = all(t is tans[0] for t in tans)
    return torch.stack(Ps), torch.stack(sts), (tans[0] if same else torch.stack(tans))


class _DDot(torch.autograd.Function):
    """sigma[q,e] = C[e] : eps[q,e] on kernel K17 (`tfem_ddot`); backward: d eps = C^T : g (the same kernel,
    transposed), dC = sum_q g (x) eps (`tfem_ddot_outer`)."""

    @staticmethod
    def forward(ctx, C: Tensor, e: Tensor):
        from . import _lib as L

        d = e.shape[-1]
        m = d * d
        n_elem = C.shape[0]
        Cm = C.detach().reshape(n_elem, m, m).contiguous()
        em = e.detach().reshape(-1, n_elem, m).contiguous()
        out = torch.empty_like(em)
        L.check(L.lib.tfem_ddot(m, em.shape[0], n_elem, L.ptr(Cm), L.ptr(em), 0, L.ptr(out), L.stream()))
        ctx.save_for_backward(Cm, em)
        ctx.shapes = (C.shape, e.shape)
        return out.reshape(e.shape)

    @staticmethod
    def backward(ctx, g: Tensor):
        from . import _lib as L

        Cm, em = ctx.saved_tensors
        n_elem, m = Cm.shape[0], Cm.shape[1]
        gm = g.reshape(-1, n_elem, m).contiguous()
        gC = ge = None
        if ctx.needs_input_grad[1]:
            ge = torch.empty_like(gm)
            L.check(L.lib.tfem_ddot(m, gm.shape[0], n_elem, L.ptr(Cm), L.ptr(gm), 1, L.ptr(ge), L.stream()))
            ge = ge.reshape(ctx.shapes[1])
        if ctx.needs_input_grad[0]:
            gC = torch.empty_like(Cm)
            L.check(L.lib.tfem_ddot_outer(m, gm.shape[0], n_elem, L.ptr(gm), L.ptr(em), L.ptr(gC), L.stream()))
            gC = gC.reshape(ctx.shapes[0])
        return gC, ge


def _ddot(C: Tensor, e: Tensor) -> Tensor:
    """C_ijkl e_kl batched over leading dims. Per-element tangents on the device go through kernel K17 (one pass:
    C read once for all Gauss points); everything else (a single unvectorised tangent, CPU tensors, higher-order
    differentiation) uses one broadcast multiply + reduction. (`torch.einsum` lowers this to a batched matrix-vector
    product whose cuBLAS kernels took 8.7 ms per Gauss point at 3.4 M elements; the broadcast form writes and reads
    a [n_int, n_elem, d^4] temporary — 17.5 GB at config B, 41 ms per solve.)"""
    d = e.shape[-1]
    if (C.is_cuda and e.is_cuda and C.dtype == torch.float64 and e.dtype == torch.float64 and d in (1, 2, 3)
            and C.dim() == 5 and e.dim() in (3, 4) and e.shape[-3] == C.shape[0] and C.shape[1:] == (d, d, d, d)
            and not torch._C._functorch.is_functorch_wrapped_tensor(e)
            and not torch._C._functorch.is_functorch_wrapped_tensor(C)):
        return _DDot.apply(C, e)
    return (C.reshape(*C.shape[:-2], d * d) * e.reshape(*e.shape[:-2], 1, 1, d * d)).sum(-1)


def _isotropic_tensor(lbd: Tensor, G: Tensor, d: int) -> Tensor:
    """C_ijkl = lbd d_ij d_kl + G (d_ik d_jl + d_il d_jk), batched over the leading dims of lbd."""
    eye = torch.eye(d, dtype=lbd.dtype, device=lbd.device)
    vol = torch.einsum("ij,kl->ijkl", eye, eye)
    sym = torch.einsum("ik,jl->ijkl", eye, eye) + torch.einsum("il,jk->ijkl", eye, eye)
    return lbd[..., None, None, None, None] * vol + G[..., None, None, None, None] * sym


class _LinearElasticity(Material):
    """Small-strain linear elasticity with a constant stiffness tensor `C` [..., d,d,d,d]: the stress update shared by
    the isotropic and orthotropic materials."""

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        """sigma_{n+1} = sigma_n + C : (sym(dH) - de0); tangent = C (elasticity.py:119-127)."""
        de = 0.5 * (H_inc + H_inc.transpose(-1, -2)) - de0
        return stress + _ddot(self.C, de), state, self.C

    def step_points(self, H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent: bool = True):
        """The same update at all Gauss points in one pass: the per-element tangent is read once (kernel K17)
        instead of once per point."""
        if self.C.dim() != 5:            # one tangent for the whole model: the per-point broadcast is already cheap
            return step_points_loop(self, H_all, F_all, stress_all, state_all, de0, cl, iter)
        de = 0.5 * (H_all + H_all.transpose(-1, -2)) - de0
        return stress_all + _ddot(self.C, de), state_all, self.C


class IsotropicElasticity3D(_LinearElasticity):
    """Small-strain isotropic elasticity; `C` is [..., 3,3,3,3] (reference elasticity.py:11-127)."""

    def __init__(self, E, nu, rho=1.0):
        self.E = torch.as_tensor(E)
        self.nu = torch.as_tensor(nu)
        self.rho = torch.as_tensor(rho)
        self.n_state = 0
        self.is_vectorized = self.E.dim() > 0
        self.lbd = self.E * self.nu / ((1.0 + self.nu) * (1.0 - 2.0 * self.nu))
        self.G = self.E / (2.0 * (1.0 + self.nu))
        self.C = _isotropic_tensor(self.lbd, self.G, 3)


class IsotropicElasticityPlaneStress(IsotropicElasticity3D):
    """sigma_33 = 0: C_0000 = C_1111 = E/(1-nu^2), C_0011 = nu E/(1-nu^2), shear E/(2(1+nu))
    (reference elasticity.py:130-185)."""

    def __init__(self, E, nu, rho=1.0):
        super().__init__(E, nu, rho)
        f = self.E / (1.0 - self.nu ** 2)
        self.C = _isotropic_tensor(f * self.nu, 0.5 * f * (1.0 - self.nu), 2)


class IsotropicElasticityPlaneStrain(IsotropicElasticity3D):
    """eps_33 = 0: the in-plane block of the 3-D tensor (reference elasticity.py:188-241)."""

    def __init__(self, E, nu, rho=1.0):
        super().__init__(E, nu, rho)
        self.C = _isotropic_tensor(self.lbd, self.G, 2)


_GRAPH_MAX_POINTS = 4_000_000       # static buffers of a graph: 81 doubles of tangent per point
_GRAPH_OFF = os.environ.get("TFEM_MATERIAL_GRAPH", "1") == "0"
_GRAPH_CACHE: dict = {}


class _GraphedDerivatives:
    """P = d psi / dF and (optionally) d2 psi / dF2 for a fixed batch shape, captured once into a CUDA graph and replayed
    on static buffers (inputs copied in, outputs cloned out: the callers keep stresses across iterations)."""

    def __init__(self, psi, F: Tensor, params: Tensor, need_tangent: bool):
        self.F, self.params = F.detach().clone(), params.detach().clone()
        self.need_tangent = need_tangent

        def run():
            with torch.enable_grad(), small_matrix_mode(self.F):
                P = vmap(jacrev(psi))(self.F, self.params)
                T = vmap(jacrev(jacrev(psi)))(self.F, self.params) if need_tangent else None
            return P.detach(), (None if T is None else T.detach())

        # capture by hand on a side stream: the `torch.cuda.graph` context manager also runs gc.collect() and
        # torch.cuda.empty_cache(), which hands every cached block back to the driver — measured: the whole solve after a
        # capture got slower than the capture saved (0.87 vs 0.68 s forward at 63 k DOFs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for _ in range(2):
                run()
            self.graph.capture_begin()
            try:
                self.P, self.T = run()
            finally:
                self.graph.capture_end()
        torch.cuda.current_stream().wait_stream(side)

    def __call__(self, F: Tensor, params: Tensor):
        self.F.copy_(F.detach())
        self.params.copy_(params.detach())
        self.graph.replay()
        return self.P.clone(), (None if self.T is None else self.T.clone())


class Hyperelastic3D(Material):
    """Hyperelasticity from a strain-energy density psi(F, params): P = dpsi/dF and the tangent
    d2psi/dF2 by forward-over-reverse autodiff, batched with vmap (reference hyperelasticity.py:66-127)."""

    def __init__(self, psi: Callable, params, rho=1.0):
        self.psi = psi
        self.params = torch.as_tensor(params)
        self.n_state = 0
        self.rho = torch.as_tensor(rho)
        self.is_vectorized = self.params.dim() > 1

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        with torch.enable_grad(), small_matrix_mode(F):
            F_new = (F + H_inc).requires_grad_(True)
            P = vmap(jacrev(self.psi))(F_new, self.params)
            tangent = vmap(jacrev(jacrev(self.psi)))(F_new, self.params)
        return P, state, tangent

    def _psi_point(self):
        return self.psi

    def _graphed_update(self, psi, F_new: Tensor, params: Tensor, need_tangent: bool):
        """(P, tangent) through a cached CUDA graph of the batched derivative evaluation; None if this energy function
        cannot be captured (host synchronisation or data-dependent control flow inside psi) — decided once."""
        # keyed by the user's energy FUNCTION (models are rebuilt per load case / design iteration with the same psi):
        # a small process-wide cache, oldest entry dropped first
        cache = _GRAPH_CACHE
        key = (self.psi, type(self).__name__, tuple(F_new.shape), tuple(params.shape), bool(need_tangent), F_new.device,
               F_new.dtype)
        entry = cache.get(key)
        if entry is False:
            return None
        if entry is None:
            try:
                entry = _GraphedDerivatives(psi, F_new, params, need_tangent)
            except Exception:   # noqa: BLE001 — any capture failure: stay on the eager path for this shape
                cache[key] = False
                torch.cuda.synchronize()
                return None
            while len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = entry
        return entry(F_new, params)

    def step_points(self, H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent: bool = True):
        """All Gauss points in ONE vmap call (the per-point calls of the reference loop are launch-bound: a few
        hundred tiny kernels each); the tangent is skipped when the caller does not integrate a stiffness."""
        n_int, n_elem = H_all.shape[:2]
        psi = self._psi_point()
        params = self.params
        if params.dim() == 1:
            params = params.expand(n_elem, -1)
        params = params.expand(n_int, *params.shape).reshape(n_int * n_elem, -1)
        tracked = torch.is_grad_enabled() and (H_all.requires_grad or F_all.requires_grad or params.requires_grad)
        if F_all.is_cuda and not tracked and n_int * n_elem <= _GRAPH_MAX_POINTS and not _GRAPH_OFF:
            # forward Newton iterations (no autograd graph wanted): the update is a few hundred tiny elementwise kernels
            # under vmap(jacrev(jacrev(psi))) — dispatch-bound (9.6 ms eager vs 4.4 ms as ONE replayed CUDA graph at
            # 131 k points, bitwise equal: tools/graph_probe.py). One graph per (psi, shape, need_tangent); anything that
            # cannot be captured falls back to eager.
            out = self._graphed_update(psi, (F_all + H_all).reshape(n_int * n_elem, *H_all.shape[2:]), params, need_tangent)
            if out is not None:
                P, tangent = out
                P = P.reshape(H_all.shape)
                if tangent is not None:
                    tangent = tangent.reshape(n_int, n_elem, *tangent.shape[1:])
                return P, state_all, tangent
        with torch.enable_grad(), small_matrix_mode(F_all):
            F_new = (F_all + H_all).reshape(n_int * n_elem, *H_all.shape[2:]).requires_grad_(True)
            P = vmap(jacrev(psi))(F_new, params)
            tangent = vmap(jacrev(jacrev(psi)))(F_new, params) if need_tangent else None
        P = P.reshape(H_all.shape)
        if tangent is not None:
            tangent = tangent.reshape(n_int, n_elem, *tangent.shape[1:])
        return P, state_all, tangent


class HyperelasticPlaneStrain(Hyperelastic3D):
    """Plane strain: psi is evaluated on the 3x3 embedding diag(F2d, 1), so stress and tangent are the
    in-plane derivatives (reference hyperelasticity.py:309-355)."""

    def _psi_point(self):
        def psi2(F2, p):
            F3 = torch.zeros(3, 3, dtype=F2.dtype, device=F2.device)
            F3 = F3 + torch.nn.functional.pad(F2, (0, 1, 0, 1))
            F3 = F3 + torch.diag(torch.tensor([0.0, 0.0, 1.0], dtype=F2.dtype, device=F2.device))
            return self.psi(F3, p)

        return psi2

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        psi2 = self._psi_point()
        with torch.enable_grad(), small_matrix_mode(F):
            F_new = (F + H_inc).requires_grad_(True)
            P = vmap(jacrev(psi2))(F_new, self.params)
            tangent = vmap(jacrev(jacrev(psi2)))(F_new, self.params)
        return P, state, tangent


class HyperelasticPlaneStress(Hyperelastic3D):
    """Plane stress: the thickness stretch lambda_z is a state variable found by a local Newton iteration on
    P_33(F_2d, lambda_z) = 0, and the in-plane tangent is condensed, C_abcd - C_ab33 C_33cd / C_3333 (reference
    hyperelasticity.py:130-269; same iteration: every point is updated until all satisfy |P_33| < tolerance, at most
    `max_iter` times). `n_state = 1`: state = lambda_z - 1."""

    def __init__(self, psi: Callable, params, rho=1.0, tolerance: float = 1e-5, max_iter: int = 10):
        super().__init__(psi, params, rho)
        self.tolerance, self.max_iter = tolerance, max_iter
        self.n_state = 1

    def _evaluate(self, F2: Tensor, stretch: Tensor, params: Tensor):
        """P and tangent of the 3-D energy at diag(F2, stretch), batched over the leading axis."""
        F3 = torch.nn.functional.pad(F2, (0, 1, 0, 1))
        F3 = F3 + stretch[:, None, None] * torch.nn.functional.one_hot(torch.tensor(8, device=F2.device), 9).reshape(3, 3).to(F2)
        with torch.enable_grad(), small_matrix_mode(F3):
            F3 = F3.requires_grad_(True) if not F3.requires_grad else F3
            P = vmap(jacrev(self.psi))(F3, params)
            tangent = vmap(jacrev(jacrev(self.psi)))(F3, params)
        return P, tangent

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        batch = H_inc.shape[:-2]
        F2 = (F + H_inc).reshape(-1, 2, 2)
        stretch = (1.0 + state[..., 0]).reshape(-1)
        params = self.params
        if params.dim() == 1:
            params = params.expand(F2.shape[0], -1)
        else:   # one set per element, shared by the Gauss points when the batch is [n_int, n_elem]
            params = params.expand(*batch, params.shape[-1]).reshape(-1, params.shape[-1])
        residual = None
        for _ in range(self.max_iter):
            P, tangent = self._evaluate(F2, stretch, params)
            residual = P[:, 2, 2]
            stretch = stretch - residual / tangent[:, 2, 2, 2, 2]
            if bool((residual.abs() < self.tolerance).all()):
                break
        if bool((residual.abs() > self.tolerance).any()):
            print("Local Newton iteration did not converge.")
        condensed = tangent[:, :2, :2, :2, :2] - (tangent[:, :2, :2, 2, 2, None, None] * tangent[:, None, None, 2, 2, :2, :2]
                                                  / tangent[:, 2, 2, 2, 2, None, None, None, None])
        return (P[:, :2, :2].reshape(*batch, 2, 2), (stretch - 1.0).reshape(*batch, 1),
                condensed.reshape(*batch, 2, 2, 2, 2))

    def step_points(self, H_all, F_all, stress_all, state_all, de0, cl, iter, need_tangent: bool = True):
        """All Gauss points in one batch. (The reference iterates each Gauss point's batch until that batch has
        converged; here the convergence test spans all points, so a point may receive an extra Newton update —
        a difference below the tolerance the iteration stops at.)"""
        P, state, tangent = self.step(H_all, F_all, stress_all, state_all, de0, cl, iter)
        return P, state, tangent


# ------------------------------------------------------------------------------------------------ orthotropy
_VOIGT_PAIRS = {2: ((0, 0), (1, 1), (0, 1)), 3: ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))}


def _voigt_matrix(C: Tensor) -> Tensor:
    """[..., d,d,d,d] -> Voigt matrix [..., m, m] (m = 3 or 6; component order 11, 22, (33, 23, 13,) 12 as in the
    reference's utils.stiffness2voigt)."""
    pairs = _VOIGT_PAIRS[C.shape[-1]]
    return torch.stack([torch.stack([C[..., i, j, k, l] for (k, l) in pairs], -1) for (i, j) in pairs], -2)


def _tensor_from_voigt(V: Tensor, d: int) -> Tensor:
    """Voigt matrix [..., m, m] -> tensor [..., d,d,d,d] with minor symmetries."""
    pairs = _VOIGT_PAIRS[d]
    lookup = {}
    for a, (i, j) in enumerate(pairs):
        lookup[(i, j)] = lookup[(j, i)] = a
    rows = []
    for i in range(d):
        for j in range(d):
            rows.append(torch.stack([V[..., lookup[(i, j)], lookup[(k, l)]] for k in range(d) for l in range(d)], -1))
    return torch.stack(rows, -2).reshape(*V.shape[:-2], d, d, d, d)


def _rotate4(C: Tensor, R: Tensor) -> Tensor:
    """C'_mnop = R_mi R_nj R_ok R_pl C_ijkl, two indices at a time."""
    half = torch.einsum("...mi,...nj,...ijkl->...mnkl", R, R, C)
    return torch.einsum("...ok,...pl,...mnkl->...mnop", R, R, half)


class OrthotropicElasticity3D(_LinearElasticity):
    """Small-strain orthotropic elasticity from nine engineering constants (reference elasticity.py:324-522). The
    normal block of the stiffness is the inverse of the 3x3 compliance block (1/E_i on the diagonal, -nu_ij/E_i off
    it); the shear moduli fill the rest. `rotate(R)` turns the material axes; the engineering constants of the
    result are the apparent ones along the global axes, read off the rotated compliance."""

    _dim = 3

    def __init__(self, E_1, E_2, E_3, nu_12, nu_13, nu_23, G_12, G_13, G_23, rho=1.0):
        t = torch.as_tensor
        self.E_1, self.E_2, self.E_3 = t(E_1), t(E_2), t(E_3)
        self.nu_12, self.nu_13, self.nu_23 = t(nu_12), t(nu_13), t(nu_23)
        self.nu_21 = self.E_2 / self.E_1 * self.nu_12
        self.nu_31 = self.E_3 / self.E_1 * self.nu_13
        self.nu_32 = self.E_3 / self.E_2 * self.nu_23
        self.G_12, self.G_13, self.G_23 = t(G_12), t(G_13), t(G_23)
        self.rho = t(rho)
        self.n_state = 0
        self.is_vectorized = self.E_1.dim() > 0
        one = torch.ones_like(self.E_1 * self.E_2 * self.E_3 * self.nu_12 * self.nu_13 * self.nu_23, dtype=torch.get_default_dtype())
        S = torch.stack([
            torch.stack([one / self.E_1, -self.nu_12 / self.E_1 * one, -self.nu_13 / self.E_1 * one], -1),
            torch.stack([-self.nu_12 / self.E_1 * one, one / self.E_2, -self.nu_23 / self.E_2 * one], -1),
            torch.stack([-self.nu_13 / self.E_1 * one, -self.nu_23 / self.E_2 * one, one / self.E_3], -1)], -2)
        normal = torch.linalg.inv(S)
        V = torch.zeros(*one.shape, 6, 6, dtype=normal.dtype, device=normal.device)
        V[..., :3, :3] = normal
        V[..., 3, 3], V[..., 4, 4], V[..., 5, 5] = self.G_23 * one, self.G_13 * one, self.G_12 * one
        self.C = _tensor_from_voigt(V, 3)

    def rotate(self, R: Tensor):
        d = self._dim
        if R.shape[-2] != d or R.shape[-1] != d:
            raise ValueError(f"Rotation matrix must be a {d}x{d} tensor.")
        out = copy.copy(self)
        out.C = _rotate4(self.C, R.to(self.C))
        S = torch.linalg.inv(_voigt_matrix(out.C))
        out.E_1, out.E_2 = 1.0 / S[..., 0, 0], 1.0 / S[..., 1, 1]
        out.nu_12 = -S[..., 0, 1] / S[..., 0, 0]
        if d == 3:
            out.E_3 = 1.0 / S[..., 2, 2]
            out.nu_13, out.nu_23 = -S[..., 0, 2] / S[..., 0, 0], -S[..., 1, 2] / S[..., 1, 1]
            out.G_23, out.G_13, out.G_12 = 1.0 / S[..., 3, 3], 1.0 / S[..., 4, 4], 1.0 / S[..., 5, 5]
        else:
            out.G_12 = 1.0 / S[..., 2, 2]
        return out


class TransverseIsotropicElasticity3D(OrthotropicElasticity3D):
    """Axis 1 is the symmetry axis, the 2-3 plane is isotropic: five constants (reference elasticity.py:525-571)."""

    def __init__(self, E_L, E_T, nu_L, nu_T, G_L, rho=1.0):
        if bool(torch.any(torch.as_tensor(G_L) > torch.as_tensor(E_L) / (2 * (1 + torch.as_tensor(nu_L))))):
            raise ValueError("G must be less than E_L/(2*(1+nu_L))")
        G_T = torch.as_tensor(E_T) / (2 * (1 + torch.as_tensor(nu_T)))
        super().__init__(E_L, E_T, E_T, nu_L, nu_L, nu_T, G_L, G_L, G_T, rho)


class OrthotropicElasticityPlaneStress(OrthotropicElasticity3D):
    """sigma_33 = 0 (reference elasticity.py:574-676): the in-plane stiffness is the inverse of the 2x2 in-plane
    compliance. `G_13` / `G_23` do not enter it; they are kept when given (the reference's shells read them)."""

    _dim = 2

    def __init__(self, E_1, E_2, nu_12, G_12, G_13=None, G_23=None, rho=1.0):
        t = torch.as_tensor
        self.E_1, self.E_2, self.nu_12, self.G_12 = t(E_1), t(E_2), t(nu_12), t(G_12)
        self.nu_21 = self.E_2 / self.E_1 * self.nu_12
        if G_13 is not None:
            self.G_13 = t(G_13)
        if G_23 is not None:
            self.G_23 = t(G_23)
        self.rho = t(rho)
        self.n_state = 0
        self.is_vectorized = self.E_1.dim() > 0
        one = torch.ones_like(self.E_1 * self.E_2 * self.nu_12 * self.G_12, dtype=torch.get_default_dtype())
        S = torch.stack([torch.stack([one / self.E_1, -self.nu_12 / self.E_1 * one], -1),
                         torch.stack([-self.nu_12 / self.E_1 * one, one / self.E_2], -1)], -2)
        V = torch.zeros(*one.shape, 3, 3, dtype=S.dtype, device=S.device)
        V[..., :2, :2] = torch.linalg.inv(S)
        V[..., 2, 2] = self.G_12 * one
        self.C = _tensor_from_voigt(V, 2)


class OrthotropicElasticityPlaneStrain(OrthotropicElasticity3D):
    """eps_33 = 0 (reference elasticity.py:679-793): the in-plane part of the 3-D orthotropic stiffness."""

    _dim = 2

    def __init__(self, E_1, E_2, E_3, nu_12, nu_13, nu_23, G_12, G_13=None, G_23=None, rho=1.0):
        full = OrthotropicElasticity3D(E_1, E_2, E_3, nu_12, nu_13, nu_23, G_12,
                                       G_12 if G_13 is None else G_13, G_12 if G_23 is None else G_23, rho)
        for name in ("E_1", "E_2", "E_3", "nu_12", "nu_21", "nu_13", "nu_31", "nu_23", "nu_32", "G_12", "rho"):
            setattr(self, name, getattr(full, name))
        if G_13 is not None:
            self.G_13 = full.G_13
        if G_23 is not None:
            self.G_23 = full.G_23
        self.n_state = 0
        self.is_vectorized = full.is_vectorized
        self.C = full.C[..., :2, :2, :2, :2].clone()


class _IsotropicConductivity(Material):
    """Fourier conduction q = kappa grad T; `KAPPA` is [..., d, d] (reference conductivity.py:11-240)."""

    _dim = 3

    def __init__(self, kappa, rho=1.0):
        self.kappa = torch.as_tensor(kappa)
        self.rho = torch.as_tensor(rho)
        self.n_state = 0
        self.is_vectorized = self.kappa.dim() > 0
        self.KAPPA = self.kappa[..., None, None] * torch.eye(self._dim)

    def step(self, H_inc, F, stress, state, de0, cl, iter):
        """flux_{n+1} = flux_n + KAPPA (dgrad - de0) on [..., 1, d] rows; tangent = KAPPA."""
        flux = stress + torch.einsum("...ij,...kj->...ki", self.KAPPA, H_inc - de0)
        return flux, state, self.KAPPA


class IsotropicConductivity3D(_IsotropicConductivity):
    _dim = 3


class IsotropicConductivity2D(_IsotropicConductivity):
    _dim = 2


class IsotropicConductivity1D(_IsotropicConductivity):
    _dim = 1


class _OrthotropicConductivity(_IsotropicConductivity):
    """Conductivities along the material axes, KAPPA = diag(kappa_i); `rotate(R)` gives R KAPPA R^T (reference
    conductivity.py:143-242)."""

    def __init__(self, *kappas, rho=1.0):
        if len(kappas) != self._dim:
            raise TypeError(f"expected {self._dim} conductivities")
        ks = [torch.as_tensor(k) for k in kappas]
        for i, k in enumerate(ks):
            setattr(self, f"kappa_{i + 1}", k)
        self.rho = torch.as_tensor(rho)
        self.n_state = 0
        self.is_vectorized = ks[0].dim() > 0
        self.KAPPA = torch.diag_embed(torch.stack(torch.broadcast_tensors(*ks), -1).to(torch.get_default_dtype()))

    def rotate(self, R: Tensor):
        d = self._dim
        if R.shape[-2] != d or R.shape[-1] != d:
            raise ValueError(f"Rotation matrix must be a {d}x{d} tensor.")
        out = copy.copy(self)
        R = R.to(self.KAPPA)
        out.KAPPA = R @ self.KAPPA @ R.transpose(-1, -2)
        return out


class OrthotropicConductivity3D(_OrthotropicConductivity):
    _dim = 3

    def __init__(self, kappa_1, kappa_2, kappa_3, rho=1.0):
        super().__init__(kappa_1, kappa_2, kappa_3, rho=rho)


class OrthotropicConductivity2D(_OrthotropicConductivity):
    _dim = 2

    def __init__(self, kappa_1, kappa_2, rho=1.0):
        super().__init__(kappa_1, kappa_2, rho=rho)


__all__ = [
    "Material", "IsotropicElasticity3D", "IsotropicElasticityPlaneStress", "IsotropicElasticityPlaneStrain",
    "Hyperelastic3D", "HyperelasticPlaneStrain", "HyperelasticPlaneStress", "IsotropicConductivity3D", "IsotropicConductivity2D",
    "IsotropicConductivity1D", "OrthotropicElasticity3D", "TransverseIsotropicElasticity3D",
    "OrthotropicElasticityPlaneStress", "OrthotropicElasticityPlaneStrain", "OrthotropicConductivity3D",
    "OrthotropicConductivity2D",
]
