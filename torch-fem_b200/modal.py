"""Modal analysis on the device CSR — the drop-in for `modal_eigsolve`, `Eigensolve` and
`differentiable_modal_eigsolve` of the reference (src/torchfem/sparse.py:798-1011), used by `Mechanics.solve_modes`
(base.py:1097-1129):  K phi = omega^2 M phi  for the `n_modes` lowest eigenpairs on the free-DOF subspace.

The reference calls shift-invert Lanczos (`scipy_eigsh(..., sigma=0.0)` / `cupy_eigsh`, sparse.py:820,858), i.e. a
sparse LU of K. Neither exists on this path (no CuPy, no sparse direct solver), and an LU is the wrong tool at B200
scale anyway. Here:

  * small systems (free DOFs <= DENSE_LIMIT) — dense generalized `eigh` on the GPU, the same policy as
    `method="spsolve"` (a library call on a tiny problem);
  * everything else — LOBPCG (Knyazev 2001) on a block of vectors, preconditioned with one AMG V cycle per vector (4 vectors per pass over
    the finest matrix)
    (kernels K11-K16, `amg.AMGPreconditioner`), products with K and M through the SELL-32 SpMV kernels. The basis
    [X, W, P] is kept M-orthonormal block by block (Cholesky), the small Rayleigh-Ritz problems are dense `eigh` calls.

Eigenvectors come back M-normalised with exact zeros on the constrained rows, eigenvalues ascending, like the
reference. Gradients: Rayleigh-quotient sensitivities d lambda / dK_ij = phi_i phi_j, d lambda / dM_ij = -lambda phi_i
phi_j on the matrix pattern (sparse.py:944-982) for torch sparse inputs; for the element-matrix route of
`solve_modes` they are contracted per element instead (`ModesFromElements`), so the global gradient matrix is never
formed.
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from .amg import AMGPreconditioner
from .csr import CSRMatrix

DENSE_LIMIT = 4000     # free DOFs up to which the eigenproblem is solved densely
LOBPCG_TOL = 1e-7      # relative residual ||K x - lambda M x|| / (||K x|| + lambda ||M x||): eigenvalues ~ tol^2
LOBPCG_MAXITER = 400


def _as_csr(A) -> CSRMatrix:
    if isinstance(A, CSRMatrix):
        return A
    if not A.is_cuda:
        raise RuntimeError("torch-fem_b200 has no CPU path: pass CUDA tensors")
    return CSRMatrix.from_coo(A)


def _spmm(A: CSRMatrix, X: Tensor, mask: Tensor) -> Tensor:
    """mask * (A @ X) with the multi-vector SELL product (X is [n, m] row-major: the matrix is read once per 4 columns,
    not once per column)."""
    return A.matmat(X) * mask[:, None]


def _b_orthonormalize(V: Tensor, BV: Tensor, *others: Tensor):
    """V <- V T with T = chol(V^T B V)^-T so that V^T B V = I; the same T is applied to BV and `others` (A V)."""
    G = V.T @ BV
    G = 0.5 * (G + G.T)
    Lc, info = torch.linalg.cholesky_ex(G)
    if int(info) != 0:
        return None
    Tm = torch.linalg.solve_triangular(Lc, torch.eye(G.shape[0], dtype=G.dtype, device=G.device), upper=False).T
    return (V @ Tm, BV @ Tm, *(o @ Tm for o in others))


def lobpcg(K: CSRMatrix, M: CSRMatrix, mask: Tensor, n_modes: int, precondition, tol: float = LOBPCG_TOL,
           maxiter: int = LOBPCG_MAXITER, seed: int = 0):
    """Lowest `n_modes` eigenpairs of (K, M) restricted to the DOFs where mask == 1. Returns (eigenvalues,
    eigenvectors [n, n_modes] M-orthonormal, iterations)."""
    n = K.n
    m = n_modes + max(2, n_modes // 2)                     # guard vectors speed up the wanted ones
    m = min(m, int(mask.sum().item()))
    gen = torch.Generator(device=mask.device).manual_seed(seed)
    X = torch.randn(n, m, dtype=torch.float64, device=mask.device, generator=gen) * mask[:, None]
    out = _b_orthonormalize(X, _spmm(M, X, mask))
    if out is None:
        raise RuntimeError("modal_eigsolve: the mass matrix is not positive definite on the free DOFs")
    X, MX = out
    KX = _spmm(K, X, mask)
    lam, C = torch.linalg.eigh(0.5 * (X.T @ KX + (X.T @ KX).T))
    X, KX, MX = X @ C, KX @ C, MX @ C
    P = KP = MP = None
    it = 0
    for it in range(1, maxiter + 1):
        R = KX - MX * lam[None, :]
        rel = R.norm(dim=0) / (KX.norm(dim=0) + lam.abs() * MX.norm(dim=0))
        if bool((rel[:n_modes] <= tol).all()):
            break
        W = precondition(R) * mask[:, None]
        W = W - X @ (MX.T @ W)                              # M-orthogonal to the current block
        out = _b_orthonormalize(W, _spmm(M, W, mask))
        if out is None:                                      # preconditioned residuals became dependent: converged
            break
        W, MW = out
        KW = _spmm(K, W, mask)
        blocks = [(X, KX, MX), (W, KW, MW)]
        if P is not None:
            Pn = P - X @ (MX.T @ P) - W @ (MW.T @ P)
            # K and M are linear: transform K P, M P with the same coefficients instead of new products
            KPn = KP - KX @ (MX.T @ P) - KW @ (MW.T @ P)
            MPn = MP - MX @ (MX.T @ P) - MW @ (MW.T @ P)
            out = _b_orthonormalize(Pn, MPn, KPn)
            if out is not None:
                Pn, MPn, KPn = out
                blocks.append((Pn, KPn, MPn))
        S = torch.cat([b[0] for b in blocks], dim=1)
        KS = torch.cat([b[1] for b in blocks], dim=1)
        MS = torch.cat([b[2] for b in blocks], dim=1)
        gK = S.T @ KS
        gM = S.T @ MS
        gK, gM = 0.5 * (gK + gK.T), 0.5 * (gM + gM.T)
        Lc, info = torch.linalg.cholesky_ex(gM)
        if int(info) != 0:                                   # ill-conditioned basis: restart without P
            P = KP = MP = None
            continue
        Li = torch.linalg.solve_triangular(Lc, torch.eye(gM.shape[0], dtype=gM.dtype, device=gM.device), upper=False)
        ev, Z = torch.linalg.eigh(Li @ gK @ Li.T)
        Cfull = Li.T @ Z[:, :m]                              # coefficients of the m lowest Ritz vectors
        lam = ev[:m]
        Cx = Cfull.clone()
        Cx[:m] = 0.0                                         # the part outside span(X): the new search directions
        P, KP, MP = S @ Cx, KS @ Cx, MS @ Cx
        X, KX, MX = S @ Cfull, KS @ Cfull, MS @ Cfull
    # final M-normalisation of the wanted vectors
    X, lam = X[:, :n_modes], lam[:n_modes]
    MXn = MX[:, :n_modes]
    X = X / (X * MXn).sum(0).abs().sqrt()[None, :]
    return lam, X, it


def modal_eigsolve(K, M, n_modes: int, free_indices: Tensor, tol: float = LOBPCG_TOL, method: str | None = None):
    """`modal_eigsolve(K, M, n_modes, free_indices) -> (eigenvalues [n_modes] ascending, eigenvectors [n_dofs, n_modes]
    with zero constrained rows)` — reference sparse.py:868-902. `method`: None (policy above), "dense", "lobpcg"."""
    Kc, Mc = _as_csr(K), _as_csr(M)
    n = Kc.n
    free = free_indices.to(device=Kc.device, dtype=torch.int64)
    if method is None:
        method = "dense" if free.numel() <= DENSE_LIMIT else "lobpcg"
    if method == "dense":
        Kd = Kc.to_dense()[free][:, free]
        Md = Mc.to_dense()[free][:, free]
        Lc = torch.linalg.cholesky(0.5 * (Md + Md.T))
        Li = torch.linalg.solve_triangular(Lc, torch.eye(len(free), dtype=Kd.dtype, device=Kd.device), upper=False)
        ev, Z = torch.linalg.eigh(Li @ (0.5 * (Kd + Kd.T)) @ Li.T)
        vecs = torch.zeros(n, n_modes, dtype=torch.float64, device=Kc.device)
        vecs[free] = (Li.T @ Z[:, :n_modes])
        return ev[:n_modes].clone(), vecs
    if method != "lobpcg":
        raise ValueError(f"Method {method} is not supported. Choose from 'dense' or 'lobpcg'.")
    mask = torch.zeros(n, dtype=torch.float64, device=Kc.device)
    mask[free] = 1.0
    amg = AMGPreconditioner(Kc)

    def precondition(R: Tensor) -> Tensor:
        return amg.apply_block(R)

    lam, X, _ = lobpcg(Kc, Mc, mask, n_modes, precondition, tol=tol)
    return lam, X


class Eigensolve(Function):
    """Differentiable eigenvalues of K v = omega^2 M v for torch sparse K, M (reference sparse.py:905-982): only
    eigenvalue gradients, evaluated on the sparsity pattern of the respective matrix."""

    @staticmethod
    def forward(K, M, n_modes, free_indices):
        return modal_eigsolve(K, M, n_modes, free_indices)

    @staticmethod
    def setup_context(ctx, inputs, output) -> None:
        K, M, n_modes, free_indices = inputs
        lambdas, phis = output
        ctx.K, ctx.M = K, M
        ctx.save_for_backward(lambdas, phis)

    @staticmethod
    def backward(ctx, *grad_outputs):
        grad_lambdas = grad_outputs[0]
        K, M = ctx.K, ctx.M
        lambdas, phis = ctx.saved_tensors
        Mc = _as_csr(M)
        mask = torch.ones(Mc.n, dtype=torch.float64, device=phis.device)
        M_phis = _spmm(Mc, phis.contiguous(), mask)
        phi_hat = phis / (phis * M_phis).sum(0).abs().sqrt().unsqueeze(0)
        grad_K = grad_M = None
        if grad_lambdas is not None:
            for which, A in (("K", K), ("M", M)):
                if not (isinstance(A, Tensor) and A.requires_grad):
                    continue
                idx = A._indices()
                w = grad_lambdas if which == "K" else -(lambdas * grad_lambdas)
                vals = (phi_hat[idx[0]] * phi_hat[idx[1]] * w.unsqueeze(0)).sum(-1)
                with torch.sparse.check_sparse_tensor_invariants(False):
                    g = torch.sparse_coo_tensor(idx, vals.to(A.dtype), A.shape, is_coalesced=A.is_coalesced())
                if which == "K":
                    grad_K = g
                else:
                    grad_M = g
        return grad_K, grad_M, None, None


def differentiable_modal_eigsolve(K, M, n_modes: int, free_indices: Tensor):
    """(lambdas differentiable w.r.t. sparse K / M values, phis detached) — reference sparse.py:985-1011."""
    lambdas, phis = Eigensolve.apply(K, M, n_modes, free_indices)
    if lambdas is None:
        raise RuntimeError("Eigensolve.apply returned None.")
    return lambdas, phis.detach()


class ModesFromElements(Function):
    """omega^2 and mode shapes from ELEMENT matrices k, m (the route of `solve_modes`): the global matrices are
    assembled by the kernels (no autograd), the Rayleigh-quotient sensitivities are contracted per element,
    d lambda / dk_e = phi_e phi_e^T, d lambda / dm_e = -lambda phi_e phi_e^T — what the reference obtains by
    back-propagating the sparse gradients through `index_add_` (base.py:407-419 + sparse.py:944-982)."""

    @staticmethod
    def forward(ctx, k: Tensor, m: Tensor, model, n_modes: int):
        con = torch.nonzero(model.constraints.ravel(), as_tuple=False).ravel()
        free = torch.nonzero(~model.constraints.ravel(), as_tuple=False).ravel()
        K = model.assemble_matrix(k.detach(), con)
        M = model.assemble_matrix(m.detach(), con)
        lambdas, phis = modal_eigsolve(K, M, n_modes, free)
        ctx.save_for_backward(lambdas, phis)
        ctx.idx = model.idx
        ctx.needs = (k.requires_grad, m.requires_grad)
        ctx.mark_non_differentiable(phis)
        return lambdas, phis

    @staticmethod
    def backward(ctx, grad_lambdas, _grad_phis):
        lambdas, phis = ctx.saved_tensors
        pe = phis[ctx.idx.long()]                               # [n_elem, nd, n_modes]; M-normalised, zero at constraints
        gk = gm = None
        if grad_lambdas is not None:
            if ctx.needs[0]:
                gk = torch.einsum("eak,ebk,k->eab", pe, pe, grad_lambdas)
            if ctx.needs[1]:
                gm = -torch.einsum("eak,ebk,k->eab", pe, pe, grad_lambdas * lambdas)
        return gk, gm, None, None
