"""Jacobi-PCG on a symmetric matrix with a few very long rows, kept out of the SELL-32 copy ("bordered" solve).

Why: a reference point that drives a whole face (reference assembly.py:180-246) puts rows with thousands of entries
into the reduced tangent `T^T K T`. SELL-32 pads the slice of such a row to its length and one warp walks it alone:
4.9 ms per CG iteration instead of 0.25 ms at 1.35 M unknowns. Here the matrix is split once,

        A = [ A11  B ]      A11: every short row (the long rows reduced to a unit diagonal), streamed by kernel K5
            [ B^T  C ]      B  : the long rows' entries as a dense [n, k] block, C their k x k coupling

and `y = A x` becomes one SELL SpMV plus two dense products with k columns. The Krylov recurrences (the algorithm of
scipy `cg` with the preconditioner diag(A)^-1, reference sparse.py:406-419, same stopping rule
`||r|| < max(atol, rtol ||b||)`, `maxiter = 10 n`) run as torch operations on the device, polled every
`check_every` iterations.

STATUS: opt-in (`Assembly.long_row_threshold`); the host logic is checked on the CPU (tests/test_bordered_cpu.py), the
GPU run has not been measured yet — the long-row side path belongs inside the SELL kernels.
"""
from __future__ import annotations

import torch
from torch import Tensor


class BorderSplit:
    """Pattern-level part of the split: which CSR entries go where. Built once per pattern from its COO indices
    (`idx` int64 [2, nnz], row-major sorted) and the long rows `border` (int64 [k], ascending)."""

    def __init__(self, idx: Tensor, n: int, border: Tensor):
        dev = idx.device
        row, col = idx[0], idx[1]
        self.n, self.k = int(n), int(border.shape[0])
        self.border = border
        slot = torch.full((n,), -1, dtype=torch.int64, device=dev)
        slot[border] = torch.arange(self.k, device=dev)
        in_row, in_col = slot[row] >= 0, slot[col] >= 0
        diagonal = row == col
        # A11: short rows entirely (their entries in border columns are zeroed), long rows as a unit diagonal
        self.keep = torch.nonzero(~in_row | diagonal).ravel()
        kept_row, kept_col = row[self.keep], col[self.keep]
        self.zeroed = torch.nonzero(in_col[self.keep] & ~diagonal[self.keep]).ravel()
        self.unit = torch.nonzero(in_row[self.keep]).ravel()
        self.indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        self.indptr[1:] = torch.cumsum(torch.bincount(kept_row, minlength=n), 0)
        self.indices = kept_col.to(torch.int32).contiguous()
        # B: entries (short row, border column); C: entries (border row, border column)
        b = torch.nonzero(in_col & ~in_row).ravel()
        self.b_src, self.b_row, self.b_slot = b, row[b], slot[col[b]]
        c = torch.nonzero(in_col & in_row).ravel()
        self.c_src, self.c_row, self.c_col = c, slot[row[c]], slot[col[c]]
        self.diag_src = torch.nonzero(diagonal).ravel()
        self.diag_row = row[self.diag_src]
        self.interior_template = None   # first A11 (a CSRMatrix); later ones share its structures through `_like`


class BorderedOperator:
    """`y = A x` through the split; `matrix_type(indptr, indices, values, n, symmetric=True)` builds A11 (the
    package's `CSRMatrix` on the device)."""

    def __init__(self, split: BorderSplit, values: Tensor, matrix_type):
        s = self.split = split
        v11 = values[s.keep].clone()
        v11[s.zeroed] = 0.0
        v11[s.unit] = 1.0
        if s.interior_template is None:
            s.interior_template = matrix_type(s.indptr, s.indices, v11, s.n, symmetric=True)
            self.A11 = s.interior_template
        else:
            self.A11 = s.interior_template._like(v11)
        self.B = torch.zeros(s.n, s.k, dtype=values.dtype, device=values.device)
        self.B[s.b_row, s.b_slot] = values[s.b_src]
        self.C = torch.zeros(s.k, s.k, dtype=values.dtype, device=values.device)
        self.C[s.c_row, s.c_col] = values[s.c_src]
        self.diagonal = torch.zeros(s.n, dtype=values.dtype, device=values.device)
        self.diagonal[s.diag_row] = values[s.diag_src]

    def matvec(self, x: Tensor) -> Tensor:
        s = self.split
        xp = x[s.border]
        y = self.A11.matvec(x, fmt="sell") + self.B @ xp     # rows of B at the border are zero
        y[s.border] = self.B.T @ x + self.C @ xp            # B^T ignores x at the border for the same reason
        return y


def bordered_pcg(op: BorderedOperator, b: Tensor, rtol: float = 1e-10, atol: float = 0.0, x0: Tensor | None = None,
                 maxiter: int = 0, check_every: int = 16):
    """Jacobi-preconditioned CG on the bordered operator. Returns (x, info) with info = {"iterations", "resnorm",
    "bnorm", "converged"}; raises RuntimeError("CG failed with exit code ...") like `csr.krylov_solve`."""
    n = b.shape[0]
    maxiter = int(maxiter) if maxiter else 10 * n
    dinv = 1.0 / op.diagonal
    bnorm = float(torch.linalg.norm(b))
    tol = max(float(atol), float(rtol) * bnorm)
    x = torch.zeros_like(b) if x0 is None else x0.to(b).clone()
    r = b - op.matvec(x) if x0 is not None else b.clone()
    rnorm = float(torch.linalg.norm(r))
    it = 0
    if rnorm >= tol:
        z = dinv * r
        p = z.clone()
        rz = torch.dot(r, z)
        while it < maxiter:
            for _ in range(min(check_every, maxiter - it)):
                q = op.matvec(p)
                alpha = rz / torch.dot(p, q)
                x = x + alpha * p
                r = r - alpha * q
                z = dinv * r
                rz_new = torch.dot(r, z)
                p = z + (rz_new / rz) * p
                rz = rz_new
                it += 1
            rnorm = float(torch.linalg.norm(r))     # the one host synchronisation per batch
            if not rnorm == rnorm:                   # NaN: breakdown (indefinite or singular matrix)
                raise RuntimeError("CG failed with exit code -1")
            if rnorm < tol:
                break
    info = {"iterations": it, "resnorm": rnorm, "bnorm": bnorm, "converged": rnorm < tol or bnorm == 0.0}
    if not info["converged"]:
        raise RuntimeError(f"CG failed with exit code {it}")
    return x, info
