"""Host logic of torch-fem_b200/modal.py on the CPU (no kernel call): the LOBPCG driver with scipy products standing in
for the SpMV kernels and the numpy AMG oracle as preconditioner, against scipy's shift-invert `eigsh` — the routine
the reference's `modal_eigsolve` calls (src/torchfem/sparse.py:820)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from oracle import amg_oracle as AM
from oracle import fem_oracle as O


class _ScipyOperator:
    def __init__(self, A):
        self.A, self.n = A.tocsr(), A.shape[0]

    def matvec(self, x, fmt=None):
        return torch.from_numpy(self.A @ x.numpy())

    def matmat(self, X):
        return torch.from_numpy(self.A @ X.numpy())


def test_lobpcg_with_amg_preconditioner_matches_shift_invert_lanczos():
    from torchfem_b200.modal import lobpcg

    nodes, elements = O.cube_hexa(9, 9, 9)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con, disp = O.cube_extension_bcs(nodes)
    r = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con, disp, rtol=1e-8)
    K = r["A"]
    n = K.shape[0]
    free = ~con.ravel()
    md = 1.0 + 0.3 * np.sin(np.arange(n))          # a lumped (diagonal) mass, identity on the constrained DOFs
    md[~free] = 1.0
    M = sp.diags(md).tocsr()
    lv = AM.build_hierarchy(K, 3, max_coarse=300)

    def precondition(R):
        return torch.from_numpy(np.stack([AM.vcycle(lv, R[:, j].numpy()) for j in range(R.shape[1])], axis=1))

    mask = torch.from_numpy(free.astype(np.float64))
    lam, X, its = lobpcg(_ScipyOperator(K), _ScipyOperator(M), mask, 5, precondition, tol=1e-7)
    fi = np.nonzero(free)[0]
    ref = np.sort(spla.eigsh(K[fi][:, fi].tocsc(), k=5, M=M[fi][:, fi].tocsc(), sigma=0.0, return_eigenvectors=False))
    assert its < 100
    assert np.allclose(lam.numpy(), ref, rtol=1e-9)
    Xn = X.numpy()
    assert np.abs(Xn[~free]).max() == 0.0                                        # constrained rows stay exactly zero
    assert np.allclose(Xn.T @ (M @ Xn), np.eye(5), atol=1e-8)                   # M-orthonormal
    assert np.abs(K @ Xn - (M @ Xn) * lam.numpy()).max() <= 1e-5 * np.abs(K @ Xn).max()
