"""Host logic of torch-fem_b200/bordered.py on the CPU: the split of a symmetric matrix with a few long rows into
A11 (short rows), B, C reproduces `A x`; the bordered Jacobi-PCG takes the iterations of the oracle's CG (scipy's
algorithm) rounded up to the polling interval and reaches the same solution; and `Assembly` hands the split to
`sparse_solve` when `long_row_threshold` is set (opt-in). A11's products are the SELL SpMV kernel on the device; here a
host matrix stands in for it."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import load_case
from host_standins import HostMatrix, host_model
from oracle import fem_oracle as O


def _arrowhead(n=400, k=3, seed=0):
    """SPD: a banded matrix plus k dense rows / columns (what a reference point makes of a reduced tangent)."""
    rng = np.random.default_rng(seed)
    A = sp.diags([rng.uniform(-1, 0, n - 1), rng.uniform(-1, 0, n - 2)], [1, 2], format="lil")
    border = np.sort(rng.choice(n, size=k, replace=False))
    for r in border:
        cols = rng.choice(n, size=n // 2, replace=False)
        A[r, cols] = rng.uniform(-0.02, 0.02, len(cols))
    A = sp.csr_matrix(A)
    A = A + A.T
    A = (A + sp.diags(np.abs(A).sum(1).A1 + 1.0)).tocsr()     # diagonally dominant
    A.sort_indices()
    return A, border


def _host_csr(A):
    return HostMatrix(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)),
                      torch.from_numpy(A.data.copy()), A.shape[0], True)


def test_split_reproduces_the_operator_and_the_solve():
    from torchfem_b200.bordered import BorderedOperator, BorderSplit, bordered_pcg

    A, border = _arrowhead()
    n = A.shape[0]
    H = _host_csr(A)
    lengths = np.diff(A.indptr)
    assert set(np.nonzero(lengths > 100)[0]) == set(border)
    split = BorderSplit(H._indices(), n, torch.from_numpy(border))
    op = BorderedOperator(split, H._values(), HostMatrix)
    assert int((op.A11.indptr[1:] - op.A11.indptr[:-1]).max()) <= 5 + len(border)      # the long rows are gone
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal(n))
    assert np.abs(op.matvec(x).numpy() - A @ x.numpy()).max() <= 1e-13 * np.abs(A @ x.numpy()).max()
    # a second matrix on the same pattern reuses the structures
    op2 = BorderedOperator(split, 2.0 * H._values(), HostMatrix)
    assert op2.A11.indptr is op.A11.indptr
    assert np.abs(op2.matvec(x).numpy() - 2.0 * (A @ x.numpy())).max() <= 1e-12 * np.abs(A @ x.numpy()).max()

    b = rng.standard_normal(n)
    x_ref, flag, its_ref = O.jacobi_cg(A, b, rtol=1e-10)
    assert flag == 0
    xs, info = bordered_pcg(op, torch.from_numpy(b), rtol=1e-10, check_every=4)
    assert 0 <= info["iterations"] - its_ref < 4
    assert np.abs(xs.numpy() - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    assert np.linalg.norm(A @ xs.numpy() - b) <= 1e-10 * np.linalg.norm(b)
    # warm start from the solution: no iteration
    _, info0 = bordered_pcg(op, torch.from_numpy(b), rtol=1e-8, x0=xs)
    assert info0["iterations"] == 0
    with pytest.raises(RuntimeError, match="CG failed with exit code"):
        bordered_pcg(op, torch.from_numpy(b), rtol=1e-14, maxiter=3, check_every=2)


def test_assembly_hands_the_split_to_the_solver(monkeypatch):
    """A reference point driving a face of 49 nodes: with the long-row option on, the reduced matrix carries the split
    (the six point rows, 153 entries each against 81 for mesh rows) and the bordered PCG reproduces the solution of
    the dense solve."""
    import test_assembly_host_cpu as H
    import torchfem_b200 as T
    import torchfem_b200.assembly as A
    from host_standins import dense_sparse_solve
    from torchfem_b200.bordered import BorderedOperator, bordered_pcg
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    seen = []

    def bordered(K, b, B=None, stol=1e-10, device=None, method=None, M=None, x0=None):
        split = getattr(K, "border_split", None)
        assert split is not None
        seen.append(split.k)
        x, info = bordered_pcg(BorderedOperator(split, K._values(), HostMatrix), b.detach(), rtol=1e-13)
        return x, None

    monkeypatch.setattr(A.L, "require_cuda", lambda *t: None)
    monkeypatch.setattr(A, "spmv_plan", lambda *a: None)
    monkeypatch.setattr(A, "spgemm", H._spgemm)
    monkeypatch.setattr(A, "sell_structure", lambda *a: None)
    monkeypatch.setattr(A, "CSRMatrix", HostMatrix)
    monkeypatch.setattr(A._RectCSR, "matvec", H._host_matvec)

    nodes, elements = cube_hexa(7, 7, 3)
    results = []
    for threshold, solver in ((None, dense_sparse_solve), (100, bordered)):
        monkeypatch.setattr(T.sparse, "sparse_solve", solver)
        solid = host_model(T.Solid, nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
        solid.constraints[nodes[:, 2] == 0.0] = True
        point = A.ReferencePoint([0.5, 0.5, 2.0])
        point.forces[0, 3], point.forces[0, 5], point.forces[0, 0] = 50.0, -20.0, 10.0
        asm = A.Assembly([solid, point])
        asm.long_row_threshold = threshold
        asm.coupling(solid, nodes[:, 2] == 1.0, point)
        results.append(asm.solve())
        if threshold is None:
            assert getattr(asm._elimination._last[2], "border_split", None) is None      # off by default
    assert seen and all(k == 6 for k in seen)
    for a, b in zip(results[0][0] + results[0][1], results[1][0] + results[1][1]):
        assert float((a - b).abs().max()) <= 1e-9 * max(float(a.abs().max()), 1.0)
