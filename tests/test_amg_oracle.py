"""CPU checks of oracle/amg_oracle.py (the restatement the AMG kernels are compared with): the hierarchy is a valid
SPD preconditioner and AMG-PCG reproduces the reference's solution (golden `u` of config A generated from the
unmodified reference, tests/golden/config_a.npz) at the tolerance the reference tests use (tests/test_sparse.py:49-91)."""
import numpy as np
import scipy.sparse as sp

from conftest import load_case
from oracle import amg_oracle as M
from oracle import fem_oracle as O


def _system(N):
    nodes, elements = O.cube_hexa(N, N, N)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con, disp = O.cube_extension_bcs(nodes)
    r = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con, disp, rtol=1e-10)
    return r, con, disp


def test_aggregates_are_a_partition_of_radius_one():
    r, _, _ = _system(8)
    ptr, adj, _ = M.block_graph(r["A"], 3)
    agg, n_agg, rounds = M.mis_aggregate(ptr, adj)
    assert agg.min() == 0 and agg.max() == n_agg - 1 and len(np.unique(agg)) == n_agg
    assert rounds < 20
    # every aggregate has a root adjacent to all of its members: the aggregate id of the root is shared
    G = sp.csr_matrix((np.ones(len(adj)), adj, ptr))
    same = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)))     # node x aggregate
    reach = (G @ same).tocsr()                                                # aggregates adjacent to a node
    assert all(reach[i, agg[i]] > 0 for i in range(len(agg)))


def test_vcycle_is_symmetric_positive_definite():
    r, _, _ = _system(7)
    lv = M.build_hierarchy(r["A"], 3, max_coarse=200)
    assert len(lv) >= 2
    n = r["A"].shape[0]
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    Mu, Mv = M.vcycle(lv, u), M.vcycle(lv, v)
    assert abs(v @ Mu - u @ Mv) <= 1e-12 * abs(v @ Mu)
    assert u @ Mu > 0 and v @ Mv > 0


def test_amg_pcg_reproduces_the_reference_solution_of_config_a():
    g = load_case("config_a.npz")
    r, con, disp = _system(11)
    lv = M.build_hierarchy(r["A"], 3)
    x, info, its = M.amg_pcg(r["A"], r["res"], lv, rtol=1e-10)
    assert info == 0 and its < 40 < r["iterations"]          # Jacobi-CG needs 100+ iterations here
    u = -x
    c = np.nonzero(con.ravel())[0]
    u[c] = disp.ravel()[c]
    assert np.linalg.norm(u.reshape(-1, 3) - g["u"]) <= 1e-8 * np.linalg.norm(g["u"])


def test_isolated_rows_stay_out_of_the_coarse_space():
    r, con, _ = _system(6)
    lv = M.build_hierarchy(r["A"], 3, max_coarse=100)
    L0 = lv[0]
    assert np.array_equal(L0.iso, con.ravel())
    # prolongator rows of Dirichlet DOFs are empty, coarse operators keep a non-zero diagonal
    P = L0.P.tocsr()
    assert abs(P[np.nonzero(L0.iso)[0]]).sum() == 0.0
    assert all((L.A.diagonal() != 0).all() for L in lv)


def test_block_operator_layout_matches_scalar_csr():
    """Host logic of torch-fem_b200/amg.py (no kernel call): the block-CSR value layout (entry (a, s, c) of block row
    I at d*d*bptr[I] + (a*m + s)*d + c) is the scalar CSR order of the assembled matrix, so `indptr`, `to_scipy` and
    the dense copy agree with scipy's CSR of the same matrix."""
    import torch

    from torchfem_b200.amg import BlockOperator, _dense

    r, _, _ = _system(4)
    A = r["A"].tocsr()
    A.sort_indices()
    d = 3
    ptr, adj, _ = M.block_graph(A, d)
    op = BlockOperator(d, len(ptr) - 1, len(ptr) - 1, torch.as_tensor(ptr), torch.as_tensor(adj),
                       torch.as_tensor(A.data.copy()))
    assert np.array_equal(op.indptr.numpy(), A.indptr)
    assert abs(op.to_scipy() - A).max() == 0.0
    assert np.array_equal(_dense(op).numpy(), A.toarray())
    assert op.use_bcsr        # few rows: the coarse-level layout
