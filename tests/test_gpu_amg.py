"""GPU parity of the aggregation-AMG kernels K11-K16 (csrc/amg.cu, through the C ABI) against oracle/amg_oracle.py,
level by level: aggregates and patterns bit-exact (integer work), operator values <= 1e-11 relative (different but
fixed summation orders), and the solution of AMG-PCG against the reference's golden displacement (<= 1e-8)."""
import numpy as np
import pytest
import torch

from conftest import load_case
from oracle import amg_oracle as M
from oracle import fem_oracle as O
from test_gpu_kernels import T, _cube_system, build_pattern, dev  # noqa: F401

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def same_pattern(a, b):
    a, b = a.tocsr(), b.tocsr()
    a.sort_indices(), b.sort_indices()
    return np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


def _compare_hierarchy(amg, lv_ref, tol=1e-11):
    assert amg.n_levels == len(lv_ref)
    for lv, ref in zip(amg.levels, lv_ref):
        assert lv.n == ref.n
        assert np.array_equal(lv.iso.cpu().numpy()[: lv.n].astype(bool), ref.iso)
        assert rel(lv.dinv.cpu().numpy()[: lv.n], ref.dinv) <= tol
        if ref is lv_ref[-1]:
            break
        assert abs(lv.rho - ref.rho) <= 1e-9 * ref.rho
        assert lv.agg_distance == ref.agg_distance
        assert np.array_equal(lv.agg.cpu().numpy()[: lv.op.nbr], ref.agg) and lv.n_agg == ref.n_agg
        P, R = lv.P.to_scipy(), lv.R.to_scipy()
        Pd, Pr = P.toarray(), ref.P.toarray()
        assert np.abs(Pd - Pr).max() <= tol * np.abs(Pr).max()
        assert np.array_equal(R.toarray(), Pd.T)                     # the transpose is a pure permutation
    for lv, ref in zip(amg.levels[1:], lv_ref[1:]):
        Ad, Ar = lv.op.to_scipy().toarray(), ref.A.toarray()
        assert np.abs(Ad - Ar).max() <= tol * np.abs(Ar).max()
    inv = amg.levels[-1].inv.cpu().numpy()
    assert np.abs(inv - lv_ref[-1].inv).max() <= 1e-8 * np.abs(lv_ref[-1].inv).max()


@pytest.mark.parametrize("layout", ["auto", "sell"])
@pytest.mark.parametrize("N,max_coarse", [(7, 150), (11, 300)])
def test_hierarchy_matches_oracle_mechanics(T, tables, N, max_coarse, layout, monkeypatch):
    from torchfem_b200 import amg as amg_mod
    from torchfem_b200.amg import AMGPreconditioner

    if layout == "sell":   # force SELL-32 for every operator (small problems default to block CSR on coarse levels)
        monkeypatch.setattr(amg_mod, "BCSR_MAX_ROWS", 0)
        monkeypatch.setattr(amg_mod, "BCSR_MIN_AVG", 10**9)

    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, N, tables)
    A = p.matrix(A.values_)
    amg = AMGPreconditioner(A, max_coarse=max_coarse)
    A_ref = O.to_csr(A.values_.cpu().numpy(), p.glob_idx.cpu().numpy(), p.n_dofs)
    lv_ref = M.build_hierarchy(A_ref, 3, max_coarse=max_coarse)
    assert amg.n_levels >= 2
    assert amg.levels[0].P.use_bcsr == (layout == "auto")
    _compare_hierarchy(amg, lv_ref)
    # one V cycle
    r = np.random.default_rng(0).standard_normal(p.n_dofs)
    z = amg.apply(dev(r)).cpu().numpy()
    z_ref = M.vcycle(lv_ref, r)
    assert rel(z, z_ref) <= 1e-10
    assert np.array_equal(z, amg.apply(dev(r)).cpu().numpy())      # bitwise reproducible


@pytest.mark.parametrize("layout", ["auto", "sell"])
@pytest.mark.parametrize("N,max_coarse,m", [(7, 150, 3), (11, 300, 4), (11, 300, 9), (4, 10**6, 5)])
def test_block_vcycle_equals_single_cycles(T, tables, N, max_coarse, m, layout, monkeypatch):
    """`apply_block` (the eigensolver's preconditioner step: finest-level sweeps on 4 vectors per pass over the matrix) ==
    one `apply` per column, bit for bit; 2- and 3-level hierarchies, block-CSR and SELL coarse operators, and the
    single-level (dense) case."""
    from torchfem_b200 import amg as amg_mod
    from torchfem_b200.amg import AMGPreconditioner

    if layout == "sell":
        monkeypatch.setattr(amg_mod, "BCSR_MAX_ROWS", 0)
        monkeypatch.setattr(amg_mod, "BCSR_MIN_AVG", 10**9)
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, N, tables)
    A = p.matrix(A.values_)
    amg = AMGPreconditioner(A, max_coarse=max_coarse)
    assert (amg.n_levels == 1) == (max_coarse == 10**6)
    R = dev(np.random.default_rng(m).standard_normal((p.n_dofs, m)))
    Z = amg.apply_block(R)
    assert Z.shape == R.shape
    for j in range(m):
        assert torch.equal(Z[:, j], amg.apply(R[:, j].contiguous()))


def _spd_on_pattern(c, n, seed=3):
    """A symmetric, strictly diagonally dominant matrix (SPD) with random entries on the pattern of a fixture
    (the fixtures' own K are unconstrained / geometrically nonlinear tangents, not all positive definite)."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    row, col = c["glob_idx"]
    B = sp.csr_matrix((rng.uniform(0.1, 1.0, len(row)), (row, col)), shape=(n, n))
    W = ((B + B.T) * 0.5).tocsr()
    W.setdiag(0.0)                                  # keeps the stored diagonal as an explicit zero
    off = np.asarray(W.sum(axis=1)).ravel()
    A = (-W).tocsr()
    A.setdiag(off + 1.0)
    A.sort_indices()
    assert np.array_equal(A.indices, col) and A.nnz == len(col)
    return A


@pytest.mark.parametrize("tag,max_coarse", [("heat_hexa1", 10), ("heat_hexa1", 100), ("quad1", 12), ("tetra2", 40),
                                            ("hexa2", 40), ("hexa1_orphan", 20), ("heat_tetra2", 20), ("tetra1", 20),
                                            ("tria1", 10)])
@pytest.mark.parametrize("layout", ["auto", "sell"])
def test_hierarchy_other_block_sizes(T, tag, max_coarse, layout, monkeypatch):
    """d = 1 (heat, and the scalar fall-back for patterns with unreferenced nodes), d = 2 (planar), longer rows, and a
    single-level hierarchy (max_coarse above the size: dense solve only)."""
    from torchfem_b200 import amg as amg_mod
    from torchfem_b200.amg import AMGPreconditioner

    if layout == "sell":
        monkeypatch.setattr(amg_mod, "BCSR_MAX_ROWS", 0)
        monkeypatch.setattr(amg_mod, "BCSR_MIN_AVG", 10**9)
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    A_ref = _spd_on_pattern(c, n)
    A = p.matrix(dev(np.ascontiguousarray(A_ref.data)))
    d = dpn if (dpn in (2, 3) and tag != "hexa1_orphan") else 1
    amg = AMGPreconditioner(A, max_coarse=max_coarse)
    assert amg.levels[0].d == d
    lv_ref = M.build_hierarchy(A_ref, d, max_coarse=max_coarse)
    assert (amg.n_levels == 1) == (n <= max_coarse)
    _compare_hierarchy(amg, lv_ref)
    b = np.random.default_rng(2).standard_normal(n)
    z = amg.apply(dev(b)).cpu().numpy()
    assert rel(z, M.vcycle(lv_ref, b)) <= 1e-10
    x, info = amg.solve(dev(b), rtol=1e-10)
    x_ref, _, its_ref = M.amg_pcg(A_ref, b, lv_ref, rtol=1e-10)
    assert abs(info["iterations"] - its_ref) <= 1
    assert np.linalg.norm(A_ref @ x.cpu().numpy() - b) <= 1e-9 * np.linalg.norm(b)
    assert rel(x.cpu().numpy(), x_ref) <= 1e-8


@pytest.mark.parametrize("aggregation", ["mis1", "mis2"])
def test_aggregation_variants_match_oracle(T, tables, aggregation):
    """Radius-1 and radius-2 aggregates forced on the Hexa1 cube (the automatic rule picks radius 2 only where
    radius 1 leaves fewer than 6 nodes per aggregate, e.g. on the Tetra1 fixture above)."""
    from torchfem_b200.amg import AMGPreconditioner

    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 9, tables)
    A = p.matrix(A.values_)
    amg = AMGPreconditioner(A, max_coarse=100, aggregation=aggregation)
    A_ref = O.to_csr(A.values_.cpu().numpy(), p.glob_idx.cpu().numpy(), p.n_dofs)
    lv_ref = M.build_hierarchy(A_ref, 3, max_coarse=100, aggregation=aggregation)
    assert amg.levels[0].agg_distance == (1 if aggregation == "mis1" else 2)
    _compare_hierarchy(amg, lv_ref)
    r = np.random.default_rng(1).standard_normal(p.n_dofs)
    assert rel(amg.apply(dev(r)).cpu().numpy(), M.vcycle(lv_ref, r)) <= 1e-10


def test_amgx_method_reproduces_reference_solution_config_a(T, tables):
    """sparse_solve(method="amgx") — the reference's GPU AMG method (sparse.py:422-442) served by the in-house
    kernels — against the golden displacement of the unmodified reference (config A)."""
    g = load_case("config_a.npz")
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    A = p.matrix(A.values_)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    b = dev(ref["res"])
    x, Mp = T.sparse.sparse_solve(A, b, stol=1e-10, method="amgx")
    con = np.nonzero(con_mask.ravel())[0]
    u = -x.cpu().numpy()
    u[con] = disp.ravel()[con]
    assert np.linalg.norm(u.reshape(-1, 3) - g["u"]) <= 1e-8 * np.linalg.norm(g["u"])
    # iteration count equals the oracle's; far fewer than Jacobi-CG
    lv_ref = M.build_hierarchy(ref["A"], 3)
    _, _, its_ref = M.amg_pcg(ref["A"], ref["res"], lv_ref, rtol=1e-10)
    x2, st = Mp.solve(b, rtol=1e-10)
    assert abs(st["iterations"] - its_ref) <= 1 and st["iterations"] < ref["iterations"] // 2
    assert torch.equal(x2, x)
    # passing M back = coefficient refresh on the stored patterns (AmgX resetup): scaled matrix, same solution / 2
    A2 = p.matrix(A.values_ * 2.0)
    n_levels = Mp.n_levels
    x3, M3 = T.sparse.sparse_solve(A2, b, stol=1e-10, method="amgx", M=Mp)
    assert M3 is Mp and Mp.n_levels == n_levels
    assert float((x3 - 0.5 * x).abs().max()) <= 1e-8 * float(x.abs().max())
    # warm start from the solution: converged at once
    x4, st4 = Mp.solve(b, x0=x3, rtol=1e-8)
    assert st4["iterations"] == 0 and torch.equal(x4, x3)


@pytest.mark.parametrize("d,nx,m,ny,px,py", [(2, 300, 200, 150, 9, 7), (3, 40, 600, 4000, 48, 40), (1, 64, 300, 9000, 60, 90)])
def test_spgemm_against_scipy(T, d, nx, m, ny, px, py):
    """K15 on random block operators: pattern bit-exact, values to round-off. The second and third shapes have product
    rows of 1,000-3,000 distinct columns: the symbolic kernel overflows its 1,024-slot hash set and reruns with the
    16,384-slot one, the numeric kernel takes the CTA-per-row variant."""
    import scipy.sparse as sp
    from torchfem_b200.amg import BlockOperator, spgemm

    rng = np.random.default_rng(5)

    def rand_block(nr, nc, per_row):
        ptr, col, val = [0], [], []
        for _ in range(nr):
            c = np.sort(rng.choice(nc, size=rng.integers(1, per_row + 1), replace=False))
            col.extend(c)
            ptr.append(len(col))
        ptr, col = np.array(ptr, dtype=np.int64), np.array(col, dtype=np.int32)
        vals = rng.standard_normal(d * d * len(col))
        return BlockOperator(d, nr, nc, dev(ptr), dev(col), dev(vals))

    X, Y = rand_block(nx, m, px), rand_block(m, ny, py)
    C, _ = spgemm(d, X, Y)
    ref = (X.to_scipy() @ Y.to_scipy()).tocsr()
    got = C.to_scipy()
    # structural product pattern (no numerical cancellation in random data)
    assert same_pattern(got, ref)
    assert np.abs(got.toarray() - ref.toarray()).max() <= 1e-13 * np.abs(ref.toarray()).max()


def test_full_size_property_amg_beats_jacobi(T, tables):
    """Size-independent property at 64^3 elements (823,875 DOFs): AMG-PCG reaches the same true residual as
    Jacobi-PCG in far fewer iterations, and both solutions agree."""
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 65, tables)
    del k
    A = p.matrix(A.values_)
    ubc = dev((disp * con_mask).ravel())
    Af = T.csr.assemble(p, T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(bref), torch.as_tensor(w), dev(nodes),
                                             dev(elements), dev(C)), None)
    b = p.matrix(Af).matvec(ubc)
    b[dev(con_mask.ravel())] = 0.0
    xj, _, sj = T.csr.krylov_solve(A, b, method="cg", rtol=1e-8)
    xa, Mp = T.sparse.sparse_solve(A, b, stol=1e-8, method="amgx")
    xa2, sa = Mp.solve(b, rtol=1e-8)
    assert sa["iterations"] * 5 < sj["iterations"]
    res = float((A.matvec(xa) - b).norm() / b.norm())
    assert res <= 1.5e-8
    assert float((xa - xj).norm() / xj.norm()) <= 1e-6
    assert Mp.operator_complexity < 1.6


def test_auto_selected_amg_falls_back_to_minres_on_failure(T, tables, monkeypatch):
    """Size policy picks AMG; if AMG-PCG fails (indefinite / singular system) the solve is retried with the reference's
    default Krylov method instead of failing — but an explicitly requested method="amgx" still raises."""
    from torchfem_b200.amg import AMGPreconditioner

    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    A = p.matrix(A.values_)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    b = dev(ref["res"])
    monkeypatch.setattr(T.sparse, "AMG_MIN_DOFS", 1000)
    monkeypatch.setattr(T.sparse, "DIRECT_LIMIT", 100)
    assert T.sparse.resolve_method(A.n, "cuda", None) == "amgx"
    x_amg, M1 = T.sparse.sparse_solve(A, b, stol=1e-10)
    assert isinstance(M1, AMGPreconditioner)

    def boom(self, *a, **k):
        raise RuntimeError("CG failed with exit code -1")

    monkeypatch.setattr(AMGPreconditioner, "solve", boom)
    x_mr, M2 = T.sparse.sparse_solve(A, b, stol=1e-10)
    assert not isinstance(M2, AMGPreconditioner)
    assert float((x_mr - x_amg).norm() / x_amg.norm()) <= 1e-8
    with pytest.raises(RuntimeError, match="CG failed"):
        T.sparse.sparse_solve(A, b, stol=1e-10, method="amgx")
