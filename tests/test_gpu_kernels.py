"""GPU parity: the sm_100a kernels, called through the C ABI (torchfem_b200.csr -> libtfem_b200.so),
against the CPU oracle and the fixtures generated from the unmodified reference.

Bars (BASELINE.json): integer CSR structure bit-exact; element matrices and assembled K <= 1e-12
relative (max|d| / max|ref| per tensor); displacements <= 1e-8 relative at equal tolerance.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import HEAT_CASES, MECH_CASES, load_case
from oracle import fem_oracle as O

pytestmark = pytest.mark.gpu

TOL_K = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def T():
    import torchfem_b200 as T

    assert torch.cuda.is_available()
    return T


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    return t.to(dtype) if dtype is not None else t


def build_pattern(T, c, dpn):
    return T.csr.Pattern(dev(c["elements"]), c["nodes"].shape[0], dpn)


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_pattern_bit_exact(T, tag):
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    p = build_pattern(T, c, dpn)
    assert p.nnz == c["glob_idx"].shape[1]
    g = p.glob_idx.cpu().numpy()
    assert g.dtype == np.int64 and np.array_equal(g, c["glob_idx"])
    km = p.k_map.cpu().numpy()
    assert km.dtype == np.int32 and np.array_equal(km, c["k_map"])
    dm = p.diag_map.cpu().numpy()
    assert dm.dtype == np.int32 and np.array_equal(dm, c["diag_map"])
    indptr, indices = O.csr_from_glob_idx(c["glob_idx"], dpn * c["nodes"].shape[0])
    assert np.array_equal(p.indptr.cpu().numpy(), indptr)
    assert np.array_equal(p.indices.cpu().numpy(), indices)


@pytest.mark.parametrize("tag", MECH_CASES)
def test_integrate_k_mech(T, tag, tables):
    c = load_case(f"case_{tag}.npz")
    et = str(c["etype"])
    th = dev(c["thickness"]) if "thickness" in c else None
    k = T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(tables[f"{et}.B_ip"]),
                          torch.as_tensor(tables[f"{et}.iweights"]), dev(c["nodes"]),
                          dev(c["elements"]), dev(c["C"]), th)
    assert rel(k.cpu().numpy(), c["k"]) <= TOL_K


@pytest.mark.parametrize("tag", HEAT_CASES)
def test_integrate_k_heat(T, tag, tables):
    c = load_case(f"case_{tag}.npz")
    et = str(c["etype"])
    th = dev(c["thickness"]) if "thickness" in c else None
    k = T.csr.integrate_k(T._lib.KIND_HEAT, torch.as_tensor(tables[f"{et}.B_ip"]),
                          torch.as_tensor(tables[f"{et}.iweights"]), dev(c["nodes"]),
                          dev(c["elements"]), dev(c["kappa"]), th)
    assert rel(k.cpu().numpy(), c["k"]) <= TOL_K


def test_integrate_k_per_gauss_point_tangent(T, tables):
    c = load_case("case_hyper_hexa1.npz")
    k = T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(tables["Hexa1.B_ip"]),
                          torch.as_tensor(tables["Hexa1.iweights"]), dev(c["nodes"]),
                          dev(c["elements"]), dev(c["C"]))
    assert rel(k.cpu().numpy(), c["k"]) <= TOL_K


def test_negative_jacobian_raises(T, tables):
    c = load_case("case_hexa1.npz")
    el = c["elements"].copy()
    el[3] = el[3][[1, 0, 3, 2, 5, 4, 7, 6]]
    with pytest.raises(ValueError, match="Negative Jacobian. Check element numbering."):
        T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(tables["Hexa1.B_ip"]),
                          torch.as_tensor(tables["Hexa1.iweights"]), dev(c["nodes"]), dev(el),
                          dev(c["C"]))


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_assemble(T, tag):
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n_dofs = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    is_con = np.zeros(n_dofs, dtype=np.uint8)
    is_con[c["con"]] = 1
    vals = T.csr.assemble(p, dev(c["k"]), dev(is_con))
    v = vals.cpu().numpy()
    assert rel(v, c["K_val"]) <= TOL_K
    assert np.array_equal(v == 1.0, c["K_val"] == 1.0)
    masked = (is_con[c["glob_idx"][0]] | is_con[c["glob_idx"][1]]).astype(bool)
    assert np.all((v[masked] == 0.0) | (v[masked] == 1.0))
    # bitwise run-to-run determinism (no FP atomics)
    v2 = T.csr.assemble(p, dev(c["k"]), dev(is_con)).cpu().numpy()
    assert np.array_equal(v, v2)
    # unconstrained assembly (reference passes con=EMPTY, assembly.py:19,509)
    v3 = T.csr.assemble(p, dev(c["k"]), None).cpu().numpy()
    ref3 = O.assemble_values(c["k"], c["k_map"], c["glob_idx"], c["diag_map"],
                             np.zeros(0, dtype=np.int64), n_dofs)
    assert rel(v3, ref3) <= TOL_K
    # Dirichlet lifting fused into the assembly == K_unconstrained @ u_bc on the free rows (base.py:708-741)
    rng = np.random.default_rng(3)
    ubc = rng.standard_normal(n_dofs)
    lift = torch.full((n_dofs,), np.nan, dtype=torch.float64, device="cuda")
    v4 = T.csr.assemble(p, dev(c["k"]), dev(is_con), ubc=dev(ubc), lift=lift).cpu().numpy()
    assert np.array_equal(v4, v)
    A3 = O.to_csr(ref3, c["glob_idx"], n_dofs)
    want = A3 @ (ubc * is_con)
    want[is_con.astype(bool)] = 0.0
    got = lift.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    assert np.all(got[is_con.astype(bool)] == 0.0)


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_assemble_writes_solver_order(T, tag):
    """The assembly's SELL-32 values and 1/diagonal == the CSR -> SELL copy and the Jacobi setup of its CSR values,
    bit for bit (padding and the rows past the last one written as 0), with and without the CSR output."""
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n_dofs = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    is_con = np.zeros(n_dofs, dtype=np.uint8)
    is_con[c["con"]] = 1
    k = dev(c["k"])
    vals = T.csr.assemble(p, k, dev(is_con))
    A = p.matrix(vals)
    want_sell = A.sell()                      # tfem_sell_fill of the CSR values
    want = A._sell_vals.cpu().numpy()
    want_dinv = T.csr.JacobiPreconditioner(A).dinv.cpu().numpy()
    st = p.sell_structure
    # both are pinned by the numpy restatement of the layout (oracle/fem_oracle.py::sell32_values)
    sp_ref, sell_ref = O.sell32_values(p.indptr.cpu().numpy(), vals.cpu().numpy())
    assert np.array_equal(st.slice_ptr.cpu().numpy(), sp_ref) and st.padded == len(sell_ref)
    assert np.array_equal(want[: st.padded], sell_ref)
    with np.errstate(divide="ignore"):     # an unreferenced, unconstrained node has a zero diagonal (base.py:419)
        assert np.array_equal(want_dinv, 1.0 / vals.cpu().numpy()[c["diag_map"]])
    for csr_too in (True, False):
        sv = torch.full((max(st.padded, 2),), np.nan, dtype=torch.float64, device="cuda")
        dinv = torch.full((n_dofs,), np.nan, dtype=torch.float64, device="cuda")
        v, sv2, dinv2 = T.csr.assemble(p, k, dev(is_con), csr=csr_too, sell_out=sv, dinv_out=dinv)
        assert sv2 is sv and dinv2 is dinv
        assert (v is None) == (not csr_too)
        if csr_too:
            assert torch.equal(v, vals)
        assert np.array_equal(sv.cpu().numpy()[: st.padded], want[: st.padded])
        assert np.array_equal(dinv.cpu().numpy(), want_dinv)
    # the solver-only matrix: same product and same preconditioner bit for bit, hence the same Krylov solve
    rng = np.random.default_rng(5)
    _, sv, dinv = T.csr.assemble(p, k, dev(is_con), csr=False, sell_out=True, dinv_out=True)
    S = p.matrix(None, sell_vals=sv)
    xv = dev(rng.standard_normal(n_dofs))
    assert torch.equal(S.matvec(xv), A.matvec(xv, fmt="sell"))
    M = T.csr.JacobiPreconditioner(dinv=dinv)
    assert M.shape == (n_dofs, n_dofs) and torch.equal(M.dinv, T.csr.JacobiPreconditioner(A).dinv)
    with pytest.raises(ValueError, match="assembled for the solver only"):
        S.values_
    del want_sell


def test_assemble_solver_order_argument_errors(T):
    c = load_case("case_hexa1.npz")
    p = build_pattern(T, c, 3)
    k = dev(c["k"])
    with pytest.raises(ValueError, match="nothing to write"):
        T.csr.assemble(p, k, None, csr=False)
    with pytest.raises(ValueError, match="sell_out must be a contiguous float64 tensor"):
        T.csr.assemble(p, k, None, sell_out=torch.empty(8, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError, match="needs is_con and ubc"):
        T.csr.assemble(p, k, None, lift=torch.empty(p.n_dofs, dtype=torch.float64, device="cuda"), sell_out=True)
    # the C entry point itself refuses a call without any output
    L = T._lib
    rc = L.lib.tfem_assemble_solve(p.n_nod, p.nn, p.dpn, L.ptr(p.node_ptr), L.ptr(p.adj), L.ptr(p.indptr),
                                   L.ptr(p.src_ptr), L.ptr(p.src), L.ptr(k), None, None, None, None, None, None, None,
                                   L.stream())
    assert rc == L.ERR_INVALID


@pytest.mark.parametrize("tag", ["hexa1", "tetra2", "quad2", "heat_hexa1", "hexa1_orphan"])
def test_assemble_rhs_deterministic_gather(T, tag):
    """a9 `assemble_rhs` (reference base.py:428-445) on the gather kernel: equals the oracle's index_add to round-off,
    two runs are bitwise equal, the backward is the transposed gather."""
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    p = build_pattern(T, c, dpn)
    nd = c["elements"].shape[1] * dpn
    f = np.random.default_rng(3).standard_normal((len(c["elements"]), nd))
    idx = O.dof_map(c["elements"], dpn)
    ref = O.assemble_rhs(f, idx, dpn * c["nodes"].shape[0])
    fd = dev(f).requires_grad_(True)
    F = T.csr.assemble_rhs(p, fd)
    assert np.abs(F.detach().cpu().numpy() - ref).max() <= 1e-14 * np.abs(ref).max()
    assert torch.equal(F.detach(), T.csr.assemble_rhs(p, dev(f)))
    g = dev(np.random.default_rng(4).standard_normal(ref.shape))
    (F * g).sum().backward()
    assert torch.equal(fd.grad, g[dev(idx.astype(np.int64))])


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_residual_contractions(T, tag, tables):
    """K9/K10 (`tfem_elem_grad` / `tfem_elem_force`) against the reference formulas evaluated with torch on the
    shape gradients B = J^-1 b (base.py:306-314, 1052, 1082-1083; solid.py:56-58), for every element type; and
    the pair is each other's transpose (that is their autograd backward)."""
    from torchfem_b200 import residual as R

    c = load_case(f"case_{tag}.npz")
    et = str(c["etype"])
    heat = tag.startswith("heat")
    nodes, elements = dev(c["nodes"]), dev(c["elements"])
    bref, w = torch.as_tensor(tables[f"{et}.B_ip"]), torch.as_tensor(tables[f"{et}.iweights"]).to(torch.float64)
    n_int, dim, nn = bref.shape
    dpn = 1 if heat else dim
    g = R.Geometry(bref, w, nodes, elements, dpn)
    X = nodes[elements]                                            # [n_elem, nn, dim]
    J = torch.einsum("qiN,eNj->qeij", bref.cuda(), X)
    B = torch.einsum("qeij,qjN->qeiN", torch.linalg.inv(J), bref.cuda())   # [q, e, dim, nn]
    detJ = torch.linalg.det(J)
    gen = torch.Generator(device="cuda").manual_seed(5)
    u_e = torch.randn(len(elements), nn, dpn, device="cuda", generator=gen, requires_grad=True)
    P = torch.randn(n_int, len(elements), dpn, dim, device="cuda", generator=gen, requires_grad=True)
    H = R.elem_grad(g, u_e)
    H_ref = torch.einsum("eni,qeJn->qeiJ", u_e, B)
    assert float((H - H_ref).detach().abs().max()) <= 1e-12 * float(H_ref.detach().abs().max())
    f = R.elem_force(g, P)
    f_ref = torch.einsum("q,qe,qeJn,qeiJ->eni", w.cuda(), detJ, B, P)
    assert float((f - f_ref).detach().abs().max()) <= 1e-12 * float(f_ref.detach().abs().max())
    g.check()
    # autograd: gradients equal those of the torch formulas
    gH, gf = torch.randn_like(H), torch.randn_like(f)
    (gu,) = torch.autograd.grad(H, u_e, gH)
    (gu_ref,) = torch.autograd.grad(H_ref, u_e, gH)
    assert float((gu - gu_ref).abs().max()) <= 1e-12 * float(gu_ref.abs().max())
    (gP,) = torch.autograd.grad(f, P, gf)
    (gP_ref,) = torch.autograd.grad(f_ref, P, gf)
    assert float((gP - gP_ref).abs().max()) <= 1e-12 * float(gP_ref.abs().max())


def _cube_system(T, N, tables):
    nodes, elements = O.cube_hexa(N, N, N)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con_mask, disp = O.cube_extension_bcs(nodes)
    p = T.csr.Pattern(dev(elements), nodes.shape[0], 3)
    k = T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(bref), torch.as_tensor(w), dev(nodes),
                          dev(elements), dev(C))
    vals = T.csr.assemble(p, k, dev(con_mask.ravel().astype(np.uint8)))
    A = T.csr.CSRMatrix(p.indptr, p.indices, vals, p.n_dofs, chunk_rows=p.chunk_rows,
                        diag_pos=p.diag_pos, symmetric=True)
    return nodes, elements, bref, w, C, con_mask, disp, p, k, A


def test_config_a_structure_and_values(T, tables):
    """BASELINE config[0] (benchmarks/cubes.py, N=11) against the reference's golden vectors."""
    g = load_case("config_a.npz")
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    assert p.nnz == int(g["nnz"]) == 268119
    assert sha(p.glob_idx.cpu().numpy()) == str(g["sha_glob_idx"])
    assert sha(p.k_map.cpu().numpy()) == str(g["sha_k_map"])
    assert sha(p.diag_map.cpu().numpy()) == str(g["sha_diag_map"])
    kk = k.cpu().numpy()
    assert rel(kk[0], g["k_e0"]) <= TOL_K and rel(kk[777], g["k_e777"]) <= TOL_K
    assert abs(np.linalg.norm(kk) - g["k_fro"]) <= TOL_K * g["k_fro"]
    v = A.values_.cpu().numpy()
    assert np.abs(v[:4096] - g["val_head"]).max() <= TOL_K * g["val_absmax"]
    assert abs(np.linalg.norm(v) - g["val_norm"]) <= TOL_K * g["val_norm"]
    assert int((v == 1.0).sum()) == int(g["val_n_one"])


@pytest.mark.parametrize("N", [4, 11, 24])
def test_spmv_matches_oracle(T, tables, N):
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, N, tables)
    A_ref = O.to_csr(A.values_.cpu().numpy(), p.glob_idx.cpu().numpy(), p.n_dofs)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(p.n_dofs, generator=g, dtype=torch.float64)
    y = A.matvec(x.cuda()).cpu().numpy()
    y_ref = A_ref @ x.numpy()
    assert np.abs(y - y_ref).max() <= 1e-13 * np.abs(y_ref).max() * 81
    y2 = A.matvec(x.cuda()).cpu().numpy()
    assert np.array_equal(y, y2)
    ys = A.matvec(x.cuda(), fmt="sell-scalar").cpu().numpy()
    assert np.abs(ys - y_ref).max() <= 1e-13 * np.abs(y_ref).max() * 81
    assert np.array_equal(ys, A.matvec(x.cuda(), fmt="sell-scalar").cpu().numpy())
    # node-block column indices (pattern-built matrix): same products, same order -> bitwise equal
    Ab = p.matrix(A.values_)
    assert Ab.sell().block
    yb = Ab.matvec(x.cuda(), fmt="sell").cpu().numpy()
    assert np.array_equal(yb, ys)


@pytest.mark.parametrize("tag", ["hexa2", "tetra2", "quad1", "heat_quad2", "hexa1_orphan"])
def test_spmv_other_row_lengths(T, tag):
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    A = T.csr.CSRMatrix(p.indptr, p.indices, dev(c["K_val"]), n, chunk_rows=p.chunk_rows)
    A_ref = O.to_csr(c["K_val"], c["glob_idx"], n)
    x = np.random.default_rng(1).standard_normal(n)
    y = A.matvec(dev(x)).cpu().numpy()
    y_ref = A_ref @ x
    assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    ys = A.matvec(dev(x), fmt="sell").cpu().numpy()
    assert np.abs(ys - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    Ab = p.matrix(dev(c["K_val"]), symmetric=False)
    assert Ab.sell().block == (dpn in (2, 3) and tag != "hexa1_orphan")
    yb = Ab.matvec(dev(x), fmt="sell").cpu().numpy()
    assert np.array_equal(yb, ys)


@pytest.mark.parametrize("tag", ["hexa1", "hexa2", "tetra2", "quad1", "heat_quad2", "hexa1_orphan"])
@pytest.mark.parametrize("m", [1, 4, 7, 13])
def test_multi_vector_product_equals_single_products(T, tag, m):
    """Y = A X with the multi-vector SELL kernel (the eigensolver's block products): every column bit-equal to the
    single-vector SELL product, for node-block (dpn 2, 3) and scalar (heat, orphan) column indices, full and ragged
    blocks of 4."""
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    A = p.matrix(dev(c["K_val"]), symmetric=False)
    X = dev(np.random.default_rng(m).standard_normal((n, m)))
    Y = A.matmat(X)
    assert Y.shape == X.shape
    for j in range(m):
        assert torch.equal(Y[:, j], A.matvec(X[:, j].contiguous(), fmt="sell"))
    with pytest.raises(ValueError):
        A.matmat(X, out=X)


@pytest.mark.parametrize("method", ["cg", "minres"])
def test_config_a_solve(T, tables, method):
    """Displacements of config A vs the reference (spsolve golden and the reference's own Jacobi-CG /
    MINRES run at the same stol=1e-10): <= 1e-8 relative."""
    g = load_case("config_a.npz")
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10,
                                        method=method)
    b = dev(ref["res"])
    x, M, info = T.csr.krylov_solve(A, b, method=method, rtol=1e-10)
    con = np.nonzero(con_mask.ravel())[0]
    u = -x.cpu().numpy()
    u[con] = disp.ravel()[con]
    u = u.reshape(-1, 3)
    nrm = np.linalg.norm(g["u"])
    assert np.linalg.norm(u - g["u"]) / nrm <= 1e-8
    assert np.linalg.norm(u - g["u_cg" if method == "cg" else "u_minres"]) / nrm <= 1e-8
    assert np.linalg.norm(u - ref["u"]) / nrm <= 1e-8
    # same algorithm -> same iteration count as the oracle (scipy restatement), give or take round-off
    assert abs(info["iterations"] - ref["iterations"]) <= 2, (info, ref["iterations"])
    assert info["converged"]


def test_fused_peer_cg_world1_equals_single_gpu_driver(T, tables):
    """The peer-to-peer multi-GPU CG (`tfem_dcg_solve`) degenerates to the single-GPU driver for one rank
    (self-published reductions, no halo): same iterates, same iteration count. A sub-range of owned rows
    (identity rows outside it play the halo) exercises the slice / row masking."""
    from torchfem_b200 import distributed as D

    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    b = dev(ref["res"])
    x, M, info = T.csr.krylov_solve(A, b, method="cg", rtol=1e-10)
    cg = D.FusedCG(p.indptr, p.indices, A.n, 0, A.n, D.HaloPlan(neighbours=[]), b.device)
    try:
        assert cg.interior == (0, A.n)
        x2, info2 = cg.solve(A, M.dinv, b, rtol=1e-10)
        assert info2["iterations"] == info["iterations"] and info2["converged"]
        assert float((x2 - x).abs().max()) <= 1e-12 * float(x.abs().max())
        x3, info3 = cg.solve(A, M.dinv, b, rtol=1e-10)   # epochs carry over between solves
        assert torch.equal(x3, x2) and info3["iterations"] == info2["iterations"]
        with pytest.raises(RuntimeError, match="CG failed with exit code"):
            cg.solve(A, M.dinv, b, rtol=1e-14, maxiter=3)
        x4, info4 = cg.solve(A, M.dinv, b, rtol=1e-10)
        assert torch.equal(x4, x2)
    finally:
        cg.close()


@pytest.mark.parametrize("tag", ["hexa1", "hexa2", "tetra2", "quad1", "heat_hexa1", "heat_quad2"])
def test_element_operator_equals_assembled_matrix(T, tag):
    """K8 matrix-free operator == the assembled, Dirichlet-masked matrix (product and diagonal)."""
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n_dofs = dpn * c["nodes"].shape[0]
    p = build_pattern(T, c, dpn)
    is_con = np.zeros(n_dofs, dtype=np.uint8)
    is_con[c["con"]] = 1
    k = dev(c["k"])
    A = p.matrix(T.csr.assemble(p, k, dev(is_con)))
    x = dev(np.random.default_rng(1).standard_normal(n_dofs))
    for con in (dev(is_con), None):
        Aop = T.csr.ElementOperator(p, k, con)
        Aref = A if con is not None else p.matrix(T.csr.assemble(p, k, None))
        y, yr = Aop.matvec(x), Aref.matvec(x, fmt="csr")
        assert float((y - yr).abs().max()) <= 1e-12 * float(yr.abs().max())
        assert float((Aop.diagonal() - Aref.diagonal()).abs().max()) <= 1e-12 * float(Aref.diagonal().abs().max())
    assert torch.equal(Aop.matvec(x), Aop.matvec(x))  # deterministic


def test_element_operator_cg_matches_assembled_cg(T, tables):
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 11, tables)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    b = dev(ref["res"])
    x, M, info = T.csr.krylov_solve(A, b, method="cg", rtol=1e-10)
    Aop = T.csr.ElementOperator(p, k, dev(con_mask.ravel().astype(np.uint8)))
    for method in ("cg", "minres"):
        x2, M2, info2 = T.csr.krylov_solve(Aop, b, method=method, rtol=1e-10)
        assert float((x2 - x).abs().max()) <= 1e-8 * float(x.abs().max())
        if method == "cg":
            assert abs(info2["iterations"] - info["iterations"]) <= 1


def test_cg_warm_start_and_zero_rhs(T, tables):
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 6, tables)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    b = dev(ref["res"])
    x, M, info = T.csr.krylov_solve(A, b, method="cg", rtol=1e-10)
    x2, _, info2 = T.csr.krylov_solve(A, b, method="cg", rtol=1e-10, x0=x, M=M)
    assert info2["iterations"] <= 1
    assert torch.allclose(x, x2, atol=1e-8)
    x3, _, info3 = T.csr.krylov_solve(A, torch.zeros_like(b), method="cg", rtol=1e-10)
    assert info3["iterations"] == 0 and float(x3.abs().max()) == 0.0


def test_cg_maxiter_raises_runtime_error(T, tables):
    nodes, elements, bref, w, C, con_mask, disp, p, k, A = _cube_system(T, 6, tables)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    with pytest.raises(RuntimeError, match="CG failed with exit code"):
        T.csr.krylov_solve(A, dev(ref["res"]), method="cg", rtol=1e-14, maxiter=3)


def test_transpose_and_general_matvec(T):
    rng = np.random.default_rng(3)
    n = 200
    import scipy.sparse as sp

    Am = sp.random(n, n, density=0.05, random_state=3, format="csr") + sp.eye(n, format="csr") * 4
    Am = Am.tocsr()
    Am.sort_indices()
    A = T.csr.CSRMatrix(dev(Am.indptr.astype(np.int64)), dev(Am.indices.astype(np.int32)),
                        dev(Am.data), n)
    x = rng.standard_normal(n)
    assert np.allclose(A.matvec(dev(x)).cpu().numpy(), Am @ x, atol=1e-12)
    assert np.allclose(A.matvec(dev(x), fmt="sell").cpu().numpy(), Am @ x, atol=1e-12)
    assert np.allclose(A.T.matvec(dev(x)).cpu().numpy(), Am.T @ x, atol=1e-12)
    At = Am.T.tocsr()
    At.sort_indices()
    assert np.array_equal(A.T.indices.cpu().numpy(), At.indices)
    assert np.array_equal(A.T.values_.cpu().numpy(), At.data)


def test_full_size_properties(T, tables):
    """Size-independent checks at a size the CPU oracle does not touch (N=64, 786k DOFs):
    rigid-body translations are in the null space of the unconstrained K (row sums of each DOF
    direction vanish), K is symmetric in action (x.Ay == y.Ax), CG converges and the residual it
    reports is the true residual."""
    N = 64
    nodes, elements = O.cube_hexa(N, N, N)
    bref, w = O.hexa1_tables()
    p = T.csr.Pattern(dev(elements), nodes.shape[0], 3)
    assert p.nnz == 9 * (3 * N - 2) ** 3
    Cd = dev(O.isotropic_C3d(1000.0, 0.3, 1)).expand(len(elements), 3, 3, 3, 3).contiguous()
    k = T.csr.integrate_k(T._lib.KIND_MECH, torch.as_tensor(bref), torch.as_tensor(w), dev(nodes),
                          dev(elements), Cd)
    vals = T.csr.assemble(p, k, None)
    A = T.csr.CSRMatrix(p.indptr, p.indices, vals, p.n_dofs, chunk_rows=p.chunk_rows, symmetric=True)
    scale = float(vals.abs().max())
    for d in range(3):
        t = torch.zeros(p.n_nod, 3, dtype=torch.float64, device="cuda")
        t[:, d] = 1.0
        assert float(A.matvec(t.ravel()).abs().max()) <= 1e-12 * scale * 81
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(p.n_dofs, generator=g, dtype=torch.float64, device="cuda")
    y = torch.randn(p.n_dofs, generator=g, dtype=torch.float64, device="cuda")
    a, b = float(x @ A.matvec(y)), float(y @ A.matvec(x))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), scale)
    con_mask, disp = O.cube_extension_bcs(nodes)
    is_con = dev(con_mask.ravel().astype(np.uint8))
    vals_c = T.csr.assemble(p, k, is_con)
    Ac = T.csr.CSRMatrix(p.indptr, p.indices, vals_c, p.n_dofs, chunk_rows=p.chunk_rows,
                         diag_pos=p.diag_pos, symmetric=True)
    du = dev(disp.ravel()) * is_con
    rhs = A.matvec(du)
    rhs[is_con.bool()] = 0.0
    xs, M, info = T.csr.krylov_solve(Ac, rhs, method="cg", rtol=1e-8)
    true_res = float(torch.linalg.norm(rhs - Ac.matvec(xs)) / torch.linalg.norm(rhs))
    assert info["converged"] and true_res <= 2e-8
    assert 250 <= info["iterations"] <= 400  # ~5(N-1)+2


@pytest.mark.parametrize("d,n_q", [(3, 8), (3, 1), (2, 4)])
def test_tangent_contraction_kernel(T, d, n_q):
    """K17 `tfem_ddot` / `tfem_ddot_outer` behind `materials._ddot` vs the reference's einsum
    (materials/elasticity.py:119-127), forward and both gradients, <= 1e-14 relative."""
    from torchfem_b200.materials import _ddot

    gen = torch.Generator(device="cuda").manual_seed(d * 10 + n_q)
    n_elem = 1237
    C = torch.randn(n_elem, d, d, d, d, device="cuda", generator=gen, requires_grad=True)
    shape = (n_q, n_elem, d, d) if n_q > 1 else (n_elem, d, d)
    e = torch.randn(*shape, device="cuda", generator=gen, requires_grad=True)
    out = _ddot(C, e)
    ref = torch.einsum("eijkl,...ekl->...eij", C, e)
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-14 * float(ref.abs().max()) * d * d
    g = torch.randn_like(out)
    gC, ge = torch.autograd.grad(out, (C, e), g)
    gC_ref, ge_ref = torch.autograd.grad(ref, (C, e), g)
    assert float((gC - gC_ref).abs().max()) <= 1e-13 * float(gC_ref.abs().max())
    assert float((ge - ge_ref).abs().max()) <= 1e-13 * float(ge_ref.abs().max())
