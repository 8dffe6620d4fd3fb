"""`Assembly` on the GPU: the cases of reference tests/test_assembly.py:54-77, 128-184, 215-272,
300-306, 351-392, 415-455 run through the kernels (K1-K3 per part, K15 SpGEMM for T^T K T, K5 SpMV for T q / T^T f,
the Krylov / dense solvers behind `newton_solve`) and are compared with fixtures from the unmodified reference
(tests/golden/assembly.npz, `oracle/make_golden.py::assembly_cases`). Tolerance: 1e-8 relative for fields (the
north-star bound for displacements at equal solver tolerance), 1e-7 for the adjoint gradient."""
import os

import numpy as np
import pytest
import torch

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def cuda_default():
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda")
    yield
    torch.set_default_device("cpu")


@pytest.fixture(scope="module")
def T():
    import torchfem_b200 as T

    return T


@pytest.fixture(scope="module")
def gold():
    return load_case("assembly.npz")


def _close(res, gold, tag, tol=1e-8):
    u, f, flux, grad, _ = res
    for j in range(len(u)):
        for name, got in (("u", u[j]), ("f", f[j]), ("flux", flux[j]), ("grad", grad[j])):
            ref = gold[f"{tag}.{name}{j}"]
            assert tuple(got.shape) == ref.shape, (tag, name, j)
            if ref.size:
                scale = max(np.abs(gold[f"{tag}.{name}0"]).max(), np.abs(ref).max(), 1e-300)
                assert np.abs(got.cpu().numpy() - ref).max() <= tol * scale, (tag, name, j)


def _tie(T):
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    n_a, e_a = cube_hexa(4, 4, 3, 1.0, 1.0, 1.0)
    n_b, e_b = cube_hexa(4, 4, 4, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    a, b = T.Solid(n_a, e_a, mat), T.Solid(n_b, e_b, mat)
    a.constraints[n_a[:, 2] == 0.0] = True
    b.forces[n_b[:, 2] == 2.0, 2] = 25.0 / 16
    b.forces[n_b[:, 2] == 2.0, 0] = 5.0 / 16
    asm = T.Assembly([a, b])
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    return asm, n_a, n_b


@pytest.mark.parametrize("node_blocks", [True, False])
def test_elimination_operators_match_the_reference(T, gold, node_blocks):
    """T and T^T on the device (CSR, kernel K5) against the reference's COO map, and T^T K T against a dense product,
    with the products on 3 x 3 node blocks and on scalar entries."""
    asm, n_a, n_b = _tie(T)
    Tcoo, retained = asm._build_T()
    assert np.array_equal(Tcoo._indices().cpu().numpy(), gold["tie.T_idx"])
    assert np.abs(Tcoo._values().cpu().numpy() - gold["tie.T_val"]).max() <= 1e-15
    assert np.array_equal(retained.cpu().numpy(), gold["tie.retained"])
    from torchfem_b200.assembly import EMPTY, _Elimination

    elim = _Elimination(asm, node_blocks)
    assert elim.d == (3 if node_blocks else 1)
    Td = Tcoo.to_dense()
    q = torch.randn(elim.n_retained, generator=torch.Generator(device="cuda").manual_seed(0))
    v = torch.randn(asm.n_dofs, generator=torch.Generator(device="cuda").manual_seed(1))
    assert float((elim.T @ q - Td @ q).abs().max()) <= 1e-14 * float(q.abs().max())
    assert float((elim.Tt @ v - Td.T @ v).abs().max()) <= 1e-13 * float(v.abs().max())
    # autograd through both products
    qg = q.clone().requires_grad_(True)
    (elim.T @ qg).dot(v).backward()
    assert float((qg.grad - Td.T @ v).abs().max()) <= 1e-13 * float(v.abs().max())
    blocks = [p.assemble_matrix(p.k0(), EMPTY) for p in asm.parts]
    con = torch.tensor([0, 5, 7])
    K = elim.reduced(blocks, con)
    Kd = torch.block_diag(*[b.to_dense() for b in blocks])
    ref = Td.T @ Kd @ Td
    ref[con, :] = 0.0
    ref[:, con] = 0.0
    ref[con, con] = 1.0
    assert float((K.to_dense() - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    assert elim.reduced(blocks, con) is K              # same part matrices, same constraints: reused
    assert (K._sell_struct.block is not None) == node_blocks     # node-block column indices for the Krylov kernels
    x = torch.randn(elim.n_retained, generator=torch.Generator(device="cuda").manual_seed(2))
    for fmt in ("sell", "csr"):
        assert float((K.matvec(x, fmt=fmt) - ref @ x).abs().max()) <= 1e-12 * float(ref.abs().max()) * float(x.abs().max())
    K2 = elim.reduced([blocks[0] * 2.0, blocks[1]], con)   # numeric phase only, on the stored patterns
    ref2 = Td.T @ torch.block_diag(2.0 * blocks[0].to_dense(), blocks[1].to_dense()) @ Td
    ref2[con, :] = 0.0
    ref2[:, con] = 0.0
    ref2[con, con] = 1.0
    assert float((K2.to_dense() - ref2).abs().max()) <= 1e-12 * float(ref2.abs().max())


@pytest.mark.parametrize("method,node_blocks", [(None, True), (None, False), ("cg", True), ("cg", False),
                                                ("minres", True), ("amgx", True)])
def test_tied_solids(T, gold, method, node_blocks):
    asm, n_a, n_b = _tie(T)
    asm.node_blocks = node_blocks
    res = asm.solve(method=method)
    assert asm._elimination.d == (3 if node_blocks else 1)
    _close(res, gold, "tie", 1e-8 if method is None else 1e-6)    # dense LU / Krylov at the default stol = 1e-10
    assert torch.equal(res[0][1][n_b[:, 2] == 1.0], res[0][0][n_a[:, 2] == 1.0])      # the tie is exact
    if method is None and node_blocks:
        inc = torch.linspace(0.0, 1.0, 4)
        every = asm.solve(increments=inc, return_intermediate=True)
        assert [x.shape[0] for x in every[0]] == [4, 4]
        for got, key in ((every[0][1], "tie.every_u1"), (every[1][0], "tie.every_f0")):
            assert np.abs(got.cpu().numpy() - gold[key]).max() <= 1e-8 * np.abs(gold[key]).max()
        final = asm.solve(increments=inc)
        assert all(torch.allclose(x[-1], y) for x, y in zip(every[0], final[0]))


def test_reference_point_moment_and_rigid_relation(T, gold):
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    nodes, elements = cube_hexa(4, 4, 4)
    solid = T.Solid(nodes, elements, mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = T.ReferencePoint([0.5, 0.5, 2.0])
    point.forces[0, 3], point.forces[0, 5], point.forces[0, 0] = 50.0, -20.0, 10.0
    asm = T.Assembly([solid, point])
    top = nodes[:, 2] == 1.0
    asm.coupling(solid, top, point)
    res = asm.solve()
    assert asm._elimination.d == 3          # the point's six DOFs are two node blocks (translations, rotations)
    _close(res, gold, "point")
    u, f = res[0], res[1]
    u_p, theta = u[1][0, :3], u[1][0, 3:]
    expected = u_p + torch.cross(theta.expand(int(top.sum()), 3), nodes[top] - point.nodes[0], dim=-1)
    assert torch.allclose(u[0][top], expected, atol=1e-12)
    assert float(f[1][0, 3]) == pytest.approx(50.0)
    assert res[2][1].shape == (0,) and res[4][1].shape == (0,)
    _close(asm.solve(method="cg"), gold, "point", 1e-6)     # Jacobi-PCG on the reduced system (six long rows)

    # only u_z coupled, the point fully prescribed
    solid = T.Solid(nodes, elements, mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = T.ReferencePoint([0.5, 0.5, 2.0])
    point.constraints[0, :] = True
    point.displacements[0, 2] = 0.1
    asm = T.Assembly([solid, point])
    asm.coupling(solid, top, point, dofs=[2])
    res = asm.solve()
    assert asm._elimination.d == 1          # partial nodes are eliminated: scalar operators
    _close(res, gold, "subset")
    assert torch.allclose(res[0][0][top][:, 2], torch.full((int(top.sum()),), 0.1))
    solid.constraints[top] = True
    with pytest.raises(ValueError, match="constrained DOF is eliminated"):
        asm.solve()


def test_quadratic_part_tied_to_a_linear_one(T, gold):
    from torchfem_b200.elements import linear_to_quadratic
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    n_q, e_q = linear_to_quadratic(*cube_hexa(3, 3, 3))   # mid-side nodes in the reference's CPU order
    n_l, e_l = cube_hexa(3, 3, 3)
    n_l = n_l + torch.tensor([0.0, 0.0, 1.0])
    q, l = T.Solid(n_q, e_q, mat), T.Solid(n_l, e_l, mat)
    q.constraints[n_q[:, 2] == 0.0] = True
    l.forces[n_l[:, 2] == 2.0, 1] = 1.0
    asm = T.Assembly([q, l])
    asm.coupling(l, n_l[:, 2] == 1.0, q, n_q[:, 2] == 1.0)
    _close(asm.solve(), gold, "mixed")


def test_heat_tie_and_isothermal_surface(T, gold):
    from torchfem_b200.materials import IsotropicConductivity3D
    from torchfem_b200.mesh import cube_hexa

    cond = IsotropicConductivity3D(1.5)
    n_a, e_a = cube_hexa(3, 3, 3)
    n_b, e_b = cube_hexa(3, 3, 4)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    ha, hb = T.SolidHeat(n_a, e_a, cond), T.SolidHeat(n_b, e_b, cond)
    ha.constraints[n_a[:, 2] == 0.0] = True
    hp = T.ReferencePointHeat([0.5, 0.5, 2.5])
    hp.heat_flux[0, 0] = 4.0
    asm = T.Assembly([ha, hb, hp])
    asm.coupling(hb, n_b[:, 2] == 1.0, ha, n_a[:, 2] == 1.0)
    asm.coupling(hb, n_b[:, 2] == 2.0, hp)
    res = asm.solve()
    _close(res, gold, "heat")
    top = n_b[:, 2] == 2.0
    assert torch.allclose(res[0][1][top], res[0][2][0, 0].expand(int(top.sum()), 1))   # one temperature on the surface
    with pytest.raises(ValueError, match="mechanical or thermal"):
        T.Assembly([ha, T.ReferencePoint([0.0, 0.0, 1.0])])


def test_planar_tie_and_point_rotation(T, gold):
    from torchfem_b200.materials import IsotropicElasticityPlaneStress
    from torchfem_b200.mesh import rect_quad

    plane = IsotropicElasticityPlaneStress(1000.0, 0.3)
    n_a, e_a = rect_quad(4, 4, 1.0, 1.0)
    n_b, e_b = rect_quad(4, 4, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0])
    pa, pb = T.Planar(n_a, e_a, plane), T.Planar(n_b, e_b, plane)
    pa.constraints[n_a[:, 0] == 0.0] = True
    pp = T.ReferencePoint([2.5, 0.5])
    pp.forces[0, 2], pp.forces[0, 1] = 20.0, -1.0
    asm = T.Assembly([pa, pb, pp])
    asm.coupling(pb, n_b[:, 0] == 1.0, pa, n_a[:, 0] == 1.0)
    asm.coupling(pb, n_b[:, 0] == 2.0, pp)
    res = asm.solve()
    _close(res, gold, "planar")
    assert float(res[1][2][0, 2]) == pytest.approx(20.0)
    with pytest.raises(ValueError, match="one spatial dimension"):
        T.Assembly([pa, T.ReferencePoint([0.0, 0.0, 0.0])])


@pytest.mark.parametrize("method", [None, "cg"])
def test_adjoint_through_the_constrained_solve(T, gold, method):
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    n_a, e_a = cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b, e_b = cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0, 0.0])
    rho = torch.tensor(gold["adjoint.rho"], requires_grad=True)
    mat = IsotropicElasticity3D(E=1000.0, nu=0.3).vectorize(len(e_a))
    mat.C = (rho ** 3.0)[:, None, None, None, None] * mat.C
    a, b = T.Solid(n_a, e_a, mat), T.Solid(n_b, e_b, IsotropicElasticity3D(1000.0, 0.3))
    a.constraints[n_a[:, 0] == 0.0] = True
    point = T.ReferencePoint([2.5, 0.5, 0.5])
    point.forces[0, 2], point.forces[0, 3] = -3.0, 1.0
    asm = T.Assembly([a, b, point])
    asm.coupling(b, n_b[:, 0] == 1.0, a, n_a[:, 0] == 1.0)
    asm.coupling(b, n_b[:, 0] == 2.0, point)
    u, *_ = asm.solve(differentiable_parameters=rho, method=method, stol=1e-12)
    work = torch.inner(point.forces.ravel(), u[2].ravel())
    work.backward()
    tol = 1e-8 if method is None else 1e-7
    assert abs(float(work.detach()) - float(gold["adjoint.work"])) <= tol * abs(float(gold["adjoint.work"]))
    g = rho.grad.cpu().numpy()
    assert np.linalg.norm(g - gold["adjoint.grad_rho"]) <= 10 * tol * np.linalg.norm(gold["adjoint.grad_rho"])


def test_larger_assembly_iterative_equals_monolithic(T):
    """Size-independent property: two 16^3-element blocks tied at a face of 289 nodes (28,611 retained DOFs, Jacobi-MINRES
    by the size policy, SpGEMM over 29,478 rows) reproduce the monolithic 16x16x32-element bar."""
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    N = 17
    nodes, elements = cube_hexa(N, N, 2 * N - 1, 1.0, 1.0, 2.0)
    mono = T.Solid(nodes, elements, mat)
    mono.constraints[nodes[:, 2] == 0.0] = True
    mono.forces[nodes[:, 2] == 2.0, 2] = 1.0 / N ** 2
    mono.forces[nodes[:, 2] == 2.0, 0] = 0.2 / N ** 2
    u_ref = mono.solve()[0]

    n_a, e_a = cube_hexa(N, N, N, 1.0, 1.0, 1.0)
    n_b = n_a + torch.tensor([0.0, 0.0, 1.0])
    a, b = T.Solid(n_a, e_a, mat), T.Solid(n_b, e_a.clone(), mat)
    a.constraints[n_a[:, 2] == 0.0] = True
    asm = T.Assembly([a, b])
    asm.coupling(b, torch.isclose(n_b[:, 2], torch.tensor(1.0)), a, torch.isclose(n_a[:, 2], torch.tensor(1.0)))
    b.forces[n_b[:, 2] == 2.0, 2] = 1.0 / N ** 2
    b.forces[n_b[:, 2] == 2.0, 0] = 0.2 / N ** 2
    u, f, *_ = asm.solve()
    assert asm._elimination.d == 3
    scale = float(u_ref.abs().max())
    lower = nodes[:, 2] <= 1.0 + 1e-12
    upper = nodes[:, 2] >= 1.0 - 1e-12
    assert float((u[0] - u_ref[lower]).abs().max()) <= 1e-6 * scale
    assert float((u[1] - u_ref[upper]).abs().max()) <= 1e-6 * scale
    top, interface = torch.isclose(n_a[:, 2], torch.tensor(1.0)), torch.isclose(n_b[:, 2], torch.tensor(1.0))
    assert torch.equal(u[1][interface], u[0][top])
    # the eliminated interface DOFs of b carry b's own internal force, which balances the load on its top face
    assert float(f[1][interface][:, 2].sum() + f[1][n_b[:, 2] == 2.0][:, 2].sum()) == pytest.approx(0.0, abs=1e-6)
    # the reduced matrix keeps 3 x 3 node blocks, so the AMG kernels coarsen it like a single model's tangent
    ua = asm.solve(method="amgx")[0]
    assert float((ua[0] - u_ref[lower]).abs().max()) <= 1e-6 * scale and float((ua[1] - u_ref[upper]).abs().max()) <= 1e-6 * scale


def test_reference_point_long_rows_take_the_side_path(T):
    """A reference point driving a face of 441 nodes puts six rows of 1,329 entries into the reduced tangent (10,590
    retained DOFs; reference assembly.py:295-335). Rows beyond TFEM_SELL_LONG_ROW stay out of the SELL-32 slices and are
    computed by the side path of the SpMV (one CTA per row, fixed order): the product equals the CSR kernel's, CG /
    MINRES reproduce the dense solution, two solves are bitwise equal, and the slices are not padded to 1,329."""
    from torchfem_b200 import _lib as L
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    nodes, elements = cube_hexa(21, 21, 9, 1.0, 1.0, 0.4)
    top = nodes[:, 2] == nodes[:, 2].max()

    def build():
        solid = T.Solid(nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
        solid.constraints[nodes[:, 2] == 0.0] = True
        point = T.ReferencePoint([0.5, 0.5, 1.0])
        point.forces[0, 2], point.forces[0, 3], point.forces[0, 0] = -5.0, 2.0, 1.0
        asm = T.Assembly([solid, point])
        asm.coupling(solid, top, point)
        return asm

    asm = build()
    ref = asm.solve(method="spsolve")
    K = asm._elimination._last[2]
    lens = K.indptr[1:] - K.indptr[:-1]
    assert int(lens.max()) > L.SELL_LONG_ROW and int((lens > L.SELL_LONG_ROW).sum()) == 6
    st = K._sell_struct
    K.sell()
    assert st.long_rows is not None and st.long_rows.numel() == 6
    assert st.padded < 1.2 * K.nnz                      # not padded to the longest row (that would be > 10x)
    x = torch.randn(K.n, dtype=torch.float64, generator=torch.Generator(device="cuda").manual_seed(0))
    y_sell, y_csr = K.matvec(x, fmt="sell"), K.matvec(x, fmt="csr")
    assert float((y_sell - y_csr).abs().max()) <= 1e-12 * float(y_csr.abs().max())
    assert torch.equal(y_sell, K.matvec(x, fmt="sell"))
    for method in ("cg", "minres", None):
        sol = build().solve(method=method, stol=1e-12)
        for a, b in zip(ref[0] + ref[1], sol[0] + sol[1]):
            assert float((a - b).abs().max()) <= 1e-6 * max(float(a.abs().max()), 1e-300)
    again = build().solve(method="cg", stol=1e-12)
    first = build().solve(method="cg", stol=1e-12)
    assert all(torch.equal(a, b) for a, b in zip(first[0], again[0]))
