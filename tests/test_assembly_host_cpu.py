"""`Assembly.solve` end to end on the CPU with every kernel call replaced by a numpy stand-in: the orchestration
(stacking of the parts, elimination, Dirichlet handling, carried state, coupling forces, the adjoint through the
constrained Newton solve) is checked against fixtures from the unmodified reference
(`oracle/make_golden.py::assembly_cases`). The kernels themselves are checked on the GPU
(tests/test_gpu_assembly.py runs the same cases through them; tests/test_gpu_amg.py the SpGEMM, tests/test_gpu_kernels.py
the SpMV). Stand-ins: element matrices and assembly from `oracle/fem_oracle.py`, dense numpy products for K15 / K5,
a dense solve for `sparse_solve`."""
import numpy as np
import pytest
import torch

from conftest import load_case
from host_standins import HostMatrix, dense_sparse_solve, host_model


def _block_pattern(op):
    mask = np.zeros((op.nbr, op.nbc), dtype=np.int64)
    mask[np.repeat(np.arange(op.nbr), np.diff(op.bptr.numpy())), op.bcol.numpy()] = 1
    return mask


def _spgemm(d, X, Y, structure=None, out_vals=None):
    """Structural product pattern over d x d blocks (explicit zeros kept) and values in the block layout, like K15."""
    from torchfem_b200.amg import BlockOperator

    pat = (_block_pattern(X) @ _block_pattern(Y)) > 0
    rows, cols = np.nonzero(pat)
    if structure is not None:
        assert np.array_equal(cols, structure[1].numpy())
    cnt = pat.sum(1)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    dense = (X.to_scipy() @ Y.to_scipy()).toarray()
    vals = np.zeros(d * d * len(cols))
    slot = np.arange(len(cols)) - ptr[rows]
    for a in range(d):
        for c in range(d):
            vals[d * d * ptr[rows] + (a * cnt[rows] + slot) * d + c] = dense[rows * d + a, cols * d + c]
    cptr, ccol = torch.from_numpy(ptr.astype(np.int64)), torch.from_numpy(cols.astype(np.int32))
    return BlockOperator(d, X.nbr, Y.nbc, cptr, ccol, torch.from_numpy(vals)), (cptr, ccol, int(cnt.max()))


def _host_matvec(self, x):
    rows = torch.repeat_interleave(torch.arange(self.n_rows), self.indptr[1:] - self.indptr[:-1])
    y = torch.zeros(self.n_rows, dtype=torch.float64)
    return y.index_add_(0, rows, self.values * x.detach()[self.indices.to(torch.int64)])


@pytest.fixture()
def host(monkeypatch):
    """torchfem_b200.assembly with the kernels swapped for the stand-ins, and a factory for CPU models."""
    import torchfem_b200 as T
    import torchfem_b200.assembly as A

    monkeypatch.setattr(A.L, "require_cuda", lambda *t: None)
    monkeypatch.setattr(A, "spmv_plan", lambda *a: None)
    monkeypatch.setattr(A, "spgemm", _spgemm)
    monkeypatch.setattr(A, "sell_structure", lambda *a: None)
    monkeypatch.setattr(A, "CSRMatrix", HostMatrix)
    monkeypatch.setattr(A._RectCSR, "matvec", _host_matvec)
    monkeypatch.setattr(T.sparse, "sparse_solve", dense_sparse_solve)

    def model(cls, nodes, elements, material):
        return host_model(cls, nodes, elements, material)

    return T, A, model


@pytest.fixture(scope="module")
def gold():
    return load_case("assembly.npz")


def _close(res, gold, tag, tol=1e-9):
    u, f, flux, grad, _ = res
    for j in range(len(u)):
        for name, got in (("u", u[j]), ("f", f[j]), ("flux", flux[j]), ("grad", grad[j])):
            ref = gold[f"{tag}.{name}{j}"]
            assert tuple(got.shape) == ref.shape, (tag, name, j)
            if ref.size:
                scale = max(np.abs(gold[f"{tag}.{name}0"]).max(), np.abs(ref).max(), 1e-300)
                assert np.abs(got.numpy() - ref).max() <= tol * scale, (tag, name, j)


@pytest.mark.parametrize("node_blocks", [True, False])
def test_tied_solids_and_increments(host, gold, node_blocks):
    T, A, model = host
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    n_a, e_a = cube_hexa(4, 4, 3, 1.0, 1.0, 1.0)
    n_b, e_b = cube_hexa(4, 4, 4, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    a, b = model(T.Solid, n_a, e_a, mat), model(T.Solid, n_b, e_b, mat)
    a.constraints[n_a[:, 2] == 0.0] = True
    b.forces[n_b[:, 2] == 2.0, 2] = 25.0 / 16
    b.forces[n_b[:, 2] == 2.0, 0] = 5.0 / 16
    asm = A.Assembly([a, b])
    asm.node_blocks = node_blocks
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    res = asm.solve()
    assert asm._elimination.d == (3 if node_blocks else 1)
    _close(res, gold, "tie")
    assert torch.equal(res[0][1][n_b[:, 2] == 1.0], res[0][0][n_a[:, 2] == 1.0])   # the tie is exact
    every = asm.solve(increments=torch.linspace(0.0, 1.0, 4), return_intermediate=True)
    assert np.abs(every[0][1].numpy() - gold["tie.every_u1"]).max() <= 1e-9 * np.abs(gold["tie.every_u1"]).max()
    assert np.abs(every[1][0].numpy() - gold["tie.every_f0"]).max() <= 1e-9 * np.abs(gold["tie.every_f0"]).max()


def test_reference_point_load_and_prescribed_motion(host, gold):
    T, A, model = host
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    mat = IsotropicElasticity3D(1000.0, 0.3)
    nodes, elements = cube_hexa(4, 4, 4)
    solid = model(T.Solid, nodes, elements, mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = A.ReferencePoint([0.5, 0.5, 2.0])
    point.forces[0, 3], point.forces[0, 5], point.forces[0, 0] = 50.0, -20.0, 10.0
    asm = A.Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point)
    res = asm.solve()
    assert asm._elimination.d == 3        # the point's six DOFs are two node blocks
    _close(res, gold, "point")
    assert res[2][1].shape == (0,) and res[4][1].shape == (0,)
    assert float(res[1][1][0, 3]) == pytest.approx(50.0)      # the point's force is what the coupling transmits

    solid = model(T.Solid, nodes, elements, mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = A.ReferencePoint([0.5, 0.5, 2.0])
    point.constraints[0, :] = True
    point.displacements[0, 2] = 0.1
    asm = A.Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point, dofs=[2])
    _close(asm.solve(), gold, "subset")
    assert asm._elimination.d == 1        # partial nodes are eliminated: scalar operators

    solid.constraints[nodes[:, 2] == 1.0] = True
    with pytest.raises(ValueError, match="constrained DOF is eliminated"):
        asm.solve()


def test_heat_and_planar(host, gold):
    T, A, model = host
    from torchfem_b200.materials import IsotropicConductivity3D, IsotropicElasticityPlaneStress
    from torchfem_b200.mesh import cube_hexa, rect_quad

    cond = IsotropicConductivity3D(1.5)
    n_a, e_a = cube_hexa(3, 3, 3)
    n_b, e_b = cube_hexa(3, 3, 4)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    ha, hb = model(T.SolidHeat, n_a, e_a, cond), model(T.SolidHeat, n_b, e_b, cond)
    ha.constraints[n_a[:, 2] == 0.0] = True
    hp = A.ReferencePointHeat([0.5, 0.5, 2.5])
    hp.heat_flux[0, 0] = 4.0
    asm = A.Assembly([ha, hb, hp])
    asm.coupling(hb, n_b[:, 2] == 1.0, ha, n_a[:, 2] == 1.0)
    asm.coupling(hb, n_b[:, 2] == 2.0, hp)
    _close(asm.solve(), gold, "heat")

    plane = IsotropicElasticityPlaneStress(1000.0, 0.3)
    n_a, e_a = rect_quad(4, 4, 1.0, 1.0)
    n_b, e_b = rect_quad(4, 4, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0])
    pa, pb = model(T.Planar, n_a, e_a, plane), model(T.Planar, n_b, e_b, plane)
    pa.constraints[n_a[:, 0] == 0.0] = True
    pp = A.ReferencePoint([2.5, 0.5])
    pp.forces[0, 2], pp.forces[0, 1] = 20.0, -1.0
    asm = A.Assembly([pa, pb, pp])
    asm.coupling(pb, n_b[:, 0] == 1.0, pa, n_a[:, 0] == 1.0)
    asm.coupling(pb, n_b[:, 0] == 2.0, pp)
    _close(asm.solve(), gold, "planar")
    assert asm._elimination.d == 1        # a 2-D point has three DOFs: not a whole number of 2-DOF nodes
    # the tie alone is 2 x 2-blocked: both layouts give the same solution
    pb.forces[n_b[:, 0] == 2.0, 1] = -0.25
    sols = []
    for node_blocks in (True, False):
        asm = A.Assembly([pa, pb])
        asm.node_blocks = node_blocks
        asm.coupling(pb, n_b[:, 0] == 1.0, pa, n_a[:, 0] == 1.0)
        sols.append(asm.solve()[0])
        assert asm._elimination.d == (2 if node_blocks else 1)
    assert all(torch.allclose(x, y, rtol=0, atol=1e-13) for x, y in zip(*sols))


def test_adjoint_through_the_constrained_solve(host, gold):
    T, A, model = host
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    n_a, e_a = cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b, e_b = cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0, 0.0])
    rho = torch.tensor(gold["adjoint.rho"], requires_grad=True)
    mat = IsotropicElasticity3D(E=1000.0, nu=0.3).vectorize(len(e_a))
    mat.C = (rho ** 3.0)[:, None, None, None, None] * mat.C
    a, b = model(T.Solid, n_a, e_a, mat), model(T.Solid, n_b, e_b, IsotropicElasticity3D(1000.0, 0.3))
    a.constraints[n_a[:, 0] == 0.0] = True
    point = A.ReferencePoint([2.5, 0.5, 0.5])
    point.forces[0, 2], point.forces[0, 3] = -3.0, 1.0
    asm = A.Assembly([a, b, point])
    asm.coupling(b, n_b[:, 0] == 1.0, a, n_a[:, 0] == 1.0)
    asm.coupling(b, n_b[:, 0] == 2.0, point)
    u, *_ = asm.solve(differentiable_parameters=rho)
    assert asm._elimination.d == 3
    work = torch.inner(point.forces.ravel(), u[2].ravel())
    work.backward()
    assert abs(float(work.detach()) - float(gold["adjoint.work"])) <= 1e-9 * abs(float(gold["adjoint.work"]))
    assert np.abs(u[2].detach().numpy() - gold["adjoint.u2"]).max() <= 1e-9 * np.abs(gold["adjoint.u2"]).max()
    g = rho.grad.numpy()
    assert np.linalg.norm(g - gold["adjoint.grad_rho"]) <= 1e-8 * np.linalg.norm(gold["adjoint.grad_rho"])
