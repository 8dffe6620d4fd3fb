"""Orthotropic / transversely isotropic elasticity and orthotropic conductivity (the tangent
producers of the material interface) against fixtures from the unmodified reference
(`oracle/make_golden.py::orthotropic`, reference elasticity.py:324-793, conductivity.py:143-242), and a block with a
per-element rotated orthotropic material solved through the host stand-ins."""
import numpy as np
import pytest
import torch

from conftest import load_case
from torchfem_b200 import materials as M


P3 = dict(E_1=150.0, E_2=12.0, E_3=9.0, nu_12=0.3, nu_13=0.25, nu_23=0.4, G_12=5.0, G_13=4.0, G_23=3.0)


def _rel(a, b):
    return float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())


def test_orthotropic_elasticity_matches_reference():
    """Stiffness tensors, rotation and re-extracted engineering constants, stress update (reference
    elasticity.py:324-793; fixtures from `oracle/make_golden.py::orthotropic`)."""
    g = load_case("orthotropic.npz")
    m = M.OrthotropicElasticity3D(**P3)
    assert m.n_state == 0 and not m.is_vectorized and m.C.shape == (3, 3, 3, 3)
    assert _rel(m.C.numpy(), g["o3.C"]) <= 1e-13
    C = m.C
    assert torch.allclose(C, C.permute(2, 3, 0, 1)) and torch.allclose(C, C.permute(1, 0, 2, 3))    # major / minor
    R = torch.tensor(g["o3.R"])
    mv = m.vectorize(4).rotate(R)
    assert _rel(mv.C.numpy(), g["o3.C_rot"]) <= 1e-13
    names = ["E_1", "E_2", "E_3", "nu_12", "nu_13", "nu_23", "G_12", "G_13", "G_23"]
    assert _rel(np.stack([getattr(mv, k).numpy() for k in names]), g["o3.consts_rot"]) <= 1e-12
    sig, state, dd = mv.step(torch.tensor(g["o3.H"]), torch.eye(3).expand(4, 3, 3), torch.tensor(g["o3.s0"]),
                             torch.zeros(4, 0), torch.tensor(g["o3.de0"]), torch.ones(4, 1), 0)
    assert _rel(sig.numpy(), g["o3.sig"]) <= 1e-13 and dd is mv.C and state.shape == (4, 0)
    E1 = torch.tensor(g["o3.E1_batch"])
    mb = M.OrthotropicElasticity3D(E1, 0.1 * E1, 0.08 * E1, 0.3, 0.25, 0.4, 0.04 * E1, 0.03 * E1, 0.02 * E1)
    assert mb.is_vectorized and _rel(mb.C.numpy(), g["o3.C_batch"]) <= 1e-13
    with pytest.raises(ValueError, match="3x3"):
        m.rotate(torch.eye(2))
    # rotating by the identity changes nothing; the constants come back
    same = m.rotate(torch.eye(3))
    assert torch.allclose(same.C, m.C) and float(same.E_1) == pytest.approx(150.0) and float(same.G_23) == pytest.approx(3.0)

    ti = M.TransverseIsotropicElasticity3D(E_L=140.0, E_T=10.0, nu_L=0.28, nu_T=0.42, G_L=5.5)
    assert _rel(ti.C.numpy(), g["ti.C"]) <= 1e-13
    with pytest.raises(ValueError, match="G must be less"):
        M.TransverseIsotropicElasticity3D(E_L=10.0, E_T=10.0, nu_L=0.3, nu_T=0.3, G_L=50.0)

    R2 = torch.tensor(g["ps.R"])
    ps = M.OrthotropicElasticityPlaneStress(E_1=150.0, E_2=12.0, nu_12=0.3, G_12=5.0)
    psr = ps.vectorize(3).rotate(R2)
    assert ps.C.shape == (2, 2, 2, 2) and not hasattr(ps, "G_13")
    assert _rel(ps.C.numpy(), g["ps.C"]) <= 1e-13 and _rel(psr.C.numpy(), g["ps.C_rot"]) <= 1e-13
    assert _rel(np.stack([getattr(psr, k).numpy() for k in ["E_1", "E_2", "nu_12", "G_12"]]), g["ps.consts_rot"]) <= 1e-12
    pe = M.OrthotropicElasticityPlaneStrain(E_1=150.0, E_2=12.0, E_3=9.0, nu_12=0.3, nu_13=0.25, nu_23=0.4, G_12=5.0)
    per = pe.vectorize(3).rotate(R2)
    assert _rel(pe.C.numpy(), g["pe.C"]) <= 1e-13 and _rel(per.C.numpy(), g["pe.C_rot"]) <= 1e-13
    assert _rel(np.stack([getattr(per, k).numpy() for k in ["E_1", "E_2", "nu_12", "G_12"]]), g["pe.consts_rot"]) <= 1e-12
    with pytest.raises(ValueError, match="2x2"):
        pe.rotate(torch.eye(3))


def test_orthotropic_conductivity_matches_reference():
    g = load_case("orthotropic.npz")
    k3 = M.OrthotropicConductivity3D(10.0, 2.0, 0.5)
    assert _rel(k3.KAPPA.numpy(), g["k3.K"]) <= 1e-15
    assert _rel(k3.vectorize(4).rotate(torch.tensor(g["o3.R"])).KAPPA.numpy(), g["k3.K_rot"]) <= 1e-13
    k2 = M.OrthotropicConductivity2D(torch.tensor([10.0, 4.0, 1.0]), torch.tensor([2.0, 1.0, 0.5]))
    assert k2.is_vectorized and _rel(k2.KAPPA.numpy(), g["k2.K"]) <= 1e-15
    assert _rel(k2.rotate(torch.tensor(g["ps.R"])).KAPPA.numpy(), g["k2.K_rot"]) <= 1e-13
    flux, _, tangent = k3.step(torch.tensor([[1.0, 2.0, 3.0]]), None, torch.zeros(1, 3), torch.zeros(0), torch.zeros(1, 3),
                               None, 0)
    assert torch.allclose(flux, torch.tensor([[10.0, 4.0, 1.5]])) and tangent is k3.KAPPA
    with pytest.raises(ValueError, match="3x3"):
        k3.rotate(torch.eye(2))


def test_block_with_rotated_orthotropic_material(monkeypatch):
    """Clamped 3x2x2-element block, fibre direction rotated element by element: displacements and stresses of the
    reference (host stand-ins for the kernels; tests/test_gpu_models.py runs the same case on the GPU)."""
    import torchfem_b200 as T
    from host_standins import dense_sparse_solve, host_model
    from torchfem_b200.mesh import cube_hexa

    monkeypatch.setattr(T.sparse, "sparse_solve", dense_sparse_solve)
    g = load_case("orthotropic.npz")
    nodes, elements = cube_hexa(4, 3, 3, 1.5, 1.0, 1.0)
    material = M.OrthotropicElasticity3D(**P3).vectorize(len(elements)).rotate(torch.tensor(g["solid.R"]))
    model = host_model(T.Solid, nodes, elements, material)
    model.constraints[nodes[:, 0] == 0.0, :] = True
    model.forces[nodes[:, 0] == 1.5, 2] = -0.1
    u, f, sigma, eps, _ = model.solve()
    assert _rel(u.numpy(), g["solid.u"]) <= 1e-9 and _rel(sigma.numpy(), g["solid.sigma"]) <= 1e-8
