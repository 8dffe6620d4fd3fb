"""Material tangents and stress updates against the reference's classes (run in the build container, where
/root/reference exists; the golden model fixtures cover the same ground on the GPU box)."""
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def mods():
    torch.set_default_device("cpu")
    ref_import.load()
    from torchfem import materials as R

    from torchfem_b200 import materials as M

    return M, R


@pytest.mark.parametrize("name,args,d", [
    ("IsotropicElasticity3D", (1000.0, 0.3), 3), ("IsotropicElasticityPlaneStress", (1000.0, 0.3), 2),
    ("IsotropicElasticityPlaneStrain", (1000.0, 0.3), 2),
    ("IsotropicElasticity3D", (torch.tensor([1.0, 2.0, 3.0]), torch.tensor([0.1, 0.2, 0.3])), 3),
    ("IsotropicElasticityPlaneStress", (torch.tensor([1.0, 2.0]), torch.tensor([0.1, 0.2])), 2),
])
def test_elastic_tangent_and_step(mods, name, args, d):
    M, R = mods
    a, b = getattr(M, name)(*args), getattr(R, name)(*args)
    assert a.is_vectorized == b.is_vectorized and a.n_state == b.n_state == 0
    assert torch.allclose(a.C, b.C, atol=1e-13, rtol=1e-14)
    n = 4
    av, bv = a.vectorize(n) if not a.is_vectorized else a, b.vectorize(n) if not b.is_vectorized else b
    n = av.C.shape[0]
    assert av.C.shape == bv.C.shape == (n, d, d, d, d)
    g = torch.Generator().manual_seed(0)
    H, F, S, de0 = (torch.rand(n, d, d, generator=g) for _ in range(4))
    st = torch.zeros(n, 0)
    for x, y in zip(av.step(H, F, S, st, de0, None, 0), bv.step(H, F, S, st, de0, None, 0)):
        assert torch.allclose(x, y, atol=1e-12)


def test_hyperelastic_and_conductivity(mods):
    M, R = mods

    def psi(F, p):
        C = F.transpose(-1, -2) @ F
        logJ = 0.5 * torch.logdet(C)
        return p[0] / 2 * (torch.trace(C) - 3.0) - p[0] * logJ + p[1] / 2 * logJ ** 2

    p = torch.tensor([384.6, 576.9])
    a, b = M.Hyperelastic3D(psi, p).vectorize(5), R.Hyperelastic3D(psi, p).vectorize(5)
    g = torch.Generator().manual_seed(1)
    H = 0.1 * torch.rand(5, 3, 3, generator=g)
    F = torch.eye(3).repeat(5, 1, 1)
    z = torch.zeros(5, 3, 3)
    for x, y in zip(a.step(H, F, z, torch.zeros(5, 0), z, None, 0), b.step(H, F, z, torch.zeros(5, 0), z, None, 0)):
        assert torch.allclose(x, y, atol=1e-10)
    P0, _, _ = a.step(z, F, z, torch.zeros(5, 0), z, None, 0)
    assert torch.allclose(P0, z, atol=1e-10)  # stress-free at identity (reference tests/test_materials.py:190-226)
    a2, b2 = M.HyperelasticPlaneStrain(psi, p).vectorize(3), R.HyperelasticPlaneStrain(psi, p).vectorize(3)
    H2 = 0.1 * torch.rand(3, 2, 2, generator=g)
    F2 = torch.eye(2).repeat(3, 1, 1)
    z2 = torch.zeros(3, 2, 2)
    for x, y in zip(a2.step(H2, F2, z2, torch.zeros(3, 0), z2, None, 0), b2.step(H2, F2, z2, torch.zeros(3, 0), z2, None, 0)):
        assert torch.allclose(x, y, atol=1e-10)
    for cls, d in (("IsotropicConductivity3D", 3), ("IsotropicConductivity2D", 2)):
        ka, kb = getattr(M, cls)(400.0).vectorize(4), getattr(R, cls)(400.0).vectorize(4)
        assert torch.equal(ka.KAPPA, kb.KAPPA)
        gi, f0 = torch.rand(4, 1, d, generator=g), torch.rand(4, 1, d, generator=g)
        for x, y in zip(ka.step(gi, f0, f0, torch.zeros(4, 0), 0 * gi, None, 0), kb.step(gi, f0, f0, torch.zeros(4, 0), 0 * gi, None, 0)):
            assert torch.allclose(x, y, atol=1e-12)

