"""Host logic of `FEM.solve` / `newton_solve` on the CPU (`-m "not gpu"`): load increments, Newton loop, Dirichlet
handling, viscous stabilisation, the implicit-function-theorem adjoint chained across increments — the package's own
code, with the kernel-backed pieces replaced by the stand-ins of tests/host_standins.py (oracle element matrices and
assembly, dense solve). The same test bodies run on the GPU through the kernels (tests/test_gpu_reference_suite.py,
tests/test_gpu_models.py); here they guard the host code where no GPU is available. Fixtures: tests/golden/ (generated
from the unmodified reference)."""
import types

import numpy as np
import pytest
import torch

import test_gpu_models as GM
import test_gpu_reference_suite as RS
from conftest import load_case
from host_standins import dense_sparse_solve, host_model


@pytest.fixture()
def T(monkeypatch):
    """A namespace that builds CPU stand-in models where the GPU tests build `torchfem_b200` models."""
    import torchfem_b200 as TT

    monkeypatch.setattr(TT.sparse, "sparse_solve", dense_sparse_solve)
    monkeypatch.setattr(TT.sparse, "_as_csr", lambda A: A)      # `Solve.backward` converts torch tensors only

    def make(cls):
        return lambda nodes, elements, material, thickness=1.0: host_model(cls, nodes, elements, material, thickness)

    return types.SimpleNamespace(Solid=make(TT.Solid), SolidHeat=make(TT.SolidHeat), Planar=make(TT.Planar),
                                 PlanarHeat=make(TT.PlanarHeat), sparse=TT.sparse)


# the self-checking reference tests (stabilisation, README known answer, load-side adjoint, detached outputs, thermal
# topology gradient) — identical bodies, CPU models
for _name in dir(RS):
    if _name.startswith("test_"):
        globals()[_name] = getattr(RS, _name)


def test_config_a_displacements_forces_and_adjoint(T):
    """BASELINE configs[0] (benchmarks/cubes.py, N = 11) against the reference's vectors."""
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    g = load_case("config_a.npz")
    nodes, elements = cube_hexa(11, 11, 11)
    cube = T.Solid(nodes, elements, IsotropicElasticity3D(E=1000.0, nu=0.3))
    cube.constraints[nodes[:, 0] == 0.0, :] = True
    cube.constraints[nodes[:, 0] == 1.0, 0] = True
    cube.displacements[nodes[:, 0] == 1.0, 0] = 0.1
    cube.forces.requires_grad = True
    u, f, sigma, eps, state = cube.solve(differentiable_parameters=cube.forces)
    assert np.linalg.norm(u.detach().numpy() - g["u"]) <= 1e-9 * np.linalg.norm(g["u"])
    assert np.abs(f.detach().numpy() - g["f"]).max() <= 1e-8 * np.abs(g["f"]).max()
    assert np.abs(sigma.detach().numpy() - g["sigma"]).max() <= 1e-8 * np.abs(g["sigma"]).max()
    u.sum().backward()
    gf = cube.forces.grad.numpy()
    assert np.linalg.norm(gf - g["grad_forces"]) <= 1e-8 * np.linalg.norm(g["grad_forces"])


def test_topology_compliance_gradient(T):
    GM.TestGradients().test_topology_compliance_gradient(T)


def test_hyperelastic_increments_and_parameter_gradient(T):
    GM.TestGradients().test_hyperelastic_increments_and_parameter_gradient(T)


def test_thickness_gradient_through_increments(T):
    GM.TestBase().test_planar_thickness_gradient_incremental_equals_single(T)


@pytest.mark.parametrize("method", [None])
def test_heat_transient_plate_and_flux_gradient(T, method):
    GM.TestHeatTransient().test_plate_matches_reference_and_heat_flux_gradient(T, method)


def test_heat_transient_material_parameter_gradients(T):
    GM.TestHeatTransient().test_material_parameter_gradients(T)


def test_heat_transient_late_start_and_solid_cube(T):
    GM.TestHeatTransient().test_late_start_single_time_and_invalid_times(T)
    GM.TestHeatTransient().test_solid_heat_cube(T)
