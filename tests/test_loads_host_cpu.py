"""Consistent nodal loads (`integrate_body_load`, `integrate_surface_load`, `integrate_line_load`, reference
base.py:446-568) on CPU stand-in models against fixtures from the unmodified reference (`oracle/make_golden.py::loads`):
distorted Hexa1/Hexa2/Tetra1/Tetra2 blocks, Quad/Tria plates with per-element thickness, a heat model; plus the
self-checking properties of reference tests/test_loads.py:121-145, 211-222. Pure torch host code, same on the GPU."""
import numpy as np
import pytest
import torch

from conftest import load_case
from host_standins import host_model


@pytest.fixture(scope="module")
def gold():
    return load_case("loads.npz")


def _close(got, ref, tol=1e-12):
    assert tuple(got.shape) == ref.shape
    assert np.abs(got.numpy() - ref).max() <= tol * max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("tag", ["hexa1", "hexa2", "tetra1", "tetra2"])
def test_solid_loads(gold, tag):
    import torchfem_b200 as T
    from torchfem_b200.materials import IsotropicElasticity3D

    nodes, elements = torch.tensor(gold[f"{tag}.nodes"]), torch.tensor(gold[f"{tag}.elements"])
    model = host_model(T.Solid, nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
    top, boundary = torch.tensor(gold[f"{tag}.top"]), torch.tensor(gold[f"{tag}.boundary"])
    assert np.array_equal(model._boundary_facets(top).numpy(), gold[f"{tag}.facets_top"])
    _close(model.integrate_body_load(torch.tensor([0.0, 0.0, -9.81])), gold[f"{tag}.body"])
    _close(model.integrate_surface_load(top, -2.5), gold[f"{tag}.pressure_top"])
    _close(model.integrate_surface_load(top, torch.tensor([1.0, 0.5, -0.25])), gold[f"{tag}.traction_top"])
    closed = model.integrate_surface_load(boundary, 1.0)
    _close(closed, gold[f"{tag}.pressure_all"], 1e-11)
    assert float(closed.sum(0).abs().max()) <= 1e-12            # pressure on a closed surface: no net force
    assert float(model.integrate_body_load(torch.tensor([0.0, 0.0, -9.81])).sum(0)[2]) == pytest.approx(-9.81 * 1.5)
    with pytest.raises(NotImplementedError, match="no edges to load"):
        model.integrate_line_load(top, 1.0)


@pytest.mark.parametrize("tag", ["quad1", "quad2", "tria1", "tria2"])
def test_planar_loads(gold, tag):
    import torchfem_b200 as T
    from torchfem_b200.materials import IsotropicElasticityPlaneStress

    nodes, elements = torch.tensor(gold[f"{tag}.nodes"]), torch.tensor(gold[f"{tag}.elements"])
    model = host_model(T.Planar, nodes, elements, IsotropicElasticityPlaneStress(1000.0, 0.3),
                       thickness=torch.tensor(gold[f"{tag}.thickness"]))
    right = torch.tensor(gold[f"{tag}.right"])
    _close(model.integrate_body_load(torch.tensor([0.0, -9.81])), gold[f"{tag}.body"])
    _close(model.integrate_line_load(right, 3.0), gold[f"{tag}.pressure_right"])
    _close(model.integrate_line_load(right, torch.tensor([1.0, -2.0])), gold[f"{tag}.traction_right"])
    with pytest.raises(NotImplementedError, match="no surfaces to load"):
        model.integrate_surface_load(right, 1.0)


def test_heat_flux_and_source(gold):
    import torchfem_b200 as T
    from torchfem_b200.materials import IsotropicConductivity3D
    from torchfem_b200.mesh import cube_hexa

    nodes, elements = cube_hexa(3, 3, 3)
    heat = host_model(T.SolidHeat, nodes, elements, IsotropicConductivity3D(1.0))
    flux = heat.integrate_surface_load(nodes[:, 2] == 1.0, 4.0)
    _close(flux, gold["heat.flux_top"])
    assert flux.shape == (27, 1) and float(flux.sum()) == pytest.approx(4.0)
    _close(heat.integrate_body_load(2.0), gold["heat.source"])


def test_load_is_differentiable():
    import torchfem_b200 as T
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    nodes, elements = cube_hexa(3, 3, 3)
    model = host_model(T.Solid, nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
    p = torch.tensor(2.0, requires_grad=True)
    f = model.integrate_surface_load(nodes[:, 2] == 1.0, p)
    f[:, 2].sum().backward()
    assert float(p.grad) == pytest.approx(1.0)       # unit area of the top face
