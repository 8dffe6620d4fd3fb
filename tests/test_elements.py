"""Element tables of torch-fem_b200 against fixtures dumped from the reference's torchfem.elements
(oracle/make_golden.py -> tests/golden/element_tables.npz). Mirrors reference tests/test_elements.py:21-64."""
import numpy as np
import pytest
import torch

NAMES = ["Tria1", "Tria2", "Quad1", "Quad2", "Tetra1", "Tetra2", "Hexa1", "Hexa2"]


@pytest.fixture(scope="module")
def E():
    torch.set_default_device("cpu")
    from torchfem_b200 import elements

    return elements


@pytest.mark.parametrize("name", NAMES)
def test_tables_identical_to_reference(E, tables, name):
    c = getattr(E, name)
    assert [c.nodes, c.iso_dim] == tables[f"{name}.meta"].tolist()
    assert c.iso_volume == float(tables[f"{name}.iso_volume"])
    assert np.array_equal(c.iso_coords.numpy(), tables[f"{name}.iso_coords"])
    assert np.array_equal(c.ipoints.numpy(), tables[f"{name}.ipoints"])  # bit-exact, incl. 8-digit literals
    w = c.iweights.numpy()
    assert w.dtype == tables[f"{name}.iweights"].dtype and np.array_equal(w, tables[f"{name}.iweights"])
    assert np.array_equal(c.edges.numpy(), tables[f"{name}.edges"])


@pytest.mark.parametrize("name", NAMES)
def test_shape_functions_match_reference(E, tables, name):
    c = getattr(E, name)
    for pts, tagN, tagB in [(c.ipoints, "N_ip", "B_ip"), (torch.as_tensor(tables[f"{name}.xi"]), "N_xi", "B_xi")]:
        N = c.N(pts).numpy()
        B = c.B(pts).numpy()
        assert N.shape == tables[f"{name}.{tagN}"].shape and B.shape == tables[f"{name}.{tagB}"].shape
        assert np.abs(N - tables[f"{name}.{tagN}"]).max() <= 4e-16
        assert np.abs(B - tables[f"{name}.{tagB}"]).max() <= 1e-15
    # single point (reference tests/test_gradients.py:149 evaluates at one xi)
    one = torch.as_tensor(tables[f"{name}.xi"][0])
    assert c.N(one).shape == (c.nodes,) and c.B(one).shape == (c.iso_dim, c.nodes)


@pytest.mark.parametrize("name", NAMES)
def test_partition_of_unity_kronecker_and_gradient(E, name):
    c = getattr(E, name)
    xi = c.iso_coords.double()
    assert torch.allclose(c.N(xi), torch.eye(c.nodes).double(), atol=1e-14)
    pts = c.ipoints.double().clone().requires_grad_(True)
    N = c.N(pts)
    assert torch.allclose(N.sum(-1), torch.ones(len(pts)).double())
    for a in range(c.nodes):
        (g,) = torch.autograd.grad(N[:, a].sum(), pts, retain_graph=True)
        assert torch.allclose(g, c.B(pts)[:, :, a].detach(), atol=1e-12)
    assert abs(float(c.iweights.sum()) - c.iso_volume) < 1e-6


def test_linear_to_quadratic_matches_reference_numbering(E):
    from conftest import load_case

    for lin, quad in [("hexa1", "hexa2"), ("tetra1", "tetra2"), ("quad1", "quad2"), ("tria1", "tria2")]:
        a, b = load_case(f"case_{lin}.npz"), load_case(f"case_{quad}.npz")
        n2, e2 = E.linear_to_quadratic(torch.as_tensor(a["nodes"]), torch.as_tensor(a["elements"]))
        assert np.array_equal(e2.numpy(), b["elements"])
        assert np.array_equal(n2.numpy(), b["nodes"])
