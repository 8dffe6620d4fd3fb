"""`HyperelasticPlaneStress` (reference hyperelasticity.py:130-269) against fixtures from the
unmodified reference (`oracle/make_golden.py::hyper_plane_stress`): the material update with its local Newton iteration
on the thickness stretch, and a Neo-Hookean strip stretched by 30 % in three `nlgeom` increments through `Planar.solve`
(host stand-ins for the kernels; the state variable is carried by the solve)."""
import numpy as np
import torch

from conftest import load_case
from host_standins import dense_sparse_solve, host_model

MU, LBD = 384.6153846153846, 576.9230769230769


def psi(F, params):
    Cg = F.transpose(-1, -2) @ F
    logJ = 0.5 * torch.logdet(Cg)
    return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2


def _rel(a, b):
    return float(np.abs(a.detach().numpy() - b).max() / np.abs(b).max())


def test_material_update():
    from torchfem_b200.materials import HyperelasticPlaneStress

    g = load_case("hyper_plane_stress.npz")
    mat = HyperelasticPlaneStress(psi, torch.tensor([MU, LBD])).vectorize(6)
    assert mat.n_state == 1
    P, state, tangent = mat.step(torch.tensor(g["H"]), torch.tensor(g["F"]), torch.zeros(6, 2, 2),
                                 torch.tensor(g["state"]), torch.zeros(6, 2, 2), torch.ones(6, 1), 0)
    assert _rel(P, g["P"]) <= 1e-12 and _rel(state, g["state_new"]) <= 1e-12 and _rel(tangent, g["ddsdde"]) <= 1e-12
    # all Gauss points at once (the batched entry the device path uses): same values on a [n_int, n_elem] batch
    stack = lambda a: torch.tensor(a).expand(2, *a.shape)
    P2, state2, tangent2 = mat.step_points(stack(g["H"]), stack(g["F"]), torch.zeros(2, 6, 2, 2), stack(g["state"]),
                                           torch.zeros(6, 2, 2), torch.ones(6, 1), 0)
    assert P2.shape == (2, 6, 2, 2) and state2.shape == (2, 6, 1) and tangent2.shape == (2, 6, 2, 2, 2, 2)
    assert _rel(P2[1], g["P"]) <= 1e-12 and _rel(tangent2[0], g["ddsdde"]) <= 1e-12


def test_stretched_strip(monkeypatch):
    import torchfem_b200 as T
    from torchfem_b200.materials import HyperelasticPlaneStress
    from torchfem_b200.mesh import rect_quad

    monkeypatch.setattr(T.sparse, "sparse_solve", dense_sparse_solve)
    g = load_case("hyper_plane_stress.npz")
    nodes, elements = rect_quad(5, 3, 2.0, 1.0)
    strip = host_model(T.Planar, nodes, elements, HyperelasticPlaneStress(psi, torch.tensor([MU, LBD])))
    left, right = nodes[:, 0] == 0.0, nodes[:, 0] == 2.0
    strip.constraints[left, 0] = True
    strip.constraints[right, 0] = True
    strip.constraints[nodes[:, 1] == 0.5, 1] = True
    strip.displacements[right, 0] = 0.6
    u, f, sigma, F, alpha = strip.solve(increments=torch.linspace(0.0, 1.0, 4), nlgeom=True)
    assert _rel(u, g["strip.u"]) <= 1e-8 and _rel(f, g["strip.f"]) <= 1e-7
    assert _rel(sigma, g["strip.sigma"]) <= 1e-7 and _rel(alpha, g["strip.state"]) <= 1e-7
    assert alpha.shape == (len(elements), 1) and float(alpha.mean()) < 0.0       # the strip thins
