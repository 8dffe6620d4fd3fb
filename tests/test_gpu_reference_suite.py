"""Reference tests of the solve path that check themselves (no fixture needed), run unchanged in meaning against the
B200 models: viscous stabilisation (reference tests/test_stabilization.py:38-118), the README's known answer
(reference README.md:101-113), load-side adjoint across increments (tests/test_gradients.py:52-76), the detached-output
guard and the thermal topology gradient (tests/test_gradients.py:131-176)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

INCREMENTS = [0.0, 0.25, 0.5, 0.75, 1.0]


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


def _cantilever(T):
    from torchfem_b200.materials import IsotropicElasticityPlaneStress
    from torchfem_b200.mesh import rect_quad

    model = T.Planar(*rect_quad(5, 3, 4.0, 2.0), IsotropicElasticityPlaneStress(E=1000.0, nu=0.3))
    west = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].min())
    east = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].max())
    model.constraints[west] = True
    model.forces[east, 1] = 1.0
    return model


def _minimal(T):
    from torchfem_b200.materials import IsotropicElasticityPlaneStress

    nodes = torch.tensor([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [0.0, 1.0], [1.0, 1.0], [2.0, 1.0]])
    elements = torch.tensor([[0, 1, 4, 3], [1, 2, 5, 4]])
    cantilever = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(E=1000.0, nu=0.3))
    cantilever.forces[5, 1] = -1.0
    cantilever.constraints[[0, 3], :] = True
    return cantilever


def _dense(model, k):
    K = torch.zeros(model.n_dofs, model.n_dofs)
    idx = model.idx.long()
    for e in range(model.n_elem):
        i = idx[e]
        K[i[:, None], i[None, :]] += k[e]
    return K


def test_readme_known_answer(T):
    cantilever = _minimal(T)
    cantilever.thickness.requires_grad = True
    u, f, _, _, _ = cantilever.solve(differentiable_parameters=cantilever.thickness)
    compliance = torch.inner(f.ravel(), u.ravel())
    g = torch.autograd.grad(compliance, cantilever.thickness)[0]
    assert [round(v, 4) for v in g.tolist()] == [-0.0208, -0.0053]


def test_stabilization_is_off_by_default_and_matches_the_abaqus_formulation(T):
    inc = torch.tensor(INCREMENTS)
    plain = _cantilever(T).solve(increments=inc)
    explicit = _cantilever(T).solve(increments=inc, alpha=0.0)
    # bitwise equality, as the reference asserts on its CPU path: every reduction on the path has a fixed order (the
    # nodal force sum `assemble_rhs` is the deterministic gather kernel tfem_assemble_rhs, not index_add_'s atomics)
    assert all(torch.equal(a, b) for a, b in zip(plain, explicit))
    model = _cantilever(T)
    model.solve(increments=inc)
    assert bool(torch.all(model.stabilization_energy == 0.0))

    alpha = 2.0
    model = _cantilever(T)
    u, f, _, _, _ = model.solve(increments=inc, alpha=alpha, return_intermediate=True)
    # incremental solves of (K + alpha/dt M) du = F_ext - K u_prev with dense matrices
    K, M = _dense(model, model.k0()), _dense(model, model.integrate_mass())
    free = torch.nonzero(~model.constraints.ravel()).ravel()
    u_ref = torch.zeros(len(INCREMENTS), model.n_dofs)
    energy_ref = torch.zeros(len(INCREMENTS))
    for n in range(1, len(INCREMENTS)):
        c = alpha / (INCREMENTS[n] - INCREMENTS[n - 1])
        A = (K + c * M)[free[:, None], free[None, :]]
        b = (INCREMENTS[n] * model.forces.ravel() - K @ u_ref[n - 1])[free]
        du = torch.zeros(model.n_dofs)
        du[free] = torch.linalg.solve(A, b)
        u_ref[n] = u_ref[n - 1] + du
        energy_ref[n] = energy_ref[n - 1] + c * du @ (M @ du)
    assert torch.allclose(u.reshape(len(INCREMENTS), -1), u_ref)
    assert torch.allclose(model.stabilization_energy, energy_ref)
    # the returned nodal forces include the viscous part: free DOFs stay in balance
    for n in range(1, len(INCREMENTS)):
        residual = f[n].ravel() - INCREMENTS[n] * model.forces.ravel()
        assert torch.allclose(residual[free], torch.zeros_like(residual[free]), atol=1e-8)


def test_stabilization_vanishes_for_small_damping(T):
    inc = torch.tensor(INCREMENTS)
    u_ref = _cantilever(T).solve(increments=inc)[0]
    errors = [float((_cantilever(T).solve(increments=inc, alpha=a)[0] - u_ref).abs().max()) for a in (1e-2, 1e-3, 1e-4)]
    assert errors[0] > errors[1] > errors[2]
    assert errors[-1] < 1e-4 * float(u_ref.abs().max())


def test_force_gradient_through_increments_equals_single_step(T):
    grads = []
    for increments in (None, torch.linspace(0.1, 1.0, 5)):
        cantilever = _minimal(T)
        forces = torch.zeros_like(cantilever.nodes)
        forces[5, 1] = -1.0
        cantilever.forces = forces
        cantilever.forces.requires_grad = True
        if increments is None:
            u = cantilever.solve(differentiable_parameters=cantilever.forces)[0]
        else:
            u = cantilever.solve(increments=increments, return_intermediate=True,
                                 differentiable_parameters=cantilever.forces)[0][-1]
        grads.append(torch.autograd.grad(u.sum(), cantilever.forces)[0])
    assert torch.allclose(grads[1], grads[0], atol=1e-9, rtol=1e-7)


def test_outputs_are_detached_without_differentiable_parameters(T):
    cantilever = _minimal(T)
    rho = torch.ones(cantilever.n_elem, requires_grad=True)
    cantilever.thickness = rho ** 3
    out = cantilever.solve()
    assert not any(t.requires_grad for t in out)
    assert not torch.inner(out[1].ravel(), out[0].ravel()).requires_grad


def test_thermal_topology_gradient_is_finite(T):
    from torchfem_b200.materials import IsotropicConductivity2D
    from torchfem_b200.mesh import rect_quad

    model = T.PlanarHeat(*rect_quad(5, 5, 1.0, 1.0), IsotropicConductivity2D(kappa=400.0))
    west = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].min())
    north = torch.isclose(model.nodes[:, 1], model.nodes[:, 1].max())
    model.constraints[west | north] = True
    model.temperatures[west | north] = 0.0
    volume = model.integrate_field()
    model.heat_flux[:, 0] = model.assemble_rhs(
        (1000.0 * volume / volume.sum()).unsqueeze(1).repeat(1, model.etype.nodes)) / model.etype.nodes
    rho_nodes = 0.4 * torch.ones(len(model.nodes), requires_grad=True)
    N, _, _ = model.eval_shape_functions(model.etype.ipoints.sum(dim=0))
    model.thickness = torch.einsum("EN, N -> E", rho_nodes[model.elements], N) ** 3.0
    temperature, internal_force, _, _, _ = model.solve(differentiable_parameters=rho_nodes)
    compliance = torch.inner(internal_force.ravel(), temperature.ravel())
    sensitivity = torch.autograd.grad(compliance, rho_nodes)[0]
    assert bool(torch.isfinite(sensitivity).all()) and float(sensitivity.abs().max()) > 0.0
