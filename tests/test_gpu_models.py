"""Model-level GPU parity: `Solid` / `Planar` / `SolidHeat` / `PlanarHeat` and `torchfem_b200.sparse`
used exactly like the reference's classes (these tests mirror reference tests/test_models.py:54-71,
tests/test_sparse.py:49-206, tests/test_base.py:77-128,209-220, tests/test_gradients.py) and are checked
against fixtures generated from the unmodified reference (tests/golden/, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def cuda_default():
    """Run like the reference's GPU benchmarks do: default device cuda, float64
    (reference benchmarks/utils.py:59-60)."""
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda")
    yield
    torch.set_default_device("cpu")


@pytest.fixture(scope="module")
def T():
    import torchfem_b200 as T

    return T


ETYPES = ["Tria1", "Tria2", "Quad1", "Quad2", "Tetra1", "Tetra2", "Hexa1", "Hexa2"]


def _build(T, etype):
    from torchfem_b200.elements import linear_to_quadratic
    from torchfem_b200.materials import IsotropicElasticity3D, IsotropicElasticityPlaneStress
    from torchfem_b200.mesh import cube_hexa, cube_tetra, rect_quad, rect_tri

    planar = IsotropicElasticityPlaneStress(1000.0, 0.3)
    solid = IsotropicElasticity3D(1000.0, 0.3)
    cases = {
        "Tria1": (rect_tri(4, 4), T.Planar, planar, False), "Tria2": (rect_tri(4, 4), T.Planar, planar, True),
        "Quad1": (rect_quad(4, 4), T.Planar, planar, False), "Quad2": (rect_quad(4, 4), T.Planar, planar, True),
        "Tetra1": (cube_tetra(3, 3, 3), T.Solid, solid, False), "Tetra2": (cube_tetra(3, 3, 3), T.Solid, solid, True),
        "Hexa1": (cube_hexa(3, 3, 3), T.Solid, solid, False), "Hexa2": (cube_hexa(3, 3, 3), T.Solid, solid, True),
    }
    mesh, model, material, quadratic = cases[etype]
    nodes, elements = linear_to_quadratic(*mesh) if quadratic else mesh
    return model(nodes, elements, material)


def _on_boundary(nodes):
    mask = torch.zeros(len(nodes), dtype=torch.bool)
    for dim in range(nodes.shape[1]):
        c = nodes[:, dim]
        mask |= torch.isclose(c, c.min()) | torch.isclose(c, c.max())
    return mask


class TestPatch:
    @pytest.mark.parametrize("etype", ETYPES)
    @pytest.mark.parametrize("method", [None, "cg", "amgx"])
    def test_reproduces_a_linear_displacement_field(self, T, etype, method):
        """First-order patch test for all 8 element types (reference tests/test_models.py:54-71),
        through the default direct method and through the Jacobi-CG kernels at stol 1e-14."""
        model = _build(T, etype)
        assert model.etype.__name__ == etype
        g2 = torch.tensor([[1.0e-3, 2.0e-4], [3.0e-4, -1.0e-3]])
        g3 = torch.tensor([[1.0e-3, 2.0e-4, -1.0e-4], [3.0e-4, -1.0e-3, 5.0e-5], [1.0e-4, 2.0e-4, 7.0e-4]])
        gradient = g2 if model.n_dim == 2 else g3
        u_exact = model.nodes @ gradient.T
        model.constraints = _on_boundary(model.nodes)[:, None].repeat(1, model.n_dof_per_node)
        model.displacements = u_exact
        u, _, sigma, _, _ = model.solve(method=method, stol=1e-14)
        assert torch.allclose(u, u_exact, atol=1e-12)
        assert torch.allclose(sigma, sigma[0].expand_as(sigma), atol=1e-12)

    def test_rejects_unsupported_connectivity(self, T):
        from torchfem_b200.materials import IsotropicElasticity3D

        with pytest.raises(ValueError, match="Element type not supported."):
            T.Solid(torch.rand(5, 3), torch.tensor([[0, 1, 2, 3, 4]]), IsotropicElasticity3D(1.0, 0.3))


class TestConfigA:
    """BASELINE configs[0]: benchmarks/cubes.py at N=11 through the model API vs the reference's vectors."""

    def _cube(self, T):
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(11, 11, 11)
        cube = T.Solid(nodes, elements, IsotropicElasticity3D(E=1000.0, nu=0.3))
        cube.forces = torch.zeros_like(nodes, requires_grad=True)
        cube.constraints[nodes[:, 0] == 0.0, :] = True
        cube.constraints[nodes[:, 0] == 1.0, 0] = True
        cube.displacements[nodes[:, 0] == 1.0, 0] = 0.1
        return cube

    @pytest.mark.parametrize("method", ["spsolve", "cg", "minres", "amgx"])
    def test_forward_and_adjoint(self, T, method):
        g = load_case("config_a.npz")
        cube = self._cube(T)
        assert cube.idx.dtype == torch.int32 and cube.k_map.dtype == torch.int32
        import hashlib
        assert hashlib.sha256(cube.idx.cpu().numpy().tobytes()).hexdigest()[:16] == str(g["sha_idx"])
        u, f, sigma, eps, state = cube.solve(differentiable_parameters=cube.forces, method=method, stol=1e-10)
        nrm = np.linalg.norm(g["u"])
        assert np.linalg.norm(u.detach().cpu().numpy() - g["u"]) / nrm <= 1e-8
        assert np.abs(f.detach().cpu().numpy() - g["f"]).max() <= 1e-6 * np.abs(g["f"]).max()
        assert np.abs(sigma.detach().cpu().numpy() - g["sigma"]).max() <= 1e-6 * np.abs(g["sigma"]).max()
        assert np.abs(eps.detach().cpu().numpy() - g["eps"]).max() <= 1e-7
        assert state.shape == (cube.n_elem, 0)
        u.sum().backward()
        gf = cube.forces.grad.cpu().numpy()
        assert np.linalg.norm(gf - g["grad_forces"]) / np.linalg.norm(g["grad_forces"]) <= 1e-7

    def test_k0_and_assembled_matrix(self, T):
        g = load_case("config_a.npz")
        cube = self._cube(T)
        k = cube.k0()
        assert np.abs(k[0].cpu().numpy() - g["k_e0"]).max() <= 1e-12 * g["k_absmax"]
        con = torch.nonzero(cube.constraints.ravel()).ravel()
        K = cube.assemble_matrix(k, con)
        v = K._values().cpu().numpy()
        assert abs(np.linalg.norm(v) - g["val_norm"]) <= 1e-12 * g["val_norm"]
        assert K.shape == (3993, 3993) and K._indices().shape == (2, 268119)
        dense = K.to_dense()
        assert torch.allclose(dense, dense.T, atol=1e-10)


class TestGradients:
    def test_topology_compliance_gradient(self, T):
        """benchmarks/topopt.py at N=3: compliance and d(compliance)/d(rho) vs the reference."""
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        g = load_case("topopt_n3.npz")
        nodes, elements = cube_hexa(7, 4, 4, 2.0, 1.0, 1.0)
        rho = torch.tensor(g["rho"], requires_grad=True)
        material = IsotropicElasticity3D(E=70000.0, nu=0.3).vectorize(len(elements))
        scale = 1e-3 + (1.0 - 1e-3) * rho ** 3.0
        material.C = scale[:, None, None, None, None] * material.C
        model = T.Solid(nodes, elements, material)
        model.constraints[nodes[:, 0] == 0.0, :] = True
        model.forces = torch.tensor(g["forces"])
        for method, tol in [("spsolve", 1e-9), ("cg", 1e-7)]:
            rho.grad = None
            u, *_ = model.solve(differentiable_parameters=rho, method=method, stol=1e-12)
            c = torch.inner(model.forces.ravel(), u.ravel())
            c.backward()
            assert abs(float(c) - float(g["compliance"])) <= tol * abs(float(g["compliance"]))
            gr = rho.grad.cpu().numpy()
            assert np.linalg.norm(gr - g["grad_rho"]) / np.linalg.norm(g["grad_rho"]) <= tol * 10

    def test_hyperelastic_increments_and_parameter_gradient(self, T):
        """benchmarks/hyperelasticity.py at N=3, 3 increments, nlgeom: reaction force and
        d(reaction)/d(mu, lambda) through the chained adjoint vs the reference."""
        import math

        from torchfem_b200.materials import Hyperelastic3D
        from torchfem_b200.mesh import cube_hexa

        g = load_case("hyper_n3.npz")
        En, NU = 1000.0, 0.3
        LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
        MU = En / (2.0 * (1.0 + NU))

        def psi(F, params):
            Cg = F.transpose(-1, -2) @ F
            logJ = 0.5 * torch.logdet(Cg)
            return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2

        lx = 2.0
        nodes, elements = cube_hexa(5, 3, 3, lx, 1.0, 1.0)
        params = torch.tensor([MU, LBD], requires_grad=True)
        box = T.Solid(nodes, elements, Hyperelastic3D(psi, params))
        left, right = nodes[:, 0] == 0.0, nodes[:, 0] == lx
        box.constraints[left, 0] = True
        box.constraints[right, 0] = True
        box.constraints[nodes[:, 1] == 0.5, 1] = True
        box.constraints[nodes[:, 2] == 0.5, 2] = True
        box.displacements[right, 0] = lx
        increments = torch.tensor(g["increments"])
        u, f, *_ = box.solve(increments=increments, nlgeom=True, differentiable_parameters=params,
                             method="spsolve")
        reaction = f[right, 0].sum()
        reaction.backward()
        assert np.abs(u.detach().cpu().numpy() - g["u"]).max() <= 1e-8 * np.abs(g["u"]).max()
        assert abs(float(reaction) - float(g["reaction"])) <= 1e-8 * abs(float(g["reaction"]))
        assert np.allclose(params.grad.cpu().numpy(), g["grad_params"], rtol=1e-6)

    def test_outputs_detached_without_parameters(self, T):
        model = _build(T, "Quad1")
        model.constraints[model.nodes[:, 0] == 0.0, :] = True
        model.forces[model.nodes[:, 0] == 1.0, 0] = 1.0
        u, f, s, e, a = model.solve()
        assert not any(t.requires_grad for t in (u, f, s, e, a))


class TestSparse:
    """reference tests/test_sparse.py:49-206 against torchfem_b200.sparse."""

    def _spd(self, n, seed):
        gen = torch.Generator(device="cpu").manual_seed(seed)
        A = torch.rand(n, n, generator=gen, device="cpu")
        return (A @ A.T + n * torch.eye(n, device="cpu")).cuda()

    def _sp(self, dense):
        return dense.to_sparse_coo().coalesce()

    @pytest.mark.parametrize("method", [None, "spsolve", "cg", "minres", "amgx"])
    def test_matches_dense_solution(self, T, method):
        A = self._spd(6, 0)
        b = torch.randn(6)
        x, M = T.sparse.sparse_solve(self._sp(A), b, method=method)
        assert torch.allclose(x, torch.linalg.solve(A, b), atol=1e-8)
        assert (M is None) == (method in (None, "spsolve"))

    def test_errors(self, T):
        A = torch.sparse_coo_tensor(torch.tensor([[0, 1], [0, 1]]), torch.ones(2), (2, 3)).coalesce()
        with pytest.raises(ValueError, match="square 2D matrix"):
            T.sparse.sparse_solve(A, torch.ones(2))
        with pytest.raises(ValueError, match="is not supported"):
            T.sparse.sparse_solve(self._sp(self._spd(6, 0)), torch.randn(6), method="not-a-solver")

    def test_initial_guess_does_not_change_the_solution(self, T):
        A = self._spd(6, 0)
        b = torch.randn(6)
        ref = torch.linalg.solve(A, b)
        x, _ = T.sparse.sparse_solve(self._sp(A), b, method="cg", x0=ref + 0.1 * torch.randn(6))
        assert torch.allclose(x, ref, atol=1e-8)

    @pytest.mark.parametrize("tag", ["spd", "nonsym"])
    @pytest.mark.parametrize("method", ["spsolve", "cg"])
    def test_adjoint_matches_reference(self, T, tag, method):
        """Forward, dL/db and dL/dA (on A's pattern) for an SPD and a NON-symmetric matrix — pins the
        transpose in the adjoint (reference tests/test_sparse.py:94-158); CG only for the SPD one."""
        if tag == "nonsym" and method == "cg":
            pytest.skip("CG needs an SPD matrix")
        g = load_case("sparse_small.npz")
        Ad = torch.tensor(g[f"{tag}.A"])
        A = Ad.to_sparse_coo().detach().requires_grad_(True)
        b = torch.tensor(g["b"]).requires_grad_(True)
        x = T.sparse.differentiable_sparse_solve(A, b, method=method, stol=1e-13)
        x.sum().backward()
        assert np.allclose(x.detach().cpu().numpy(), g[f"{tag}.x"], atol=1e-10)
        assert np.allclose(b.grad.cpu().numpy(), g[f"{tag}.gb"], atol=1e-10)
        gA = A.grad.coalesce()
        assert np.array_equal(gA.indices().cpu().numpy(), g[f"{tag}.gA_idx"])
        assert np.allclose(gA.values().cpu().numpy(), g[f"{tag}.gA_val"], atol=1e-10)

    def test_cached_solve_warm_start(self, T):
        A = self._sp(self._spd(6, 1))
        b = torch.randn(6)
        cache = T.sparse.CachedSolve()
        x1 = T.sparse.differentiable_sparse_solve(A, b, method="cg", cached_solve=cache, update_cache=True)
        assert cache.previous_x is not None and not cache.previous_x.requires_grad
        x2 = T.sparse.differentiable_sparse_solve(A, b, method="cg", cached_solve=cache)
        assert torch.allclose(x1, x2, atol=1e-8)

    def test_resolve_method_policy(self, T):
        assert T.sparse.resolve_method(9999, "cuda", None) == "spsolve"
        assert T.sparse.resolve_method(10000, "cuda", None) == "minres"
        assert T.sparse.resolve_method(5, "cuda", "cg") == "cg"
        assert T.sparse.describe_method(10000, "cuda", None) == "minres | iterative | jacobi | tfem_b200 | cuda"
        # AMG (the reference's CUDA default when its AMG backend is installed) above the measured crossover
        assert T.sparse.resolve_method(T.sparse.AMG_MIN_DOFS, "cuda", None) == "amgx"
        assert T.sparse.resolve_method(T.sparse.AMG_MIN_DOFS - 1, "cuda", None) == "minres"
        assert T.sparse.describe_method(3_000_000, "cuda", None) == "amgx | iterative | amg | tfem_b200 | cuda"
        assert "amgx" in T.sparse.available_backends


class TestBase:
    @pytest.mark.parametrize("kind", ["solid", "planar", "heat"])
    def test_compute_B_and_idx_match_reference(self, T, kind):
        """`FEM.compute_B` (rigid-body near-null space, reference base.py:316-344) and the DOF map `idx`
        (base.py:73-76,119) against the unmodified reference (fixture compute_B.npz, oracle/make_golden.py)."""
        from torchfem_b200.materials import (IsotropicConductivity3D, IsotropicElasticity3D,
                                             IsotropicElasticityPlaneStress)

        g = load_case("compute_B.npz")
        src = "planar" if kind == "planar" else "solid"
        nodes, elements = torch.as_tensor(g[f"{src}.nodes"]).cuda(), torch.as_tensor(g[f"{src}.elements"]).cuda()
        if kind == "solid":
            model = T.Solid(nodes, elements, IsotropicElasticity3D(E=1000.0, nu=0.3))
        elif kind == "planar":
            model = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(E=1000.0, nu=0.3))
        else:
            model = T.SolidHeat(nodes, elements, IsotropicConductivity3D(kappa=2.0, rho=1.0))
        B = model.compute_B()
        assert B.shape == g[f"{kind}.B"].shape and B.dtype == torch.float64
        assert np.array_equal(B.cpu().numpy(), g[f"{kind}.B"])          # copies of coordinates: exact
        assert model.idx.dtype == torch.int32 and np.array_equal(model.idx.cpu().numpy(), g[f"{kind}.idx"])

    def test_inverted_element_raises(self, T):
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(2, 2, 2)
        elements = elements[:, [1, 0, 3, 2, 5, 4, 7, 6]]
        model = T.Solid(nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
        with pytest.raises(ValueError, match="Negative Jacobian"):
            model.k0()

    def test_heat_k0_symmetric_zero_row_sums_and_scaling(self, T):
        from torchfem_b200.materials import IsotropicConductivity2D, IsotropicConductivity3D
        from torchfem_b200.mesh import cube_hexa, rect_quad

        m = T.PlanarHeat(*rect_quad(3, 3), IsotropicConductivity2D(kappa=400.0))
        k = m.k0()
        assert k.shape == (m.n_elem, 4, 4)
        assert torch.allclose(k, k.transpose(-1, -2))
        assert torch.allclose(k.sum(-1), torch.zeros(m.n_elem, 4), atol=1e-10)
        k2 = T.PlanarHeat(*rect_quad(3, 3), IsotropicConductivity2D(800.0)).k0()
        assert torch.allclose(k2, 2.0 * k)
        s = T.SolidHeat(*cube_hexa(3, 3, 3), IsotropicConductivity3D(400.0))
        ks = s.k0()
        assert ks.shape == (s.n_elem, 8, 8) and torch.allclose(ks.sum(-1), torch.zeros(s.n_elem, 8), atol=1e-10)

    def test_heat_solve_linear_temperature(self, T):
        """1-D conduction through a slab: prescribed T on two faces gives a linear profile."""
        from torchfem_b200.materials import IsotropicConductivity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(5, 3, 3)
        m = T.SolidHeat(nodes, elements, IsotropicConductivity3D(10.0))
        m.constraints[nodes[:, 0] == 0.0, 0] = True
        m.constraints[nodes[:, 0] == 1.0, 0] = True
        m.temperatures[nodes[:, 0] == 1.0, 0] = 100.0
        Tn, q, flux, grad, _ = m.solve()
        assert torch.allclose(Tn[:, 0], 100.0 * nodes[:, 0], atol=1e-9)
        assert flux.shape == (m.n_elem, 3) and torch.allclose(flux[:, 0], torch.full((m.n_elem,), 1000.0), atol=1e-7)

    def test_output_shapes(self, T):
        """reference tests/test_base.py:209-220: integration-point axis and increment axis."""
        model = _build(T, "Hexa1")
        model.constraints[model.nodes[:, 0] == 0.0, :] = True
        model.forces[model.nodes[:, 0] == 1.0, 0] = 1.0
        u, f, s, e, a = model.solve(aggregate_integration_points=False)
        assert s.shape == (8, model.n_elem, 3, 3) and u.shape == (model.n_nod, 3)
        inc = torch.tensor([0.0, 0.5, 1.0])
        u, f, s, e, a = model.solve(increments=inc, return_intermediate=True)
        assert u.shape == (3, model.n_nod, 3) and s.shape == (3, model.n_elem, 3, 3)
        assert torch.allclose(u[1] * 2, u[2], atol=1e-10)

    def test_planar_thickness_gradient_incremental_equals_single(self, T):
        """reference tests/test_gradients.py:27-49: adjoint through increments == single step."""
        from torchfem_b200.materials import IsotropicElasticityPlaneStress
        from torchfem_b200.mesh import rect_quad

        nodes, elements = rect_quad(4, 3)
        grads = []
        for inc in (torch.tensor([0.0, 1.0]), torch.tensor([0.0, 0.3, 0.7, 1.0])):
            th = torch.full((len(elements),), 0.5, requires_grad=True)
            m = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(1000.0, 0.3), thickness=th)
            m.constraints[nodes[:, 0] == 0.0, :] = True
            m.forces[nodes[:, 0] == 1.0, 1] = -1.0
            u, *_ = m.solve(increments=inc, differentiable_parameters=th)
            (u ** 2).sum().backward()
            grads.append(th.grad.clone())
        assert torch.allclose(grads[0], grads[1], atol=1e-9, rtol=1e-7)


class TestHeatTransient:
    """`Heat.time_integration` (reference base.py:1288-1553, tests/test_time_integration.py) against vectors
    generated from the unmodified reference (oracle/make_golden.py::heat_transient)."""

    def _plate(self, T):
        from torchfem_b200.materials import IsotropicConductivity2D
        from torchfem_b200.mesh import rect_quad

        model = T.PlanarHeat(*rect_quad(5, 5, 1.0, 1.0), IsotropicConductivity2D(kappa=400.0, rho=1.0e5))
        west = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].min())
        east = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].max())
        model.constraints[west | east] = True
        model.temperatures[west, 0] = 5.0
        model.temperatures[east, 0] = 20.0
        return model

    @pytest.mark.parametrize("method", [None, "cg"])
    def test_plate_matches_reference_and_heat_flux_gradient(self, T, method):
        g = load_case("heat_transient.npz")
        model = self._plate(T)
        model.heat_flux = torch.tensor(g["plate.heat_flux"]).requires_grad_(True)
        temp, rfl, flux, grad, state = model.time_integration(torch.tensor(g["plate.t_out"]), delta_t=1.0,
                                                              method=method, stol=1e-12)
        for got, key in ((temp, "plate.temp"), (rfl, "plate.rfl"), (flux, "plate.flux"), (grad, "plate.grad")):
            ref = g[key]
            assert got.shape == ref.shape
            assert np.abs(got.detach().cpu().numpy() - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
        assert state.shape[0] == 5
        temp[-1].sum().backward()
        gh = model.heat_flux.grad.cpu().numpy()
        assert np.abs(gh - g["plate.grad_heat_flux"]).max() <= 1e-8 * np.abs(g["plate.grad_heat_flux"]).max()

    def test_material_parameter_gradients(self, T):
        """d(sum T_end^2)/d(kappa_e, rho_e): flows through the system matrix M + dt/2 K of every step and through
        M @ du (reference: differentiable assemble_matrix + Solve.backward's gradA, sparse.py:212-216)."""
        from torchfem_b200.materials import IsotropicConductivity2D
        from torchfem_b200.mesh import rect_quad

        g = load_case("heat_transient.npz")
        kappa = torch.tensor(g["het.kappa"]).requires_grad_(True)
        rho = torch.tensor(g["het.rho"]).requires_grad_(True)
        model = T.PlanarHeat(*rect_quad(5, 5, 1.0, 1.0), IsotropicConductivity2D(kappa=kappa, rho=rho))
        west = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].min())
        east = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].max())
        model.constraints[west | east] = True
        model.temperatures[west, 0] = 5.0
        model.temperatures[east, 0] = 20.0
        model.heat_flux = torch.tensor(g["plate.heat_flux"])
        temp, *_ = model.time_integration(torch.tensor(g["het.t_out"]), delta_t=1.0, stol=1e-13,
                                          differentiable_parameters=[kappa, rho])
        assert np.abs(temp.detach().cpu().numpy() - g["het.temp"]).max() <= 1e-8 * np.abs(g["het.temp"]).max()
        (temp[-1] ** 2).sum().backward()
        for got, key in ((kappa.grad, "het.grad_kappa"), (rho.grad, "het.grad_rho")):
            ref = g[key]
            assert np.abs(got.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max(), key

    def test_late_start_single_time_and_invalid_times(self, T):
        g = load_case("heat_transient.npz")
        model = self._plate(T)
        model.heat_flux = torch.tensor(g["plate.heat_flux"])
        late, *_ = model.time_integration(torch.tensor([5.0, 10.0]), delta_t=0.5)
        assert np.abs(late.cpu().numpy() - g["plate.late"]).max() <= 1e-8 * np.abs(g["plate.late"]).max()
        one, _, fl, gr, _ = model.time_integration(torch.tensor([10.0]), delta_t=1.0)
        assert one.shape[0] == fl.shape[0] == gr.shape[0] == 1
        for bad in (torch.tensor([]), torch.tensor([-1.0, 1.0]), torch.tensor([0.0, 2.0, 2.0])):
            with pytest.raises(ValueError, match="t_output must"):
                model.time_integration(bad)
        assert bool(model.constraints.sum() == 10)  # boundary conditions restored

    def test_solid_heat_cube(self, T):
        from torchfem_b200.materials import IsotropicConductivity3D
        from torchfem_b200.mesh import cube_hexa

        g = load_case("heat_transient.npz")
        nodes, elements = cube_hexa(4, 4, 4)
        cube = T.SolidHeat(nodes, elements, IsotropicConductivity3D(kappa=2.0, rho=30.0))
        cube.constraints[nodes[:, 0] == 0.0] = True
        cube.temperatures[nodes[:, 0] == 0.0, 0] = 1.0
        cube.heat_flux[nodes[:, 0] == 1.0, 0] = 0.05
        temp, _, flux, grad, _ = cube.time_integration(torch.tensor(g["cube.t_out"]), delta_t=0.25)
        for got, key in ((temp, "cube.temp"), (flux, "cube.flux"), (grad, "cube.grad")):
            assert np.abs(got.cpu().numpy() - g[key]).max() <= 1e-8 * max(1.0, np.abs(g[key]).max())


class TestModelsWithAMG:
    """method="amgx" through the model API on problems large enough for a multi-level hierarchy, against the
    Jacobi-CG kernels at the same tolerance (both validated against the reference on config A)."""

    def test_heat_cube_scalar_blocks(self, T):
        from torchfem_b200.materials import IsotropicConductivity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(23, 23, 23)          # 12,167 unknowns, d = 1
        out = {}
        for method in ("cg", "amgx"):
            m = T.SolidHeat(nodes, elements, IsotropicConductivity3D(10.0))
            m.constraints[nodes[:, 0] == 0.0, 0] = True
            m.constraints[nodes[:, 0] == 1.0, 0] = True
            m.temperatures[nodes[:, 0] == 1.0, 0] = 100.0
            m.heat_flux[nodes[:, 1] == 1.0, 0] = 3.0
            out[method], *_ = m.solve(method=method, stol=1e-12)
        assert float((out["amgx"] - out["cg"]).abs().max()) <= 1e-8 * float(out["cg"].abs().max())

    def test_hierarchy_is_cached_with_the_pattern_across_load_cases(self, T):
        """Second solve on the same model with other Dirichlet conditions: the hierarchy kept with the pattern is
        refreshed (new isolated-DOF mask, same aggregates), the result still equals the Jacobi-CG path."""
        from torchfem_b200.amg import AMGPreconditioner
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(13, 9, 9, 2.0, 1.0, 1.0)
        box = T.Solid(nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
        cases = [(nodes[:, 0] == 0.0, nodes[:, 0] == 2.0, 0), (nodes[:, 2] == 0.0, nodes[:, 2] == 1.0, 2)]
        cached = []
        for fixed, pulled, axis in cases:
            out = {}
            for method in ("cg", "amgx"):
                box.constraints = torch.zeros_like(nodes, dtype=torch.bool)
                box.displacements = torch.zeros_like(nodes)
                box.constraints[fixed, :] = True
                box.constraints[pulled, axis] = True
                box.displacements[pulled, axis] = 0.05
                out[method], *_ = box.solve(method=method, stol=1e-11)
            assert float((out["amgx"] - out["cg"]).abs().max()) <= 1e-8 * float(out["cg"].abs().max())
            cached.append(box.pattern.sell_structure.amg_cache)
            assert isinstance(cached[-1], AMGPreconditioner)
        assert cached[0] is cached[1]

    def test_planar_two_dofs_per_node_and_gradient(self, T):
        from torchfem_b200.materials import IsotropicElasticityPlaneStress
        from torchfem_b200.mesh import rect_quad

        nodes, elements = rect_quad(81, 41, 2.0, 1.0)      # 6,642 DOFs, d = 2
        res = {}
        for method in ("cg", "amgx"):
            th = torch.full((len(elements),), 0.5, requires_grad=True)
            m = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(1000.0, 0.3), thickness=th)
            m.constraints[nodes[:, 0] == 0.0, :] = True
            m.forces[nodes[:, 0] == 2.0, 1] = -0.01
            u, *_ = m.solve(differentiable_parameters=th, method=method, stol=1e-12)
            (u ** 2).sum().backward()
            res[method] = (u.detach(), th.grad.clone())
        assert float((res["amgx"][0] - res["cg"][0]).abs().max()) <= 1e-8 * float(res["cg"][0].abs().max())
        assert float((res["amgx"][1] - res["cg"][1]).abs().max()) <= 1e-7 * float(res["cg"][1].abs().max())

    def test_newton_loop_refreshes_the_hierarchy(self, T):
        """Neo-Hookean block, nlgeom, two increments: every Newton iteration hands the solver object back
        (reference sparse.py:438-441 `resetup`) and the tangent changes — the numeric phase is repeated on the
        stored aggregates; result equal to the Jacobi-CG path."""
        from torchfem_b200.materials import Hyperelastic3D
        from torchfem_b200.mesh import cube_hexa

        En, NU = 1000.0, 0.3
        LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
        MU = En / (2.0 * (1.0 + NU))

        def psi(F, params):
            Cg = F.transpose(-1, -2) @ F
            logJ = 0.5 * torch.logdet(Cg)
            return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2

        nodes, elements = cube_hexa(17, 9, 9, 2.0, 1.0, 1.0)     # 4,131 DOFs
        out = {}
        for method in ("cg", "amgx"):
            box = T.Solid(nodes, elements, Hyperelastic3D(psi, torch.tensor([MU, LBD])))
            left, right = nodes[:, 0] == 0.0, nodes[:, 0] == 2.0
            box.constraints[left, :] = True
            box.constraints[right, 0] = True
            box.displacements[right, 0] = 0.4
            u, f, *_ = box.solve(increments=torch.tensor([0.0, 0.5, 1.0]), nlgeom=True, method=method, stol=1e-11)
            out[method] = u
        assert float((out["amgx"] - out["cg"]).abs().max()) <= 1e-7 * float(out["cg"].abs().max())


class TestModal:
    """`modal_eigsolve` / `differentiable_modal_eigsolve` (reference tests/test_sparse.py:208-306) and
    `solve_modes` against vectors generated from the unmodified reference (oracle/make_golden.py::modal,
    scipy eigsh shift-invert): eigenvalues <= 1e-8 relative, sensitivities <= 1e-6."""

    N_MODES = 3

    @staticmethod
    def _spd(n, seed):
        torch.manual_seed(seed)
        A = torch.randn(n, n, device="cpu")
        return (A @ A.T + n * torch.eye(n, device="cpu")).cuda()

    def _problem(self, n=8):
        K_dense, M_dense = self._spd(n, 0), self._spd(n, 1)
        return K_dense, M_dense, K_dense.to_sparse_coo().coalesce(), M_dense.to_sparse_coo().coalesce(), torch.arange(2, n)

    def test_matches_dense_generalized_eigenproblem(self, T):
        import scipy.linalg

        K_dense, M_dense, K, M, free = self._problem()
        values, vectors = T.sparse.modal_eigsolve(K, M, self.N_MODES, free)
        f = free.cpu().numpy()
        ref = scipy.linalg.eigh(K_dense.cpu().numpy()[f][:, f], M_dense.cpu().numpy()[f][:, f], eigvals_only=True)
        assert np.allclose(values.cpu().numpy(), ref[: self.N_MODES], rtol=1e-8)
        assert bool((values[1:] >= values[:-1]).all())
        assert vectors.shape == (8, self.N_MODES) and bool((vectors[:2] == 0).all())
        # M-normalised like eigsh
        assert torch.allclose((vectors.T @ M_dense @ vectors), torch.eye(self.N_MODES), atol=1e-10)

    @pytest.mark.parametrize("which", ["K", "M"])
    def test_eigenvalue_gradients_match_finite_differences(self, T, which):
        n, eps = 8, 1e-6
        K_dense, M_dense, K, M, free = self._problem(n)
        torch.manual_seed(7)
        D = torch.randn(n, n)
        D = 0.5 * (D + D.T)
        base = K_dense if which == "K" else M_dense
        idx = base.to_sparse_coo().coalesce().indices()
        values = base[idx[0], idx[1]].clone().requires_grad_(True)
        A = torch.sparse_coo_tensor(idx, values, (n, n)).coalesce()
        lambdas, phis = T.sparse.differentiable_modal_eigsolve(A if which == "K" else K, M if which == "K" else A,
                                                               self.N_MODES, free)
        assert lambdas.requires_grad and not phis.requires_grad
        (grad,) = torch.autograd.grad(lambdas.sum(), [values])

        def solve(pert):
            P = (base + pert).to_sparse_coo().coalesce()
            v, _ = T.sparse.modal_eigsolve(P if which == "K" else K, M if which == "K" else P, self.N_MODES, free)
            return v.sum()

        fd = (solve(eps * D) - solve(-eps * D)) / (2 * eps)
        assert torch.allclose(torch.dot(grad, D[idx[0], idx[1]]), fd, rtol=1e-5)

    @staticmethod
    def _aligned(a, b):
        return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)

    def test_solve_modes_solid_matches_reference(self, T):
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        g = load_case("modal.npz")
        nodes, elements = cube_hexa(6, 4, 4, 2.0, 1.0, 1.0)
        n_elem = len(elements)
        E_ = (1000.0 * (1.0 + 0.2 * torch.sin(torch.arange(n_elem, dtype=torch.float64)))).requires_grad_(True)
        rho = (2.0 + 0.5 * torch.cos(torch.arange(n_elem, dtype=torch.float64))).requires_grad_(True)
        box = T.Solid(nodes, elements, IsotropicElasticity3D(E=E_, nu=torch.full((n_elem,), 0.3), rho=rho))
        box.constraints[nodes[:, 0] == 0.0, :] = True
        omega_sq, modes = box.solve_modes(n_modes=6)
        assert modes.shape == (6, box.n_nod, 3) and not modes.requires_grad
        assert np.allclose(omega_sq.detach().cpu().numpy(), g["solid.omega_sq"], rtol=1e-8)
        for i in range(6):
            assert self._aligned(modes[i].cpu().numpy(), g["solid.modes"][i]) <= 1e-6
        omega_sq.sum().backward()
        assert np.allclose(E_.grad.cpu().numpy(), g["solid.grad_E"], rtol=1e-6, atol=1e-9 * np.abs(g["solid.grad_E"]).max())
        assert np.allclose(rho.grad.cpu().numpy(), g["solid.grad_rho"], rtol=1e-6, atol=1e-9 * np.abs(g["solid.grad_rho"]).max())

    def test_solve_modes_planar_matches_reference(self, T):
        from torchfem_b200.materials import IsotropicElasticityPlaneStress
        from torchfem_b200.mesh import rect_quad

        g = load_case("modal.npz")
        nodes, elements = rect_quad(9, 4, 2.0, 0.5)
        n_elem = len(elements)
        E_ = (500.0 * (1.0 + 0.1 * torch.cos(torch.arange(n_elem, dtype=torch.float64)))).requires_grad_(True)
        strip = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(E=E_, nu=torch.full((n_elem,), 0.25),
                                                                          rho=torch.full((n_elem,), 3.0)),
                         thickness=torch.full((n_elem,), 0.1))
        strip.constraints[nodes[:, 0] == 0.0, :] = True
        omega_sq, _ = strip.solve_modes(n_modes=5)
        assert np.allclose(omega_sq.detach().cpu().numpy(), g["planar.omega_sq"], rtol=1e-8)
        omega_sq[0].backward()
        assert np.allclose(E_.grad.cpu().numpy(), g["planar.grad_E"], rtol=1e-6, atol=1e-9 * np.abs(g["planar.grad_E"]).max())

    def test_lobpcg_with_amg_matches_the_dense_path(self, T):
        """The large-system route (AMG-preconditioned LOBPCG on the device CSR) forced on a model small enough for the
        dense route: same eigenvalues (<= 1e-8), same invariant subspace."""
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(13, 7, 7, 2.0, 1.0, 1.0)           # 1,911 DOFs
        box = T.Solid(nodes, elements, IsotropicElasticity3D(E=1000.0, nu=0.3, rho=2.0))
        box.constraints[nodes[:, 0] == 0.0, :] = True
        con = torch.nonzero(box.constraints.ravel()).ravel()
        free = torch.nonzero(~box.constraints.ravel()).ravel()
        K = box.assemble_matrix(box.k0(), con)
        M = box.assemble_matrix(box.integrate_mass(), con)
        vd, Xd = T.sparse.modal_eigsolve(K, M, 6, free, method="dense")
        vl, Xl = T.sparse.modal_eigsolve(K, M, 6, free, method="lobpcg")
        assert torch.allclose(vl, vd, rtol=1e-8)
        assert bool((Xl[con] == 0).all())
        # the bending pair is nearly degenerate: compare the subspaces through M-inner products
        G = Xd.T @ torch.stack([M.matvec(Xl[:, j].contiguous()) for j in range(6)], dim=1)
        assert torch.allclose(G @ G.T, torch.eye(6), atol=1e-6)

    def test_lobpcg_at_scale(self, T):
        """Beyond the dense limit (46,875 DOFs): residuals of the returned pairs and agreement with the analytic
        scaling omega^2 ~ E / rho."""
        from torchfem_b200.materials import IsotropicElasticity3D
        from torchfem_b200.mesh import cube_hexa

        nodes, elements = cube_hexa(41, 21, 21, 2.0, 1.0, 1.0)
        vals = []
        for E_ in (1000.0, 4000.0):
            box = T.Solid(nodes, elements, IsotropicElasticity3D(E=E_, nu=0.3, rho=2.0))
            box.constraints[nodes[:, 0] == 0.0, :] = True
            omega_sq, modes = box.solve_modes(n_modes=4)
            vals.append(omega_sq)
            con = torch.nonzero(box.constraints.ravel()).ravel()
            K = box.assemble_matrix(box.k0(), con)
            Mm = box.assemble_matrix(box.integrate_mass(), con)
            for i in range(4):
                x = modes[i].reshape(-1).contiguous()
                r = K.matvec(x) - omega_sq[i] * Mm.matvec(x)
                r[con] = 0.0
                assert float(r.norm()) <= 1e-6 * float(K.matvec(x).norm())
        assert torch.allclose(vals[1], 4.0 * vals[0], rtol=1e-7)


def test_block_with_rotated_orthotropic_material(T):
    """an orthotropic tangent rotated element by element goes through K1 / K17 like any per-element
    `ddsdde`; displacements and stresses of the reference (`oracle/make_golden.py::orthotropic`)."""
    from torchfem_b200.materials import OrthotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    g = load_case("orthotropic.npz")
    nodes, elements = cube_hexa(4, 3, 3, 1.5, 1.0, 1.0)
    material = OrthotropicElasticity3D(E_1=150.0, E_2=12.0, E_3=9.0, nu_12=0.3, nu_13=0.25, nu_23=0.4, G_12=5.0,
                                       G_13=4.0, G_23=3.0).vectorize(len(elements)).rotate(torch.tensor(g["solid.R"]))
    model = T.Solid(nodes, elements, material)
    model.constraints[nodes[:, 0] == 0.0, :] = True
    model.forces[nodes[:, 0] == 1.5, 2] = -0.1
    for method, tol in (("spsolve", 1e-8), ("cg", 1e-6)):
        u, f, sigma, eps, _ = model.solve(method=method)
        assert np.abs(u.cpu().numpy() - g["solid.u"]).max() <= tol * np.abs(g["solid.u"]).max()
        assert np.abs(sigma.cpu().numpy() - g["solid.sigma"]).max() <= 10 * tol * np.abs(g["solid.sigma"]).max()


@pytest.mark.parametrize("tag", ["hexa2", "tetra1", "tria2", "quad1"])
def test_consistent_nodal_loads(T, tag):
    """`integrate_body_load` / `integrate_surface_load` / `integrate_line_load` (reference base.py:446-568) on the
    device against the reference's vectors (`oracle/make_golden.py::loads`); tests/test_loads_host_cpu.py covers all
    element types on the CPU."""
    from torchfem_b200.materials import IsotropicElasticity3D, IsotropicElasticityPlaneStress

    g = load_case("loads.npz")
    nodes, elements = torch.tensor(g[f"{tag}.nodes"]), torch.tensor(g[f"{tag}.elements"])

    def close(got, ref):
        assert np.abs(got.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()

    if nodes.shape[1] == 3:
        model = T.Solid(nodes, elements, IsotropicElasticity3D(1000.0, 0.3))
        top = torch.tensor(g[f"{tag}.top"])
        assert np.array_equal(model._boundary_facets(top).cpu().numpy(), g[f"{tag}.facets_top"])
        close(model.integrate_body_load(torch.tensor([0.0, 0.0, -9.81])), g[f"{tag}.body"])
        close(model.integrate_surface_load(top, -2.5), g[f"{tag}.pressure_top"])
        close(model.integrate_surface_load(top, torch.tensor([1.0, 0.5, -0.25])), g[f"{tag}.traction_top"])
        model.forces = model.integrate_surface_load(top, -2.5)          # a load vector is what `forces` takes
        model.constraints[nodes[:, 2] == 0.0, :] = True
        u = model.solve()[0]
        assert float(u[top][:, 2].mean()) < 0.0
    else:
        model = T.Planar(nodes, elements, IsotropicElasticityPlaneStress(1000.0, 0.3),
                         thickness=torch.tensor(g[f"{tag}.thickness"]))
        right = torch.tensor(g[f"{tag}.right"])
        close(model.integrate_body_load(torch.tensor([0.0, -9.81])), g[f"{tag}.body"])
        close(model.integrate_line_load(right, 3.0), g[f"{tag}.pressure_right"])
        close(model.integrate_line_load(right, torch.tensor([1.0, -2.0])), g[f"{tag}.traction_right"])


def test_hyperelastic_plane_stress_strip(T):
    """`HyperelasticPlaneStress` (reference hyperelasticity.py:130-269): Neo-Hookean strip stretched
    by 30 % in three `nlgeom` increments, thickness stretch carried as a state variable; the device path updates all
    Gauss points in one batch (`step_points`), the reference one Gauss point at a time — the local Newton iteration
    stops at |P_33| < 1e-5, so the two agree to that tolerance's effect, not to round-off."""
    from torchfem_b200.materials import HyperelasticPlaneStress
    from torchfem_b200.mesh import rect_quad

    def psi(F, params):
        Cg = F.transpose(-1, -2) @ F
        logJ = 0.5 * torch.logdet(Cg)
        return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2

    g = load_case("hyper_plane_stress.npz")
    nodes, elements = rect_quad(5, 3, 2.0, 1.0)
    strip = T.Planar(nodes, elements, HyperelasticPlaneStress(psi, torch.tensor([384.6153846153846, 576.9230769230769])))
    left, right = nodes[:, 0] == 0.0, nodes[:, 0] == 2.0
    strip.constraints[left, 0] = True
    strip.constraints[right, 0] = True
    strip.constraints[nodes[:, 1] == 0.5, 1] = True
    strip.displacements[right, 0] = 0.6
    u, f, sigma, F, alpha = strip.solve(increments=torch.linspace(0.0, 1.0, 4), nlgeom=True, method="spsolve")
    assert np.abs(u.cpu().numpy() - g["strip.u"]).max() <= 1e-6 * np.abs(g["strip.u"]).max()
    assert np.abs(alpha.cpu().numpy() - g["strip.state"]).max() <= 1e-5 * np.abs(g["strip.state"]).max()
    assert abs(float(f[right, 0].sum()) - float(g["strip.f"][np.isclose(nodes.cpu().numpy()[:, 0], 2.0), 0].sum())) <= 1e-5 * 250.0
