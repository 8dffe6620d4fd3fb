"""TEST INFRASTRUCTURE ONLY — generates `tests/golden/*.npz` by running the UNMODIFIED reference
(`/root/reference/src/torchfem`, imported through `oracle/ref_import.py`) in the build container.

    python oracle/make_golden.py            # rewrites tests/golden/

The fixtures are committed; the GPU box (which has no /root/reference) reads only them.
Everything is float64 / CPU, seeds are fixed, so re-running reproduces the files bit for bit
(torch 2.11.0, numpy 2.3, scipy 1.18.1).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

tf = ref_import.load()
import torch  # noqa: E402
from torchfem import Planar, PlanarHeat, Solid, SolidHeat  # noqa: E402
from torchfem import elements as E  # noqa: E402
from torchfem import materials as M  # noqa: E402
from torchfem import mesh  # noqa: E402
from torchfem.sparse import sparse_solve  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


# ---------------------------------------------------------------------------------- element tables
def element_tables():
    out = {}
    g = torch.Generator().manual_seed(1234)
    for name in ["Tria1", "Tria2", "Quad1", "Quad2", "Tetra1", "Tetra2", "Hexa1", "Hexa2"]:
        c = getattr(E, name)
        ip = c.ipoints
        out[f"{name}.iso_coords"] = npy(c.iso_coords)
        out[f"{name}.ipoints"] = npy(ip)
        out[f"{name}.iweights"] = npy(c.iweights)
        out[f"{name}.B_ip"] = npy(c.B(ip))
        out[f"{name}.N_ip"] = npy(c.N(ip))
        out[f"{name}.edges"] = npy(c.edges)
        out[f"{name}.facets"] = npy(c.facets)
        xi = torch.rand(5, c.iso_dim, generator=g) * 0.5
        out[f"{name}.xi"] = npy(xi)
        out[f"{name}.N_xi"] = npy(c.N(xi))
        out[f"{name}.B_xi"] = npy(c.B(xi))
        out[f"{name}.meta"] = np.array([c.nodes, c.iso_dim], dtype=np.int64)
        out[f"{name}.iso_volume"] = np.array(c.iso_volume)
    np.savez_compressed(os.path.join(OUT, "element_tables.npz"), **out)


# ---------------------------------------------------------------------------------- small meshes
def distort(nodes, amp, seed):
    g = torch.Generator().manual_seed(seed)
    return nodes + amp * (torch.rand(nodes.shape, generator=g) - 0.5)


def random_C(n_elem, d, seed):
    """Random tangent with the minor/major structure of nothing in particular — the kernel must be
    exact for ANY ddsdde, so no symmetry is assumed."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n_elem, d, d, d, d, generator=g) + 0.1


def small_cases():
    cases = {}
    n3, e3 = mesh.cube_hexa(4, 3, 3, 1.5, 1.0, 1.0)
    n3 = distort(n3, 0.08, 1)
    nt, et = mesh.cube_tetra(3, 3, 3)
    nt = distort(nt, 0.05, 2)
    n2, e2 = mesh.rect_quad(5, 4, 2.0, 1.0)
    n2 = distort(n2, 0.06, 3)
    ntr, etr = mesh.rect_tri(4, 4)
    ntr = distort(ntr, 0.05, 4)

    def mech(tag, nodes, elements, planar):
        n_elem = len(elements)
        d = nodes.shape[1]
        if planar:
            mat = M.IsotropicElasticityPlaneStress(E=1000.0, nu=0.3).vectorize(n_elem)
        else:
            mat = M.IsotropicElasticity3D(E=1000.0, nu=0.3).vectorize(n_elem)
        mat.C = random_C(n_elem, d, 7) * 100.0
        if planar:
            g = torch.Generator().manual_seed(9)
            th = torch.rand(n_elem, generator=g) + 0.5
            model = Planar(nodes, elements, mat, thickness=th)
        else:
            th = None
            model = Solid(nodes, elements, mat)
        k = model.k0()
        con_mask = torch.zeros(model.n_nod, d, dtype=torch.bool)
        con_mask[nodes[:, 0] < 0.05, :] = True
        con_mask[-1, 0] = True
        con = torch.nonzero(con_mask.ravel()).ravel()
        K = model.assemble_matrix(k, con)
        cases[tag] = dict(
            nodes=npy(nodes), elements=npy(elements), C=npy(mat.C), k=npy(k),
            idx=npy(model.idx), glob_idx=npy(model.glob_idx), k_map=npy(model.k_map),
            diag_map=npy(model.diag_map), con=npy(con), K_val=npy(K._values()),
            etype=np.array(model.etype.__name__),
        )
        if th is not None:
            cases[tag]["thickness"] = npy(th)

    def heat(tag, nodes, elements, planar):
        n_elem = len(elements)
        d = nodes.shape[1]
        g = torch.Generator().manual_seed(11)
        A = torch.rand(n_elem, d, d, generator=g)
        kappa = A @ A.transpose(-1, -2) + torch.eye(d)
        if planar:
            mat = M.IsotropicConductivity2D(kappa=1.0).vectorize(n_elem)
            mat.KAPPA = kappa
            th = torch.rand(n_elem, generator=g) + 0.5
            model = PlanarHeat(nodes, elements, mat, thickness=th)
        else:
            mat = M.IsotropicConductivity3D(kappa=1.0).vectorize(n_elem)
            mat.KAPPA = kappa
            th = None
            model = SolidHeat(nodes, elements, mat)
        k = model.k0()
        con_mask = torch.zeros(model.n_nod, 1, dtype=torch.bool)
        con_mask[nodes[:, 0] < 0.05, :] = True
        con = torch.nonzero(con_mask.ravel()).ravel()
        K = model.assemble_matrix(k, con)
        cases[tag] = dict(
            nodes=npy(nodes), elements=npy(elements), kappa=npy(kappa), k=npy(k),
            idx=npy(model.idx), glob_idx=npy(model.glob_idx), k_map=npy(model.k_map),
            diag_map=npy(model.diag_map), con=npy(con), K_val=npy(K._values()),
            etype=np.array(model.etype.__name__),
        )
        if th is not None:
            cases[tag]["thickness"] = npy(th)

    mech("hexa1", n3, e3, False)
    mech("hexa2", *E.linear_to_quadratic(n3, e3), False)
    mech("tetra1", nt, et, False)
    mech("tetra2", *E.linear_to_quadratic(nt, et), False)
    mech("quad1", n2, e2, True)
    mech("quad2", *E.linear_to_quadratic(n2, e2), True)
    mech("tria1", ntr, etr, True)
    mech("tria2", *E.linear_to_quadratic(ntr, etr), True)
    heat("heat_hexa1", n3, e3, False)
    heat("heat_tetra2", *E.linear_to_quadratic(nt, et), False)
    heat("heat_quad1", n2, e2, True)
    heat("heat_quad2", *E.linear_to_quadratic(n2, e2), True)

    # orphan node (not referenced by any element) + element order shuffled: pins the lone-diagonal
    # rule of base.py:89-91 and that the pattern does not depend on element order
    n_or = torch.cat([n3, torch.tensor([[5.0, 5.0, 5.0]])])
    mech("hexa1_orphan", n_or, e3.flip(0), False)

    # per-Gauss-point tangent (hyperelastic, nlgeom): capture ddsdde of Material.step
    nh, eh = mesh.cube_hexa(3, 3, 3)
    nh = distort(nh, 0.05, 5)

    def psi(F, params):
        Cg = F.transpose(-1, -2) @ F
        logJ = 0.5 * torch.logdet(Cg)
        return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ**2

    params = torch.tensor([384.6153846153846, 576.9230769230769])
    box = Solid(nh, eh, M.Hyperelastic3D(psi, params))
    captured = []
    orig_step = box.material.step

    def spy(*a, **k):
        r = orig_step(*a, **k)
        captured.append(r[2].detach().clone())
        return r

    box.material.step = spy
    g = torch.Generator().manual_seed(6)
    du = 0.05 * (torch.rand(box.n_nod, 3, generator=g) - 0.5)
    grad = torch.zeros(box.n_int, box.n_elem, 3, 3)
    grad[:] = torch.eye(3)
    flux = torch.zeros(box.n_int, box.n_elem, 3, 3)
    state = torch.zeros(box.n_int, box.n_elem, 0)
    box.K = torch.empty(0)
    k, f, *_ = box.integrate_material(
        torch.zeros(box.n_nod, 3), grad, flux, state, du, torch.zeros(box.n_elem, 3, 3), 0, True)
    cases["hyper_hexa1"] = dict(
        nodes=npy(nh), elements=npy(eh), C=npy(torch.stack(captured)), k=npy(k), f=npy(f),
        du=npy(du), etype=np.array("Hexa1"))

    for tag, d in cases.items():
        np.savez_compressed(os.path.join(OUT, f"case_{tag}.npz"), **d)


# ---------------------------------------------------------------------------------- config A
def config_a():
    """BASELINE config[0]: benchmarks/cubes.py with N=11 (SURVEY §8c golden vectors)."""
    nodes, elements = mesh.cube_hexa(11, 11, 11)
    cube = Solid(nodes, elements, M.IsotropicElasticity3D(E=1000.0, nu=0.3))
    cube.forces = torch.zeros_like(nodes, requires_grad=True)
    cube.constraints[nodes[:, 0] == 0.0, :] = True
    cube.constraints[nodes[:, 0] == 1.0, 0] = True
    cube.displacements[nodes[:, 0] == 1.0, 0] = 0.1
    k = cube.k0()
    con = torch.nonzero(cube.constraints.ravel()).ravel()
    K = cube.assemble_matrix(k, con)
    val = K._values()
    u, f, sigma, eps, _ = cube.solve(differentiable_parameters=cube.forces, method="spsolve")
    u.sum().backward()
    gF = cube.forces.grad.clone()
    # Jacobi-CG / MINRES through the reference's own sparse_solve CPU code path (Jacobi stand-in)
    u_cg, *_ = cube.solve(method="cg", stol=1e-10)
    u_mr, *_ = cube.solve(method="minres", stol=1e-10)
    # linear system of the first Newton step, for solver-level parity
    du0 = torch.zeros(cube.n_dofs)
    du0[con] = cube.displacements.ravel()[con]
    g = torch.Generator().manual_seed(0)
    probe = torch.randn(val.shape[0], generator=g)
    kprobe = torch.randn(k.numel(), generator=g)
    out = dict(
        sha_idx=np.array(sha(npy(cube.idx))), sha_glob_idx=np.array(sha(npy(cube.glob_idx))),
        sha_k_map=np.array(sha(npy(cube.k_map))), sha_diag_map=np.array(sha(npy(cube.diag_map))),
        nnz=np.array(val.shape[0]), k_fro=npy(torch.linalg.norm(k)), k_000=npy(k[0, 0, :]),
        k_probe=npy((k.ravel() * kprobe).sum()), k_absmax=npy(k.abs().max()),
        k_e0=npy(k[0]), k_e777=npy(k[777]),
        val_norm=npy(torch.linalg.norm(val)), val_sum=npy(val.sum()),
        val_probe=npy((val * probe).sum()), val_absmax=npy(val.abs().max()),
        val_n_zero=np.array(int((val == 0).sum())), val_n_one=np.array(int((val == 1).sum())),
        val_head=npy(val[:4096]), con=npy(con),
        u=npy(u), f=npy(f), sigma=npy(sigma), eps=npy(eps), grad_forces=npy(gF),
        u_cg=npy(u_cg), u_minres=npy(u_mr),
    )
    np.savez_compressed(os.path.join(OUT, "config_a.npz"), **out)
    print("config A: nnz", val.shape[0], "sha glob_idx", out["sha_glob_idx"], "k_map", out["sha_k_map"])


# ---------------------------------------------------------------------------------- gradients
def topopt_small():
    """benchmarks/topopt.py at N=3 (6x3x3 elements): compliance and its density gradient."""
    N = 3
    nx, ny, nz = 2 * N, N, N
    nodes, elements = mesh.cube_hexa(nx + 1, ny + 1, nz + 1, 2.0, 1.0, 1.0)
    rng = np.random.default_rng(0)
    values = np.clip(0.5 + 0.3 * rng.standard_normal(len(elements)), 0.05, 0.95)
    rho = torch.tensor(values, requires_grad=True)
    material = M.IsotropicElasticity3D(E=70000.0, nu=0.3).vectorize(len(elements))
    scale = 1e-3 + (1.0 - 1e-3) * rho**3.0
    material.C = scale[:, None, None, None, None] * material.C
    model = Solid(nodes, elements, material)
    model.constraints[nodes[:, 0] == 0.0, :] = True
    right = nodes[:, 0] == 2.0
    wy = torch.full((model.n_nod,), 1.0 / ny)
    wy[(nodes[:, 1] == 0.0) | (nodes[:, 1] == 1.0)] /= 2.0
    wz = torch.full((model.n_nod,), 1.0 / nz)
    wz[(nodes[:, 2] == 0.0) | (nodes[:, 2] == 1.0)] /= 2.0
    model.forces[right, 2] = -1.0 * wy[right] * wz[right]
    u, *_ = model.solve(differentiable_parameters=rho, method="spsolve")
    c = torch.inner(model.forces.ravel(), u.ravel())
    c.backward()
    np.savez_compressed(os.path.join(OUT, "topopt_n3.npz"), rho=values, u=npy(u),
                        compliance=npy(c), grad_rho=npy(rho.grad), forces=npy(model.forces))


def hyper_small():
    """benchmarks/hyperelasticity.py at N=3 with 3 increments: reaction force and d/d(mu,lambda)."""
    import math

    En, NU = 1000.0, 0.3
    LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
    MU = En / (2.0 * (1.0 + NU))

    def psi(F, params):
        Cg = F.transpose(-1, -2) @ F
        logJ = 0.5 * torch.logdet(Cg)
        return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ**2

    N = 3
    lx = 4.0 / (N - 1)
    nodes, elements = mesh.cube_hexa(5, N, N, lx, 1.0, 1.0)
    params = torch.tensor([MU, LBD], requires_grad=True)
    box = Solid(nodes, elements, M.Hyperelastic3D(psi, params))
    left = nodes[:, 0] == 0.0
    right = nodes[:, 0] == lx
    box.constraints[left, 0] = True
    box.constraints[right, 0] = True
    box.constraints[nodes[:, 1] == 0.5, 1] = True
    box.constraints[nodes[:, 2] == 0.5, 2] = True
    STRETCH = 2.0
    box.displacements[right, 0] = (STRETCH - 1.0) * lx
    lam = torch.logspace(0, math.log10(STRETCH), 4)
    increments = (lam - 1.0) / (STRETCH - 1.0)
    u, f, *_ = box.solve(increments=increments, nlgeom=True, differentiable_parameters=params,
                         method="spsolve")
    reaction = f[right, 0].sum()
    reaction.backward()
    np.savez_compressed(os.path.join(OUT, "hyper_n3.npz"), u=npy(u), f=npy(f),
                        reaction=npy(reaction), grad_params=npy(params.grad),
                        increments=npy(increments))


def heat_transient():
    """`Heat.time_integration` (base.py:1288-1553; tests/test_time_integration.py:9-18 heated plate) and a small
    SolidHeat cube: temperatures / fluxes at the output times, and the gradient of the final temperatures w.r.t.
    the nodal heat flux (flows through `differentiable_sparse_solve`)."""
    from torchfem import PlanarHeat, SolidHeat

    out = {}
    model = PlanarHeat(*mesh.rect_quad(5, 5, 1.0, 1.0), M.IsotropicConductivity2D(kappa=400.0, rho=1.0e5))
    west = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].min())
    east = torch.isclose(model.nodes[:, 0], model.nodes[:, 0].max())
    model.constraints[west | east] = True
    model.temperatures[west, 0] = 5.0
    model.temperatures[east, 0] = 20.0
    hf = torch.zeros(model.n_nod, 1)
    hf[12, 0] = 3.0
    model.heat_flux = hf.clone().requires_grad_(True)
    t_out = torch.tensor([0.0, 6.0, 12.0, 18.0, 24.0])
    temp, rfl, flux, grad, state = model.time_integration(t_out, delta_t=1.0)
    temp[-1].sum().backward()
    out.update({"plate.t_out": npy(t_out), "plate.temp": npy(temp), "plate.rfl": npy(rfl), "plate.flux": npy(flux),
                "plate.grad": npy(grad), "plate.heat_flux": npy(hf), "plate.grad_heat_flux": npy(model.heat_flux.grad)})
    late, *_ = model.time_integration(torch.tensor([5.0, 10.0]), delta_t=0.5)
    out["plate.late"] = npy(late)

    nodes, elements = mesh.cube_hexa(4, 4, 4)
    cube = SolidHeat(nodes, elements, M.IsotropicConductivity3D(kappa=2.0, rho=30.0))
    cube.constraints[nodes[:, 0] == 0.0] = True
    cube.temperatures[nodes[:, 0] == 0.0, 0] = 1.0
    cube.heat_flux[nodes[:, 0] == 1.0, 0] = 0.05
    t3 = torch.tensor([0.0, 1.0, 3.0])
    temp3, _, flux3, grad3, _ = cube.time_integration(t3, delta_t=0.25)
    out.update({"cube.t_out": npy(t3), "cube.temp": npy(temp3), "cube.flux": npy(flux3), "cube.grad": npy(grad3)})

    # sensitivities w.r.t. per-element MATERIAL parameters: they flow through the system matrix M + dt/2 K of every step
    # (Solve.backward's sparse gradA, sparse.py:212-216, back through assemble_matrix) and through M @ du (base.py:1483)
    n_el = model.n_elem
    gen = torch.Generator().manual_seed(5)
    kappa = (300.0 + 200.0 * torch.rand(n_el, generator=gen)).requires_grad_(True)
    rho = (0.5e5 + 1.0e5 * torch.rand(n_el, generator=gen)).requires_grad_(True)
    het = PlanarHeat(*mesh.rect_quad(5, 5, 1.0, 1.0), M.IsotropicConductivity2D(kappa=kappa, rho=rho))
    het.constraints[west | east] = True
    het.temperatures[west, 0] = 5.0
    het.temperatures[east, 0] = 20.0
    het.heat_flux = hf.clone()
    t_het = torch.tensor([0.0, 4.0, 8.0])
    temp_h, *_ = het.time_integration(t_het, delta_t=1.0, differentiable_parameters=[kappa, rho])
    (temp_h[-1] ** 2).sum().backward()
    out.update({"het.kappa": npy(kappa), "het.rho": npy(rho), "het.t_out": npy(t_het), "het.temp": npy(temp_h),
                "het.grad_kappa": npy(kappa.grad), "het.grad_rho": npy(rho.grad)})
    np.savez_compressed(os.path.join(OUT, "heat_transient.npz"), **out)


def sparse_small():
    """tests/test_sparse.py-style systems: SPD and non-symmetric 6x6 COO with dense solutions."""
    g = torch.Generator().manual_seed(42)
    Ad = torch.rand(6, 6, generator=g)
    spd = Ad @ Ad.T + 6 * torch.eye(6)
    nonsym = torch.rand(6, 6, generator=g) + 6 * torch.eye(6)
    b = torch.rand(6, generator=g)
    out = {}
    for tag, Am in [("spd", spd), ("nonsym", nonsym)]:
        A = Am.to_sparse_coo().detach().requires_grad_(True)
        bb = b.clone().requires_grad_(True)
        x = tf.sparse.differentiable_sparse_solve(A, bb, method="spsolve")
        x.sum().backward()
        out[f"{tag}.A"] = npy(Am)
        out[f"{tag}.x"] = npy(x)
        out[f"{tag}.gb"] = npy(bb.grad)
        out[f"{tag}.gA_idx"] = npy(A.grad.coalesce().indices())
        out[f"{tag}.gA_val"] = npy(A.grad.coalesce().values())
    out["b"] = npy(b)
    np.savez_compressed(os.path.join(OUT, "sparse_small.npz"), **out)


def modal():
    """`solve_modes` (reference base.py:1097-1129; scipy eigsh shift-invert) for a clamped solid block and a planar
    strip with per-element E and rho: omega^2, and d(sum omega^2)/dE, /drho through the Rayleigh-quotient adjoint."""
    out = {}
    nodes, elements = mesh.cube_hexa(6, 4, 4, 2.0, 1.0, 1.0)
    n_elem = len(elements)
    E_ = (1000.0 * (1.0 + 0.2 * torch.sin(torch.arange(n_elem, dtype=torch.float64)))).requires_grad_(True)
    rho = (2.0 + 0.5 * torch.cos(torch.arange(n_elem, dtype=torch.float64))).requires_grad_(True)
    box = Solid(nodes, elements, M.IsotropicElasticity3D(E=E_, nu=torch.full((n_elem,), 0.3), rho=rho))
    box.constraints[nodes[:, 0] == 0.0, :] = True
    omega_sq, modes = box.solve_modes(n_modes=6)
    omega_sq.sum().backward()
    out.update({"solid.omega_sq": npy(omega_sq), "solid.modes": npy(modes), "solid.grad_E": npy(E_.grad),
                "solid.grad_rho": npy(rho.grad)})
    nodes, elements = mesh.rect_quad(9, 4, 2.0, 0.5)
    n_elem = len(elements)
    E_ = (500.0 * (1.0 + 0.1 * torch.cos(torch.arange(n_elem, dtype=torch.float64)))).requires_grad_(True)
    strip = Planar(nodes, elements, M.IsotropicElasticityPlaneStress(E=E_, nu=torch.full((n_elem,), 0.25),
                                                                      rho=torch.full((n_elem,), 3.0)),
                   thickness=torch.full((n_elem,), 0.1))
    strip.constraints[nodes[:, 0] == 0.0, :] = True
    omega_sq, modes = strip.solve_modes(n_modes=5)
    omega_sq[0].backward()
    out.update({"planar.omega_sq": npy(omega_sq), "planar.grad_E": npy(E_.grad)})
    np.savez_compressed(os.path.join(OUT, "modal.npz"), **out)
    print("modal: solid omega^2", out["solid.omega_sq"], "planar", out["planar.omega_sq"])


def assembly_cases():
    """`Assembly` (reference assembly.py:113-608; the set-ups follow reference tests/test_assembly.py:54-77, 128-184,
    255-272, 351-392, 415-455): per-part displacement / force, the retained-to-all map T of `_build_T`, the rigid
    modes, and an adjoint gradient through the constrained solve."""
    from torchfem import Assembly, ReferencePoint, ReferencePointHeat
    from torchfem.elements import linear_to_quadratic

    solid_mat = M.IsotropicElasticity3D(1000.0, 0.3)
    plane_mat = M.IsotropicElasticityPlaneStress(1000.0, 0.3)
    out = {}

    def store(tag, res, asm=None):
        u, f, flux, grad, _ = res
        for j in range(len(u)):
            out[f"{tag}.u{j}"], out[f"{tag}.f{j}"] = npy(u[j]), npy(f[j])
            out[f"{tag}.flux{j}"], out[f"{tag}.grad{j}"] = npy(flux[j]), npy(grad[j])
        if asm is not None:
            T, retained = asm._build_T()
            out[f"{tag}.T_idx"], out[f"{tag}.T_val"] = npy(T._indices()), npy(T._values())
            out[f"{tag}.retained"], out[f"{tag}.modes"] = npy(retained), npy(asm._rigid_modes())

    # two distorted 3x3x2-element blocks stacked in z and tied at z = 1
    n_a, e_a = mesh.cube_hexa(4, 4, 3, 1.0, 1.0, 1.0)
    n_b, e_b = mesh.cube_hexa(4, 4, 4, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    a, b = Solid(n_a, e_a, solid_mat), Solid(n_b, e_b, solid_mat)
    a.constraints[n_a[:, 2] == 0.0] = True
    b.forces[n_b[:, 2] == 2.0, 2] = 25.0 / 16
    b.forces[n_b[:, 2] == 2.0, 0] = 5.0 / 16
    asm = Assembly([a, b])
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    store("tie", asm.solve(), asm)
    inc = torch.linspace(0.0, 1.0, 4)
    every = asm.solve(increments=inc, return_intermediate=True)
    out["tie.every_u1"], out["tie.every_f0"] = npy(every[0][1]), npy(every[1][0])

    # clamped cube driven by a reference point: force and moment
    nodes, elements = mesh.cube_hexa(4, 4, 4)
    solid = Solid(nodes, elements, solid_mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = ReferencePoint([0.5, 0.5, 2.0])
    point.forces[0, 3] = 50.0
    point.forces[0, 5] = -20.0
    point.forces[0, 0] = 10.0
    asm = Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point)
    store("point", asm.solve(), asm)

    # only u_z coupled, point fully prescribed
    solid = Solid(nodes, elements, solid_mat)
    solid.constraints[nodes[:, 2] == 0.0] = True
    point = ReferencePoint([0.5, 0.5, 2.0])
    point.constraints[0, :] = True
    point.displacements[0, 2] = 0.1
    asm = Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point, dofs=[2])
    store("subset", asm.solve(), asm)

    # quadratic solid (Hexa2) tied to a linear one through the nearest-node pairing
    n_q, e_q = linear_to_quadratic(*mesh.cube_hexa(3, 3, 3))
    n_l, e_l = mesh.cube_hexa(3, 3, 3)
    n_l = n_l + torch.tensor([0.0, 0.0, 1.0])
    q, l = Solid(n_q, e_q, solid_mat), Solid(n_l, e_l, solid_mat)
    q.constraints[n_q[:, 2] == 0.0] = True
    l.forces[n_l[:, 2] == 2.0, 1] = 1.0
    asm = Assembly([q, l])
    asm.coupling(l, n_l[:, 2] == 1.0, q, n_q[:, 2] == 1.0)
    store("mixed", asm.solve(), asm)

    # heat: two bars tied, the top held isothermal by a thermal point with a heat source
    cond = M.IsotropicConductivity3D(1.5)
    n_a, e_a = mesh.cube_hexa(3, 3, 3)
    n_b, e_b = mesh.cube_hexa(3, 3, 4)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    ha, hb = SolidHeat(n_a, e_a, cond), SolidHeat(n_b, e_b, cond)
    ha.constraints[n_a[:, 2] == 0.0] = True
    hp = ReferencePointHeat([0.5, 0.5, 2.5])
    hp.heat_flux[0, 0] = 4.0
    asm = Assembly([ha, hb, hp])
    asm.coupling(hb, n_b[:, 2] == 1.0, ha, n_a[:, 2] == 1.0)
    asm.coupling(hb, n_b[:, 2] == 2.0, hp)
    store("heat", asm.solve(), asm)

    # planar: two quad meshes tied, and a planar point carrying a moment about z
    n_a, e_a = mesh.rect_quad(4, 4, 1.0, 1.0)
    n_b, e_b = mesh.rect_quad(4, 4, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0])
    pa, pb = Planar(n_a, e_a, plane_mat), Planar(n_b, e_b, plane_mat)
    pa.constraints[n_a[:, 0] == 0.0] = True
    pp = ReferencePoint([2.5, 0.5])
    pp.forces[0, 2] = 20.0
    pp.forces[0, 1] = -1.0
    asm = Assembly([pa, pb, pp])
    asm.coupling(pb, n_b[:, 0] == 1.0, pa, n_a[:, 0] == 1.0)
    asm.coupling(pb, n_b[:, 0] == 2.0, pp)
    store("planar", asm.solve(), asm)

    # adjoint: d(work)/d(rho) with a SIMP-scaled stiffness in part a (the reference differentiates a shell thickness,
    # tests/test_assembly.py:187-212; shells are outside the hot path)
    n_a, e_a = mesh.cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b, e_b = mesh.cube_hexa(4, 3, 3, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0, 0.0])
    rng = np.random.default_rng(3)
    values = np.clip(0.6 + 0.2 * rng.standard_normal(len(e_a)), 0.1, 1.0)
    rho = torch.tensor(values, requires_grad=True)
    mat = M.IsotropicElasticity3D(E=1000.0, nu=0.3).vectorize(len(e_a))
    mat.C = (rho ** 3.0)[:, None, None, None, None] * mat.C
    a, b = Solid(n_a, e_a, mat), Solid(n_b, e_b, solid_mat)
    a.constraints[n_a[:, 0] == 0.0] = True
    point = ReferencePoint([2.5, 0.5, 0.5])
    point.forces[0, 2] = -3.0
    point.forces[0, 3] = 1.0
    asm = Assembly([a, b, point])
    asm.coupling(b, n_b[:, 0] == 1.0, a, n_a[:, 0] == 1.0)
    asm.coupling(b, n_b[:, 0] == 2.0, point)
    u, *_ = asm.solve(differentiable_parameters=rho)
    work = torch.inner(point.forces.ravel(), u[2].ravel())
    work.backward()
    out["adjoint.rho"], out["adjoint.work"], out["adjoint.grad_rho"] = values, npy(work), npy(rho.grad)
    out["adjoint.u2"] = npy(u[2])
    np.savez_compressed(os.path.join(OUT, "assembly.npz"), **out)
    print("assembly: point u", out["point.u1"], "work", out["adjoint.work"])


def orthotropic():
    """Orthotropic / transversely isotropic elasticity and orthotropic conductivity (reference elasticity.py:324-793,
    conductivity.py:143-242): stiffness tensors, rotated tensors and re-extracted engineering constants, one `step`,
    and a small clamped block solved with a rotated per-element orthotropic material."""
    g = torch.Generator().manual_seed(7)
    out = {}

    def rot3(n):
        q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
        return q * torch.sign(torch.linalg.det(q))[:, None, None]

    def rot2(n):
        a = torch.rand(n, generator=g) * 3.0
        return torch.stack([torch.stack([a.cos(), -a.sin()], -1), torch.stack([a.sin(), a.cos()], -1)], -2)

    def consts(m, names):
        return np.stack([npy(getattr(m, k)) for k in names])

    p3 = dict(E_1=150.0, E_2=12.0, E_3=9.0, nu_12=0.3, nu_13=0.25, nu_23=0.4, G_12=5.0, G_13=4.0, G_23=3.0)
    m = M.OrthotropicElasticity3D(**p3)
    R = rot3(4)
    out["o3.C"], out["o3.R"] = npy(m.C), npy(R)
    mv = m.vectorize(4).rotate(R)
    out["o3.C_rot"] = npy(mv.C)
    out["o3.consts_rot"] = consts(mv, ["E_1", "E_2", "E_3", "nu_12", "nu_13", "nu_23", "G_12", "G_13", "G_23"])
    H = 1e-3 * torch.randn(4, 3, 3, generator=g)
    s0 = torch.randn(4, 3, 3, generator=g)
    de0 = 1e-4 * torch.randn(4, 3, 3, generator=g)
    sig, _, dd = mv.step(H, torch.eye(3).expand(4, 3, 3), s0, torch.zeros(4, 0), de0, torch.ones(4, 1), 0)
    out["o3.H"], out["o3.s0"], out["o3.de0"], out["o3.sig"] = npy(H), npy(s0), npy(de0), npy(sig)
    # batched parameters
    E1 = torch.tensor([150.0, 80.0, 40.0])
    mb = M.OrthotropicElasticity3D(E1, 0.1 * E1, 0.08 * E1, 0.3, 0.25, 0.4, 0.04 * E1, 0.03 * E1, 0.02 * E1)
    out["o3.E1_batch"], out["o3.C_batch"] = npy(E1), npy(mb.C)
    ti = M.TransverseIsotropicElasticity3D(E_L=140.0, E_T=10.0, nu_L=0.28, nu_T=0.42, G_L=5.5)
    out["ti.C"] = npy(ti.C)
    ps = M.OrthotropicElasticityPlaneStress(E_1=150.0, E_2=12.0, nu_12=0.3, G_12=5.0)
    R2 = rot2(3)
    psr = ps.vectorize(3).rotate(R2)
    out["ps.C"], out["ps.R"], out["ps.C_rot"] = npy(ps.C), npy(R2), npy(psr.C)
    out["ps.consts_rot"] = consts(psr, ["E_1", "E_2", "nu_12", "G_12"])
    pe = M.OrthotropicElasticityPlaneStrain(E_1=150.0, E_2=12.0, E_3=9.0, nu_12=0.3, nu_13=0.25, nu_23=0.4, G_12=5.0)
    per = pe.vectorize(3).rotate(R2)
    out["pe.C"], out["pe.C_rot"] = npy(pe.C), npy(per.C)
    out["pe.consts_rot"] = consts(per, ["E_1", "E_2", "nu_12", "G_12"])
    k3 = M.OrthotropicConductivity3D(10.0, 2.0, 0.5)
    out["k3.K"], out["k3.K_rot"] = npy(k3.KAPPA), npy(k3.vectorize(4).rotate(R).KAPPA)
    k2 = M.OrthotropicConductivity2D(torch.tensor([10.0, 4.0, 1.0]), torch.tensor([2.0, 1.0, 0.5]))
    out["k2.K"], out["k2.K_rot"] = npy(k2.KAPPA), npy(k2.rotate(R2).KAPPA)

    # model level: clamped 3x2x2-element block, fibre direction rotated element by element
    nodes, elements = mesh.cube_hexa(4, 3, 3, 1.5, 1.0, 1.0)
    Re = rot3(len(elements))
    model = Solid(nodes, elements, M.OrthotropicElasticity3D(**p3).vectorize(len(elements)).rotate(Re))
    model.constraints[nodes[:, 0] == 0.0, :] = True
    model.forces[nodes[:, 0] == 1.5, 2] = -0.1
    u, f, sigma, eps, _ = model.solve(method="spsolve")
    out["solid.R"], out["solid.u"], out["solid.sigma"] = npy(Re), npy(u), npy(sigma)
    np.savez_compressed(os.path.join(OUT, "orthotropic.npz"), **out)
    print("orthotropic: C_1111", out["o3.C"][0, 0, 0, 0], "u max", np.abs(out["solid.u"]).max())


def loads():
    """Consistent nodal loads (reference base.py:452-568; set-ups after reference tests/test_loads.py): body, surface
    (scalar pressure and traction vector) and line loads on distorted linear / quadratic meshes, heat flux on a face."""
    from torchfem.elements import linear_to_quadratic

    out = {}
    solid_mat = M.IsotropicElasticity3D(1000.0, 0.3)
    plane_mat = M.IsotropicElasticityPlaneStress(1000.0, 0.3)
    for tag, gen, quadratic in (("hexa1", mesh.cube_hexa, False), ("hexa2", mesh.cube_hexa, True),
                                ("tetra1", mesh.cube_tetra, False), ("tetra2", mesh.cube_tetra, True)):
        nodes, elements = gen(4, 3, 3, 1.5, 1.0, 1.0)
        on_top = nodes[:, 2] == 1.0
        boundary = (nodes[:, 0] == 0.0) | (nodes[:, 0] == 1.5) | (nodes[:, 1] == 0.0) | (nodes[:, 1] == 1.0) \
            | (nodes[:, 2] == 0.0) | (nodes[:, 2] == 1.0)
        inner = ~boundary
        nodes = nodes.clone()
        nodes[inner] = distort(nodes, 0.1, 11)[inner]
        if quadratic:
            top_xy = nodes[on_top][:, :2]
            nodes, elements = linear_to_quadratic(nodes, elements)
            on_top = torch.isclose(nodes[:, 2], torch.tensor(1.0))
            boundary = torch.zeros(len(nodes), dtype=torch.bool)
            for c, hi in ((0, 1.5), (1, 1.0), (2, 1.0)):
                boundary |= torch.isclose(nodes[:, c], torch.tensor(0.0)) | torch.isclose(nodes[:, c], torch.tensor(hi))
        model = Solid(nodes, elements, solid_mat)
        out[f"{tag}.nodes"], out[f"{tag}.elements"] = npy(nodes), npy(elements)
        out[f"{tag}.top"], out[f"{tag}.boundary"] = npy(on_top), npy(boundary)
        out[f"{tag}.body"] = npy(model.integrate_body_load(torch.tensor([0.0, 0.0, -9.81])))
        out[f"{tag}.pressure_top"] = npy(model.integrate_surface_load(on_top, -2.5))
        out[f"{tag}.traction_top"] = npy(model.integrate_surface_load(on_top, torch.tensor([1.0, 0.5, -0.25])))
        out[f"{tag}.pressure_all"] = npy(model.integrate_surface_load(boundary, 1.0))
        out[f"{tag}.facets_top"] = npy(model._boundary_facets(on_top))
    for tag, gen, quadratic in (("quad1", mesh.rect_quad, False), ("quad2", mesh.rect_quad, True),
                                ("tria1", mesh.rect_tri, False), ("tria2", mesh.rect_tri, True)):
        nodes, elements = gen(5, 4, 2.0, 1.0)
        if quadratic:
            nodes, elements = linear_to_quadratic(nodes, elements)
        right = torch.isclose(nodes[:, 0], torch.tensor(2.0))
        thickness = 0.5 + 0.1 * torch.arange(len(elements)) / len(elements)
        model = Planar(nodes, elements, plane_mat, thickness=thickness)
        out[f"{tag}.nodes"], out[f"{tag}.elements"], out[f"{tag}.right"] = npy(nodes), npy(elements), npy(right)
        out[f"{tag}.thickness"] = npy(thickness)
        out[f"{tag}.body"] = npy(model.integrate_body_load(torch.tensor([0.0, -9.81])))
        out[f"{tag}.pressure_right"] = npy(model.integrate_line_load(right, 3.0))
        out[f"{tag}.traction_right"] = npy(model.integrate_line_load(right, torch.tensor([1.0, -2.0])))
    nodes, elements = mesh.cube_hexa(3, 3, 3)
    heat = SolidHeat(nodes, elements, M.IsotropicConductivity3D(1.0))
    out["heat.flux_top"] = npy(heat.integrate_surface_load(nodes[:, 2] == 1.0, 4.0))
    out["heat.source"] = npy(heat.integrate_body_load(2.0))
    np.savez_compressed(os.path.join(OUT, "loads.npz"), **out)
    print("loads: hexa1 pressure_top sum", out["hexa1.pressure_top"].sum(0), "tria2 line", out["tria2.pressure_right"].sum(0))


def hyper_plane_stress():
    """`HyperelasticPlaneStress` (reference hyperelasticity.py:130-269): one material update on random states, and a
    Neo-Hookean strip stretched by 30 % in three increments with `nlgeom=True` (state = thickness stretch - 1)."""
    g = torch.Generator().manual_seed(21)
    mu, lbd = 384.6153846153846, 576.9230769230769

    def psi(F, params):
        Cg = F.transpose(-1, -2) @ F
        logJ = 0.5 * torch.logdet(Cg)
        return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2

    out = {}
    n = 6
    mat = M.HyperelasticPlaneStress(psi, torch.tensor([mu, lbd])).vectorize(n)
    F = torch.eye(2).expand(n, 2, 2) + 0.1 * torch.randn(n, 2, 2, generator=g)
    H = 0.02 * torch.randn(n, 2, 2, generator=g)
    state = 0.01 * torch.randn(n, 1, generator=g)
    P, st, dd = mat.step(H, F, torch.zeros(n, 2, 2), state, torch.zeros(n, 2, 2), torch.ones(n, 1), 0)
    out["F"], out["H"], out["state"] = npy(F), npy(H), npy(state)
    out["P"], out["state_new"], out["ddsdde"] = npy(P), npy(st), npy(dd)

    nodes, elements = mesh.rect_quad(5, 3, 2.0, 1.0)
    strip = Planar(nodes, elements, M.HyperelasticPlaneStress(psi, torch.tensor([mu, lbd])))
    left, right = nodes[:, 0] == 0.0, nodes[:, 0] == 2.0
    strip.constraints[left, 0] = True
    strip.constraints[right, 0] = True
    strip.constraints[nodes[:, 1] == 0.5, 1] = True
    strip.displacements[right, 0] = 0.6
    inc = torch.linspace(0.0, 1.0, 4)
    u, f, sigma, Fd, alpha = strip.solve(increments=inc, nlgeom=True, method="spsolve")
    out["strip.u"], out["strip.f"], out["strip.sigma"], out["strip.state"] = npy(u), npy(f), npy(sigma), npy(alpha)
    np.savez_compressed(os.path.join(OUT, "hyper_plane_stress.npz"), **out)
    print("hyper plane stress: reaction", float(f[right, 0].sum()), "thickness stretch", float(alpha.mean()) + 1.0)


def near_null_space():
    """`FEM.compute_B` (base.py:316-344) of a distorted solid, a planar and a thermal model: the rigid-body modes the
    reference hands to its AMG back ends (sparse.py:493-512)."""
    n3, e3 = mesh.cube_hexa(4, 3, 3, 1.5, 1.0, 1.0)
    n3 = distort(n3, 0.08, 1)
    n2, e2 = mesh.rect_quad(5, 4, 2.0, 1.0)
    n2 = distort(n2, 0.06, 3)
    solid = Solid(n3, e3, M.IsotropicElasticity3D(E=1000.0, nu=0.3))
    planar = Planar(n2, e2, M.IsotropicElasticityPlaneStress(E=1000.0, nu=0.3))
    heat = SolidHeat(n3, e3, M.IsotropicConductivity3D(kappa=2.0, rho=1.0))
    out = {"solid.nodes": npy(n3), "solid.elements": npy(e3), "solid.B": npy(solid.compute_B()), "solid.idx": npy(solid.idx),
           "planar.nodes": npy(n2), "planar.elements": npy(e2), "planar.B": npy(planar.compute_B()),
           "planar.idx": npy(planar.idx), "heat.B": npy(heat.compute_B()), "heat.idx": npy(heat.idx)}
    np.savez_compressed(os.path.join(OUT, "compute_B.npz"), **out)
    print("compute_B:", {k: v.shape for k, v in out.items() if k.endswith(".B")})


if __name__ == "__main__":
    if len(sys.argv) > 1:          # regenerate selected fixtures only: python oracle/make_golden.py modal
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    element_tables()
    small_cases()
    config_a()
    topopt_small()
    hyper_small()
    sparse_small()
    heat_transient()
    modal()
    assembly_cases()
    orthotropic()
    loads()
    hyper_plane_stress()
    near_null_space()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
