"""The reference's three benchmark problems through the drop-in model API on one B200, at the sizes the reference
publishes: setup / forward / backward wall-clock (device-synchronised), as `benchmarks/run.py`
reports them (setup = model constructor incl. the sparsity pattern, fwd = `solve`, bwd = adjoint `backward`).

    python tools/run_workloads.py --cube 80 --topopt 60 --hyper 65 [--method cg]

Problem definitions follow the reference's generators: benchmarks/cubes.py:9-37 (cube extension, forces as the
differentiable parameter), benchmarks/topopt.py:23-63 (SIMP cantilever, rho as parameter, compliance gradient),
benchmarks/hyperelasticity.py:24-68 (Neo-Hookean stretch, 10 geometric increments, Lame parameters as parameters).
The reference's published runs use AMG back ends (AmgX / pyamg); here the linear solver is the Jacobi-PCG of the
hot path (method="cg", stol as the reference default 1e-10), so the comparison is end to end, not per iteration.
Prints one JSON line per problem.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t, out


def cube(T, N, method, stol):
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    nodes, elements = cube_hexa(N, N, N)

    def build():
        m = T.Solid(nodes, elements, IsotropicElasticity3D(E=1000.0, nu=0.3))
        m.forces = torch.zeros_like(nodes, requires_grad=True)
        m.constraints[nodes[:, 0] == 0.0, :] = True
        m.constraints[nodes[:, 0] == 1.0, 0] = True
        m.displacements[nodes[:, 0] == 1.0, 0] = 0.1
        return m

    t_setup, m = timed(build)
    t_fwd, out = timed(lambda: m.solve(differentiable_parameters=m.forces, method=method, stol=stol))
    u = out[0]
    t_bwd, _ = timed(lambda: u.sum().backward())
    return {"problem": "cube_hexa_extension", "N": N, "dofs": 3 * m.n_nod, "setup_s": t_setup, "fwd_s": t_fwd,
            "bwd_s": t_bwd, "max_u": float(u.max()), "grad_forces_norm": float(m.forces.grad.norm())}


def topopt(T, N, method, stol):
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    nx, ny, nz = 2 * N, N, N
    nodes, elements = cube_hexa(nx + 1, ny + 1, nz + 1, 2.0, 1.0, 1.0)
    rng = np.random.default_rng(0)
    values = np.clip(0.5 + 0.3 * rng.standard_normal(len(elements)), 0.05, 0.95)
    rho = torch.tensor(values, requires_grad=True)

    def build():
        material = IsotropicElasticity3D(E=70000.0, nu=0.3).vectorize(len(elements))
        scale = 1e-3 + (1.0 - 1e-3) * rho ** 3.0
        material.C = scale[:, None, None, None, None] * material.C
        m = T.Solid(nodes, elements, material)
        m.constraints[nodes[:, 0] == 0.0, :] = True
        right = nodes[:, 0] == 2.0
        wy = torch.full((m.n_nod,), 1.0 / ny)
        wy[(nodes[:, 1] == 0.0) | (nodes[:, 1] == 1.0)] /= 2.0
        wz = torch.full((m.n_nod,), 1.0 / nz)
        wz[(nodes[:, 2] == 0.0) | (nodes[:, 2] == 1.0)] /= 2.0
        m.forces[right, 2] = -1.0 * wy[right] * wz[right]
        return m

    t_setup, m = timed(build)
    t_fwd, out = timed(lambda: m.solve(differentiable_parameters=rho, method=method, stol=stol))
    u = out[0]
    compliance = torch.inner(m.forces.ravel(), u.ravel())
    t_bwd, _ = timed(lambda: compliance.backward())
    return {"problem": "structural_cantilever_simp", "N": N, "dofs": 3 * m.n_nod, "setup_s": t_setup, "fwd_s": t_fwd,
            "bwd_s": t_bwd, "compliance": float(compliance), "grad_rho_norm": float(rho.grad.norm())}


def neo_hooke_psi(F, params):
    """Strain energy of benchmarks/hyperelasticity.py:24-30 — a module-level function there too."""
    Cg = F.transpose(-1, -2) @ F
    logJ = 0.5 * torch.logdet(Cg)
    return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2


def hyper(T, N, method, stol):
    from torchfem_b200.materials import Hyperelastic3D
    from torchfem_b200.mesh import cube_hexa

    En, NU, STRETCH = 1000.0, 0.3, 10.0
    LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
    MU = En / (2.0 * (1.0 + NU))

    psi = neo_hooke_psi

    lx = 4.0 / (N - 1)
    nodes, elements = cube_hexa(5, N, N, lx, 1.0, 1.0)
    params = torch.tensor([MU, LBD], requires_grad=True)
    right = nodes[:, 0] == lx

    def build():
        box = T.Solid(nodes, elements, Hyperelastic3D(psi, params))
        box.constraints[nodes[:, 0] == 0.0, 0] = True
        box.constraints[right, 0] = True
        box.constraints[nodes[:, 1] == 0.5, 1] = True
        box.constraints[nodes[:, 2] == 0.5, 2] = True
        box.displacements[right, 0] = (STRETCH - 1.0) * lx
        return box

    lam = torch.logspace(0, math.log10(STRETCH), 11)
    increments = (lam - 1.0) / (STRETCH - 1.0)
    t_setup, box = timed(build)
    t_fwd, out = timed(lambda: box.solve(increments=increments, nlgeom=True, differentiable_parameters=params,
                                         method=method, stol=stol))
    reaction = out[1][right, 0].sum()
    t_bwd, _ = timed(lambda: reaction.backward())
    return {"problem": "hyperelasticity_stretch", "N": N, "dofs": 3 * box.n_nod, "setup_s": t_setup, "fwd_s": t_fwd,
            "bwd_s": t_bwd, "reaction": float(reaction), "grad_params": [float(v) for v in params.grad]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=80)
    ap.add_argument("--topopt", type=int, default=60)
    ap.add_argument("--hyper", type=int, default=65)
    ap.add_argument("--method", default="cg")
    ap.add_argument("--stol", type=float, default=1e-10)
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda")  # as benchmarks/utils.py:59-60
    import torchfem_b200 as T

    T.Solid(*__import__("torchfem_b200").mesh.cube_hexa(4, 4, 4), T.materials.IsotropicElasticity3D(1.0, 0.3))  # warm up
    for name, fn, n in (("cube", cube, a.cube), ("topopt", topopt, a.topopt), ("hyper", hyper, a.hyper)):
        if n <= 0:
            continue
        torch.cuda.reset_peak_memory_stats()
        cold = fn(T, n, a.method, a.stol)    # first run in the process, as benchmarks/run.py measures (lazy CUDA
        row = fn(T, n, a.method, a.stol)     # module loads, allocator growth, torch.func tracing); then a warm run
        row.update({f"cold_{k}": cold[k] for k in ("setup_s", "fwd_s", "bwd_s")})
        row.update(method=a.method, stol=a.stol, peak_vram_mb=torch.cuda.max_memory_allocated() / 2 ** 20,
                   hardware=torch.cuda.get_device_name(0))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
