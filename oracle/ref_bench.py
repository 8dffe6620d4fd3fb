"""TEST INFRASTRUCTURE ONLY — the reference's OWN CPU implementation of the hot path, timed phase by phase.

Runs the unmodified `torchfem` (from `/root/reference/src` or the staged `oracle/_ref`, see `oracle/build_ref.py`) on
the cube-extension problem of `benchmarks/cubes.py:9-24`, exactly the calls `FEM.solve` makes for one linear solve
(reference base.py:704-743 -> sparse.py:271-347, 447-514), with the phases BASELINE.md §2 names:

    setup      Solid(nodes, elements, material)                         base.py:42-132      (reported, not in the metric)
    integrate  model.integrate_material(..., du_bc, ...) -> k, f        base.py:982-1092
    assemble   model.assemble_matrix(k, con)                            base.py:398-426
    rhs        model.assemble_rhs(f) - F_ext ; res[con] = 0             base.py:428-445, 736-741
    solve      sparse_solve(K, res, method="cg", stol=rtol, M=Jacobi)   sparse.py:500-512 (scipy cg)

The preconditioner is the `LinearOperator(x -> x / diag(K))` the reference builds on its GPU path (sparse.py:408-409);
its CPU default (pyamg smoothed aggregation) is not installable offline. Only `bench.py` (`--impl reference`,
`cpu_baseline`) and `tests/` may import this file.
"""
from __future__ import annotations

import time

import numpy as np


def available() -> bool:
    from . import ref_import

    return ref_import.available()


def cube_extension_model(E: int):
    """The model of benchmarks/cubes.py:9-24 with E elements per edge, on the CPU, float64."""
    import torch

    from . import ref_import

    tf = ref_import.load()
    from torchfem.materials import IsotropicElasticity3D
    from torchfem.mesh import cube_hexa

    with torch.device("cpu"):
        t0 = time.perf_counter()
        nodes, elements = cube_hexa(E + 1, E + 1, E + 1)
        material = IsotropicElasticity3D(E=1000.0, nu=0.3)
        t1 = time.perf_counter()
        model = tf.Solid(nodes, elements, material)
        t_setup = time.perf_counter() - t1
        model.constraints[nodes[:, 0] == 0.0, :] = True
        model.constraints[nodes[:, 0] == 1.0, 0] = True
        model.displacements[nodes[:, 0] == 1.0, 0] = 0.1
    return model, {"t_mesh": t1 - t0, "t_setup": t_setup}


def linear_solve(model, rtol: float = 1e-8):
    """One pass of the hot path through the reference's own functions. Returns (u, phases dict)."""
    import torch
    from scipy.sparse.linalg import LinearOperator
    from torchfem.sparse import sparse_solve

    with torch.device("cpu"):
        n_dofs = model.n_dofs
        con = torch.nonzero(model.constraints.ravel(), as_tuple=False).ravel()
        DU = model.displacements.clone().ravel()
        F_ext = model.forces.ravel()
        u = torch.zeros(model.n_nod, model.n_dof_per_node)
        grad = torch.zeros(model.n_int, model.n_elem, *model.n_flux)
        grad[:] = model.initial_grad
        flux = torch.zeros(model.n_int, model.n_elem, *model.n_flux)
        state = torch.zeros(model.n_int, model.n_elem, model.n_state)
        de0 = torch.zeros(model.n_elem, *model.n_flux)
        du_bc = torch.zeros(n_dofs)
        du_bc[con] = DU[con]
        model.K = torch.empty(0)

        t0 = time.perf_counter()
        k, f_i, _, _, _ = model.integrate_material(u, grad, flux, state, du_bc, de0, 0, False)
        t1 = time.perf_counter()
        K = model.assemble_matrix(k, con)
        model.K = K
        t2 = time.perf_counter()
        res = model.assemble_rhs(f_i) - F_ext
        res[con] = 0.0
        t3 = time.perf_counter()
        dinv = 1.0 / K.to_dense().diagonal().numpy() if n_dofs < 2000 else None
        if dinv is None:
            idx = K._indices()
            on_diag = idx[0] == idx[1]
            d = torch.zeros(n_dofs)
            d[idx[0][on_diag]] = K._values()[on_diag]
            dinv = (1.0 / d).numpy()
        its = [0]

        def apply_m(x):
            its[0] += 1
            return dinv * x

        M = LinearOperator((n_dofs, n_dofs), matvec=apply_m)
        t4 = time.perf_counter()
        x, _ = sparse_solve(K, res, None, rtol, "cpu", "cg", M, None)
        t5 = time.perf_counter()
        u_out = du_bc - x
        u_out[con] = DU[con]
    phases = {"t_integrate": t1 - t0, "t_assemble": t2 - t1, "t_rhs": t3 - t2, "t_jacobi": t4 - t3,
              "t_solve": t5 - t4, "iterations": max(0, its[0] - 1), "n_dofs": int(n_dofs)}
    return u_out.numpy().reshape(-1, 3), phases


def hot_path_seconds(ph: dict) -> float:
    return ph["t_integrate"] + ph["t_assemble"] + ph["t_rhs"] + ph["t_jacobi"] + ph["t_solve"]


def spmv_seconds(model, reps: int = 10) -> float:
    """scipy CSR `A @ x` on the assembled matrix (what runs inside scipy's cg), median of `reps`."""
    import scipy.sparse as sp

    K = model.K
    A = sp.coo_matrix((K._values().numpy(), (K._indices()[0].numpy(), K._indices()[1].numpy())), shape=tuple(K.shape)).tocsr()
    x = np.random.default_rng(0).standard_normal(A.shape[0])
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        A @ x
        ts.append(time.perf_counter() - t)
    return float(np.median(ts))
