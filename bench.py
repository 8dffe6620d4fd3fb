#!/usr/bin/env python
"""bench.py — DOFs solved/sec (integrate + assemble + Jacobi-PCG to 1e-8) on the linear-elastic Hexa1 cube.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--edge E]

One "step" = one pass of the hot path over the synthetic cube-extension problem of the reference's
benchmarks/cubes.py (cube_hexa(E+1,E+1,E+1), E=1000, nu=0.3, x=0 clamped, u_x=0.1 at x=1):
element matrices (K1) -> deterministic assembly with Dirichlet masking (K2/K3) -> right-hand side ->
Jacobi setup (K4) -> PCG to ||r|| <= 1e-8 ||b|| (K5/K6). The sparsity pattern (K0) is setup and is reported
separately, as the reference does (benchmarks/run.py:56-58: "setup" vs "fwd").

N=1 runs BASELINE.json configs[1]: 150^3 elements, 10,328,853 DOFs, nnz 825,604,659 (matrix 10.1 GB >> L2).
`--impl reference` times the reference's OWN CPU code (staged by `__graft_entry__.build()` into oracle/_ref; the numpy
oracle port only if that is missing) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DOFs solved/sec (assembly+PCG)"
UNIT = "DOF/s"
RTOL = 1e-8
NCU_SPMV_DRAM_BYTES_CONFIG_B = 7233596000 + 85989632   # profiles/r2_x_cg_iteration_ncu.txt


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_step(E, state=None):
    """One pass of the hot path on the host cores over a bounded E^3 sample. Runs the reference's OWN code (staged
    copy `oracle/_ref`, or /root/reference in the build container: kind "reference"); the numpy/scipy oracle port is
    only the fallback when neither exists (kind "port"). Returns (seconds, iterations, n_dofs, kind, detail, state)."""
    from oracle import ref_bench as R

    if R.available():
        if state is None:
            model, setup = R.cube_extension_model(E)
            state = (model, setup)
        model, setup = state
        _, ph = R.linear_solve(model, rtol=RTOL)
        detail = (f"setup {setup['t_setup']:.2f}s (excluded) integrate_material {ph['t_integrate']:.2f}s assemble_matrix "
                  f"{ph['t_assemble']:.2f}s rhs {ph['t_rhs']:.3f}s sparse_solve(cg, Jacobi M) {ph['t_solve']:.2f}s; unmodified "
                  f"torchfem (src/torchfem/base.py:982-1092, 398-445; sparse.py:447-514 -> scipy cg), torch "
                  f"{_torch_threads()} intra-op threads, scipy CSR SpMV single-threaded")
        return R.hot_path_seconds(ph), ph["iterations"], ph["n_dofs"], "reference", detail, state
    from oracle import fem_oracle as O

    if state is None:
        N = E + 1
        nodes, elements = O.cube_hexa(N, N, N)
        bref, w = O.hexa1_tables()
        C = O.isotropic_C3d(1000.0, 0.3, len(elements))
        con_mask, disp = O.cube_extension_bcs(nodes)
        state = (nodes, elements, bref, w, C, con_mask, disp)
    nodes, elements, bref, w, C, con_mask, disp = state
    out = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=RTOL)
    t = out["t_integrate"] + out["t_assemble"] + out["t_rhs"] + out["t_solve"]
    detail = (f"integrate {out['t_integrate']:.2f}s assemble {out['t_assemble']:.2f}s solve {out['t_solve']:.2f}s; "
              f"numpy/scipy oracle PORT of the reference CPU path (oracle/_ref not staged), scipy-CSR SpMV single-threaded")
    return t, out["iterations"], int(nodes.size), "port", detail, state


def _torch_threads():
    try:
        import torch

        return torch.get_num_threads()
    except Exception:
        return 1


def _use_all_host_threads():
    """The reference arm runs with all the host threads it can use. torchrun exports OMP_NUM_THREADS=1 to every rank
    unless the variable is set, which would leave the reference's torch ops on one core (integrate_material 4.3 s instead
    of 0.6 s on the 40^3 sample); only rank 0 runs the reference, so it takes the whole box."""
    try:
        import torch

        n = os.cpu_count() or 1
        if torch.get_num_threads() < n:
            torch.set_num_threads(n)
        return torch.get_num_threads()
    except Exception:
        return 1


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, every step a
    bounded sample (E = --cpu-edge) of the workload; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    E = args.cpu_edge
    _use_all_host_threads()
    times, state = [], None
    its = n_dofs = 0
    kind = detail = ""
    for s in range(args.warmup + args.steps):
        t, its, n_dofs, kind, detail, state = cpu_reference_step(E, state)
        if s >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = n_dofs / (ms / 1e3)
    cores = _torch_threads()
    sample = (f"cube_hexa({E + 1},{E + 1},{E + 1}) = {E}^3 Hexa1 elements, {n_dofs} DOFs, {its} Jacobi-CG iterations to "
              f"1e-8 (last step: {detail}); os.cpu_count()={cores}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"linear-elastic Hexa1 cube {E}^3 elements (bounded CPU sample of configs[1])",
                   "n_dofs": int(n_dofs), "rtol": RTOL, "cg_iterations": int(its)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def element_tables(etype):
    """(bref [n_q, dim, nn], w [n_q]) of an element type of the product package, on the host in float64: what
    `csr.integrate_k` takes (the timing tools under tools/ use this too; nothing on the `ours` arm comes from oracle/)."""
    import torch

    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)   # the Gauss points are created in the default dtype, like the reference's
    try:
        ip = etype.ipoints.to(torch.float64).cpu()
        return etype.B(ip), etype.iweights.to(torch.float64).cpu()
    finally:
        torch.set_default_dtype(prev)


def build_problem(T, torch, E, device):
    """Synthetic inputs exactly as the reference generators make them (mesh.py:8-46, cubes.py:9-24),
    built with torch on the host, then moved to the device."""
    N = E + 1
    X = torch.linspace(0, 1.0, N, dtype=torch.float64)
    x, y, z = torch.meshgrid(X, X, X, indexing="ij")
    nodes = torch.stack([x.ravel(), y.ravel(), z.ravel()], dim=1).contiguous()
    ind = torch.arange(N * N * N, dtype=torch.int64).reshape(N, N, N)
    n0 = ind[:-1, :-1, :-1].ravel()
    n1 = ind[1:, :-1, :-1].ravel()
    n2 = ind[:-1, 1:, :-1].ravel()
    n3 = ind[1:, 1:, :-1].ravel()
    n4 = ind[:-1, :-1, 1:].ravel()
    n5 = ind[1:, :-1, 1:].ravel()
    n6 = ind[:-1, 1:, 1:].ravel()
    n7 = ind[1:, 1:, 1:].ravel()
    elements = torch.stack([n0, n1, n3, n2, n4, n5, n7, n6], dim=1).contiguous()
    con = torch.zeros(N * N * N, 3, dtype=torch.bool)
    disp = torch.zeros(N * N * N, 3, dtype=torch.float64)
    con[nodes[:, 0] == 0.0, :] = True
    con[nodes[:, 0] == 1.0, 0] = True
    disp[nodes[:, 0] == 1.0, 0] = 0.1
    return nodes, elements, con, disp


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200 import csr

    if world > 1 or args.config == "C" or args.dist_path:
        return run_multi_gpu(args)

    E = args.edge
    t0 = time.perf_counter()
    nodes_h, elements_h, con_h, disp_h = build_problem(T, torch, E, device)
    n_elem = elements_h.shape[0]
    n_dofs = nodes_h.numel()
    from torchfem_b200.elements import Hexa1
    from torchfem_b200.materials import IsotropicElasticity3D

    # reference-element tables of the PRODUCT package (elements.py; pinned to the reference in tests/test_elements.py)
    bref, w = element_tables(Hexa1)

    # per-element material parameters (heterogeneous materials are the general case, cf. benchmarks/topopt.py);
    # the [n_elem,3,3,3,3] tangent is built from them ON THE DEVICE, as the reference's vectorised material does
    host = [nodes_h, elements_h, torch.full((n_elem,), 1000.0), torch.full((n_elem,), 0.3),
            con_h.ravel().to(torch.uint8), disp_h.ravel().contiguous()]
    host = [t.pin_memory() for t in host]
    nodes_h, elements_h, E_h, nu_h, iscon_h, disp_h = host
    t_gen = time.perf_counter() - t0

    # ---- setup (pattern), timed separately
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    elements = elements_h.to(device, non_blocking=True)
    ev0.record()
    pattern = csr.Pattern(elements, nodes_h.shape[0], 3)
    ev1.record()
    torch.cuda.synchronize()
    t_setup_ms = ev0.elapsed_time(ev1)
    nnz = pattern.nnz

    nodes = nodes_h.to(device)
    E_mod, nu = E_h.to(device), nu_h.to(device)
    is_con = iscon_h.to(device)
    disp = disp_h.to(device)
    sell_vals = torch.empty(max(pattern.sell_structure.padded, 2), dtype=torch.float64, device=device)
    dinv_buf = torch.empty(n_dofs, dtype=torch.float64, device=device)
    rhs_buf = torch.empty(n_dofs, dtype=torch.float64, device=device)
    state = {}

    def hot_path(nodes, elements, E_mod, nu, is_con, disp):
        """material tangent -> integrate -> assemble (Dirichlet rows/cols masked; the entries being masked give
        the right-hand side K[free, con] u_con in the same pass) -> Jacobi -> PCG."""
        C = IsotropicElasticity3D(E_mod, nu).C
        k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False)
        del C
        # residual of the first Newton step: F_int(du_bc) with du_bc = prescribed increment (base.py:708-741)
        rhs = rhs_buf
        # the assembly writes what the solve streams: values in SELL-32 order and 1/diagonal (no CSR copy, no
        # CSR -> SELL pass, no Jacobi setup pass)
        csr.assemble(pattern, k, is_con, ubc=disp, lift=rhs, csr=False, sell_out=sell_vals, dinv_out=dinv_buf)
        del k
        A = pattern.matrix(None, sell_vals=sell_vals)
        x, M, info = csr.krylov_solve(A, rhs, method="cg", rtol=RTOL, M=csr.JacobiPreconditioner(dinv=dinv_buf))
        u = torch.where(is_con.bool(), disp, -x)
        state.update(info=info, A=A, rhs=rhs, x=x)
        return u

    def step_resident():
        return hot_path(nodes, elements, E_mod, nu, is_con, disp)

    u_pinned = torch.empty(n_dofs, dtype=torch.float64).pin_memory()   # the user's result buffer (pinned like the inputs)

    def step_e2e():
        d = [t.to(device, non_blocking=True) for t in host]
        u = hot_path(*d)
        u_pinned.copy_(u, non_blocking=True)      # stream-ordered: inside the timed region
        return u_pinned

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, u = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    info = dict(state["info"])
    value = n_dofs / (ms / 1e3)

    ms_e2e, u_h = timed(step_e2e, max(1, min(args.steps, 3)), 1)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = u_h.numel() * u_h.element_size()

    # ---- true relative residual of the last solve (checks the work was done)
    A, rhs, x = state["A"], state["rhs"], state["x"]
    true_res = float(torch.linalg.norm(rhs - A.matvec(x)) / torch.linalg.norm(rhs))
    # the side measurements below (CSR-chunk SpMV, AMG setup) read CSR values, which the timed step never writes:
    # the same matrix once more in both orders
    C_ = IsotropicElasticity3D(E_mod, nu).C
    k_ = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C_, check=False)
    vals = csr.assemble(pattern, k_, is_con)
    A = pattern.matrix(vals, sell_vals=sell_vals)
    del C_, k_

    # ---- per-phase times + dominant kernel (SpMV) measured live with CUDA events
    def ev_time(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    xs = torch.randn(n_dofs, dtype=torch.float64, device=device,
                     generator=torch.Generator(device=device).manual_seed(0))
    ys = torch.empty_like(xs)
    spmv_ms = ev_time(lambda: A.matvec(xs, out=ys, fmt="sell"), 20)
    spmv_scalar_ms = ev_time(lambda: A.matvec(xs, out=ys, fmt="sell-scalar"), 10)
    A._sell_mats.pop(False, None)  # drop the 3.3 GB scalar-column copy again
    A._sell_struct._cols = None
    spmv_csr_ms = ev_time(lambda: A.matvec(xs, out=ys, fmt="csr"), 5)
    C_ = IsotropicElasticity3D(E_mod, nu).C
    k_ = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C_, check=False)
    del C_
    Aop = csr.ElementOperator(pattern, k_, is_con)
    spmv_ebe_ms = ev_time(lambda: Aop.matvec(xs, out=ys), 3)   # K8, matrix-free on stored element matrices
    del Aop, k_
    spmv_bytes = 12 * nnz + 20 * n_dofs
    peak, peak_src = measured_peaks()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    C = IsotropicElasticity3D(E_mod, nu).C
    k_ms = ev_time(lambda: csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False), 3)
    kk = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C, check=False)
    del C
    asm_ms = ev_time(lambda: csr.assemble(pattern, kk, is_con, ubc=disp, lift=rhs_buf, csr=False, sell_out=sell_vals,
                                          dinv_out=dinv_buf), 3)          # what the step runs
    asm_csr_ms = ev_time(lambda: csr.assemble(pattern, kk, is_con, out=vals), 3)   # CSR order (round 1's step: + a
    # CSR -> SELL copy and a Jacobi pass)
    del kk
    t_s = time.perf_counter()
    x2, _, info2 = csr.krylov_solve(A, rhs, method="cg", rtol=RTOL)
    torch.cuda.synchronize()
    solve_ms = 1e3 * (time.perf_counter() - t_s)

    # ---- the same system with the AMG-preconditioned CG behind sparse_solve(method="amgx") (the reference's GPU
    # default when AmgX is installed, sparse.py:422-442): NOT the headline (BASELINE names Jacobi-PCG), reported beside it
    amg_info = None
    if not args.no_amg:
        from torchfem_b200.amg import AMGPreconditioner

        def amg_once():
            torch.cuda.synchronize()
            t_a = time.perf_counter()
            Mp = AMGPreconditioner(A)
            torch.cuda.synchronize()
            t_b = time.perf_counter()
            xa, st = Mp.solve(rhs, rtol=RTOL)
            torch.cuda.synchronize()
            return Mp, xa, st, 1e3 * (t_b - t_a), 1e3 * (time.perf_counter() - t_b)

        Mp, xa, st_a, _, _ = amg_once()       # first call grows the allocator pools
        del Mp, xa
        Mp, xa, st_a, amg_setup_ms, amg_solve_ms = amg_once()
        amg_res = float(torch.linalg.norm(rhs - A.matvec(xa)) / torch.linalg.norm(rhs))
        amg_info = {"call": "sparse_solve(method='amgx'): smoothed-aggregation AMG V(1,1) + CG, kernels K11-K16",
                    "setup_ms": amg_setup_ms, "solve_ms": amg_solve_ms, "iterations": st_a["iterations"],
                    "ms_per_iteration": amg_solve_ms / max(1, st_a["iterations"]),
                    "true_rel_residual": amg_res, "rel_diff_vs_jacobi_pcg": float((xa - x2).norm() / x2.norm()),
                    "levels": [int(lv.n) for lv in Mp.levels], "operator_complexity": Mp.operator_complexity,
                    "dofs_per_s_integrate_assemble_setup_solve": n_dofs / ((k_ms + asm_ms + amg_setup_ms + amg_solve_ms) * 1e-3),
                    "speedup_vs_jacobi_pcg_solve": solve_ms / (amg_setup_ms + amg_solve_ms)}
        # parity statement at the benchmark size: the two product solvers, both iterated to stol 1e-12, must give the
        # same displacement to <= 1e-8 (north_star's displacement bar) although they stop on residuals
        xj12, _, ij12 = csr.krylov_solve(A, rhs, method="cg", rtol=1e-12)
        xa12, sa12 = Mp.solve(rhs, rtol=1e-12)
        amg_info["cross_check_stol_1e-12"] = {
            "rel_diff_jacobi_pcg_vs_amg_pcg": float((xa12 - xj12).norm() / xj12.norm()),
            "iterations_jacobi": ij12["iterations"], "iterations_amg": sa12["iterations"],
            "rel_diff_stol_1e-8_vs_1e-12_jacobi": float((x2 - xj12).norm() / xj12.norm()),
            "bar": 1e-8}
        del Mp, xa, xj12, xa12

    # ---- the same through the public model API (`Solid.solve`, the call a torch-fem user makes):
    # the model (mesh + pattern) is setup; per step the material tangent and the boundary conditions
    # arrive from pinned host memory and the displacement field goes back to the host.
    api = None
    if not args.no_api:
        state.clear()
        del A, rhs, x, x2, xs, ys
        torch.cuda.empty_cache()
        model = T.Solid(nodes, elements, IsotropicElasticity3D(E=E_mod, nu=nu))
        model.pattern = pattern  # reuse the setup product instead of building it twice
        con_bool_h = con_h.pin_memory()
        disp2_h = disp_h.reshape(-1, 3)
        u_api_pinned = torch.empty(nodes_h.shape, dtype=torch.float64).pin_memory()
        u_amg_pinned = torch.empty(nodes_h.shape, dtype=torch.float64).pin_memory()

        def step_api():
            model.material = IsotropicElasticity3D(E_h.to(device, non_blocking=True), nu_h.to(device, non_blocking=True))
            model.constraints = con_bool_h.to(device, non_blocking=True)
            model.displacements = disp2_h.to(device, non_blocking=True)
            u, *_ = model.solve(method="cg", stol=RTOL, rtol=1e-6)
            u_api_pinned.copy_(u, non_blocking=True)     # stream-ordered, inside the timed region
            return u_api_pinned

        ms_api, u_api = timed(step_api, max(1, min(args.steps, 2)), 1)
        api_amg = None
        if not args.no_amg:
            def step_api_amg():
                model.material = IsotropicElasticity3D(E_h.to(device, non_blocking=True), nu_h.to(device, non_blocking=True))
                model.constraints = con_bool_h.to(device, non_blocking=True)
                model.displacements = disp2_h.to(device, non_blocking=True)
                u, *_ = model.solve(method="amgx", stol=RTOL, rtol=1e-6)
                u_amg_pinned.copy_(u, non_blocking=True)
                return u_amg_pinned

            ms_api_amg, u_amg = timed(step_api_amg, max(1, min(args.steps, 2)), 1)
            api_amg = {"value": n_dofs / (ms_api_amg / 1e3), "unit": UNIT, "ms_per_step": ms_api_amg,
                       "call": "Solid.solve(method='amgx', stol=1e-8)",
                       "rel_diff_vs_kernel_path": float((u_amg.ravel() - u_h.ravel()).norm() / u_h.norm())}
        err = float((u_api.ravel() - u_h.ravel()).norm() / u_h.norm())
        api = {"value": n_dofs / (ms_api / 1e3), "unit": UNIT, "ms_per_step": ms_api,
               "h2d_bytes_per_step": int(2 * E_h.numel() * 8 + con_bool_h.numel() + disp2_h.numel() * 8),
               "d2h_bytes_per_step": int(d2h), "call": "Solid.solve(method='cg', stol=1e-8)",
               "rel_diff_vs_kernel_path": err, "with_amg": api_amg}

    # ---- CPU baseline: the reference's own code (oracle/_ref) on a bounded sample, rank 0 / N=1 only
    cpu = None
    if not args.no_cpu_baseline:
        n_thr = _use_all_host_threads()
        t_cpu, its_c, n_c, kind_c, detail_c, _ = cpu_reference_step(args.cpu_edge)
        cpu = {"value": n_c / t_cpu, "unit": UNIT, "cores": n_thr, "kind": kind_c,
               "sample": f"{args.cpu_edge}^3 Hexa1 elements, {n_c} DOFs, {its_c} CG its to 1e-8: {detail_c}"}

    launches = info["launches"] + 1 + 1  # integrate, assemble (+ lifting, SELL-order values, 1/diagonal)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"linear-elastic Hexa1 cube {E}^3 elements, Jacobi-PCG to 1e-8 (BASELINE configs[1])"
                   if E == 150 else f"linear-elastic Hexa1 cube {E}^3 elements, Jacobi-PCG to 1e-8",
                   "n_dofs": int(n_dofs), "n_elem": int(n_elem), "nnz": int(nnz), "rtol": RTOL,
                   "cg_iterations": info["iterations"], "true_rel_residual": true_res,
                   "l2_policy": "inputs larger than L2 (CSR matrix 10.1 GB, k_e 15.6 GB vs 126 MB L2)",
                   "setup_ms_pattern": t_setup_ms, "phases_ms": {"integrate_k": k_ms, "assemble": asm_ms,
                                                                 "pcg_solve": solve_ms,
                                                                 "assemble_csr_order_not_in_step": asm_csr_ms},
                   "per_iteration_ms": solve_ms / max(1, info2["iterations"]),   # solve only (same key at N > 1)
                   "amg_pcg": amg_info},
        "e2e": {"value": n_dofs / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                "call": "C-ABI ops on host buffers (H2D mesh + per-element E, nu + BCs -> tangent -> integrate -> "
                        "assemble -> PCG -> D2H u)",
                "public_api": api},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_bsell_spmv<3> (SELL-32 values + node-block column indices, 8.5 B/nnz moved; "
                               "scored against the scalar-CSR algorithmic bytes 12*nnz+20*n)",
                     "sell_scalar_cols_kernel_ms": spmv_scalar_ms,
                     "sell_scalar_cols_frac": spmv_bytes / (spmv_scalar_ms * 1e-3) / 1e9 / peak,
                     "csr_chunk_kernel_ms": spmv_csr_ms,
                     "matrix_free_ebe_kernel_ms": spmv_ebe_ms,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_spec_8000": achieved / 8000.0,
                     "algorithmic_bytes": int(spmv_bytes), "ms_per_launch": spmv_ms,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at config B from the committed
                     # `ncu --set full` capture (profiles/r2_x_cg_iteration_ncu.txt); other sizes: not captured
                     "traffic": NCU_SPMV_DRAM_BYTES_CONFIG_B if E == 150 else None,
                     "traffic_source": "profiles/r2_x_cg_iteration_ncu.txt (k_sell_spmv<3,1>: 7.234 GB read + 0.086 GB written per launch, "
                                       "1.116 ms under ncu; launch list of this command: profiles/r2_fin_launches_bench_command.csv)"},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------ N > 1 / config C
def run_multi_gpu(args):
    """N > 1 (launched by torchrun, one rank per GPU): weak scaling of the cube benchmark — the global mesh is the
    cube with N times the elements of the single-GPU config (edge E*N^(1/3): 189^3, 238^3, 300^3 elements for
    N = 2, 4, 8 at E = 150), cut into N slabs of x-planes, so every rank holds ~E^3 elements. `--config C`: the Hexa2
    cube of BASELINE configs[2], one global mesh cut by x-coordinate (strong scaling). value = global DOFs /
    max-over-ranks device time of integrate + assemble + rhs + distributed Jacobi-PCG."""
    import torch
    import torch.distributed as dist

    from torchfem_b200 import _lib as L
    from torchfem_b200 import csr
    from torchfem_b200 import distributed as D
    from torchfem_b200.materials import IsotropicElasticity3D

    multi = dist.is_initialized()

    def _barrier():
        if multi:
            dist.barrier()

    def _allreduce(t, op=None):
        if multi:
            dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM)

    rank, world = (dist.get_rank(), dist.get_world_size()) if multi else (0, 1)
    dev = torch.device("cuda", torch.cuda.current_device())
    E = args.edge
    rtol = args.rtol
    config_c = args.config == "C"
    if config_c:
        # BASELINE configs[2]: Hexa2 (20-node serendipity) cube, strong scaling of ONE global mesh. The mesh is
        # made once per rank on the host with the reference-order generator, then cut by x-coordinate.
        from torchfem_b200.elements import Hexa2 as EType, linear_to_quadratic
        from torchfem_b200.mesh import cube_hexa

        with torch.device("cpu"):
            nodes_g, elements_g = linear_to_quadratic(*cube_hexa(E + 1, E + 1, E + 1))
        nodes_h, mesh, ranges, perm = D.coordinate_partition(nodes_g, elements_g, world, rank)
        n_dofs_global = 3 * nodes_g.shape[0]
        n_elem_global = elements_g.shape[0]
        del nodes_g, elements_g
        Lx = 1.0
        workload = (f"linear-elastic Hexa2 (20-node) cube {E}^3 elements ({n_dofs_global} DOFs), x-coordinate partition "
                    f"into {world} node blocks, Jacobi-PCG to {rtol:g} (BASELINE configs[2]; strong scaling)")
    else:
        from torchfem_b200.elements import Hexa1 as EType

        Eg = D.weak_scaling_edge(E, world)
        h = 1.0 / E
        nodes_h, mesh, ranges, (Nx, Ny, Nz) = D.cube_slab(Eg, Eg, Eg, h, world, rank)
        Lx = Eg * h
        n_dofs_global = Nx * Ny * Nz * 3
        n_elem_global = Eg ** 3
        workload = (f"linear-elastic Hexa1 cube {Eg}^3 elements (= {world} x {E}^3, weak scaling of BASELINE "
                    f"configs[1]), {world} slabs of x-planes, Jacobi-PCG to {rtol:g}")
    con_h = torch.zeros(mesh.n_local, 3, dtype=torch.bool)
    disp_h = torch.zeros(mesh.n_local, 3, dtype=torch.float64)
    con_h[nodes_h[:, 0] == 0.0, :] = True
    right = (nodes_h[:, 0] - Lx).abs() < 1e-12
    con_h[right, 0] = True
    disp_h[right, 0] = 0.1
    plan = D.build_halo_plan(mesh, ranges, rank, 3)
    halo = D.HaloExchanger(plan, dev)
    host = [nodes_h.contiguous(), mesh.elements.contiguous(), con_h.ravel().to(torch.uint8), disp_h.ravel().contiguous(),
            torch.full((len(mesh.elements),), 1000.0), torch.full((len(mesh.elements),), 0.3)]
    host = [t.pin_memory() for t in host]
    nodes, elements, is_con, disp, E_mod, nu = (t.to(dev) for t in host)
    ip = EType.ipoints.to(torch.float64).cpu()
    bref, w = EType.B(ip), EType.iweights.to(torch.float64).cpu()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pattern = csr.Pattern(elements, mesh.n_local, 3)
    ev1.record()
    torch.cuda.synchronize()
    setup_ms = ev0.elapsed_time(ev1)
    row_lo, n_owned = 3 * mesh.lo, 3 * mesh.n_owned
    own = slice(row_lo, row_lo + n_owned)
    vals = torch.empty(pattern.nnz, dtype=torch.float64, device=dev)
    rhs_buf = torch.empty(pattern.n_dofs, dtype=torch.float64, device=dev)
    state = {}
    fused = args.dist_cg == "fused"
    cg = D.FusedCG(pattern.indptr, pattern.indices, pattern.n_dofs, row_lo, n_owned, plan, dev) if fused else None

    def build_system(nodes, elements, is_con, disp, E_mod, nu):
        # per-element material -> tangent on the device, as the reference's vectorised material does
        C = IsotropicElasticity3D(E_mod, nu).C
        k = csr.integrate_k(L.KIND_MECH, bref, w, nodes, elements, C, check=False)
        del C
        rhs = rhs_buf   # local halo values of du_bc come from the BC data, so the lifting needs no exchange
        csr.assemble(pattern, k, is_con, out=vals, ubc=disp, lift=rhs)
        del k
        A = pattern.matrix(vals)
        return A, rhs, csr.JacobiPreconditioner(A)

    def solve(A, rhs, M, tol):
        if fused:
            return cg.solve(A, M.dinv, rhs, rtol=tol)
        return D.distributed_cg(A, M.dinv, rhs, row_lo, n_owned, halo, rtol=tol)

    def hot_path(*inputs):
        A, rhs, M = build_system(*inputs)
        x, info = solve(A, rhs, M, rtol)
        state.update(A=A, rhs=rhs, x=x, info=info, M=M)
        return x

    def step():
        return hot_path(nodes, elements, is_con, disp, E_mod, nu)

    x_pinned = torch.empty(n_owned, dtype=torch.float64).pin_memory()

    def step_e2e():
        d = [t.to(dev, non_blocking=True) for t in host]
        x = hot_path(*d)
        x_pinned.copy_(x[own], non_blocking=True)   # stream-ordered: inside the timed region
        return x_pinned

    def timed_max(fn, reps):
        """mean device ms of `reps` calls, bracketed by barrier + synchronize, max over ranks"""
        _barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        _barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        _allreduce(t, dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ms = timed_max(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # end to end: every rank's inputs come from pinned host memory, its part of the solution goes back
    step_e2e()
    ms_e2e = timed_max(step_e2e, max(1, min(args.steps, 2)))
    io = torch.tensor([sum(t_.numel() * t_.element_size() for t_ in host), 8 * n_owned], dtype=torch.int64, device=dev)
    _allreduce(io)

    # true global relative residual of the last solve
    A, rhs, x, info, M = state["A"], state["rhs"], state["x"], state["info"], state["M"]
    iterations = info["iterations"]

    def global_res(xv):
        xv = xv.clone()
        halo(xv)
        r = (rhs - A.matvec(xv))[own]
        num = torch.stack([(r * r).sum(), (rhs[own] ** 2).sum()])
        _allreduce(num)
        return float((num[0] / num[1]).sqrt())

    true_res = global_res(x)
    nnz_owned = torch.tensor([int(pattern.indptr[row_lo + n_owned] - pattern.indptr[row_lo])], device=dev)
    nnz_mine = int(nnz_owned.item())
    _allreduce(nnz_owned)

    # phases on this rank (max over ranks): the SOLVE alone gives the per-iteration figure (same key at N = 1)
    solve_ms = timed_max(lambda: solve(A, rhs, M, rtol), 1)
    build_ms = timed_max(lambda: build_system(nodes, elements, is_con, disp, E_mod, nu), 1)
    A, rhs, M = build_system(nodes, elements, is_con, disp, E_mod, nu)

    # the dominant kernel, k_dcg_spmv (owned rows, halo wait inside), timed by CUDA events INSIDE the solve on the
    # solve's stream (tfem_comm_set_trace time_spmv): mean of the 32 launches of the second batch; max over ranks
    spmv_ms, trace_summary = None, None
    if fused:
        n_tr = 256 if args.trace else 0
        tr = cg.comm.set_trace(n_tr, first_iteration=64, time_spmv=True)
        _, info_t = solve(A, rhs, M, rtol)
        cg.comm.set_trace(0)
        t = torch.tensor([info_t["spmv_ms"]], dtype=torch.float64, device=dev)
        _allreduce(t, dist.ReduceOp.MAX)
        spmv_ms = float(t.item())
        if n_tr:
            trace_summary = summarize_trace(tr.cpu().numpy(), rank)
            gathered = [None] * world
            if multi:
                dist.all_gather_object(gathered, trace_summary)
            else:
                gathered = [trace_summary]
            trace_summary = gathered
    else:
        xs = torch.randn(A.n, dtype=torch.float64, device=dev)
        ys = torch.empty_like(xs)
        spmv_ms = timed_max(lambda: A.matvec(xs, out=ys, fmt="sell"), 10)
    spmv_bytes = 12 * nnz_mine + 20 * n_owned      # §8(d) algorithmic bytes of the rows this rank multiplies
    sell_padding = float(A._sell_struct.padded) / float(pattern.nnz) - 1.0 if A._sell_struct is not None else None

    # distributed AMG-PCG beside the Jacobi headline (like the N = 1 line)
    amg_info = None
    if not args.no_amg and multi_amg_available():
        try:
            amg_info = bench_distributed_amg(args, D, csr, pattern, A, rhs, M, mesh, plan, ranges, row_lo, n_owned, x,
                                             solve, global_res, timed_max, n_dofs_global, build_ms)
        except RuntimeError as exc:      # reported beside the headline, never instead of it
            amg_info = {"error": str(exc)[:300]}

    probe = None
    if config_c or args.probe:
        # partition-independent fingerprint of the solution (compare runs at different N: tools/compare_probe.py)
        u = torch.where(is_con.bool(), disp, -x)[own]
        sums = torch.stack([u.sum(), (u * u).sum(), u.abs().max()])
        mx = sums[2:].clone()
        _allreduce(sums)
        _allreduce(mx, dist.ReduceOp.MAX)
        xyz = nodes[mesh.lo:mesh.lo + mesh.n_owned]
        wgt = torch.stack([torch.sin(3.0 * xyz[:, 0] + 1.0), torch.cos(2.0 * xyz[:, 1]) * xyz[:, 2], xyz[:, 0] * xyz[:, 1]], 1)
        mom = (u.view(-1, 3) * wgt).sum(0)
        _allreduce(mom)
        probe = {"sum_u": float(sums[0]), "norm2_u": float(sums[1].sqrt()), "max_abs_u": float(mx[0]),
                 "weighted_moments": [float(v) for v in mom], "rtol": rtol}

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": n_dofs_global / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if config_c else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "n_dofs": int(n_dofs_global), "n_elem": int(n_elem_global),
                       "nnz": int(nnz_owned.item()), "rtol": rtol,
                       "cg_iterations": iterations, "true_rel_residual": true_res,
                       "per_rank_local_dofs": int(A.n), "halo_bytes_per_exchange": plan.bytes_per_exchange(),
                       "sell32_padding_overhead_rank0": sell_padding,
                       "collectives_per_iteration": ("fused into the kernels: halo = peer stores of the direction "
                                                     "update, 2 all-reduces = LL-protocol peer stores (tfem_dcg_solve)")
                       if fused else "NCCL: 1 halo exchange (P2P send/recv) + 2 all-reduces (1 and 2 doubles)",
                       "l2_policy": f"inputs larger than L2 (per-rank SELL matrix {8.5e-9 * pattern.nnz:.1f} GB vs 126 MB L2)",
                       "setup_ms_pattern": setup_ms,
                       "phases_ms": {"integrate_assemble_jacobi": build_ms, "pcg_solve": solve_ms},
                       "per_iteration_ms": solve_ms / max(1, iterations),   # solve only, max over ranks
                       "note": "Jacobi-PCG iterations grow with the cube edge (~N^(1/3)), so DOF/s per GPU "
                               "falls with N even at perfect per-iteration scaling; per_iteration_ms (solve only) is "
                               "the kernel/communication scaling figure",
                       "amg_pcg": amg_info, "solution_probe": probe, "wait_trace": trace_summary},
            "e2e": {"value": n_dofs_global / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
                    "call": "C-ABI ops on per-rank host buffers (H2D mesh + per-element E, nu + BCs -> tangent -> "
                            "integrate -> assemble -> distributed PCG -> D2H owned u)"},
            "gpu_launches": int(info["launches"] + 6),
            "roofline": {"bound": "hbm",
                         "kernel": "k_dcg_spmv<3> (this rank's owned rows, halo wait inside; CUDA events inside the solve, "
                                   "max over ranks)" if fused else "k_sell_spmv (per rank, standalone)",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes": int(spmv_bytes), "ms_per_launch": spmv_ms, "traffic": None},
            "cpu_baseline": None, "clocks": clocks,
        }
        print(json.dumps(line))
    if cg is not None:
        cg.close()
    _barrier()
    if multi:
        dist.destroy_process_group()


def summarize_trace(tr, rank):
    """Median durations (us) between the trace points of tfem_comm_set_trace over the traced iterations."""
    tr = tr.astype(np.float64)
    ok = (tr[:, 0] > 0) & (tr[:, 9] > 0)
    tr = tr[ok]
    if len(tr) < 2:
        return {"rank": rank, "iterations": int(len(tr))}
    med = lambda v: float(np.median(v)) / 1e3  # noqa: E731
    return {
        "rank": rank, "iterations": int(len(tr)),
        "spmv_kernel_us": med(tr[:, 2] - tr[:, 0]), "spmv_longest_halo_wait_us": med(tr[:, 1]),
        "gap_spmv_to_update_us": med(tr[:, 3] - tr[:, 2]), "update_wait_allreduce_us": med(tr[:, 4] - tr[:, 3]),
        "update_stream_us": med(tr[:, 5] - tr[:, 4]), "gap_update_to_direction_us": med(tr[:, 6] - tr[:, 5]),
        "direction_wait_allreduce_us": med(tr[:, 7] - tr[:, 6]), "direction_halo_released_after_us": med(tr[:, 8] - tr[:, 7]),
        "direction_stream_us": med(tr[:, 9] - tr[:, 7]),
        "gap_direction_to_next_spmv_us": med(tr[1:, 0] - tr[:-1, 9]),
        "iteration_us": med(tr[1:, 0] - tr[:-1, 0]),
        # the host polls the convergence flag every 32 iterations: the gaps there are the outliers of the distribution
        "iteration_us_mean": float(np.mean(tr[1:, 0] - tr[:-1, 0])) / 1e3,
        "iteration_us_max": float(np.max(tr[1:, 0] - tr[:-1, 0])) / 1e3,
        "gap_direction_to_next_spmv_us_max": float(np.max(tr[1:, 0] - tr[:-1, 9])) / 1e3,
        "gaps_over_20us": int(np.sum((tr[1:, 0] - tr[:-1, 9]) > 20e3)),
    }


def multi_amg_available():
    try:
        from torchfem_b200 import damg  # noqa: F401
        return True
    except ImportError:
        return False


def bench_distributed_amg(args, D, csr, pattern, A, rhs, M, mesh, plan, ranges, row_lo, n_owned, x_jacobi, solve,
                          global_res, timed_max, n_dofs_global, build_ms):
    import torch
    import torch.distributed as dist

    from torchfem_b200 import damg

    own = slice(row_lo, row_lo + n_owned)
    state = {}

    node_plan = D.build_halo_plan(mesh, ranges, dist.get_rank() if dist.is_initialized() else 0, 1)

    def setup():
        state["H"] = damg.DistributedAMG(A, mesh.lo, mesh.n_owned, mesh.global_nodes, node_plan)

    def run(tol):
        xa, st = state["H"].solve(rhs, rtol=tol)
        state.update(x=xa, st=st)

    setup()
    run(args.rtol)                   # first call grows pools
    state["H"].close()
    state.pop("H")
    torch.cuda.synchronize()
    setup_ms = timed_max(setup, 1)
    solve_ms = timed_max(lambda: run(args.rtol), 1)
    st, xa = state["st"], state["x"]
    res = global_res(xa)
    # parity at stol 1e-12: distributed AMG-PCG vs distributed Jacobi-PCG
    xj12, _ = solve(A, rhs, M, 1e-12)
    run(1e-12)
    num = torch.stack([((state["x"] - xj12)[own] ** 2).sum(), (xj12[own] ** 2).sum()])
    if dist.is_initialized():
        dist.all_reduce(num)
    out = {"call": "damg.DistributedAMG: rank-local aggregates, distributed Galerkin levels, V(1,1) + CG over peer memory",
           "setup_ms": setup_ms, "solve_ms": solve_ms, "iterations": st["iterations"],
           "ms_per_iteration": solve_ms / max(1, st["iterations"]), "true_rel_residual": res,
           "levels": state["H"].level_sizes, "setup_phases_ms_rank0": state["H"]._timing,
           "rel_diff_vs_jacobi_pcg_stol_1e-12": float((num[0] / num[1]).sqrt()),
           "iterations_stol_1e-12": state["st"]["iterations"],
           "dofs_per_s_integrate_assemble_setup_solve": n_dofs_global / ((build_ms + setup_ms + solve_ms) * 1e-3)}
    state["H"].close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edge", type=int, default=None,
                    help="elements per cube edge (default 150 = BASELINE configs[1]; 120 for --config C)")
    ap.add_argument("--config", default="B", choices=["B", "C"],
                    help="B: Hexa1 cube, weak scaling over GPUs (default, the metric's config). C: Hexa2 cube of "
                         "BASELINE configs[2], one global mesh partitioned over the GPUs (strong scaling)")
    ap.add_argument("--cpu-edge", type=int, default=40, help="elements per edge of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the Solid.solve end-to-end leg")
    ap.add_argument("--no-amg", action="store_true", help="skip the AMG-PCG report beside the Jacobi-PCG headline")
    ap.add_argument("--dist-cg", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused peer-to-peer CG (tfem_dcg_solve) or the host-driven NCCL variant")
    ap.add_argument("--rtol", type=float, default=RTOL, help="N>1 / config C only: relative residual of the PCG "
                    "(the metric's value is 1e-8; other values are for the parity legs)")
    ap.add_argument("--trace", action="store_true", help="N>1: in-kernel %%globaltimer profile of the cross-GPU waits")
    ap.add_argument("--dist-path", action="store_true", help="run the N>1 code path (slab + tfem_dcg_solve) even at N=1: "
                    "the single-rank reference point of the wait trace")
    ap.add_argument("--probe", action="store_true", help="N>1: add the partition-independent solution fingerprint")
    args = ap.parse_args()
    if args.edge is None:
        args.edge = 120 if args.config == "C" else 150
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
