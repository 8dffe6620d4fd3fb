#!/usr/bin/env python
"""bench.py — DOFs solved/sec (integrate + assemble + Jacobi-PCG to 1e-8) on the linear-elastic Hexa1 cube.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--edge E]

One "step" = one pass of the hot path over the synthetic cube-extension problem of the reference's
benchmarks/cubes.py (cube_hexa(E+1,E+1,E+1), E=1000, nu=0.3, x=0 clamped, u_x=0.1 at x=1):
element matrices (K1) -> deterministic assembly with Dirichlet masking (K2/K3) -> right-hand side ->
Jacobi setup (K4) -> PCG to ||r|| <= 1e-8 ||b|| (K5/K6). The sparsity pattern (K0) is setup and is reported
separately, as the reference does (benchmarks/run.py:56-58: "setup" vs "fwd").

N=1 runs BASELINE.json configs[1]: 150^3 elements, 10,328,853 DOFs, nnz 825,604,659 (matrix 10.1 GB >> L2).
`--impl reference` times the CPU oracle port of the same path (oracle/fem_oracle.py) on a bounded sample.
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
NCU_SPMV_DRAM_BYTES_CONFIG_B = 7233556000 + 85765120


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
def run_reference(args):
    """CPU baseline: the oracle port of the reference path (numpy einsum + scipy-style Jacobi-CG), all
    host threads numpy/BLAS will use, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fem_oracle as O

    E = args.cpu_edge
    N = E + 1
    nodes, elements = O.cube_hexa(N, N, N)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con_mask, disp = O.cube_extension_bcs(nodes)
    n_dofs = nodes.size
    times = []
    its = 0
    for s in range(args.warmup + args.steps):
        out = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=RTOL)
        t = out["t_integrate"] + out["t_assemble"] + out["t_rhs"] + out["t_solve"]
        its = out["iterations"]
        if s >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = n_dofs / (ms / 1e3)
    cores = os.cpu_count()
    sample = (f"cube_hexa({N},{N},{N}) = {E}^3 Hexa1 elements, {n_dofs} DOFs, {its} Jacobi-CG iterations to "
              f"1e-8; oracle port (numpy einsum integrate, bincount assemble, scipy-CSR SpMV CG); "
              f"os.cpu_count()={cores}, CG SpMV is single-threaded as in scipy")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"linear-elastic Hexa1 cube {E}^3 elements (bounded CPU sample of configs[1])",
                   "n_dofs": int(n_dofs), "rtol": RTOL, "cg_iterations": int(its)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
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
    from oracle import fem_oracle as O  # tables only (bref/w are 200 numbers) + cpu_baseline leg

    if world > 1 or args.config == "C":
        from torchfem_b200 import distributed as D
        return D.bench_multi_gpu(args, METRIC, UNIT, RTOL)

    E = args.edge
    t0 = time.perf_counter()
    nodes_h, elements_h, con_h, disp_h = build_problem(T, torch, E, device)
    n_elem = elements_h.shape[0]
    n_dofs = nodes_h.numel()
    bref_np, w_np = O.hexa1_tables()
    bref, w = torch.as_tensor(bref_np), torch.as_tensor(w_np)
    from torchfem_b200.materials import IsotropicElasticity3D

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
    vals = torch.empty(nnz, dtype=torch.float64, device=device)
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
        csr.assemble(pattern, k, is_con, out=vals, ubc=disp, lift=rhs)
        del k
        A = pattern.matrix(vals)
        x, M, info = csr.krylov_solve(A, rhs, method="cg", rtol=RTOL)
        u = torch.where(is_con.bool(), disp, -x)
        state.update(info=info, A=A, rhs=rhs, x=x)
        return u

    def step_resident():
        return hot_path(nodes, elements, E_mod, nu, is_con, disp)

    def step_e2e():
        d = [t.to(device, non_blocking=True) for t in host]
        u = hot_path(*d)
        return u.cpu()

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
    asm_ms = ev_time(lambda: csr.assemble(pattern, kk, is_con, out=vals), 3)
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
        del Mp, xa

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

        def step_api():
            model.material = IsotropicElasticity3D(E_h.to(device, non_blocking=True), nu_h.to(device, non_blocking=True))
            model.constraints = con_bool_h.to(device, non_blocking=True)
            model.displacements = disp2_h.to(device, non_blocking=True)
            u, *_ = model.solve(method="cg", stol=RTOL, rtol=1e-6)
            return u.cpu()

        ms_api, u_api = timed(step_api, max(1, min(args.steps, 2)), 1)
        api_amg = None
        if not args.no_amg:
            def step_api_amg():
                model.material = IsotropicElasticity3D(E_h.to(device, non_blocking=True), nu_h.to(device, non_blocking=True))
                model.constraints = con_bool_h.to(device, non_blocking=True)
                model.displacements = disp2_h.to(device, non_blocking=True)
                u, *_ = model.solve(method="amgx", stol=RTOL, rtol=1e-6)
                return u.cpu()

            ms_api_amg, u_amg = timed(step_api_amg, max(1, min(args.steps, 2)), 1)
            api_amg = {"value": n_dofs / (ms_api_amg / 1e3), "unit": UNIT, "ms_per_step": ms_api_amg,
                       "call": "Solid.solve(method='amgx', stol=1e-8)",
                       "rel_diff_vs_kernel_path": float((u_amg.ravel() - u_h.ravel()).norm() / u_h.norm())}
        err = float((u_api.ravel() - u_h.ravel()).norm() / u_h.norm())
        api = {"value": n_dofs / (ms_api / 1e3), "unit": UNIT, "ms_per_step": ms_api,
               "h2d_bytes_per_step": int(2 * E_h.numel() * 8 + con_bool_h.numel() + disp2_h.numel() * 8),
               "d2h_bytes_per_step": int(d2h), "call": "Solid.solve(method='cg', stol=1e-8)",
               "rel_diff_vs_kernel_path": err, "with_amg": api_amg}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 / N=1 only
    cpu = None
    if not args.no_cpu_baseline:
        Ec = args.cpu_edge
        Nc = Ec + 1
        nd_c, el_c = O.cube_hexa(Nc, Nc, Nc)
        Cc = O.isotropic_C3d(1000.0, 0.3, len(el_c))
        cm, dp = O.cube_extension_bcs(nd_c)
        out = O.linear_solve_reference_flow(nd_c, el_c, bref_np, w_np, Cc, cm, dp, rtol=RTOL)
        t_cpu = out["t_integrate"] + out["t_assemble"] + out["t_rhs"] + out["t_solve"]
        cpu = {"value": nd_c.size / t_cpu, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": (f"{Ec}^3 Hexa1 elements, {nd_c.size} DOFs, {out['iterations']} CG its to 1e-8: "
                          f"setup {out['t_setup']:.2f}s (excluded) integrate {out['t_integrate']:.2f}s "
                          f"assemble {out['t_assemble']:.2f}s solve {out['t_solve']:.2f}s; numpy/scipy oracle "
                          f"port of the reference CPU path, scipy-CSR SpMV single-threaded")}

    launches = info["launches"] + 1 + 1 + 1 + 1  # integrate, assemble(+lifting), SELL fill, jacobi
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
                                                                 "per_cg_iteration": solve_ms / max(1, info2["iterations"])},
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
                     # `ncu --set full` capture (profiles/r1_t_cg_iteration_ncu.txt); other sizes: not captured
                     "traffic": NCU_SPMV_DRAM_BYTES_CONFIG_B if E == 150 else None,
                     "traffic_source": "profiles/r1_t_cg_iteration_ncu.txt (k_sell_spmv<3,1>: 7.234 GB read + 0.086 GB written per launch)"},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))


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
    args = ap.parse_args()
    if args.edge is None:
        args.edge = 120 if args.config == "C" else 150
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
