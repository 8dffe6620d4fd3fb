"""Distributed AMG-PCG parity check, launched by torchrun on N GPUs of one box (or plainly with python: world 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/damg_check.py --edge 16

Solves the cube-extension problem (N slabs) with (a) the distributed AMG-PCG (`damg.DistributedAMG`,
`tfem_damg_pcg_solve`) — once with the default gather limit and once with a tiny one, which forces a SECOND
distributed level —, (b) the fused peer-memory Jacobi-PCG and, on rank 0, (c) the single-GPU AMG-PCG on the whole
problem. Asserts: (a) == (b) <= 1e-8 relative at stol 1e-11, true residual <= stol, iterations of (a) within +30 % + 2
of (c), bitwise identical re-run. Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=16)
    ap.add_argument("--hexa2", action="store_true")
    ap.add_argument("--rtol", type=float, default=1e-11)
    a = ap.parse_args()
    multi = "RANK" in os.environ
    rank, world = (int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])) if multi else (0, 1)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    if multi:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_default_dtype(torch.float64)
    from torchfem_b200 import _lib as L, csr, damg, distributed as D
    from torchfem_b200.amg import AMGPreconditioner
    from torchfem_b200.elements import Hexa1, Hexa2, linear_to_quadratic
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    E = a.edge
    Ex = E * world
    h = 1.0 / E
    ET = Hexa2 if a.hexa2 else Hexa1
    bref = ET.B(ET.ipoints.to(torch.float64).cpu())
    w = ET.iweights.to(torch.float64).cpu()
    C1 = IsotropicElasticity3D(1000.0, 0.3).C.to(torch.float64).cpu()

    patterns = {}

    def system(nodes_h, elements_h, heterogeneous=False):
        nodes, elements = nodes_h.to(dev), elements_h.to(dev)
        con = torch.zeros(nodes_h.shape[0], 3, dtype=torch.bool)
        disp = torch.zeros(nodes_h.shape[0], 3)
        con[nodes_h[:, 0] == 0.0, :] = True
        right = (nodes_h[:, 0] - Ex * h).abs() < 1e-12
        con[right, 0] = True
        disp[right, 0] = 0.1
        is_con, disp = con.ravel().to(torch.uint8).to(dev), disp.ravel().to(dev)
        key = (nodes_h.shape[0], len(elements_h))
        if key not in patterns:
            patterns[key] = csr.Pattern(elements, nodes_h.shape[0], 3)
        p = patterns[key]
        Cd = C1.expand(len(elements_h), 3, 3, 3, 3).contiguous().to(dev)
        if heterogeneous:   # a second set of coefficients on the same pattern (stiffness varying with the element centre)
            centre = nodes[elements].mean(dim=1)
            Cd = Cd * (1.0 + 0.8 * torch.sin(7.0 * centre[:, 0] + 3.0 * centre[:, 1]) * torch.cos(5.0 * centre[:, 2]))[:, None, None, None, None]
        k = csr.integrate_k(L.KIND_MECH, bref, w, nodes, elements, Cd)
        rhs = torch.empty(p.n_dofs, device=dev)
        A = p.matrix(csr.assemble(p, k, is_con, ubc=disp, lift=rhs))
        return p, A, rhs

    perm = None
    if a.hexa2:
        with torch.device("cpu"):
            nodes_g, elements_g = linear_to_quadratic(*cube_hexa(Ex + 1, E + 1, E + 1, Ex * h, 1.0, 1.0))
        nodes_h, mesh, ranges, perm = D.coordinate_partition(nodes_g, elements_g, world, rank)
    else:
        nodes_h, mesh, ranges, dims = D.cube_slab(Ex, E, E, h, world, rank)
    plan = D.build_halo_plan(mesh, ranges, rank, 3)
    node_plan = D.build_halo_plan(mesh, ranges, rank, 1)
    p, A, rhs = system(nodes_h, mesh.elements)
    row_lo, n_owned = 3 * mesh.lo, 3 * mesh.n_owned
    own = slice(row_lo, row_lo + n_owned)
    M = csr.JacobiPreconditioner(A)
    halo = D.HaloExchanger(plan, dev)

    def allsum(t):
        if multi:
            dist.all_reduce(t)
        return t

    def true_res(x):
        x = x.clone()
        halo(x)
        r = (rhs - A.matvec(x))[own]
        num = allsum(torch.stack([(r * r).sum(), (rhs[own] ** 2).sum()]))
        return float((num[0] / num[1]).sqrt())

    def rel_diff(x, y):
        num = allsum(torch.stack([((x - y)[own] ** 2).sum(), (y[own] ** 2).sum()]))
        return float((num[0] / num[1]).sqrt())

    cg = D.FusedCG(p.indptr, p.indices, A.n, row_lo, n_owned, plan, dev)
    x_j, info_j = cg.solve(A, M.dinv, rhs, rtol=a.rtol, maxiter=50000)
    out = {"world": world, "edge": E, "hexa2": a.hexa2, "n_dofs_global": int(allsum(torch.tensor([n_owned], device=dev)).item()),
           "iters_jacobi": info_j["iterations"], "cases": []}
    ok = True
    for gather_max in (400_000, 600):
        H = damg.DistributedAMG(A, mesh.lo, mesh.n_owned, mesh.global_nodes, node_plan, gather_max=gather_max)
        x_a, st = H.solve(rhs, rtol=a.rtol)
        x_b, st2 = H.solve(rhs, rtol=a.rtol)
        case = {"gather_max": gather_max, "distributed_levels": len(H.levels), "levels": H.level_sizes,
                "setup_phases_ms": H._timing,
                "iterations": st["iterations"], "true_rel_residual": true_res(x_a),
                "rel_diff_vs_jacobi_pcg": rel_diff(x_a, x_j), "bitwise_reproducible": bool(torch.equal(x_a[own], x_b[own])),
                "launches": st["launches"]}
        out["cases"].append(case)
        ok = ok and case["rel_diff_vs_jacobi_pcg"] <= 1e-8 and case["true_rel_residual"] <= 10 * a.rtol and case["bitwise_reproducible"]
        # resetup: new coefficients on the stored hierarchy (AmgX life-cycle, reference sparse.py:440-441) must give
        # what a hierarchy built from scratch for them gives
        _, A2, rhs2 = system(nodes_h, mesh.elements, heterogeneous=True)
        M2 = csr.JacobiPreconditioner(A2)
        x_j2, _ = cg.solve(A2, M2.dinv, rhs2, rtol=a.rtol, maxiter=50000)
        H.resetup(A2)
        x_r, st_r = H.solve(rhs2, rtol=a.rtol)
        H2 = damg.DistributedAMG(A2, mesh.lo, mesh.n_owned, mesh.global_nodes, node_plan, gather_max=gather_max)
        x_f, st_f = H2.solve(rhs2, rtol=a.rtol)
        case["resetup"] = {"iterations": st_r["iterations"], "iterations_fresh_setup": st_f["iterations"],
                           "rel_diff_vs_jacobi_pcg": rel_diff(x_r, x_j2), "equals_fresh_setup_bitwise": bool(torch.equal(x_r[own], x_f[own]))}
        ok = ok and case["resetup"]["rel_diff_vs_jacobi_pcg"] <= 1e-8 and abs(st_r["iterations"] - st_f["iterations"]) <= 1
        H2.close()
        H.close()
    # single-GPU hierarchy on the whole problem (rank 0)
    if rank == 0:
        if a.hexa2:
            with torch.device("cpu"):
                nodes_g, elements_g = linear_to_quadratic(*cube_hexa(Ex + 1, E + 1, E + 1, Ex * h, 1.0, 1.0))
        else:
            with torch.device("cpu"):
                nodes_g, elements_g = cube_hexa(Ex + 1, E + 1, E + 1, Ex * h, 1.0, 1.0)
        pg, Ag, rhs_g = system(nodes_g, elements_g)
        Mg = AMGPreconditioner(Ag)
        xg, sg = Mg.solve(rhs_g, rtol=a.rtol)
        out["iters_single_gpu_amg"] = sg["iterations"]
        out["levels_single_gpu_amg"] = [int(lv.n) for lv in Mg.levels]
        for case in out["cases"]:
            ok = ok and case["iterations"] <= 1.3 * sg["iterations"] + 2
    out["ok"] = bool(ok)
    flag = torch.tensor([1 if ok else 0], device=dev)
    if multi:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["ok"] = bool(int(flag.item()) == 1)
        print(json.dumps(out), flush=True)
    cg.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
