"""Multi-GPU parity check, launched by torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py --edge 24

Solves the cube-extension problem (N slabs) with (a) the fused peer-to-peer CG (`tfem_dcg_solve`), (b) the
host-driven NCCL CG (`distributed_cg`), and on rank 0 (c) the whole problem on one GPU with the single-GPU
driver; asserts that all three agree (<= 1e-9 relative, equal iteration counts +-1) and that the fused path
is bitwise reproducible run to run. Prints one JSON line on rank 0.
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
    ap.add_argument("--edge", type=int, default=24)
    ap.add_argument("--general", action="store_true", help="force index-list halos instead of contiguous ranges")
    ap.add_argument("--hexa2", action="store_true",
                    help="20-node hexahedra from linear_to_quadratic, x-coordinate partition (config C in small)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200 import _lib as L, csr, distributed as D
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

    def system(nodes_h, elements_h):
        nodes, elements = nodes_h.to(dev), elements_h.to(dev)
        con = torch.zeros(nodes_h.shape[0], 3, dtype=torch.bool)
        disp = torch.zeros(nodes_h.shape[0], 3)
        con[nodes_h[:, 0] == 0.0, :] = True
        right = (nodes_h[:, 0] - Ex * h).abs() < 1e-12
        con[right, 0] = True
        disp[right, 0] = 0.1
        is_con, disp = con.ravel().to(torch.uint8).to(dev), disp.ravel().to(dev)
        p = csr.Pattern(elements, nodes_h.shape[0], 3)
        Cd = C1.expand(len(elements_h), 3, 3, 3, 3).contiguous().to(dev)
        k = csr.integrate_k(L.KIND_MECH, bref, w, nodes, elements, Cd)
        A = p.matrix(csr.assemble(p, k, is_con))
        rhs = p.matrix(csr.assemble(p, k, None)).matvec(disp * is_con)
        rhs.masked_fill_(is_con.bool(), 0.0)
        return p, A, rhs

    perm = None
    if a.hexa2:
        with torch.device("cpu"):
            nodes_g, elements_g = linear_to_quadratic(*cube_hexa(Ex + 1, E + 1, E + 1, Ex * h, 1.0, 1.0))
        nodes_h, mesh, ranges, perm = D.coordinate_partition(nodes_g, elements_g, world, rank)
    else:
        nodes_h, mesh, ranges, dims = D.cube_slab(Ex, E, E, h, world, rank)
    plan = D.build_halo_plan(mesh, ranges, rank, 3)
    if a.general:
        plan.contiguous.clear()
    p, A, rhs = system(nodes_h, mesh.elements)
    row_lo, n_owned = 3 * mesh.lo, 3 * mesh.n_owned
    M = csr.JacobiPreconditioner(A)
    halo = D.HaloExchanger(plan, dev)
    own = slice(row_lo, row_lo + n_owned)

    def true_res(x):
        x = x.clone()
        halo(x)
        r = (rhs - A.matvec(x))[own]
        num = torch.stack([(r * r).sum(), (rhs[own] ** 2).sum()])
        dist.all_reduce(num)
        return float((num[0] / num[1]).sqrt())

    x_n, info_n = D.distributed_cg(A, M.dinv, rhs, row_lo, n_owned, halo, rtol=1e-10, maxiter=20000)
    if rank == 0:
        print("nccl", info_n, "true_res", true_res(x_n), flush=True)
    else:
        true_res(x_n)
    cg = D.FusedCG(p.indptr, p.indices, A.n, row_lo, n_owned, plan, dev)
    if os.environ.get("TFEM_DCG_NO_INTERIOR"):
        cg.interior = (row_lo, row_lo)
    for tol in (1e-6, 1e-8, 1e-10):
        try:
            x_t, info_t = cg.solve(A, M.dinv, rhs, rtol=tol, maxiter=20000)
            tr = true_res(x_t)
            if rank == 0:
                print("fused", tol, info_t, "true_res", tr, flush=True)
        except RuntimeError as exc:
            if rank == 0:
                print("fused", tol, "FAILED", exc, flush=True)
    x_f, info_f = cg.solve(A, M.dinv, rhs, rtol=1e-10, maxiter=20000)
    x_f2, info_f2 = cg.solve(A, M.dinv, rhs, rtol=1e-10, maxiter=20000)
    d_fn = float((x_f[own] - x_n[own]).abs().max())
    nrm = float(x_n[own].abs().max())
    repro = bool(torch.equal(x_f[own], x_f2[own]))

    # gather owned parts on rank 0 and compare with a one-GPU solve of the whole problem
    longest = max(3 * (r1 - r0) for r0, r1 in ranges)
    mine = torch.zeros(longest, device=dev)
    mine[:n_owned] = x_f[own]
    padded = [torch.empty(longest, device=dev) for _ in ranges] if rank == 0 else None
    dist.gather(mine, padded, dst=0)
    parts = [padded[i][:3 * (r1 - r0)] for i, (r0, r1) in enumerate(ranges)] if rank == 0 else None
    ok, line = True, None
    if rank == 0:
        if not a.hexa2:
            nodes_g, elements_g = cube_hexa(Ex + 1, E + 1, E + 1, Ex * h, 1.0, 1.0)
        pg, Ag, rhs_g = system(nodes_g.cpu(), elements_g.cpu())
        xg, _, info_g = csr.krylov_solve(Ag, rhs_g, method="cg", rtol=1e-10)
        xf = torch.cat(parts)
        if perm is not None:  # parts are in the partition's numbering: back to the mesh's own
            xo = torch.empty_like(xf)
            xo.view(-1, 3)[perm.to(dev)] = xf.view(-1, 3)
            xf = xo
        d_g = float((xf - xg).abs().max() / xg.abs().max())
        line = {"world": world, "edge": E, "n_dofs": int(xg.numel()), "general_halo": a.general, "hexa2": a.hexa2, "neighbours": plan.neighbours,
                "iters_fused": info_f["iterations"], "iters_nccl": info_n["iterations"],
                "iters_single": info_g["iterations"], "rel_fused_vs_nccl": d_fn / nrm,
                "rel_fused_vs_single": d_g, "bitwise_reproducible": repro, "interior": list(cg.interior),
                "row_range": [row_lo, row_lo + n_owned], "halo_bytes": cg.halo_bytes}
        ok = (d_fn / nrm <= 1e-9 and d_g <= 1e-9 and repro
              and abs(info_f["iterations"] - info_g["iterations"]) <= 1
              and abs(info_f["iterations"] - info_n["iterations"]) <= 1)
        line["ok"] = ok
        print(json.dumps(line), flush=True)
    flag = torch.tensor([1 if (ok and repro and d_fn / max(nrm, 1e-300) <= 1e-9) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    cg.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
