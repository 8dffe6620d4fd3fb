"""Host-side logic of the multi-GPU path on CPU, world_size 2 over gloo: node-block partition, local
renumbering, halo plans, the halo exchange itself, and the protocol of the distributed CG (halo exchange ->
SpMV -> all-reduce -> update -> all-reduce) with the oracle doing the per-rank arithmetic in numpy.
The CUDA kernels are exercised by the `-m gpu` tests and by `bench.py --gpus N`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fem_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem(N):
    nodes, elements = O.cube_hexa(2 * (N - 1) + 1, N, N, 2.0, 1.0, 1.0)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con_mask = np.zeros((nodes.shape[0], 3), dtype=bool)
    disp = np.zeros((nodes.shape[0], 3))
    con_mask[nodes[:, 0] == 0.0, :] = True
    con_mask[nodes[:, 0] == 2.0, 0] = True
    disp[nodes[:, 0] == 2.0, 0] = 0.1
    return nodes, elements, bref, w, C, con_mask, disp


def _worker(rank, world, port, N, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    from torchfem_b200 import distributed as D

    nodes, elements, bref, w, C, con_mask, disp = _global_problem(N)
    n_nod = nodes.shape[0]
    plane = N * N
    ranges = D.node_ranges(n_nod, world, granule=plane)
    n0, n1 = ranges[rank]
    mesh = D.local_mesh(torch.as_tensor(elements), n0, n1)
    plan = D.build_halo_plan(mesh, ranges, rank, 3)
    halo = D.HaloExchanger(plan, "cpu")
    gl = mesh.global_nodes.numpy()

    # slab generator == general partition
    E = N - 1
    nodes_s, mesh_s, ranges_s, dims = D.cube_slab(2 * E, E, E, 1.0 / E, world, rank)
    assert ranges_s == ranges and dims == (2 * E + 1, N, N)
    assert torch.equal(mesh_s.global_nodes, mesh.global_nodes) and torch.equal(mesh_s.elements, mesh.elements)
    assert mesh_s.lo == mesh.lo and mesh_s.n_owned == mesh.n_owned
    assert np.allclose(nodes_s.numpy(), nodes[gl], atol=1e-14)
    assert set(plan.contiguous) == set(plan.neighbours)  # slabs exchange contiguous plane ranges

    # local system with the oracle: owned rows must equal the global rows
    el_l = mesh.elements.numpy()
    idx = O.dof_map(el_l, 3)
    n_loc = 3 * len(gl)
    glob_idx, k_map, diag_map = O.pattern(idx, n_loc)
    k = O.integrate_k_mech(nodes[gl], el_l, bref, w, C[: len(el_l)])
    con_l = con_mask[gl].ravel()
    A = O.to_csr(O.assemble_values(k, k_map, glob_idx, diag_map, np.nonzero(con_l)[0], n_loc), glob_idx, n_loc)
    A_free = O.to_csr(O.assemble_values(k, k_map, glob_idx, diag_map, np.zeros(0, np.int64), n_loc), glob_idx, n_loc)
    lo, hi = 3 * mesh.lo, 3 * (mesh.lo + mesh.n_owned)
    du = np.where(con_l, disp[gl].ravel(), 0.0)
    b = A_free @ du
    b[con_l] = 0.0

    # halo exchange: a vector that encodes the global dof id must come back consistent
    v = torch.full((n_loc,), -1.0)
    gdof = (3 * gl[:, None] + np.arange(3)).ravel().astype(np.float64)
    v[lo:hi] = torch.as_tensor(gdof[lo:hi])
    halo(v)
    assert np.array_equal(v.numpy(), gdof)

    # peer-store plan of the fused CG: what I send lands on the peer at the peer's entry of the SAME global dof,
    # for the range form (slabs) and for the general index-list form
    for force_lists in (False, True):
        pl = D.build_halo_plan(mesh, ranges, rank, 3)
        if force_lists:
            pl.contiguous.clear()
        sends = D.peer_send_plan(pl, rank)
        all_g = [None] * world
        dist.all_gather_object(all_g, gdof.tolist())
        assert sorted(sends) == pl.neighbours
        for s_, (src, dst) in sends.items():
            assert np.array_equal(gdof[src.numpy()], np.asarray(all_g[s_])[dst.numpy()])
            assert np.all((src.numpy() >= lo) & (src.numpy() < hi))          # I send owned entries only
    # interior rows: no halo column; the rows left out are exactly the planes next to a halo
    ia, ib = D.interior_rows(torch.as_tensor(A.indptr.astype(np.int64)), torch.as_tensor(A.indices.astype(np.int32)), lo, hi)
    touches = np.array([(A.indices[A.indptr[r_]:A.indptr[r_ + 1]] < lo).any()
                        or (A.indices[A.indptr[r_]:A.indptr[r_ + 1]] >= hi).any() for r_ in range(lo, hi)])
    assert lo <= ia <= ib <= hi and not touches[ia - lo:ib - lo].any()
    assert touches[:ia - lo].any() == (lo > 0) and touches[ib - lo:].any() == (hi < n_loc)
    assert (ib - ia) >= (hi - lo) - 2 * 3 * plane

    # distributed Jacobi-CG protocol (numpy arithmetic, torch.distributed collectives)
    def allreduce(*vals):
        t = torch.tensor(vals)
        dist.all_reduce(t)
        return t.tolist()

    dinv = 1.0 / A.diagonal()
    x = np.zeros(n_loc)
    r = np.zeros(n_loc)
    r[lo:hi] = b[lo:hi]
    p = torch.zeros(n_loc)
    p[lo:hi] = torch.as_tensor(dinv[lo:hi] * r[lo:hi])
    rr, rho, bb = allreduce(r[lo:hi] @ r[lo:hi], r[lo:hi] @ (dinv[lo:hi] * r[lo:hi]), b[lo:hi] @ b[lo:hi])
    tol = 1e-10 * np.sqrt(bb)
    its = 0
    while np.sqrt(rr) >= tol and its < 2000:
        halo(p)
        q = A @ p.numpy()
        (pq,) = allreduce(p.numpy()[lo:hi] @ q[lo:hi])
        alpha = rho / pq
        x[lo:hi] += alpha * p.numpy()[lo:hi]
        r[lo:hi] -= alpha * q[lo:hi]
        rr, rho_new = allreduce(r[lo:hi] @ r[lo:hi], r[lo:hi] @ (dinv[lo:hi] * r[lo:hi]))
        p[lo:hi] = torch.as_tensor(dinv[lo:hi] * r[lo:hi]) + (rho_new / rho) * p[lo:hi]
        rho = rho_new
        its += 1
    out[rank] = (n0, n1, x[lo:hi].copy(), its)
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_cg_matches_global_oracle():
    N, world = 5, 2
    # a spawned manager: forking this process (the default start method) once BLAS / OpenMP thread pools exist leaves
    # later LAPACK calls in the same pytest process hanging
    mgr = mp.get_context("spawn").Manager()
    shared = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), N, shared), nprocs=world, join=True)
    out = dict(shared)
    mgr.shutdown()
    nodes, elements, bref, w, C, con_mask, disp = _global_problem(N)
    ref = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    x = np.zeros(nodes.size)
    for rank in range(world):
        n0, n1, xo, its = out[rank]
        x[3 * n0:3 * n1] = xo
        assert abs(its - ref["iterations"]) <= 2
    u = -x
    con = np.nonzero(con_mask.ravel())[0]
    u[con] = disp.ravel()[con]
    assert np.linalg.norm(u - ref["u"].ravel()) / np.linalg.norm(ref["u"]) <= 1e-8


def test_node_ranges_and_general_partition():
    torch.set_default_dtype(torch.float64)
    from torchfem_b200 import distributed as D

    assert D.node_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert D.node_ranges(27, 2, granule=9) == [(0, 18), (18, 27)]
    r = D.node_ranges(1000, 8, granule=100)
    assert r[0][0] == 0 and r[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    # unstructured-ish mesh (tetra split): every element touching an owned node is local, local ids sorted
    nodes, hexes = O.cube_hexa(4, 3, 3)
    el = torch.as_tensor(hexes)
    m = D.local_mesh(el, 9, 27)
    g = m.global_nodes
    assert torch.equal(g, torch.unique(g)) and m.n_owned == 18
    assert torch.equal(g[m.lo:m.lo + m.n_owned], torch.arange(9, 27))
    assert torch.equal(g[m.elements], el[m.element_ids])
    touching = ((el >= 9) & (el < 27)).any(1)
    assert int(touching.sum()) == len(m.element_ids)
    plan = D.build_halo_plan(m, [(0, 9), (9, 27), (27, 36)], 1, 3)  # world==1 process: only `need` side is exercised
    assert plan.neighbours == [0, 2]


def test_coordinate_partition_of_a_quadratic_mesh():
    """Config C in small: linear_to_quadratic appends the mid-side nodes after the corner nodes, so node blocks of
    the mesh's own numbering are not slabs; the x-coordinate renumbering makes them slabs with <= 2 neighbours."""
    torch.set_default_dtype(torch.float64)
    from torchfem_b200 import distributed as D
    from torchfem_b200.elements import linear_to_quadratic
    from torchfem_b200.mesh import cube_hexa

    with torch.device("cpu"):
        n, e = linear_to_quadratic(*cube_hexa(9, 4, 4, 2.0, 1.0, 1.0))
    world = 3
    seen = torch.zeros(n.shape[0], dtype=torch.int64)
    for r in range(world):
        nh, m, ranges, perm = D.coordinate_partition(n, e, world, r)
        g = m.global_nodes
        assert torch.equal(torch.sort(perm).values, torch.arange(n.shape[0]))      # a permutation
        assert bool((n[perm][1:, 0] >= n[perm][:-1, 0]).all())                     # sorted by x
        assert torch.equal(n[perm[g]], nh)                                          # local coordinates
        assert torch.equal(perm[g][m.elements], e[m.element_ids])                   # local connectivity
        owned = perm[g[m.lo:m.lo + m.n_owned]]
        seen[owned] += 1
        plan = D.build_halo_plan(m, ranges, r, 3)
        assert plan.neighbours == [s for s in (r - 1, r + 1) if 0 <= s < world]
    assert bool((seen == 1).all())                                                  # every node owned exactly once
