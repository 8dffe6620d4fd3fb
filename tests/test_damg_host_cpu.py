"""Host logic of the distributed AMG setup (torch-fem_b200/damg.py) on the CPU, world size 2 over gloo: the row
bookkeeping of block operators (`_ragged`, `_rows_of`, `_frame_rows`) and the node-level halo plan — halo fill of
per-node data and the exchange of whole block rows that the Galerkin products are built from. The kernels (aggregation,
prolongator, SpGEMM, the cycle) are exercised on the GPUs by tests/test_gpu_multi.py (`tools/damg_check.py`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_block_row_helpers():
    from torchfem_b200 import damg as G

    starts, lens = torch.tensor([5, 0, 9]), torch.tensor([2, 0, 3])
    assert G._ragged(starts, lens).tolist() == [5, 6, 9, 10, 11]
    assert G._ragged(starts, torch.zeros(3, dtype=torch.int64)).numel() == 0
    d = 2
    bptr = torch.tensor([0, 2, 3, 3, 6])
    bcol = torch.tensor([4, 7, 1, 0, 2, 9], dtype=torch.int32)
    vals = torch.arange(d * d * 6, dtype=torch.float64)
    lens, cols, vv = G._rows_of(bptr, bcol, vals, d, torch.tensor([1, 2, 3]))          # a contiguous range
    assert lens.tolist() == [1, 0, 3] and cols.tolist() == [1, 0, 2, 9] and vv.tolist() == list(range(8, 24))
    lens, cols, vv = G._rows_of(bptr, bcol, vals, d, torch.tensor([3, 0]))             # an index list
    assert lens.tolist() == [3, 2] and cols.tolist() == [0, 2, 9, 4, 7]
    assert vv.tolist() == list(range(12, 24)) + list(range(0, 8))
    framed = G._frame_rows(torch.tensor([0, 2, 5]), 3, 7)
    assert framed.tolist() == [0, 0, 0, 0, 2, 5, 5, 5]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    from torchfem_b200 import damg as G
    from torchfem_b200 import distributed as D
    from torchfem_b200.mesh import cube_hexa

    E = 3
    nodes, elements = cube_hexa(2 * E + 1, E + 1, E + 1, 2.0, 1.0, 1.0)
    plane = (E + 1) ** 2
    ranges = D.node_ranges(nodes.shape[0], world, granule=plane)
    n0, n1 = ranges[rank]
    mesh = D.local_mesh(elements, n0, n1)
    W = G._World()
    plan = G.node_plan_from_halo_plan(D.build_halo_plan(mesh, ranges, rank, 1), "cpu", W)
    gl = mesh.global_nodes
    # halo fill: every local node ends up with its owner's value (here: a function of the GLOBAL node number)
    data = torch.full((mesh.n_local, 2), -1, dtype=torch.int64)
    own = slice(mesh.lo, mesh.lo + mesh.n_owned)
    data[own, 0], data[own, 1] = gl[own] * 10 + rank, gl[own] ** 2
    plan.fill_halo(data, W)
    owner = torch.searchsorted(torch.tensor([r[0] for r in ranges] + [ranges[-1][1]]), gl, right=True) - 1
    assert torch.equal(data[:, 0], gl * 10 + owner) and torch.equal(data[:, 1], gl ** 2)
    # dst = where my boundary nodes live in the neighbour's numbering: the neighbour's halo block is contiguous
    for s, dst in plan.dst.items():
        assert dst.numel() == plan.send[s].numel() and D._as_range(dst) is not None
    # exchange of whole block rows: row k of the owned rows has (k % 3) + 1 blocks with GLOBAL columns and values that
    # encode (global row, slot); the receiver must see exactly the rows of its halo nodes
    d = 2
    n_own = mesh.n_owned
    lens = torch.arange(n_own) % 3 + 1
    bptr = torch.zeros(n_own + 1, dtype=torch.int64)
    bptr[1:] = torch.cumsum(lens, 0)
    row_of = torch.repeat_interleave(torch.arange(n_own), lens)
    slot = torch.arange(int(bptr[-1])) - bptr[:-1][row_of]
    bcol = (gl[own][row_of] * 7 + slot).to(torch.int64)
    vals = (gl[own][row_of].double() * 100 + slot.double()).repeat_interleave(d * d) + torch.arange(d * d).double().repeat(int(bptr[-1])) / 10
    got = plan.exchange_rows(bptr, bcol, vals, d, mesh.lo, W)
    for s, (ls, cs, vs) in got.items():
        g_rows = gl[plan.recv[s]]
        k_rows = g_rows - ranges[s][0]                       # row numbers in the sender's owned block
        assert torch.equal(ls, k_rows % 3 + 1)
        rr = torch.repeat_interleave(g_rows, ls)
        sl = torch.arange(int(ls.sum())) - (torch.cumsum(ls, 0) - ls)[torch.repeat_interleave(torch.arange(len(ls)), ls)]
        assert torch.equal(cs, rr * 7 + sl)
        expect = (rr.double() * 100 + sl.double()).repeat_interleave(d * d) + torch.arange(d * d).double().repeat(int(ls.sum())) / 10
        assert torch.equal(vs, expect)
    out[rank] = 1
    dist.barrier()
    dist.destroy_process_group()


def test_node_plan_halo_fill_and_row_exchange_world2():
    world = 2
    port = _free_port()
    # a spawned manager (see tests/test_distributed_cpu.py): forking this process once BLAS / OpenMP thread pools exist
    # leaves later LAPACK calls of the same pytest process hanging
    with mp.get_context("spawn").Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert sorted(out.keys()) == [0, 1]
