"""Multi-GPU layer: one process per GPU (torchrun), contiguous node-block partition, NCCL over NVLink for
the two exchange steps the path really has — the SpMV halo exchange and the CG dot-product all-reduces.

The reference is single-process / single-GPU, so this layer has no counterpart there; it
follows SURVEY §8(e):

* rank r owns the node block [n0_r, n1_r) of the global numbering (for `cube_hexa`, x is the slowest
  index, so blocks are slabs of x-planes) and with it the matrix rows of those nodes;
* it keeps every element that touches an owned node ("ghost" elements of the one-element-thick interface
  layer are integrated redundantly on both sides), so assembly needs NO communication: every owned row is
  complete. Rows of halo nodes are incomplete and never used;
* local numbering = sorted global ids of all nodes of the local elements: [low halo | owned | high halo];
* per CG iteration: halo exchange of the search direction p (point-to-point with the ranks that own the
  halo nodes; for slabs that is one node plane = 547 kB at config B to each of <= 2 neighbours), then two
  all-reduces of 1 and 2 doubles. Scalars stay on the device; every rank runs the same scalar recurrences
  on the same reduced values, so all ranks take identical decisions without a broadcast.

Everything here is host orchestration over `torch.distributed`; the kernels are the single-GPU ones
(`tfem_cg_stage` in include/tfem_b200.h). The partition / halo bookkeeping is plain index arithmetic and
is covered by world_size-2 gloo tests on CPU (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor


# ------------------------------------------------------------------------------------------ partition
def node_ranges(n_nod: int, world: int, granule: int = 1) -> list[tuple[int, int]]:
    """Contiguous node blocks, balanced in units of `granule` nodes (e.g. one x-plane of a structured
    mesh so that blocks are whole planes)."""
    units = -(-n_nod // granule)
    base, rem = divmod(units, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        out.append((min(start * granule, n_nod), min((start + cnt) * granule, n_nod)))
        start += cnt
    return out


@dataclass
class LocalMesh:
    """What one rank needs of the global mesh."""
    global_nodes: Tensor      # [n_local] sorted global node ids (low halo | owned | high halo)
    elements: Tensor          # [n_elem_local, nn] connectivity in LOCAL node ids
    element_ids: Tensor       # [n_elem_local] global element ids
    lo: int                   # first owned local node
    n_owned: int
    n0: int                   # owned global range [n0, n1)
    n1: int

    @property
    def n_local(self) -> int:
        return int(self.global_nodes.numel())


def local_mesh(elements: Tensor, n0: int, n1: int) -> LocalMesh:
    """Elements touching the owned node block and the local renumbering (host tensors)."""
    touch = ((elements >= n0) & (elements < n1)).any(dim=1)
    eids = torch.nonzero(touch).ravel()
    el = elements[eids]
    owned = torch.arange(n0, n1, dtype=torch.int64)
    gl = torch.unique(torch.cat([el.reshape(-1), owned]))  # sorted
    el_local = torch.searchsorted(gl, el.reshape(-1)).reshape(el.shape)
    lo = int(torch.searchsorted(gl, torch.tensor(n0)))
    return LocalMesh(gl, el_local, eids, lo, n1 - n0, n0, n1)


@dataclass
class HaloPlan:
    """Who sends which local entries to whom. Index lists are in LOCAL numbering, per DOF."""
    neighbours: list[int]
    send_idx: dict = field(default_factory=dict)   # rank -> LongTensor of local dof indices I send
    recv_idx: dict = field(default_factory=dict)   # rank -> LongTensor of local dof indices I receive into
    contiguous: dict = field(default_factory=dict)  # rank -> ((s0, s1), (r0, r1)) when both lists are ranges

    def bytes_per_exchange(self) -> int:
        return 8 * sum(int(v.numel()) for v in self.send_idx.values())


def _as_range(idx: Tensor):
    if idx.numel() == 0:
        return None
    a, b = int(idx[0]), int(idx[-1]) + 1
    return (a, b) if b - a == idx.numel() else None


def build_halo_plan(mesh: LocalMesh, ranges: list[tuple[int, int]], rank: int, dpn: int,
                    group=None) -> HaloPlan:
    """Halo lists from the ownership ranges. Ranks tell each other which global nodes they need
    (one all_gather_object at setup); the answer is turned into local DOF index lists."""
    starts = torch.tensor([r[0] for r in ranges] + [ranges[-1][1]])
    gl = mesh.global_nodes
    is_halo = (gl < mesh.n0) | (gl >= mesh.n1)
    halo_local = torch.nonzero(is_halo).ravel()
    halo_global = gl[halo_local]
    owner = torch.searchsorted(starts, halo_global, right=True) - 1
    need = {int(s): halo_global[owner == s].tolist() for s in torch.unique(owner).tolist()}
    world = len(ranges)
    gathered = [None] * world
    if world > 1 and dist.is_initialized():
        dist.all_gather_object(gathered, need, group=group)
    else:  # single process (or a dry run without a process group): only my own needs are known
        gathered[rank] = need
    dofs = torch.arange(dpn)
    plan = HaloPlan(neighbours=[])
    for s in range(world):
        if s == rank:
            continue
        mine_for_s = gathered[s].get(rank, []) if gathered[s] else []   # global nodes rank s needs from me
        theirs_for_me = need.get(s, [])
        if not mine_for_s and not theirs_for_me:
            continue
        plan.neighbours.append(s)
        snd = torch.searchsorted(gl, torch.tensor(mine_for_s, dtype=torch.int64))
        rcv = torch.searchsorted(gl, torch.tensor(theirs_for_me, dtype=torch.int64))
        plan.send_idx[s] = (snd[:, None] * dpn + dofs).reshape(-1)
        plan.recv_idx[s] = (rcv[:, None] * dpn + dofs).reshape(-1)
        rs, rr = _as_range(plan.send_idx[s]), _as_range(plan.recv_idx[s])
        if rs is not None and rr is not None:
            plan.contiguous[s] = (rs, rr)
    return plan


class HaloExchanger:
    """Executes a HaloPlan on a vector living on `device` (NCCL for CUDA tensors, gloo for CPU)."""

    def __init__(self, plan: HaloPlan, device, dtype=torch.float64, group=None):
        self.plan, self.group = plan, group
        self.send_idx = {s: v.to(device) for s, v in plan.send_idx.items()}
        self.recv_idx = {s: v.to(device) for s, v in plan.recv_idx.items()}
        self.send_buf = {s: torch.empty(v.numel(), dtype=dtype, device=device) for s, v in plan.send_idx.items()}
        self.recv_buf = {s: torch.empty(v.numel(), dtype=dtype, device=device) for s, v in plan.recv_idx.items()}

    def __call__(self, vec: Tensor) -> None:
        """In place: fills the halo entries of `vec` with the owners' values."""
        if not self.plan.neighbours:
            return
        ops, scatter = [], []
        for s in self.plan.neighbours:
            if s in self.plan.contiguous:  # slabs: send / receive straight from / into the vector
                (s0, s1), (r0, r1) = self.plan.contiguous[s]
                snd, rcv = vec[s0:s1], vec[r0:r1]
            else:
                snd = self.send_buf[s]
                torch.index_select(vec, 0, self.send_idx[s], out=snd)
                rcv = self.recv_buf[s]
                scatter.append(s)
            if snd.numel():
                ops.append(dist.P2POp(dist.isend, snd, s, group=self.group))
            if rcv.numel():
                ops.append(dist.P2POp(dist.irecv, rcv, s, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for s in scatter:
            vec.index_copy_(0, self.recv_idx[s], self.recv_buf[s])


# ------------------------------------------------------------------------------------------ distributed CG
def distributed_cg(A, dinv: Tensor, b: Tensor, row_lo: int, n_owned: int, halo: HaloExchanger,
                   rtol: float = 1e-8, atol: float = 0.0, maxiter: int = 0, check_every: int = 32,
                   group=None):
    """Jacobi-PCG over row-partitioned ranks. `A` is the local `csr.CSRMatrix` (rows in local numbering),
    `dinv`, `b` local-length vectors (owned entries meaningful). Returns (x_local, info); the owned slice
    of x_local is this rank's part of the solution. Same stopping rule and the same recurrences as the
    single-GPU driver; every reduction is a fixed-order local sum followed by an NCCL all-reduce."""
    from . import _lib as L

    n_local = A.n
    S = A.sell()
    dev = b.device
    x = torch.zeros(n_local, dtype=torch.float64, device=dev)
    work = torch.empty(int(L.lib.tfem_krylov_work_doubles(n_local)), dtype=torch.float64, device=dev)
    red = torch.zeros(4, dtype=torch.float64, device=dev)
    off_p = int(L.lib.tfem_krylov_work_offset(n_local, 1))
    p = work[off_p:off_p + n_local]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if maxiter <= 0:
        maxiter = 10 * int(_global_sum_int(n_owned, dev, group))

    def stage(k):
        L.check(L.lib.tfem_cg_stage(k, S.ref, row_lo, n_owned, L.ptr(dinv), L.ptr(b), L.ptr(x), L.ptr(work),
                                    L.ptr(red), float(rtol), float(atol), L.stream()))

    def allreduce(k):
        if world > 1:
            dist.all_reduce(red[:k], group=group)

    info = np.zeros(4)
    stage(0)
    allreduce(3)
    stage(1)
    issued = 0
    launches = 3
    while True:
        L.check(L.lib.tfem_krylov_state(n_local, L.ptr(work), info.ctypes.data, L.stream()))
        if info[3] != 0.0 or issued >= maxiter:
            break
        batch = min(check_every, maxiter - issued)
        for _ in range(batch):
            halo(p)
            stage(2)
            allreduce(1)
            stage(3)
            stage(4)
            allreduce(2)
            stage(5)
            stage(6)
        issued += batch
        launches += 5 * batch
    stats = {"iterations": int(info[0]), "resnorm": float(info[1]), "bnorm": float(info[2]),
             "converged": info[3] == 1.0, "launches": launches}
    if info[3] != 1.0:
        raise RuntimeError(f"CG failed with exit code {stats['iterations'] if info[3] == 0.0 else -1}")
    return x, stats



# ------------------------------------------------------------------------------------------ fused peer-to-peer CG
def interior_rows(indptr: Tensor, indices: Tensor, row_lo: int, row_hi: int) -> tuple[int, int]:
    """Largest contiguous range [a, b) of owned rows such that no row in it references a halo column
    (columns < row_lo or >= row_hi). Those rows can be multiplied before the halo has arrived. Works on
    host or device tensors (columns are sorted within a row, so first/last column decide)."""
    if row_hi <= row_lo:
        return row_lo, row_lo
    ptr = indptr[row_lo:row_hi + 1].to(torch.int64)
    first = indices[ptr[:-1]].to(torch.int64)
    last = indices[ptr[1:] - 1].to(torch.int64)
    low = torch.nonzero(first < row_lo).ravel()
    high = torch.nonzero(last >= row_hi).ravel()
    a = row_lo + (int(low.max()) + 1 if low.numel() else 0)
    b = row_lo + (int(high.min()) if high.numel() else row_hi - row_lo)
    return (a, b) if b > a else (a, a)


def peer_send_plan(plan: HaloPlan, rank: int, group=None):
    """For every neighbour s: (send_idx in MY local numbering, destination idx in s's local numbering).
    The destinations are s's own `recv_idx[rank]`, learnt with one all_gather_object at setup."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = {int(s): v.tolist() if s not in plan.contiguous else ("range",) + plan.contiguous[s][1]
            for s, v in plan.recv_idx.items()}
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, mine, group=group)
    else:
        gathered[rank] = mine
    out = {}
    for s in plan.neighbours:
        dst = gathered[s].get(rank) if gathered[s] else None
        if dst is None or plan.send_idx[s].numel() == 0:
            continue
        if isinstance(dst, tuple):  # the peer receives into a contiguous range
            dst_idx = torch.arange(dst[1], dst[2], dtype=torch.int64)
        else:
            dst_idx = torch.tensor(dst, dtype=torch.int64)
        assert dst_idx.numel() == plan.send_idx[s].numel(), "halo plans of the two sides disagree"
        out[s] = (plan.send_idx[s].to(torch.int64), dst_idx)
    return out


class PeerComm:
    """Owner of a `tfem_comm` (include/tfem_b200.h): this rank's communication buffer plus the CUDA-IPC
    mappings of every peer's buffer. Explicit `close()`, like the reference's AmgX handle (amgx.py:385-395)."""

    def __init__(self, vec_doubles: int, group=None):
        import ctypes

        from . import _lib as L

        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the two p buffers of every rank must sit at the same offsets: size them for the longest local vector
        if self.world > 1:
            t = torch.tensor([int(vec_doubles)], dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            vec_doubles = int(t.item())
        self.vec_doubles = int(vec_doubles)
        self.handle = ctypes.c_void_p()
        ipc = ctypes.create_string_buffer(L.IPC_HANDLE_BYTES)
        L.check(L.lib.tfem_comm_create(self.rank, self.world, self.vec_doubles, ctypes.byref(self.handle), ipc))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, ipc.raw, group=group)
            L.check(L.lib.tfem_comm_connect(self.handle, b"".join(handles)))
            dist.barrier(group=group)  # every buffer is zeroed and mapped before anybody stores into a peer

    def set_trace(self, n_iterations: int = 0, first_iteration: int = 0, time_spmv: bool = False):
        """In-kernel profile of the cross-GPU waits (`tfem_comm_set_trace`): the kernels of the next solves stamp
        %globaltimer at fixed points of iterations [first, first + n). `n_iterations = 0` switches it off. Returns
        the device buffer [n, TRACE_SLOTS] (int64 ns), to be read after the solve."""
        from . import _lib as L

        self.trace = None
        if n_iterations > 0:
            self.trace = torch.zeros(n_iterations, L.TRACE_SLOTS, dtype=torch.int64,
                                     device=torch.device("cuda", torch.cuda.current_device()))
        L.check(L.lib.tfem_comm_set_trace(self.handle, L.ptr(self.trace), int(first_iteration), int(n_iterations),
                                          1 if time_spmv else 0))
        return self.trace

    def heap(self) -> Tensor:
        """This rank's symmetric heap as a float64 tensor view (no copy)."""
        import ctypes

        from . import _lib as L

        p, n = ctypes.c_void_p(), ctypes.c_int64()
        L.check(L.lib.tfem_comm_heap(self.handle, ctypes.byref(p), ctypes.byref(n)))
        return _tensor_from_ptr(p.value, int(n.value), torch.device("cuda", torch.cuda.current_device()))

    def close(self):
        from . import _lib as L

        if self.handle:
            if self.world > 1 and dist.is_initialized():
                dist.barrier()  # nobody unmaps while a peer may still store into this buffer
            L.check(L.lib.tfem_comm_destroy(self.handle))
            self.handle = None


def _tensor_from_ptr(ptr: int, n: int, device) -> Tensor:
    """float64 tensor view of `n` doubles of device memory owned by the library (lifetime: the communicator's)."""

    class _Mem:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Mem(), device=device)


class FusedCG:
    """Jacobi-PCG over row-partitioned ranks with halo exchange and all-reduces fused into the kernels
    (`tfem_dcg_solve`). Built once per (pattern, partition); `solve` is collective."""

    def __init__(self, pattern_indptr: Tensor, pattern_indices: Tensor, n_local: int, row_lo: int, n_owned: int,
                 plan: HaloPlan, device, group=None, comm: PeerComm | None = None):
        import ctypes

        from . import _lib as L

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_local, self.row_lo, self.n_owned = int(n_local), int(row_lo), int(n_owned)
        self.comm = comm if comm is not None else PeerComm(self.n_local, group)
        self.interior = interior_rows(pattern_indptr, pattern_indices, self.row_lo, self.row_lo + self.n_owned)
        sends = peer_send_plan(plan, self.rank, group)
        if len(sends) > L.MAX_NEIGHBOURS:
            raise RuntimeError(f"more than {L.MAX_NEIGHBOURS} halo neighbours")
        self._keep = []
        self.sends = (L.HaloSendStruct * max(1, len(sends)))()
        for k, (s, (src, dst)) in enumerate(sorted(sends.items())):
            e = self.sends[k]
            e.peer, e.count = int(s), int(src.numel())
            rs, rd = _as_range(src), _as_range(dst)
            if rs is not None and rd is not None:
                e.src_idx = e.dst_idx = None
                e.src_start, e.dst_start = rs[0], rd[0]
            else:
                si, di = src.to(torch.int32).to(device), dst.to(torch.int32).to(device)
                self._keep += [si, di]
                e.src_idx, e.dst_idx = si.data_ptr(), di.data_ptr()
                e.src_start = e.dst_start = 0
        self.n_sends = len(sends)
        recv = sorted(int(s) for s, v in plan.recv_idx.items() if v.numel())
        self.recv = np.asarray(recv, dtype=np.int32)
        n_global = _global_sum_int(self.n_owned, device, group)
        self.default_maxiter = 10 * n_global
        self.halo_bytes = 8 * sum(int(self.sends[k].count) for k in range(self.n_sends))
        self.work = torch.empty(int(L.lib.tfem_krylov_work_doubles(self.n_local)), dtype=torch.float64, device=device)

    def solve(self, A, dinv: Tensor, b: Tensor, rtol: float = 1e-8, atol: float = 0.0, maxiter: int = 0,
              check_every: int = 32, timeout_s: float = 20.0):
        from . import _lib as L

        S = A.sell()
        x = torch.zeros(self.n_local, dtype=torch.float64, device=b.device)
        info = np.zeros(8)
        rc = L.lib.tfem_dcg_solve(self.comm.handle, S.ref, self.row_lo, self.n_owned, self.interior[0],
                                  self.interior[1], self.n_sends, self.sends, int(self.recv.size),
                                  self.recv.ctypes.data, L.ptr(dinv), L.ptr(b.contiguous()), L.ptr(x),
                                  L.ptr(self.work), float(rtol), float(atol),
                                  int(maxiter if maxiter > 0 else self.default_maxiter), int(check_every),
                                  float(timeout_s), info.ctypes.data, L.stream())
        stats = {"iterations": int(info[0]), "resnorm": float(info[1]), "bnorm": float(info[2]),
                 "converged": bool(info[3]), "launches": int(info[5]), "spmv_ms": float(info[7])}
        if rc in (L.ERR_NOT_CONVERGED, L.ERR_BREAKDOWN):
            raise RuntimeError(f"CG failed with exit code {stats['iterations'] if rc == L.ERR_NOT_CONVERGED else -1}")
        L.check(rc)
        return x, stats

    def close(self):
        self.comm.close()


def _global_sum_int(v: int, dev, group=None) -> int:
    t = torch.tensor([v], dtype=torch.int64, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return int(t.item())


# ------------------------------------------------------------------------------------------ slab problem
def cube_slab(Ex: int, Ey: int, Ez: int, h: float, world: int, rank: int):
    """Rank-local part of a structured Hexa1 box of Ex x Ey x Ez cubic elements of edge h, partitioned into
    slabs of x-planes (x is the slowest index of the `cube_hexa` numbering, so slabs are contiguous node
    blocks). Built directly, never materialising the global mesh, and equal entry for entry to
    `local_mesh(cube_hexa(...).elements, n0, n1)` — tests/test_distributed_cpu.py checks that."""
    from .mesh import cube_hexa

    Nx, Ny, Nz = Ex + 1, Ey + 1, Ez + 1
    plane = Ny * Nz
    ranges = node_ranges(Nx * plane, world, granule=plane)
    n0, n1 = ranges[rank]
    a, b = n0 // plane, n1 // plane            # owned planes [a, b)
    pa, pb = max(a - 1, 0), min(b + 1, Nx)     # local planes incl. halo
    with torch.device("cpu"):
        nodes, elements = cube_hexa(pb - pa, Ny, Nz, (pb - pa - 1) * h, Ey * h, Ez * h)
        nodes[:, 0] += pa * h
    mesh = LocalMesh(torch.arange(pa * plane, pb * plane), elements, None, (a - pa) * plane, n1 - n0, n0, n1)
    return nodes, mesh, ranges, (Nx, Ny, Nz)


def coordinate_partition(nodes: Tensor, elements: Tensor, world: int, rank: int, axis: int = 0):
    """Rank-local part of an arbitrary mesh (host tensors): nodes are renumbered by (coordinate along `axis`, old
    id) so that contiguous blocks of the new numbering are slabs, then cut into `world` equal node blocks.
    Returns (local node coordinates, LocalMesh in the NEW numbering, ranges, perm) with perm[new id] = old id.
    Used for meshes whose numbering is not slab-contiguous, e.g. `linear_to_quadratic` output (config C), where
    the mid-side nodes are appended after the corner nodes. Halos are index lists, not ranges."""
    nodes, elements = nodes.cpu(), elements.cpu()
    key = nodes[:, axis].contiguous()
    perm = torch.sort(key, stable=True).indices          # stable: ties keep the old order
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel())
    el_new = inv[elements]
    ranges = node_ranges(nodes.shape[0], world)
    n0, n1 = ranges[rank]
    mesh = local_mesh(el_new, n0, n1)
    return nodes[perm[mesh.global_nodes]].contiguous(), mesh, ranges, perm


def weak_scaling_edge(E: int, world: int) -> int:
    """Edge (in elements) of the global cube that gives every rank a config-sized share: E * world^(1/3)."""
    return int(round(E * world ** (1.0 / 3.0)))
