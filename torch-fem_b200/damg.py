"""Distributed aggregation AMG-PCG: the hierarchy of `amg.py` partitioned by rows over the ranks of one box
(SURVEY §8(f)-3 x §8(e)). One process per GPU; the reference is single-process and hands its systems to a third-party
AMG (AmgX on the GPU: src/torchfem/sparse.py:422-442, amgx.py:211-383 setup / resetup / solve) — this module is the
multi-GPU counterpart of that life-cycle on this library's own kernels.

Algorithm (the single-GPU one, `oracle/amg_oracle.py`, with ONE change: aggregates never straddle ranks):

* level l lives on every rank as a square block operator over the rank's LOCAL numbering of the level — sorted global
  ids, i.e. [halo of lower ranks | owned | halo of higher ranks]; owned rows are complete, the others are never used;
* aggregation = the MIS kernels on the owned subgraph (`tfem_amg_aggregate_masked`); aggregates are numbered globally
  rank by rank, the halo nodes' aggregate numbers and isolated-DOF flags come from their owners;
* P = (I - w D^-1 A) T with the GLOBAL A: owned rows by the prolongator kernel (they reference aggregates of the
  neighbour across the interface), the P rows of the halo nodes come from their owners (one halo layer of rows);
* Galerkin product: W = A P for the owned rows (K15), the W rows of the halo nodes come from their owners, then
  A_c[owned aggregates, :] = R W with R = the rows of P^T for the owned aggregates — entry for entry the global
  triple product, no approximation at the interfaces (filtering the interface couplings out of the smoother instead
  costs +35 % iterations on 2 slabs and more on 8: tools/damg_proto.py);
* a level whose global size is below `gather_max` unknowns is gathered to every rank and continued redundantly with the
  single-GPU hierarchy (`amg.AMGPreconditioner.from_operator`): at 8 x 10 M DOFs that is level 3 (9 k unknowns);
* the cycle and the CG run in ONE C call (`tfem_damg_pcg_solve`): halo entries are stored straight into the
  neighbours' vectors over NVLink, dot products are LL-protocol all-reduces (csrc/peer.cuh); NCCL is used by the SETUP
  only (row exchanges, sizes).

Everything the hierarchy needs from the mesh partition is the node-level halo plan of level 0 (`distributed.py`).
"""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L
from . import amg as _amg
from .amg import AMGPreconditioner, BlockOperator, _empty, spgemm
from .csr import CSRMatrix
from .distributed import HaloPlan, PeerComm, _as_range

# a level with at most this many unknowns (global) is gathered and solved redundantly (TFEM_DAMG_GATHER_MAX: experiments).
# Measured on 8 B200 at 81.8 M DOFs: 400 k (a 317 k-unknown tail solved redundantly) 6.98 ms per iteration, 100 k (that
# level distributed too, tail 9 k) 5.98 ms (profiles/r2_f8_*).
GATHER_MAX_DOFS = int(os.environ.get("TFEM_DAMG_GATHER_MAX", 100_000))
MAX_DIST_LEVELS = 4


# ------------------------------------------------------------------------------------------ small helpers
def _ragged(starts: Tensor, lens: Tensor) -> Tensor:
    """concatenation of arange(starts[i], starts[i] + lens[i])"""
    total = int(lens.sum().item())
    if total == 0:
        return torch.empty(0, dtype=torch.int64, device=starts.device)
    first = torch.cumsum(lens, 0) - lens
    return torch.arange(total, device=starts.device) - torch.repeat_interleave(first - starts, lens)


def _rows_of(bptr: Tensor, bcol: Tensor, vals: Tensor, d: int, rows: Tensor):
    """(lens, cols, vals) of the block rows `rows` (whole rows keep their value layout: a row is self-contained)."""
    if rows.numel() == 0:
        z = torch.empty(0, dtype=torch.int64, device=bptr.device)
        return z, bcol[:0], vals[:0]
    rng = _as_range(rows)
    if rng is not None:
        a, b = int(bptr[rng[0]]), int(bptr[rng[1]])
        return bptr[rng[0] + 1:rng[1] + 1] - bptr[rng[0]:rng[1]], bcol[a:b], vals[d * d * a:d * d * b]
    starts = bptr[rows]
    lens = bptr[rows + 1] - starts
    return lens, bcol[_ragged(starts, lens)], vals[_ragged(d * d * starts, d * d * lens)]


class _World:
    """Thin wrapper: the collectives of the SETUP (NCCL through torch.distributed; a no-op world of one rank)."""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_gather_int(self, v: int, dev) -> list[int]:
        if not self.on:
            return [int(v)]
        t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
        out = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    def max_float(self, v: float, dev) -> float:
        if not self.on:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def all_gather_object(self, obj):
        if not self.on:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def exchange(self, send: dict[int, Tensor], recv_like: dict[int, tuple]) -> dict[int, Tensor]:
        """Point-to-point: send[s] goes to rank s, recv_like[s] = (numel, dtype) of what rank s sends me."""
        out = {}
        if not self.on:
            return out
        ops = []
        for s, t in send.items():
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t.contiguous(), s, group=self.group))
        for s, (n, dt, dev) in recv_like.items():
            out[s] = torch.empty(int(n), dtype=dt, device=dev)
            if n:
                ops.append(dist.P2POp(dist.irecv, out[s], s, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out

    def broadcast_all(self, mine: Tensor, sizes: list[int]) -> list[Tensor]:
        """Every rank's tensor on every rank (variable lengths: one broadcast per rank)."""
        if not self.on:
            return [mine]
        out = []
        for s in range(self.size):
            buf = mine.contiguous() if s == self.rank else torch.empty(sizes[s], dtype=mine.dtype, device=mine.device)
            if sizes[s]:
                dist.broadcast(buf, src=dist.get_global_rank(self.group, s) if self.group is not None else s,
                               group=self.group)
            out.append(buf)
        return out


class _NodePlan:
    """Halo plan of one level in NODE indices of the local numbering: send[s] / recv[s] int64 device tensors (sorted by
    global id, the same order on both sides), dst[s] = where my send[s] entries live in rank s's numbering."""

    def __init__(self, send: dict[int, Tensor], recv: dict[int, Tensor], W: _World):
        self.send = {s: v for s, v in send.items() if v.numel()}
        self.recv = {s: v for s, v in recv.items() if v.numel()}
        got = W.exchange({s: v for s, v in self.recv.items()},
                         {s: (v.numel(), torch.int64, v.device) for s, v in self.send.items()})
        self.dst = got

    def fill_halo(self, data: Tensor, W: _World) -> None:
        """In place: data[recv nodes] <- the owners' values (data: [n_loc, ...] any dtype NCCL moves)."""
        shape = data.shape[1:]
        k = int(np.prod(shape)) if len(shape) else 1
        got = W.exchange({s: data[v].reshape(-1) for s, v in self.send.items()},
                         {s: (v.numel() * k, data.dtype, data.device) for s, v in self.recv.items()})
        for s, v in self.recv.items():
            data[v] = got[s].reshape(v.numel(), *shape)

    def exchange_rows(self, bptr, bcol_global, vals, d, own_lo, W: _World):
        """Block rows of my boundary nodes -> the ranks that hold them as halo. `bptr/bcol_global/vals` describe the
        OWNED rows (row k = local node own_lo + k), columns in global numbers. Returns {s: (lens, cols, vals)} for
        the rows of recv[s], in that order."""
        out_l, out_c, out_v = {}, {}, {}
        for s, nodes in self.send.items():
            out_l[s], out_c[s], out_v[s] = _rows_of(bptr, bcol_global, vals, d, nodes - own_lo)
        dev = vals.device
        lens = W.exchange(out_l, {s: (v.numel(), torch.int64, dev) for s, v in self.recv.items()})
        nb = {s: int(v.sum().item()) for s, v in lens.items()}
        cols = W.exchange({s: v.to(torch.int64) for s, v in out_c.items()}, {s: (nb[s], torch.int64, dev) for s in lens})
        vv = W.exchange(out_v, {s: (nb[s] * d * d, torch.float64, dev) for s in lens})
        return {s: (lens[s], cols[s], vv[s]) for s in lens}

    def to_c(self, d: int, dev):
        """(HaloSendStruct array, n_sends, recv peers int32 array, keep-alive list) in scalar (DOF) indices."""
        keep = []
        peers = sorted(self.send)
        arr = (L.HaloSendStruct * max(1, len(peers)))()
        if len(peers) > L.MAX_NEIGHBOURS:
            raise RuntimeError(f"more than {L.MAX_NEIGHBOURS} halo neighbours")
        dofs = torch.arange(d, device=dev)
        for k, s in enumerate(peers):
            src, dst = self.send[s], self.dst[s]
            e = arr[k]
            e.peer, e.count = int(s), int(src.numel()) * d
            rs, rd = _as_range(src), _as_range(dst)
            if rs is not None and rd is not None:
                e.src_idx = e.dst_idx = None
                e.src_start, e.dst_start = rs[0] * d, rd[0] * d
            else:
                si = (src[:, None] * d + dofs).reshape(-1).to(torch.int32).contiguous()
                di = (dst[:, None] * d + dofs).reshape(-1).to(torch.int32).contiguous()
                keep += [si, di]
                e.src_idx, e.dst_idx = si.data_ptr(), di.data_ptr()
                e.src_start = e.dst_start = 0
        recv = np.asarray(sorted(self.recv), dtype=np.int32)
        keep.append(recv)
        return arr, len(peers), recv, keep


def node_plan_from_halo_plan(plan: HaloPlan, dev, W: _World) -> _NodePlan:
    """Level-0 node plan from the node-level `HaloPlan` of distributed.build_halo_plan(mesh, ranges, rank, dpn=1)."""
    return _NodePlan({s: v.to(dev) for s, v in plan.send_idx.items()}, {s: v.to(dev) for s, v in plan.recv_idx.items()}, W)


class _DLevel:
    pass


def _frame_rows(bptr_rows: Tensor, lo: int, n_total: int) -> Tensor:
    """Row offsets of an operator whose rows [lo, lo + len) are the given ones and all other rows are empty."""
    n_own = bptr_rows.numel() - 1
    out = torch.empty(n_total + 1, dtype=torch.int64, device=bptr_rows.device)
    out[:lo] = 0
    out[lo:lo + n_own + 1] = bptr_rows
    out[lo + n_own + 1:] = bptr_rows[-1]
    return out


class DistributedAMG:
    """The distributed hierarchy plus its work vectors; `solve(b)` = AMG-preconditioned CG over all ranks
    (`tfem_damg_pcg_solve`). Collective: every rank constructs it and calls `resetup` / `solve` together. Life-cycle of
    the reference's AmgX solver object (amgx.py:211-383): construction = setup, `resetup(A)` = new coefficients on the
    stored aggregates / patterns / exchange plans (numeric phases and value exchanges only), `solve`, explicit `close()`
    (amgx.py:385-395)."""

    def __init__(self, A: CSRMatrix, own_lo_node: int, n_owned_nodes: int, global_nodes: Tensor, node_plan: HaloPlan,
                 group=None, gather_max: int = GATHER_MAX_DOFS, max_coarse: int = _amg.MAX_COARSE_DOFS,
                 comm: PeerComm | None = None):
        if not isinstance(A, CSRMatrix) or A._sell_struct is None or A._sell_struct.block is None:
            raise TypeError("the distributed AMG needs an assembled CSRMatrix on a mesh pattern (node blocks)")
        self.W = _World(group)
        self.device = A.device
        self.gather_max, self.max_coarse = int(gather_max), int(max_coarse)
        self.levels: list[_DLevel] = []
        self._own = (int(own_lo_node), int(own_lo_node + n_owned_nodes))
        self._plan0 = node_plan_from_halo_plan(node_plan, self.device, self.W)
        self.comm = None
        self._setup(A, symbolic=True)
        self._allocate(comm)

    def resetup(self, A: CSRMatrix) -> None:
        """New coefficients on the same sparsity pattern and partition (Newton iterations, load cases, design updates):
        aggregates, the patterns of P / R / A·P / A_c, the coarse numberings and every exchange plan are kept; row info,
        spectral radii, P values, the two value exchanges per level, the numeric SpGEMMs and the tail are refreshed in
        place (AmgX `resetup`, reference sparse.py:440-441)."""
        if (A.indptr.data_ptr(), A.indices.data_ptr(), A.n) != self._pattern_key:
            raise ValueError("resetup needs a matrix on the pattern the hierarchy was built for")
        self._setup(A, symbolic=False)
        # the new matrix brings its own values and SELL copy: level 0's operator descriptor is rebuilt, the coarser
        # levels were refreshed in place
        self._structs[0].lv.A, keep = self.levels[0].op.operator_struct()
        self._keep.append(keep)
        for i, lv in enumerate(self.levels):
            self._structs[i].lv.omega = lv.omega
        if self.W.on:
            dist.barrier(group=self.W.group)

    # ------------------------------------------------------------------------------------------ setup
    def _tick(self, name: str) -> None:
        """Per-phase wall times of the setup (TFEM_AMG_TIMING=1; synchronises; diagnostics only)."""
        if self._timing is None:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        self._timing[name] = self._timing.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def _setup(self, A: CSRMatrix, symbolic: bool) -> None:
        W, dev, st = self.W, self.device, L.stream()
        self._timing = {} if os.environ.get("TFEM_AMG_TIMING") else None
        if self._timing is not None:
            torch.cuda.synchronize()
            self._t_last = time.perf_counter()
        self._A = A
        self._pattern_key = (A.indptr.data_ptr(), A.indices.data_ptr(), A.n)
        d, n_nod, node_ptr, adj = A._sell_struct.block
        op = BlockOperator(d, n_nod, n_nod, node_ptr, adj, A.values_, sell=A.sell())
        lo, hi = self._own
        plan = self._plan0
        if symbolic:
            self.d = d
            self.level_sizes = [sum(W.all_gather_int(hi - lo, dev)) * d]
        li = 0
        while True:
            if symbolic:
                lv = _DLevel()
                self.levels.append(lv)
                lv.op, lv.lo, lv.hi, lv.plan, lv.d, lv.n = op, lo, hi, plan, d, op.n_rows
                lv.dinv, lv.iso = _empty(lv.n, torch.float64, dev), _empty(lv.n, torch.uint8, dev)
            else:
                lv = self.levels[li]
                if li == 0:
                    lv.op = op
                op, lo, hi, plan = lv.op, lv.lo, lv.hi, lv.plan
            nb, n_own = op.nbr, hi - lo
            L.check(L.lib.tfem_amg_row_info(d, nb, L.ptr(op.bptr), L.ptr(op.bcol), L.ptr(op.vals), 1 if li > 0 else 0,
                                            L.ptr(lv.dinv), L.ptr(lv.iso), st))
            if li > 0:
                op.prepare()
            # spectral radius of D^-1 A: power iteration on the owned principal block (dinv zeroed outside it: after one
            # step the iterate vanishes on the halo), maximum over the ranks — a lower bound of the global radius by
            # interlacing, within the safety factor in practice (interior modes dominate)
            dmask = torch.zeros_like(lv.dinv)
            dmask[lo * d:hi * d] = lv.dinv[lo * d:hi * d]
            work = _empty(int(L.lib.tfem_amg_work_doubles(lv.n)), torch.float64, dev)
            rho = ctypes.c_double(0.0)
            ostruct, okeep = op.operator_struct()
            L.check(L.lib.tfem_amg_rho(ctypes.byref(ostruct), L.ptr(dmask), _amg.POWER_ITS, L.ptr(work), ctypes.byref(rho), st))
            del work, okeep, dmask
            lv.rho = W.max_float(rho.value, dev) * _amg.RHO_SAFETY
            lv.omega = 4.0 / (3.0 * lv.rho)
            self._tick("row_info+rho")

            if symbolic:
                self._aggregate(lv, op, plan, lo, hi)
                self._tick("aggregate")
            iso_nodes = lv.iso.view(nb, d).clone()
            plan.fill_halo(iso_nodes, W)                   # isolated-DOF flags of the halo nodes, from their owners
            iso_ext = iso_nodes.reshape(-1).contiguous()

            # ---- P rows of the owned nodes (global aggregate numbers as columns)
            if symbolic:
                lv.max_row = int((op.bptr[1:] - op.bptr[:-1]).max().item())
                lv.pptr = _empty(n_own + 1, torch.int64, dev)
                L.check(L.lib.tfem_amg_prolongator_count_rows(d, lo, n_own, L.ptr(op.bptr), L.ptr(op.bcol),
                                                              L.ptr(lv.agg32), L.ptr(lv.pptr), st))
                npb = int(lv.pptr[-1].item())
                lv.pcol = _empty(npb, torch.int32, dev)[:npb]
                lv.pval = _empty(d * d * npb, torch.float64, dev)[: d * d * npb]
            L.check(L.lib.tfem_amg_prolongator_fill_rows(d, lo, n_own, L.ptr(op.bptr), L.ptr(op.bcol), L.ptr(op.vals),
                                                         L.ptr(lv.agg32), L.ptr(lv.dinv), L.ptr(iso_ext), lv.omega,
                                                         L.ptr(lv.pptr), L.ptr(lv.pcol), L.ptr(lv.pval), lv.max_row, st))
            self._tick("P rows")
            # ---- P_ext: + the rows of the halo nodes from their owners
            if symbolic:
                lv.P_ext, lv.p_exch = self._extend_rows(plan, lv.pptr, lv.pcol.to(torch.int64), lv.pval, d, lo, hi, nb,
                                                        lv.n_agg_global)
            else:
                self._refresh_rows(plan, lv.p_exch, lv.pptr, lv.pval, d, lo, lv.P_ext)
            self._tick("P halo rows")
            # ---- R = rows of P_ext^T for the owned aggregates
            if symbolic:
                self._transpose_structure(lv)
            R = lv.R_own
            L.check(L.lib.tfem_amg_transpose_values(d, lv.n_c_compact, L.ptr(lv.Pc_bptr), L.ptr(lv.P_ext.vals),
                                                    L.ptr(lv.t_ptr), L.ptr(lv.t_col), L.ptr(lv.t_src),
                                                    L.ptr(lv.t_vals), st))
            self._tick("R")
            # ---- W = A P_ext for the owned rows, + halo rows from the owners
            Aown = BlockOperator(d, n_own, nb, op.bptr[lo:hi + 1], op.bcol, op.vals)
            if symbolic:
                Wown, lv.w_structure = spgemm(d, Aown, lv.P_ext)
                lv.W_ext, lv.w_exch = self._extend_rows(plan, Wown.bptr, Wown.bcol.to(torch.int64), Wown.vals, d, lo, hi,
                                                        nb, lv.n_agg_global)
                lv.w_ptr = Wown.bptr
            else:
                Wown, _ = spgemm(d, Aown, lv.P_ext, lv.w_structure)
                self._refresh_rows(plan, lv.w_exch, lv.w_ptr, Wown.vals, d, lo, lv.W_ext)
            del Wown, Aown
            self._tick("A*P + halo rows")
            # ---- A_c rows of the owned aggregates, global columns
            if symbolic:
                Ac, lv.ac_structure = spgemm(d, R, lv.W_ext)
                lv.gather = lv.n_agg_global * d <= self.gather_max or li + 1 >= MAX_DIST_LEVELS
                self.level_sizes.append(lv.n_agg_global * d)
                c_off, n_agg_r = lv.c_off, lv.n_agg_r
                if lv.gather:
                    # coarse index space = global numbering of the tail
                    lv.c_lo, lv.c_hi, lv.n_c = c_off, c_off + n_agg_r, lv.n_agg_global
                    lv.P = BlockOperator(d, nb, lv.n_agg_global, _frame_rows(lv.pptr, lo, nb), lv.pcol, lv.pval)
                    lv.R = BlockOperator(d, lv.n_agg_global, nb, _frame_rows(R.bptr, c_off, lv.n_agg_global), R.bcol, R.vals)
                    lv.Ac = Ac
                    self._tick("R*(AP)")
                    self._build_tail(lv, d)
                    self._tick("tail")
                    break
                # ---- next level distributed: local numbering of the coarse nodes = sorted global ids that occur
                uniq = torch.unique(Ac.bcol.to(torch.int64))
                n_c = int(uniq.numel())
                c_lo = int(torch.searchsorted(uniq, torch.tensor(c_off, device=dev)).item())
                if not bool((uniq[c_lo:c_lo + n_agg_r] == torch.arange(c_off, c_off + n_agg_r, device=dev)).all()):
                    raise RuntimeError("distributed AMG: an owned aggregate has no diagonal block")
                lv.c_lo, lv.c_hi, lv.n_c = c_lo, c_lo + n_agg_r, n_c
                to_local = lambda g: torch.searchsorted(uniq, g.to(torch.int64)).to(torch.int32)  # noqa: E731
                lv.P = BlockOperator(d, nb, n_c, _frame_rows(lv.pptr, lo, nb), to_local(lv.pcol), lv.pval)
                lv.R = BlockOperator(d, n_c, nb, _frame_rows(R.bptr, c_lo, n_c), R.bcol, R.vals)
                lv.Ac = Ac
                op = BlockOperator(d, n_c, n_c, _frame_rows(Ac.bptr, c_lo, n_c), to_local(Ac.bcol), Ac.vals)
                plan = self._coarse_plan(uniq, lv, c_lo, n_agg_r, n_c)
                lo, hi = c_lo, c_lo + n_agg_r
                self._tick("R*(AP) + coarse numbering")
            else:
                spgemm(d, R, lv.W_ext, lv.ac_structure, out_vals=lv.Ac.vals)   # next level's operator, in place
                self._tick("R*(AP)")
                if lv.gather:
                    self._refresh_tail(lv, d)
                    self._tick("tail")
                    break
            li += 1
        if not symbolic:
            for lv in self.levels:
                lv.P.prepare(), lv.R.prepare()

    def _aggregate(self, lv, op, plan, lo, hi) -> None:
        """Aggregation on the owned subgraph; global aggregate numbers for owned and halo nodes."""
        W, dev, st = self.W, self.device, L.stream()
        nb = op.nbr
        exclude = torch.ones(nb, dtype=torch.uint8, device=dev)
        exclude[lo:hi] = 0
        agg = _empty(nb, torch.int32, dev)
        state, flag, index = _empty(nb, torch.int8, dev), _empty(nb, torch.uint8, dev), _empty(nb, torch.int32, dev)
        n_agg, rounds = ctypes.c_int64(0), ctypes.c_int32(0)
        for distance in (1, 2):
            L.check(L.lib.tfem_amg_aggregate_masked(nb, L.ptr(op.bptr), L.ptr(op.bcol), distance, L.ptr(exclude),
                                                    L.ptr(state), L.ptr(flag), L.ptr(index), L.ptr(agg),
                                                    ctypes.byref(n_agg), ctypes.byref(rounds), st))
            if (hi - lo) >= _amg.MIN_AGG_SIZE * n_agg.value:
                break
        lv.n_agg_r = int(n_agg.value)
        lv.aggs = W.all_gather_int(lv.n_agg_r, dev)
        lv.offs = [0]
        for a in lv.aggs:
            lv.offs.append(lv.offs[-1] + a)
        lv.n_agg_global = lv.offs[-1]
        if lv.n_agg_global >= 2 ** 31 - 1:
            raise RuntimeError("more than 2^31 aggregates")
        lv.c_off = lv.offs[W.rank]
        agg_g = torch.full((nb,), -1, dtype=torch.int64, device=dev)
        agg_g[lo:hi] = agg[lo:hi].to(torch.int64) + lv.c_off
        plan.fill_halo(agg_g, W)                       # aggregate numbers of the halo nodes, from their owners
        # nodes of the local numbering that no owned row references keep -1: give them a valid number
        lv.agg32 = torch.where(agg_g < 0, torch.zeros_like(agg_g), agg_g).to(torch.int32).contiguous()

    def _coarse_plan(self, uniq, lv, c_lo, n_agg_r, n_c) -> _NodePlan:
        """Halo plan of the next level: which coarse nodes I need from whom / who needs mine."""
        W, dev = self.W, self.device
        starts = torch.tensor(lv.offs, device=dev)
        halo_local = torch.cat([torch.arange(0, c_lo, device=dev), torch.arange(c_lo + n_agg_r, n_c, device=dev)])
        halo_global = uniq[halo_local]
        owner = torch.searchsorted(starts, halo_global, right=True) - 1
        need = {int(s): halo_global[owner == s].cpu().numpy() for s in torch.unique(owner).tolist()}
        gathered = W.all_gather_object(need)
        send, recv = {}, {}
        for s in range(W.size):
            if s == W.rank:
                continue
            theirs = gathered[s].get(W.rank) if gathered[s] else None
            if theirs is not None and len(theirs):
                send[s] = torch.searchsorted(uniq, torch.as_tensor(theirs, device=dev))
            if s in need and len(need[s]):
                recv[s] = torch.searchsorted(uniq, torch.as_tensor(need[s], device=dev))
        return _NodePlan(send, recv, W)

    # ------------------------------------------------------------------------------------------ setup pieces
    def _extend_rows(self, plan: _NodePlan, bptr_own, bcol_own_global, vals_own, d, lo, hi, nb, n_cols):
        """Operator over all `nb` local rows: the owned rows as given + the rows of the halo nodes from their owners.
        Local order = rank order (global ids are contiguous per rank), so the pieces are concatenated rank by rank.
        Returns (operator, exchange record for the values-only refresh)."""
        W = self.W
        got = plan.exchange_rows(bptr_own, bcol_own_global, vals_own, d, lo, W)
        dev = vals_own.device
        lens_all = torch.zeros(nb, dtype=torch.int64, device=dev)
        lens_all[lo:hi] = bptr_own[1:] - bptr_own[:-1]
        pieces = [(lo, None, bcol_own_global, vals_own)]
        for s, (lens, cols, vv) in got.items():
            nodes = plan.recv[s]
            rng = _as_range(nodes)
            if rng is None:
                raise RuntimeError("distributed AMG: the halo nodes of a neighbour are not contiguous in the local numbering")
            lens_all[nodes] = lens
            pieces.append((rng[0], s, cols, vv))
        pieces.sort(key=lambda p: p[0])
        bptr = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
        bptr[1:] = torch.cumsum(lens_all, 0)
        bcol = torch.cat([p[2].to(torch.int64) for p in pieces]).to(torch.int32).contiguous()
        vals = torch.cat([p[3] for p in pieces]).contiguous()
        # where each piece's values live in `vals` (for the values-only refresh), what to send (value index per peer)
        at, place = 0, {}
        for first_row, s, _, vv in pieces:
            place[s] = (at, at + vv.numel())
            at += vv.numel()
        send_idx = {}
        for s, nodes in plan.send.items():
            rows = nodes - lo
            rng = _as_range(rows)
            if rng is not None:
                send_idx[s] = (int(bptr_own[rng[0]]) * d * d, int(bptr_own[rng[1]]) * d * d)
            else:
                starts = bptr_own[rows]
                send_idx[s] = _ragged(d * d * starts, d * d * (bptr_own[rows + 1] - starts))
        return BlockOperator(d, nb, n_cols, bptr, bcol, vals), (place, send_idx)

    def _refresh_rows(self, plan: _NodePlan, record, bptr_own, vals_own, d, lo, ext: BlockOperator) -> None:
        """Values-only version of `_extend_rows`: own values copied, halo rows' values received, all in place."""
        place, send_idx = record
        out = {}
        for s, idx in send_idx.items():
            out[s] = vals_own[idx[0]:idx[1]] if isinstance(idx, tuple) else vals_own[idx]
        dev = vals_own.device
        got = self.W.exchange(out, {s: (place[s][1] - place[s][0], torch.float64, dev) for s in plan.recv if s in place})
        a, b = place[None]
        ext.vals[a:b] = vals_own
        for s, v in got.items():
            a, b = place[s]
            ext.vals[a:b] = v

    def _transpose_structure(self, lv) -> None:
        """Structure of R = rows of P_ext^T for the owned aggregates. The columns of P_ext are renumbered compactly
        (owned aggregates first, the halo aggregates that occur after them) so that K14 transposes the whole operator
        and the owned rows are simply the first n_agg_r rows of the result."""
        P, dev, st = lv.P_ext, self.device, L.stream()
        d, nb = P.d, P.nbr
        col = P.bcol.to(torch.int64)
        own = (col >= lv.c_off) & (col < lv.c_off + lv.n_agg_r)
        halo_ids = torch.unique(col[~own])
        compact = torch.where(own, col - lv.c_off, lv.n_agg_r + torch.searchsorted(halo_ids, col))
        lv.n_c_compact = lv.n_agg_r + int(halo_ids.numel())
        lv.Pc_bptr = P.bptr
        pc_col = compact.to(torch.int32).contiguous()
        npb = P.nblk
        lv.t_ptr = _empty(lv.n_c_compact + 1, torch.int64, dev)
        lv.t_col, lv.t_src = _empty(npb, torch.int32, dev)[:npb], _empty(npb, torch.int32, dev)[:npb]
        L.check(L.lib.tfem_amg_transpose_structure(nb, lv.n_c_compact, L.ptr(P.bptr), L.ptr(pc_col), npb, L.ptr(lv.t_ptr),
                                                   L.ptr(lv.t_col), L.ptr(lv.t_src), st))
        lv.t_vals = torch.empty_like(P.vals)
        n_r = int(lv.t_ptr[lv.n_agg_r].item())
        lv.R_own = BlockOperator(d, lv.n_agg_r, nb, lv.t_ptr[:lv.n_agg_r + 1], lv.t_col[:n_r], lv.t_vals[:d * d * n_r])

    def _build_tail(self, lv, d: int) -> None:
        """Gather the coarse operator (rows of the owned aggregates, global columns) to every rank and continue with the
        single-GPU hierarchy on it."""
        W, dev = self.W, self.device
        Ac = lv.Ac
        lens = Ac.bptr[1:] - Ac.bptr[:-1]
        lv.tail_nblk = W.all_gather_int(Ac.nblk, dev)
        all_lens = W.broadcast_all(lens.contiguous(), lv.aggs)
        all_cols = W.broadcast_all(Ac.bcol.contiguous(), lv.tail_nblk)
        all_vals = W.broadcast_all(Ac.vals.contiguous(), [n * d * d for n in lv.tail_nblk])
        n = lv.offs[-1]
        bptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        bptr[1:] = torch.cumsum(torch.cat(all_lens), 0)
        self.tail_op = BlockOperator(d, n, n, bptr, torch.cat(all_cols).contiguous(), torch.cat(all_vals).contiguous())
        self.tail = AMGPreconditioner.from_operator(self.tail_op, max_coarse=self.max_coarse)
        self.level_sizes += [int(t.n) for t in self.tail.levels[1:]]

    def _refresh_tail(self, lv, d: int) -> None:
        all_vals = self.W.broadcast_all(lv.Ac.vals.contiguous(), [n * d * d for n in lv.tail_nblk])
        self.tail_op.vals.copy_(torch.cat(all_vals))
        self.tail._setup(None, symbolic=False, op0=self.tail_op)

    def _allocate(self, comm: PeerComm | None) -> None:
        """Heap vectors (x, t of every distributed level, p, the gathered tail right-hand side) at equal offsets on all
        ranks, level structs for the C call."""
        W, dev = self.W, self.device
        pad = lambda n: (int(n) + 31) & ~31  # noqa: E731
        sizes = [lv.n for lv in self.levels]
        if W.on:
            t = torch.tensor(sizes, dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=W.group)
            sizes = [int(v) for v in t.tolist()]
        n_tail = self.tail.n
        total = sum(2 * pad(s) for s in sizes) + pad(sizes[0]) + pad(n_tail)
        self._own_comm = comm is None
        self.comm = comm if comm is not None else PeerComm((total + 1) // 2, W.group)
        heap = self.comm.heap()
        if heap.numel() < total:
            raise RuntimeError("the communicator's heap is too small for the hierarchy")
        at = 0

        def take(n_slot, n):
            nonlocal at
            v = heap[at:at + n]
            at += pad(n_slot)
            return v

        for lv, s in zip(self.levels, sizes):
            lv.x, lv.t = take(s, lv.n), take(s, lv.n)
            lv.b = torch.zeros(lv.n, dtype=torch.float64, device=dev)
        self.p = take(sizes[0], self.levels[0].n)
        self.tail_b = take(n_tail, n_tail)
        heap[:at].zero_()
        self.tail_x = torch.zeros(n_tail, dtype=torch.float64, device=dev)
        nl = len(self.levels)
        arr = (L.DamgLevelStruct * nl)()
        keep = []
        for i, lv in enumerate(self.levels):
            d = lv.d
            lv.P.prepare(), lv.R.prepare()
            arr[i].lv.A, k = lv.op.operator_struct()
            arr[i].lv.P, kp = lv.P.operator_struct()
            arr[i].lv.R, kr = lv.R.operator_struct()
            keep += [k, kp, kr]
            arr[i].lv.omega = lv.omega
            arr[i].lv.dinv = L.ptr(lv.dinv)
            arr[i].lv.x, arr[i].lv.b, arr[i].lv.t = L.ptr(lv.x), L.ptr(lv.b), L.ptr(lv.t)
            arr[i].own_lo, arr[i].own_hi = lv.lo * d, lv.hi * d
            arr[i].c_own_lo, arr[i].c_own_hi = lv.c_lo * d, lv.c_hi * d
            sends, n_sends, recv, kk = lv.plan.to_c(d, dev)
            arr[i].n_sends, arr[i].sends = n_sends, ctypes.cast(sends, ctypes.c_void_p)
            arr[i].n_recv, arr[i].recv_peers = int(recv.size), recv.ctypes.data
            keep += [sends, recv, kk]
        self._structs, self._keep = arr, keep
        self._work = torch.empty(int(L.lib.tfem_amg_work_doubles(self.levels[0].n)), dtype=torch.float64, device=dev)
        if W.on:
            dist.barrier(group=W.group)

    # ------------------------------------------------------------------------------------------ application
    @property
    def n_levels(self) -> int:
        return len(self.levels) + len(self.tail.levels)

    def solve(self, b: Tensor, rtol: float = 1e-10, atol: float = 0.0, maxiter: int = 0, timeout_s: float = 30.0):
        """AMG-preconditioned CG over all ranks (zero initial guess). b: local-length level-0 vector (owned entries
        meaningful). Returns (x_local, stats); raises RuntimeError("CG failed with exit code ...") like the reference's
        Krylov paths (sparse.py:421)."""
        L.require_cuda(b)
        b = b.to(torch.float64).contiguous()
        x = torch.zeros_like(b)
        info = np.zeros(8, dtype=np.float64)
        T = self.tail
        rc = L.lib.tfem_damg_pcg_solve(self.comm.handle, self._structs, len(self.levels), T._structs, T.n_levels,
                                       L.ptr(T.levels[-1].inv), T.n, L.ptr(self.tail_b), L.ptr(self.tail_x), L.ptr(b),
                                       L.ptr(x), L.ptr(self.p), L.ptr(self._work), float(rtol), float(atol), int(maxiter),
                                       float(timeout_s), info.ctypes.data, L.stream())
        stats = {"iterations": int(info[0]), "resnorm": float(info[1]), "bnorm": float(info[2]),
                 "converged": bool(info[3]), "spmv": int(info[4]), "launches": int(info[5])}
        if rc in (L.ERR_NOT_CONVERGED, L.ERR_BREAKDOWN):
            raise RuntimeError(f"CG failed with exit code {stats['iterations'] if rc == L.ERR_NOT_CONVERGED else -1}")
        L.check(rc)
        return x, stats

    def close(self) -> None:
        if getattr(self, "comm", None) is not None and self._own_comm:
            self.comm.close()
        self.comm = None
