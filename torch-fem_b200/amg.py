"""Aggregation algebraic multigrid on the device — the preconditioner behind `sparse_solve(method="amgx")`.

The reference preconditions its Krylov solves with third-party AMG: pyamg `smoothed_aggregation_solver(A, B,
smooth="jacobi")` on the CPU (src/torchfem/sparse.py:493-512) and AmgX aggregation AMG on the GPU
(src/torchfem/amgx.py:71-98; sparse.py:422-442: hierarchy built on the first solve, `resetup` = coefficient refresh
when the returned solver object is passed back in). Here the hierarchy is built by kernels K11-K15 of libtfem_b200.so
(include/tfem_b200.h, csrc/amg.cu) and applied by K16; this module owns the buffers and drives the setup level by level:

    row info -> rho(D^-1 A) -> aggregation -> P = (I - w D^-1 A) T -> R = P^T -> A P -> A_c = R (A P)

Smoothed aggregation on the node graph, d x d blocks (d = DOFs per node), near-null space = the d translations masked
at Dirichlet rows, V(1,1) cycle with damped Jacobi, dense inverse on the coarsest level (a library call on a system of
at most `max_coarse` unknowns, like `method="spsolve"`). The restatement every step is checked against is
oracle/amg_oracle.py (tests only).
"""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np
import torch
from torch import Tensor

from . import _lib as L
from .csr import CSRMatrix, SellMatrix, SellStructure

POWER_ITS = 8
RHO_SAFETY = 1.15
MAX_COARSE_DOFS = 1500
MAX_LEVELS = 12
DENSE_LIMIT = 12000  # largest coarsest level a dense inverse is accepted for
MIN_AGG_SIZE = 6        # nodes per aggregate below which radius-2 aggregates replace radius-1 ones (aggregation="auto")
BCSR_MAX_ROWS = int(os.environ.get("TFEM_AMG_BCSR_MAX_ROWS", 20000))   # operators with fewer block rows than this ...
BCSR_MIN_AVG = int(os.environ.get("TFEM_AMG_BCSR_MIN_AVG", 48))         # ... or more blocks per row than this are
#                                                                         streamed as block CSR instead of SELL-32


def _empty(n, dtype, device):
    return torch.empty(max(int(n), 1), dtype=dtype, device=device)


class BlockOperator:
    """Block-CSR operator over nodes: `bptr` int64 [nbr+1], `bcol` int32 (sorted per row), `vals` float64 in the
    scalar-CSR order of the assembled matrix (block row I with m blocks: entry (a, s, c) at d*d*bptr[I] + (a*m+s)*d + c).
    `nbr` x `nbc` blocks of d x d."""

    def __init__(self, d: int, nbr: int, nbc: int, bptr: Tensor, bcol: Tensor, vals: Tensor, sell: SellMatrix | None = None):
        self.d, self.nbr, self.nbc = int(d), int(nbr), int(nbc)
        self.bptr, self.bcol, self.vals = bptr, bcol, vals
        self.n_rows, self.n_cols = self.nbr * self.d, self.nbc * self.d
        self._sell = sell
        self._sell_struct = None
        self._sell_vals = None

    @property
    def nblk(self) -> int:
        return int(self.bcol.shape[0])

    @property
    def use_bcsr(self) -> bool:
        """Layout the cycle streams this operator in: SELL-32 (a row per lane) suits many short rows, the block-CSR
        kernel (8 / 32 / 256 threads per block row) few long ones — coarse levels and their restrictions."""
        if self._sell is not None and self._sell_struct is None:
            return False                                   # level 0: the assembled matrix's own SELL copy
        return self.nbr < BCSR_MAX_ROWS or self.nblk > BCSR_MIN_AVG * self.nbr

    def operator_struct(self):
        """`tfem_amg_operator_t` (plus the objects that keep its pointers alive)."""
        o = L.AmgOperatorStruct()
        if self.use_bcsr:
            o.bcsr.nb_rows, o.bcsr.n_blocks, o.bcsr.d = self.nbr, self.nblk, self.d
            o.bcsr.bptr, o.bcsr.bcol, o.bcsr.vals = L.ptr(self.bptr), L.ptr(self.bcol), L.ptr(self.vals)
            return o, (self.bptr, self.bcol, self.vals)
        S = self.sell()
        o.sell = S.struct
        return o, S

    def prepare(self) -> None:
        """Build (first call) or refresh (later calls) the layout the cycle uses."""
        if self.use_bcsr:
            return
        if self._sell is None:
            self.sell()
        else:
            self.refresh_sell()

    @property
    def indptr(self) -> Tensor:
        """Scalar CSR row offsets implied by the block layout."""
        d = self.d
        if d == 1:
            return self.bptr
        cnt = self.bptr[1:] - self.bptr[:-1]
        a = torch.arange(d, device=self.bptr.device, dtype=torch.int64)
        rows = (d * d * self.bptr[:-1])[:, None] + (a * d)[None, :] * cnt[:, None]
        return torch.cat([rows.reshape(-1), (d * d * self.bptr[-1:])])

    def sell(self) -> SellMatrix:
        """SELL-32 copy for the cycle (structure built once, values re-converted by `refresh_sell`)."""
        if self._sell is None:
            d = self.d
            indptr = self.indptr.contiguous()
            block = (d, self.nbr, self.bptr, self.bcol) if d in (2, 3) else None
            st = SellStructure(indptr, self.bcol if d == 1 else None, self.n_rows, block, n_cols=self.n_cols)
            self._sell_struct = st
            self._sell_vals = torch.empty(max(st.padded, 2), dtype=torch.float64, device=self.vals.device)
            self.refresh_sell()
            self._sell = SellMatrix(st, self._sell_vals, use_block=block is not None)
        return self._sell

    def refresh_sell(self) -> None:
        st = self._sell_struct
        if st is None:
            return
        L.check(L.lib.tfem_sell_fill(self.n_rows, L.ptr(st.indptr), None, L.ptr(self.vals), L.ptr(st.slice_ptr),
                                     None, L.ptr(self._sell_vals), L.stream()))

    def to_scipy(self):
        """Host copy as a scipy CSR matrix (tests / diagnostics)."""
        import scipy.sparse as sp

        d = self.d
        bptr = self.bptr.cpu().numpy()
        bcol = self.bcol.cpu().numpy().astype(np.int64)
        vals = self.vals.cpu().numpy()
        cnt = np.diff(bptr)
        node = np.repeat(np.arange(self.nbr), cnt)
        slot = np.arange(len(bcol)) - bptr[:-1][node]
        m, base = cnt[node], d * d * bptr[:-1][node]
        rows, cols, v = [], [], []
        for a in range(d):
            for c in range(d):
                rows.append(node * d + a)
                cols.append(bcol * d + c)
                v.append(vals[base + (a * m + slot) * d + c])
        return sp.csr_matrix((np.concatenate(v), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(self.n_rows, self.n_cols))


def _level0_operator(A: CSRMatrix) -> BlockOperator:
    st = A._sell_struct
    if st is not None and st.block is not None:
        d, n_nod, node_ptr, adj = st.block
        return BlockOperator(d, n_nod, n_nod, node_ptr, adj, A.values_, sell=A.sell())
    return BlockOperator(1, A.n, A.n, A.indptr, A.indices, A.values_, sell=A.sell())


def spgemm(d: int, X: BlockOperator, Y: BlockOperator, structure=None, out_vals: Tensor | None = None):
    """C = X Y (K15). `structure` = (cptr, ccol, max_row) of an earlier symbolic phase, `out_vals` an existing
    value buffer to overwrite."""
    dev, st = X.vals.device, L.stream()
    if structure is None:
        cptr = _empty(X.nbr + 1, torch.int64, dev)
        L.check(L.lib.tfem_amg_spgemm_count(X.nbr, L.ptr(X.bptr), L.ptr(X.bcol), L.ptr(Y.bptr), L.ptr(Y.bcol),
                                            L.ptr(cptr), st))
        nblk = int(cptr[-1].item())
        ccol = _empty(nblk, torch.int32, dev)[:nblk]
        L.check(L.lib.tfem_amg_spgemm_fill(X.nbr, L.ptr(X.bptr), L.ptr(X.bcol), L.ptr(Y.bptr), L.ptr(Y.bcol),
                                           L.ptr(cptr), L.ptr(ccol), st))
        max_row = int((cptr[1:] - cptr[:-1]).max().item())
        structure = (cptr, ccol, max_row)
    cptr, ccol, max_row = structure
    nv = d * d * ccol.shape[0]
    cvals = out_vals if out_vals is not None else _empty(nv, torch.float64, dev)[:nv]
    # a CTA per row of C when the steps over X's row are wide (blocks of Y's row x d tasks each) or the rows of C are
    # too few to fill the GPU with one warp each
    tpr = 256 if (Y.nblk * d >= 192 * Y.nbr or X.nbr < 40000) else 32
    L.check(L.lib.tfem_amg_spgemm_numeric(d, X.nbr, L.ptr(X.bptr), L.ptr(X.bcol), L.ptr(X.vals), L.ptr(Y.bptr),
                                          L.ptr(Y.bcol), L.ptr(Y.vals), L.ptr(cptr), L.ptr(ccol), L.ptr(cvals),
                                          max_row, tpr, st))
    return BlockOperator(d, X.nbr, Y.nbc, cptr, ccol, cvals), structure



class _Level:
    pass


class AMGPreconditioner:
    """The hierarchy plus its work vectors; `apply(r)` is one V cycle, `solve(b)` AMG-preconditioned CG (K16).
    Returned by `sparse_solve(method="amgx")` as `M` and accepted back like the reference's AmgX solver object
    (sparse.py:438-441): passing it back refreshes the coefficients (`resetup`) on the stored aggregates/patterns."""

    def __init__(self, A: CSRMatrix, max_coarse: int = MAX_COARSE_DOFS, max_levels: int = MAX_LEVELS,
                 power_its: int = POWER_ITS, rho_safety: float = RHO_SAFETY, aggregation: str | None = None):
        if not isinstance(A, CSRMatrix):
            raise TypeError("the AMG preconditioner needs an assembled CSRMatrix")
        self.max_coarse, self.max_levels = int(max_coarse), min(int(max_levels), 16)
        self.power_its, self.rho_safety = int(power_its), float(rho_safety)
        if aggregation is None:
            aggregation = os.environ.get("TFEM_AMG_AGGREGATION", "auto")
        if aggregation not in ("auto", "mis1", "mis2"):
            raise ValueError("aggregation must be 'auto', 'mis1' or 'mis2'")
        self.aggregation = aggregation
        self.shape = (A.n, A.n)
        self.n = A.n
        self.device = A.device
        self.levels: list[_Level] = []
        self._work = None
        self._timing = None
        self._setup(A, symbolic=True)

    @classmethod
    def from_operator(cls, op: BlockOperator, max_coarse: int = MAX_COARSE_DOFS, max_levels: int = MAX_LEVELS,
                      aggregation: str = "auto") -> "AMGPreconditioner":
        """Hierarchy on a block operator that is itself a coarse operator (the gathered tail of the distributed
        hierarchy, damg.py): zero diagonals are repaired and the layout of the cycle is chosen for level 0 too."""
        self = cls.__new__(cls)
        self.max_coarse, self.max_levels = int(max_coarse), min(int(max_levels), 16)
        self.power_its, self.rho_safety = POWER_ITS, RHO_SAFETY
        self.aggregation = aggregation
        self.n = op.n_rows
        self.shape = (self.n, self.n)
        self.device = op.vals.device
        self.levels = []
        self._work = None
        self._timing = None
        self._setup(None, symbolic=True, op0=op)
        return self

    # ------------------------------------------------------------------------------------------ setup
    def _aggregate(self, lv) -> bool:
        """K12 + the pattern of P and R for level `lv`; False if the coarsening stalled."""
        dev, st, op, d, nb = self.device, L.stream(), lv.op, lv.d, lv.op.nbr
        agg = _empty(nb, torch.int32, dev)
        state, flag, index = _empty(nb, torch.int8, dev), _empty(nb, torch.uint8, dev), _empty(nb, torch.int32, dev)
        n_agg, rounds = ctypes.c_int64(0), ctypes.c_int32(0)
        # radius-1 aggregates; on graphs of low degree (Tetra1: ~3 nodes per aggregate) radius 2 instead
        for distance in ((1, 2) if self.aggregation == "auto" else ((1,) if self.aggregation == "mis1" else (2,))):
            L.check(L.lib.tfem_amg_aggregate(nb, L.ptr(op.bptr), L.ptr(op.bcol), distance, L.ptr(state), L.ptr(flag),
                                             L.ptr(index), L.ptr(agg), ctypes.byref(n_agg), ctypes.byref(rounds), st))
            lv.agg_distance = distance
            if nb >= MIN_AGG_SIZE * n_agg.value:
                break
        if n_agg.value * d >= 0.8 * lv.n:
            return False
        lv.agg, lv.n_agg, lv.mis_rounds = agg, int(n_agg.value), int(rounds.value)
        lv.max_row = int((op.bptr[1:] - op.bptr[:-1]).max().item())     # longest block row (staging capacity of K13)
        pptr = _empty(nb + 1, torch.int64, dev)
        L.check(L.lib.tfem_amg_prolongator_count(d, nb, L.ptr(op.bptr), L.ptr(op.bcol), L.ptr(agg), L.ptr(pptr), st))
        npb = int(pptr[-1].item())
        lv.P = BlockOperator(d, nb, lv.n_agg, pptr, _empty(npb, torch.int32, dev)[:npb],
                             _empty(d * d * npb, torch.float64, dev)[: d * d * npb])
        lv.R = None
        return True

    def _tick(self, name: str) -> None:
        """Per-phase wall times of the setup (TFEM_AMG_TIMING=1; synchronises, for tools/amg_check.py only)."""
        if self._timing is None:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        self._timing[name] = self._timing.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def _setup(self, A: CSRMatrix | None, symbolic: bool, op0: BlockOperator | None = None) -> None:
        dev, st = self.device, L.stream()
        self._timing = {} if os.environ.get("TFEM_AMG_TIMING") else None
        if self._timing is not None:
            torch.cuda.synchronize()
            self._t_last = time.perf_counter()
        if op0 is None:
            self._pattern_key = (A.indptr.data_ptr(), A.indices.data_ptr(), A.n)
            self._values_key = (A.values_.data_ptr(), A.values_._version)
            self._A = A   # keeps the level-0 buffers alive
            op = _level0_operator(A)
        else:
            self._pattern_key = self._values_key = self._A = None
            op = op0
        coarse0 = op0 is not None
        if symbolic:
            self.levels = []
        li = 0
        while True:
            if symbolic:
                lv = _Level()
                self.levels.append(lv)
                lv.op, lv.d, lv.n = op, op.d, op.n_rows
                lv.dinv, lv.iso = _empty(lv.n, torch.float64, dev), _empty(lv.n, torch.uint8, dev)
                lv.x, lv.b, lv.t = (torch.zeros(lv.n, dtype=torch.float64, device=dev) for _ in range(3))
            else:
                lv = self.levels[li]
                if li == 0:
                    lv.op = op
            op, d, nb = lv.op, lv.d, lv.op.nbr
            L.check(L.lib.tfem_amg_row_info(d, nb, L.ptr(op.bptr), L.ptr(op.bcol), L.ptr(op.vals),
                                            1 if (li > 0 or coarse0) else 0, L.ptr(lv.dinv), L.ptr(lv.iso), st))
            if li > 0 or coarse0:
                op.prepare()                                   # after the zero-diagonal repair
            self._tick("row_info+layout")
            if symbolic:
                if lv.n <= self.max_coarse or len(self.levels) >= self.max_levels:
                    break
            elif li == len(self.levels) - 1:
                break
            work = _empty(int(L.lib.tfem_amg_work_doubles(lv.n)), torch.float64, dev)
            rho = ctypes.c_double(0.0)
            ostruct, okeep = op.operator_struct()
            L.check(L.lib.tfem_amg_rho(ctypes.byref(ostruct), L.ptr(lv.dinv), self.power_its, L.ptr(work),
                                       ctypes.byref(rho), st))
            del work, okeep
            self._tick("rho")
            lv.rho = rho.value * self.rho_safety
            lv.omega = 4.0 / (3.0 * lv.rho)
            if symbolic and not self._aggregate(lv):
                break                                          # coarsening stalled: this level is the coarsest
            self._tick("aggregate+P pattern")
            P = lv.P
            L.check(L.lib.tfem_amg_prolongator_fill(d, nb, L.ptr(op.bptr), L.ptr(op.bcol), L.ptr(op.vals), L.ptr(lv.agg),
                                                    L.ptr(lv.dinv), L.ptr(lv.iso), lv.omega, L.ptr(P.bptr), L.ptr(P.bcol),
                                                    L.ptr(P.vals), lv.max_row, st))
            if lv.R is None:
                npb = P.nblk
                tptr = _empty(lv.n_agg + 1, torch.int64, dev)
                tcol, lv.tsrc = _empty(npb, torch.int32, dev)[:npb], _empty(npb, torch.int32, dev)[:npb]
                L.check(L.lib.tfem_amg_transpose_structure(nb, lv.n_agg, L.ptr(P.bptr), L.ptr(P.bcol), npb, L.ptr(tptr),
                                                           L.ptr(tcol), L.ptr(lv.tsrc), st))
                lv.R = BlockOperator(d, lv.n_agg, nb, tptr, tcol, torch.empty_like(P.vals))
            R = lv.R
            L.check(L.lib.tfem_amg_transpose_values(d, lv.n_agg, L.ptr(P.bptr), L.ptr(P.vals), L.ptr(R.bptr), L.ptr(R.bcol),
                                                    L.ptr(lv.tsrc), L.ptr(R.vals), st))
            self._tick("P values + R")
            P.prepare(), R.prepare()
            self._tick("P/R layout")
            if symbolic:
                AP, lv.ap_structure = spgemm(d, op, P)
                self._tick("A*P")
                op, lv.ac_structure = spgemm(d, R, AP)
            else:
                AP, _ = spgemm(d, op, P, lv.ap_structure)
                self._tick("A*P")
                spgemm(d, R, AP, lv.ac_structure, out_vals=self.levels[li + 1].op.vals)
            del AP
            self._tick("R*(AP)")
            li += 1
        # coarsest level: dense inverse (tiny; a library call like method="spsolve")
        c = self.levels[-1]
        if c.n > DENSE_LIMIT:
            raise RuntimeError(f"AMG coarsening stalled at {c.n} unknowns; the coarsest level is solved densely "
                               f"and is limited to {DENSE_LIMIT}")
        c.inv = torch.linalg.inv(_dense(c.op)).contiguous()
        self._build_structs()
        self._tick("coarse inverse")

    def _build_structs(self) -> None:
        n_levels = len(self.levels)
        arr = (L.AmgLevelStruct * n_levels)()
        keep = []
        for i, lv in enumerate(self.levels):
            arr[i].A, k = lv.op.operator_struct()
            keep.append(k)
            if i + 1 < n_levels:
                arr[i].P, kp = lv.P.operator_struct()
                arr[i].R, kr = lv.R.operator_struct()
                keep += [kp, kr]
                arr[i].omega = lv.omega
            arr[i].dinv = L.ptr(lv.dinv)
            arr[i].x, arr[i].b, arr[i].t = L.ptr(lv.x), L.ptr(lv.b), L.ptr(lv.t)
        self._structs, self._keep = arr, keep

    def resetup(self, A: CSRMatrix) -> None:
        """New coefficients: on the same sparsity pattern only the numeric phase is repeated (aggregates, patterns of
        P / R / A_c and the SELL structures are kept) — AmgX `resetup` in the reference (sparse.py:440-441)."""
        if A is self._A and (A.values_.data_ptr(), A.values_._version) == self._values_key:
            return   # the very same matrix object and coefficients (e.g. the adjoint solve with the converged tangent)
        same = (A.indptr.data_ptr(), A.indices.data_ptr(), A.n) == self._pattern_key
        self._setup(A, symbolic=not same)

    # ------------------------------------------------------------------------------------------ application
    @property
    def n_levels(self) -> int:
        return len(self.levels)

    @property
    def operator_complexity(self) -> float:
        return sum(lv.op.nblk * lv.d * lv.d for lv in self.levels) / (self.levels[0].op.nblk * self.levels[0].d ** 2)

    def apply(self, r: Tensor) -> Tensor:
        """z = M r (one V cycle)."""
        L.require_cuda(r)
        r = r.to(torch.float64).contiguous()
        z = torch.empty_like(r)
        L.check(L.lib.tfem_amg_vcycle(self._structs, self.n_levels, L.ptr(self.levels[-1].inv), L.ptr(r), L.ptr(z),
                                      L.stream()))
        return z

    def apply_block(self, R: Tensor) -> Tensor:
        """Z = M R for a block of vectors R [n, m] (row-major): one V cycle per column, 4 columns at a time with the two
        finest-level products of the cycle reading the matrix once for the 4 (`tfem_amg_vcycle_block`; 14.2 ms per 4 vectors
        against 4 x 4.17 ms at config B — restriction, coarse levels and prolongation stay per vector). Each column equals
        `apply(R[:, j])` bit for bit."""
        L.require_cuda(R)
        if R.dim() != 2 or R.shape[0] != self.n:
            raise ValueError("apply_block expects a block of shape [n, m]")
        R = R.to(torch.float64)
        m = int(R.shape[1])
        Z = torch.empty(R.shape, dtype=torch.float64, device=R.device)
        work = torch.empty(2 * self.n * min(m, 4), dtype=torch.float64, device=R.device) if m else None
        for j0 in range(0, m, 4):
            if m - j0 == 1:   # a lone last column: the single-vector cycle is faster than a pass with one live column
                Z[:, j0] = self.apply(R[:, j0].contiguous())
                break
            Rc = R[:, j0:j0 + 4].contiguous()
            Zc = torch.empty_like(Rc)
            L.check(L.lib.tfem_amg_vcycle_block(self._structs, self.n_levels, L.ptr(self.levels[-1].inv), Rc.shape[1],
                                                L.ptr(Rc), L.ptr(Zc), L.ptr(work), L.stream()))
            Z[:, j0:j0 + 4] = Zc
        return Z

    def solve(self, b: Tensor, x0: Tensor | None = None, rtol: float = 1e-10, atol: float = 0.0, maxiter: int = 0):
        """AMG-preconditioned CG. Returns (x, stats); raises RuntimeError("CG failed with exit code …") like the
        reference's Krylov paths (sparse.py:421) so that `FEM.solve` can cut the load step back (base.py:831)."""
        L.require_cuda(b)
        b = b.to(torch.float64).contiguous()
        x = torch.empty_like(b)
        nwork = int(L.lib.tfem_amg_work_doubles(self.n))
        if self._work is None or self._work.shape[0] != nwork:
            self._work = torch.empty(nwork, dtype=torch.float64, device=b.device)
        if x0 is not None:
            x0 = x0.to(device=b.device, dtype=torch.float64).contiguous()
        info = np.zeros(8, dtype=np.float64)
        rc = L.lib.tfem_amg_pcg_solve(self._structs, self.n_levels, L.ptr(self.levels[-1].inv), L.ptr(b), L.ptr(x0),
                                      float(rtol), float(atol), int(maxiter), L.ptr(x), L.ptr(self._work),
                                      info.ctypes.data, L.stream())
        stats = {"iterations": int(info[0]), "resnorm": float(info[1]), "bnorm": float(info[2]),
                 "converged": bool(info[3]), "spmv": int(info[4]), "launches": int(info[5])}
        if rc in (L.ERR_NOT_CONVERGED, L.ERR_BREAKDOWN):
            raise RuntimeError(f"CG failed with exit code {stats['iterations'] if rc == L.ERR_NOT_CONVERGED else -1}")
        L.check(rc)
        return x, stats


def _dense(op: BlockOperator) -> Tensor:
    """Dense copy of a (small) block operator."""
    d = op.d
    cnt = (op.bptr[1:] - op.bptr[:-1])
    out = torch.zeros(op.n_rows, op.n_cols, dtype=torch.float64, device=op.vals.device)
    node = torch.repeat_interleave(torch.arange(op.nbr, device=cnt.device), cnt)          # block -> block row
    slot = torch.arange(op.nblk, device=cnt.device) - op.bptr[:-1][node]                   # position inside the row
    m = cnt[node]
    base = d * d * op.bptr[:-1][node]
    for a in range(d):
        for c in range(d):
            v = op.vals[base + (a * m + slot) * d + c]
            out[node * d + a, op.bcol.long() * d + c] = v
    return out
