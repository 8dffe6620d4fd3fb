"""Device-resident sparsity pattern and CSR matrix — the host objects that own the buffers the
C-ABI kernels work on (torch tensors; the library only borrows pointers).

`Pattern` replaces the products of the reference's `FEM.__init__` (src/torchfem/base.py:73-119):
`glob_idx` / `k_map` / `diag_map`. `CSRMatrix` replaces the `torch.sparse_coo_tensor` returned by
`FEM.assemble_matrix` (base.py:421-426) and the CuPy CSR built from it on every solve
(src/torchfem/sparse.py:385-393); it answers the handful of sparse-tensor methods the reference's
callers use (`.T`, `._indices()`, `._values()`, `.shape`, `.to_dense()`, `@`).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib as L


def _i32(n, device):
    return torch.empty(int(n), dtype=torch.int32, device=device)


def _i64(n, device):
    return torch.empty(int(n), dtype=torch.int64, device=device)


class Pattern:
    """CSR sparsity pattern of a mesh with `dpn` DOFs per node, built once on the device (kernel K0).

    Attributes (device tensors): `indptr` int64 [n_dofs+1], `indices` int32 [nnz], `diag_map` int32
    [n_dofs], `node_ptr` int64 [n_nod+1] / `adj` int32 [nnzb] (node graph), `src_ptr` int64 [nnzb+1] /
    `src` int32 [n_elem*nn*nn] (element-slot -> CSR permutation of the deterministic assembly),
    `chunk_rows` int32 (SpMV plan).
    """

    def __init__(self, elements: Tensor, n_nod: int, dpn: int):
        L.require_cuda(elements)
        if elements.dtype != torch.int64:
            elements = elements.to(torch.int64)
        elements = elements.contiguous()
        dev = elements.device
        self.device = dev
        self.n_nod = int(n_nod)
        self.dpn = int(dpn)
        self.n_elem, self.nn = (int(s) for s in elements.shape)
        self.n_dofs = self.n_nod * self.dpn
        self.elements = elements
        st = L.stream()

        inc_ptr = _i32(self.n_nod + 1, dev)
        inc_list = _i32(max(1, self.n_elem * self.nn), dev)
        blk_cnt = _i32(self.n_nod, dev)
        totals = _i64(4, dev)
        L.check(L.lib.tfem_pattern_phase1(self.n_nod, self.n_elem, self.nn, self.dpn, L.ptr(elements),
                                          L.ptr(inc_ptr), L.ptr(inc_list), L.ptr(blk_cnt),
                                          L.ptr(totals), st))
        self.inc_ptr, self.inc_list = inc_ptr, inc_list   # node -> (element, local node) incidence (kept for K8)
        nnzb, nnz, max_blk, max_inc = (int(v) for v in totals.tolist())
        self.nnzb, self.nnz = nnzb, nnz
        self.max_blocks_per_node, self.max_elements_per_node = max_blk, max_inc

        self.node_ptr = _i64(self.n_nod + 1, dev)
        self.adj = _i32(max(1, nnzb), dev)
        self.indptr = _i64(self.n_dofs + 1, dev)
        self.indices = _i32(nnz, dev)
        self.diag_map = _i32(self.n_dofs, dev)
        self.src_ptr = _i64(nnzb + 1, dev)
        self.src = _i32(max(1, self.n_elem * self.nn * self.nn), dev)
        L.check(L.lib.tfem_pattern_phase2(self.n_nod, self.n_elem, self.nn, self.dpn, L.ptr(elements),
                                          L.ptr(inc_ptr), L.ptr(inc_list), L.ptr(blk_cnt),
                                          L.ptr(self.node_ptr), L.ptr(self.adj), L.ptr(self.indptr),
                                          L.ptr(self.indices), L.ptr(self.diag_map),
                                          L.ptr(self.src_ptr), L.ptr(self.src), st))
        self.chunk_rows = spmv_plan(self.indptr, self.n_dofs, nnz)
        self._k_map = None
        self._glob_idx = None
        self._diag_pos = None

    # -- reference-compatible views, materialised only on request (they are large) --------------
    @property
    def k_map(self) -> Tensor:
        """int32 [n_elem*(nn*dpn)^2] — the reference's `k_map` (base.py:94-104), bit-identical."""
        if self._k_map is None:
            nd = self.nn * self.dpn
            if self.nnz >= 2**31:
                raise RuntimeError("k_map is int32 in the reference; nnz >= 2^31 needs a partitioned mesh")
            out = _i32(self.n_elem * nd * nd, self.device)
            L.check(L.lib.tfem_pattern_k_map(self.n_nod, self.n_elem, self.nn, self.dpn,
                                             L.ptr(self.elements), L.ptr(self.node_ptr),
                                             L.ptr(self.adj), L.ptr(self.indptr), L.ptr(out), L.stream()))
            self._k_map = out
        return self._k_map

    @property
    def glob_idx(self) -> Tensor:
        """int64 [2, nnz] — the reference's `glob_idx` (base.py:110-118), bit-identical."""
        if self._glob_idx is None:
            out = torch.empty(2, self.nnz, dtype=torch.int64, device=self.device)
            L.check(L.lib.tfem_pattern_coo_rows(self.n_dofs, L.ptr(self.indptr), out[0].data_ptr(),
                                                L.stream()))
            out[1] = self.indices
            self._glob_idx = out
        return self._glob_idx

    @property
    def diag_pos(self) -> Tensor:
        if self._diag_pos is None:
            self._diag_pos = self.diag_map.to(torch.int64)
        return self._diag_pos

    @property
    def sell_structure(self):
        """(slice_ptr, sell_cols, padded_nnz) of the solver-internal SELL-32 layout, built once."""
        if getattr(self, "_sell", None) is None:
            has_orphans = self.nnz != self.dpn * self.dpn * self.nnzb
            block = None
            if self.dpn in (2, 3) and not has_orphans:
                block = (self.dpn, self.n_nod, self.node_ptr, self.adj)
            self._sell = sell_structure(self.indptr, self.indices, self.n_dofs, block)
        return self._sell

    def matrix(self, values: Tensor | None, symmetric: bool = True, sell_vals: Tensor | None = None) -> "CSRMatrix":
        """Wrap CSR values living on this pattern. `sell_vals`: the same values in the solver's SELL-32 order when the
        assembly wrote them (`assemble(..., sell_out=)`); `values=None` is a matrix that exists for the solver only."""
        A = CSRMatrix(self.indptr, self.indices, values, self.n_dofs, chunk_rows=self.chunk_rows,
                      diag_pos=self.diag_pos, symmetric=symmetric, sell_struct=self.sell_structure,
                      coo_indices=self._glob_idx)
        A._sell_vals = sell_vals
        return A


class SellStructure:
    """Pattern-level part of the solver-internal SELL-32 layout: slice offsets, and either scalar column
    indices (4 B/nnz) or node-block column indices (4/dpn^2 B/nnz, only for patterns built from a mesh
    with dpn in {2,3} and no unreferenced nodes). Built lazily, shared by every matrix on the pattern."""

    def __init__(self, indptr: Tensor, indices: Tensor, n: int, block=None, n_cols: int | None = None,
                 long_cap: int | None = None):
        self.indptr, self.indices, self.n = indptr, indices, int(n)
        self.n_cols = int(n) if n_cols is None else int(n_cols)   # rectangular operators (AMG P / R)
        # long rows (reference-point couplings, reference assembly.py:295-335) stay out of the slices and are computed on
        # the side path of the SpMV from the CSR arrays (`tfem_sell_t`); None: every row goes into its slice (AMG operators)
        self.long_cap = None if long_cap is None else int(long_cap)
        self.long_rows = None
        if self.long_cap is not None:
            lens = indptr[1:] - indptr[:-1]
            rows = torch.nonzero(lens > self.long_cap).ravel()
            if rows.numel():
                self.long_rows = rows.to(torch.int32).contiguous()
            else:
                self.long_cap = None
        n_slices = (self.n + 31) // 32
        self.slice_ptr = _i64(n_slices + 1, indptr.device)
        if self.long_cap is None:
            L.check(L.lib.tfem_sell_slice_ptr(self.n, L.ptr(indptr), L.ptr(self.slice_ptr), L.stream()))
        else:
            L.check(L.lib.tfem_sell_slice_ptr_capped(self.n, L.ptr(indptr), self.long_cap, L.ptr(self.slice_ptr),
                                                     L.stream()))
        self.padded = int(self.slice_ptr[-1].item())
        self._cols = None
        self.block = block  # (dpn, n_nod, node_ptr, adj) or None
        self._bslice_ptr = self._bcols = None

    @property
    def cols(self) -> Tensor:
        if self._cols is None:
            self._cols = _i32(max(self.padded, 4), self.indptr.device)
            L.check(L.lib.tfem_sell_fill_capped(self.n, self.n_cols, L.ptr(self.indptr), L.ptr(self.indices), None,
                                                self.fill_cap, L.ptr(self.slice_ptr), L.ptr(self._cols), None,
                                                L.stream()))
        return self._cols

    @property
    def fill_cap(self) -> int:
        return (1 << 62) if self.long_cap is None else self.long_cap

    @property
    def bcols(self):
        if self.block is None:
            return None
        if self._bcols is None:
            dpn, n_nod, node_ptr, adj = self.block
            n_slices = (self.n + 31) // 32
            self._bslice_ptr = _i64(n_slices + 1, self.indptr.device)
            L.check(L.lib.tfem_bsell_slice_ptr(self.n, dpn, L.ptr(self.slice_ptr), L.ptr(self._bslice_ptr),
                                               L.stream()))
            total = int(self._bslice_ptr[-1].item())
            self._bcols = _i32(max(total, 4), self.indptr.device)
            L.check(L.lib.tfem_bsell_fill(self.n, dpn, n_nod, L.ptr(node_ptr), L.ptr(adj),
                                          L.ptr(self._bslice_ptr), L.ptr(self._bcols), L.stream()))
        return self._bslice_ptr, self._bcols, self.block[0]


class SellMatrix:
    """A `tfem_sell_t` plus the tensors that keep its pointers alive."""

    def __init__(self, st: SellStructure, vals: Tensor, use_block: bool = True, csr_vals: Tensor | None = None):
        blk = st.bcols if use_block else None
        self.keep = [st.slice_ptr, vals]
        self.struct = L.SellStruct()
        self.struct.n_rows = st.n
        self.struct.slice_ptr = L.ptr(st.slice_ptr)
        self.struct.vals = L.ptr(vals)
        self.struct.n_long = 0
        if st.long_rows is not None:
            if csr_vals is None:
                raise ValueError("a SELL matrix with long rows needs the CSR values")
            self.keep += [st.long_rows, st.indptr, st.indices, csr_vals]
            self.struct.n_long = int(st.long_rows.numel())
            self.struct.long_rows, self.struct.csr_indptr = L.ptr(st.long_rows), L.ptr(st.indptr)
            self.struct.csr_cols, self.struct.csr_vals = L.ptr(st.indices), L.ptr(csr_vals)
        if blk is not None:
            bslice_ptr, bcols, dpn = blk
            self.keep += [bslice_ptr, bcols]
            self.struct.bslice_ptr, self.struct.bcols, self.struct.dpn = L.ptr(bslice_ptr), L.ptr(bcols), dpn
            self.struct.cols = None
        else:
            cols = st.cols
            self.keep.append(cols)
            self.struct.cols = L.ptr(cols)
            self.struct.bslice_ptr = self.struct.bcols = None
            self.struct.dpn = 0
        self.n = st.n
        self.block = blk is not None

    @property
    def ref(self):
        import ctypes

        return ctypes.byref(self.struct)


def sell_structure(indptr: Tensor, indices: Tensor, n: int, block=None) -> SellStructure:
    """SELL-32 structure of an assembled matrix (long rows on the side path)."""
    return SellStructure(indptr, indices, n, block, long_cap=L.SELL_LONG_ROW)


def spmv_plan(indptr: Tensor, n_rows: int, nnz: int) -> Tensor:
    n_chunks = int(L.lib.tfem_spmv_num_chunks(nnz))
    chunk_rows = _i32(n_chunks + 1, indptr.device)
    L.check(L.lib.tfem_spmv_plan(n_rows, nnz, L.ptr(indptr), L.ptr(chunk_rows), L.stream()))
    return chunk_rows


class CSRMatrix:
    """Square CSR matrix on the device: int64 `indptr`, int32 `indices`, float64 `values`.

    `symmetric=True` (FEM tangents) makes `.T` free. For general matrices `.T` builds the transposed
    copy with `tfem_csr_transpose` (needed by the adjoint of `differentiable_sparse_solve` with a
    non-symmetric A, reference sparse.py:203).
    """

    def __init__(self, indptr: Tensor, indices: Tensor, values: Tensor, n: int, *, chunk_rows=None,
                 diag_pos=None, symmetric=False, coo_indices=None, sell_struct=None):
        L.require_cuda(indptr, indices)
        if values is not None:
            L.require_cuda(values)
            if values.dtype != torch.float64:
                raise TypeError("torch-fem_b200 computes in float64 only (the reference runs in float64)")
        self.indptr, self.indices, self._values_csr = indptr, indices, values
        self.n = int(n)
        self.nnz = int(indices.shape[0])
        self.symmetric = symmetric
        self.chunk_rows = chunk_rows if chunk_rows is not None else spmv_plan(indptr, self.n, self.nnz)
        self._diag_pos = diag_pos
        self._coo = coo_indices
        self._T = None
        self._sell_struct = sell_struct
        self._sell_vals = None
        self._sell_mats = {}

    @property
    def values_(self) -> Tensor:
        if self._values_csr is None:
            raise ValueError("this matrix was assembled for the solver only (SELL-32 values); "
                             "assemble with csr=True for its CSR values")
        return self._values_csr

    @values_.setter
    def values_(self, v):
        self._values_csr = v

    # ---- torch-sparse look-alike surface used by the reference's callers
    @property
    def shape(self):
        return torch.Size((self.n, self.n))

    def size(self):
        return self.shape

    @property
    def ndim(self):
        return 2

    @property
    def device(self):
        return self.indptr.device

    @property
    def dtype(self):
        return torch.float64

    @property
    def is_cuda(self):
        return True

    def numel(self):
        return self.n * self.n

    def is_coalesced(self):
        return True

    def coalesce(self):
        return self

    def _values(self):
        return self.values_

    values = _values

    def _indices(self):
        if self._coo is None:
            out = torch.empty(2, self.nnz, dtype=torch.int64, device=self.device)
            L.check(L.lib.tfem_pattern_coo_rows(self.n, L.ptr(self.indptr), out[0].data_ptr(), L.stream()))
            out[1] = self.indices
            self._coo = out
        return self._coo

    indices_coo = _indices

    def to_sparse_coo(self) -> Tensor:
        with torch.sparse.check_sparse_tensor_invariants(False):
            return torch.sparse_coo_tensor(self._indices(), self.values_, size=(self.n, self.n),
                                           is_coalesced=True)

    def to_dense(self) -> Tensor:
        return self.to_sparse_coo().to_dense()

    def detach(self):
        return self

    @property
    def diag_pos(self) -> Tensor:
        if self._diag_pos is None:
            pos = _i64(self.n, self.device)
            L.check(L.lib.tfem_csr_diag_positions(self.n, L.ptr(self.indptr), L.ptr(self.indices),
                                                  L.ptr(pos), L.stream()))
            self._diag_pos = pos
        return self._diag_pos

    def diagonal(self) -> Tensor:
        pos = self.diag_pos
        d = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        ok = pos >= 0
        d[ok] = self.values_[pos[ok]]
        return d

    @property
    def T(self) -> "CSRMatrix":
        if self.symmetric:
            return self
        if self._T is None:
            dev = self.device
            t_indptr = _i64(self.n + 1, dev)
            t_indices = _i32(max(1, self.nnz), dev)[: self.nnz]
            t_vals = torch.empty(self.nnz, dtype=torch.float64, device=dev)
            L.check(L.lib.tfem_csr_transpose(self.n, self.n, self.nnz, L.ptr(self.indptr),
                                             L.ptr(self.indices), L.ptr(self.values_),
                                             L.ptr(t_indptr), L.ptr(t_indices), L.ptr(t_vals), L.stream()))
            self._T = CSRMatrix(t_indptr, t_indices, t_vals, self.n)
            self._T._T = self
        return self._T

    def sell(self, block: bool = True) -> SellMatrix:
        """The SELL-32 copy the Krylov kernels stream. The structure is shared with the pattern; the values
        are converted once per matrix (8 B/nnz read + write). `block=False` forces scalar column indices."""
        if self._sell_struct is None:
            self._sell_struct = sell_structure(self.indptr, self.indices, self.n)
        st = self._sell_struct
        if self._sell_vals is None:
            sv = torch.empty(max(st.padded, 2), dtype=torch.float64, device=self.device)
            L.check(L.lib.tfem_sell_fill_capped(self.n, self.n, L.ptr(self.indptr), None, L.ptr(self.values_), st.fill_cap,
                                                L.ptr(st.slice_ptr), None, L.ptr(sv), L.stream()))
            self._sell_vals = sv
        key = bool(block and st.block is not None)
        if self._sell_mats.get(key) is None:
            self._sell_mats[key] = SellMatrix(st, self._sell_vals, use_block=key, csr_vals=self._values_csr)
        return self._sell_mats[key]

    def matvec(self, x: Tensor, out: Tensor | None = None, fmt: str = "auto") -> Tensor:
        """y = A x with the K5 SpMV kernels. fmt: "sell" (block columns when the pattern has them),
        "sell-scalar" (scalar columns, 12 B/nnz), "csr" (CSR-chunk kernel, no conversion; right for a
        one-off product), "auto" = "sell" if that copy already exists else "csr"."""
        L.require_cuda(x)
        x = x.contiguous()
        if x.dtype != torch.float64 or x.shape != (self.n,):
            raise ValueError("matvec expects a float64 vector of length n")
        y = out if out is not None else torch.empty_like(x)
        if fmt in ("sell", "sell-scalar") or (fmt == "auto" and self._sell_vals is not None):
            S = self.sell(block=(fmt != "sell-scalar"))
            L.check(L.lib.tfem_sell_spmv(S.ref, L.ptr(x), L.ptr(y), L.stream()))
        else:
            L.check(L.lib.tfem_spmv(self.n, self.nnz, L.ptr(self.indptr), L.ptr(self.indices),
                                    L.ptr(self.values_), L.ptr(self.chunk_rows), L.ptr(x), L.ptr(y),
                                    L.stream()))
        return y

    def matmat(self, X: Tensor, out: Tensor | None = None) -> Tensor:
        """Y = A X for a block of vectors X [n, m] (row-major) with the multi-vector SELL kernel: the matrix is read once
        per 4 columns. Each column equals `matvec(X[:, j], fmt="sell")` bit for bit."""
        L.require_cuda(X)
        if X.dtype != torch.float64 or X.dim() != 2 or X.shape[0] != self.n:
            raise ValueError("matmat expects a float64 block of shape [n, m]")
        X = X.contiguous()
        m = int(X.shape[1])
        Y = out if out is not None else torch.empty_like(X)
        if Y.shape != X.shape or not Y.is_contiguous() or Y.data_ptr() == X.data_ptr():
            raise ValueError("matmat: `out` must be a separate contiguous block of X's shape")
        S = self.sell()
        if m == 0:
            return Y
        if S.struct.n_long > 0:   # long rows live on the single-vector side path
            for j in range(m):
                Y[:, j] = self.matvec(X[:, j].contiguous(), fmt="sell")
            return Y
        if m == 1:     # the single-vector kernel is faster than a pass with one live column (1.15 vs 1.44 ms at config B)
            Y[:, 0] = self.matvec(X[:, 0].contiguous(), fmt="sell")
            return Y
        if m <= 4:
            L.check(L.lib.tfem_sell_spmm(S.ref, m, L.ptr(X), m, L.ptr(Y), m, L.stream()))
            return Y
        # wider blocks: 4 columns per pass, packed — a gathered row of X must be one full sector; with the leading
        # dimension of a wide block the product gets slower than single products (tools/time_spmm.py)
        for j0 in range(0, m, 4):
            Xc = X[:, j0:j0 + 4].contiguous()
            Yc = torch.empty_like(Xc)
            L.check(L.lib.tfem_sell_spmm(S.ref, Xc.shape[1], L.ptr(Xc), Xc.shape[1], L.ptr(Yc), Xc.shape[1], L.stream()))
            Y[:, j0:j0 + 4] = Yc
        return Y

    def __matmul__(self, x: Tensor) -> Tensor:
        """A @ x, differentiable w.r.t. x (the reference's `self.M @ du` on a sparse COO tensor, base.py:1483)."""
        shape = x.shape
        y = _Matvec.apply(x.reshape(-1), self) if x.requires_grad else self.matvec(x.reshape(-1).contiguous())
        return y.reshape(shape)

    # ---- linear combinations of matrices on the SAME pattern (e.g. `M + 0.5*dt*K`, reference base.py:1513; the
    # reference lets torch / scipy sum the duplicate COO entries, here the values are combined entry by entry)
    def _like(self, values: Tensor) -> "CSRMatrix":
        return CSRMatrix(self.indptr, self.indices, values, self.n, chunk_rows=self.chunk_rows,
                         diag_pos=self._diag_pos, symmetric=self.symmetric, coo_indices=self._coo,
                         sell_struct=self._sell_struct)

    def __mul__(self, c) -> "CSRMatrix":
        return self._like(self.values_ * float(c))

    __rmul__ = __mul__

    def __neg__(self) -> "CSRMatrix":
        return self._like(-self.values_)

    def __add__(self, other: "CSRMatrix") -> "CSRMatrix":
        if not isinstance(other, CSRMatrix):
            return NotImplemented
        if other.indices.data_ptr() != self.indices.data_ptr() and not (
                other.nnz == self.nnz and torch.equal(other.indptr, self.indptr) and torch.equal(other.indices, self.indices)):
            raise ValueError("CSRMatrix addition needs both matrices on the same sparsity pattern")
        out = self._like(self.values_ + other.values_)
        out.symmetric = self.symmetric and other.symmetric
        return out

    def __sub__(self, other: "CSRMatrix") -> "CSRMatrix":
        return self + (-other)

    # ---- construction from what the reference passes around
    @staticmethod
    def from_coo(A: Tensor) -> "CSRMatrix":
        """From a torch sparse COO tensor (possibly uncoalesced, e.g. `M + 0.5*dt*K`, or a
        transposed view `K.T`): duplicates are summed like scipy's `coo_matrix.tocsr` does on the
        reference's CPU path (sparse.py:462-464) and `A.coalesce()` on its GPU path (:379-384)."""
        L.require_cuda(A)
        if A.layout != torch.sparse_coo:
            A = A.to_sparse_coo()
        A = A.detach().coalesce()
        n = int(A.shape[0])
        row, col = A.indices()
        vals = A.values().to(torch.float64).contiguous()
        counts = torch.bincount(row, minlength=n)
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=A.device)
        indptr[1:] = torch.cumsum(counts, 0)
        return CSRMatrix(indptr, col.to(torch.int32).contiguous(), vals, n,
                         coo_indices=A.indices())


class _Matvec(torch.autograd.Function):
    """y = A x with dy/dx^T g = A^T g; A itself carries no gradient (its values come from the kernels)."""

    @staticmethod
    def forward(ctx, x: Tensor, A: "CSRMatrix"):
        ctx.A = A
        return A.matvec(x.detach().to(torch.float64).contiguous())

    @staticmethod
    def backward(ctx, g: Tensor):
        return _Matvec.apply(g, ctx.A.T) if g.requires_grad else ctx.A.T.matvec(g.contiguous()), None


def integrate_k(kind: int, bref: Tensor, w: Tensor, nodes: Tensor, elements: Tensor, tangent: Tensor,
                scale: Tensor | None = None, check: bool = True) -> Tensor:
    """Element matrices with kernel K1 (see include/tfem_b200.h `tfem_integrate_k`).

    bref: [n_int, dim, nn] = etype.B(etype.ipoints) (host or device, any float dtype);
    w: [n_int] = etype.iweights; tangent: [n_elem, d,d,d,d] | [n_int, n_elem, d,d,d,d] (mechanics) or
    [n_elem, d, d] | [n_int, n_elem, d, d] (heat). Raises the reference's
    ValueError("Negative Jacobian. Check element numbering.") (base.py:311-312) when `check`.
    """
    L.require_cuda(nodes, elements, tangent)
    if nodes.dtype != torch.float64 or tangent.dtype != torch.float64:
        raise TypeError("torch-fem_b200 computes in float64 only (the reference runs in float64)")
    n_int, dim, nn = (int(s) for s in bref.shape)
    n_elem = int(elements.shape[0])
    dpn = dim if kind == L.KIND_MECH else 1
    base_nd = 5 if kind == L.KIND_MECH else 3
    per_gp = tangent.dim() == base_nd + 1
    if tangent.dim() not in (base_nd, base_nd + 1):
        raise ValueError("tangent has the wrong number of dimensions")
    bref_h = np.ascontiguousarray(bref.detach().cpu().numpy(), dtype=np.float64)
    w_h = np.ascontiguousarray(w.detach().cpu().numpy(), dtype=np.float64)
    nodes = nodes.contiguous()
    elements = elements.to(torch.int64).contiguous()
    tangent = tangent.contiguous()
    if scale is not None:
        scale = scale.to(torch.float64).contiguous()
    nd = nn * dpn
    k = torch.empty(n_elem, nd, nd, dtype=torch.float64, device=nodes.device)
    flag = torch.zeros(1, dtype=torch.int32, device=nodes.device)
    L.check(L.lib.tfem_integrate_k(kind, dim, nn, n_int, bref_h.ctypes.data, w_h.ctypes.data,
                                   L.ptr(nodes), L.ptr(elements), n_elem, L.ptr(tangent),
                                   1 if per_gp else 0, L.ptr(scale), L.ptr(k), L.ptr(flag), L.stream()))
    if check and int(flag.item()) != 0:
        raise ValueError("Negative Jacobian. Check element numbering.")
    return k


def assemble(pattern: Pattern, k: Tensor, is_con: Tensor | None, out: Tensor | None = None,
             ubc: Tensor | None = None, lift: Tensor | None = None, *, csr: bool = True,
             sell_out: Tensor | bool | None = None, dinv_out: Tensor | bool | None = None):
    """CSR values from element matrices with kernel K2/K3 (deterministic; Dirichlet rows/cols fused).
    With `ubc` (prescribed values, [n_dofs]) and `lift` ([n_dofs] out) the kernel also returns the Dirichlet
    lifting K[free, con] @ ubc[con] (the right-hand side of the first Newton step, base.py:708-741).

    `sell_out` / `dinv_out` (a tensor, or True to allocate): the same pass also writes the values in the solver's
    SELL-32 order and 1/diagonal, so that neither the CSR -> SELL copy nor the Jacobi setup runs afterwards;
    `csr=False` skips the CSR values. Returns the CSR values, or (vals | None, sell_vals | None, dinv | None) when one
    of the solver outputs was asked for."""
    L.require_cuda(k)
    nd = pattern.nn * pattern.dpn
    if k.dtype != torch.float64:
        raise TypeError("torch-fem_b200 computes in float64 only (the reference runs in float64)")
    if tuple(k.shape) != (pattern.n_elem, nd, nd):
        raise ValueError(f"k must have shape {(pattern.n_elem, nd, nd)}")
    k = k.contiguous()
    if is_con is not None:
        is_con = is_con.to(torch.uint8).contiguous() if is_con.dtype != torch.uint8 else is_con.contiguous()
    solver_outputs = (sell_out is not None and sell_out is not False) or (dinv_out is not None and dinv_out is not False)
    if not csr and not solver_outputs:
        raise ValueError("assemble: nothing to write (csr=False needs sell_out)")
    vals = None
    if csr:
        vals = out if out is not None else torch.empty(pattern.nnz, dtype=torch.float64, device=k.device)
    if lift is not None:
        if is_con is None or ubc is None:
            raise ValueError("the Dirichlet lifting needs is_con and ubc")
        L.require_cuda(ubc, lift)
        ubc = ubc.to(torch.float64).contiguous()
    if not solver_outputs:
        L.check(L.lib.tfem_assemble_bc(pattern.n_nod, pattern.nn, pattern.dpn, L.ptr(pattern.node_ptr),
                                       L.ptr(pattern.adj), L.ptr(pattern.indptr), L.ptr(pattern.src_ptr),
                                       L.ptr(pattern.src), L.ptr(k), L.ptr(is_con),
                                       L.ptr(ubc) if lift is not None else None, L.ptr(vals), L.ptr(lift), L.stream()))
        return vals
    st = pattern.sell_structure
    sv = dinv = None
    if sell_out is not None and sell_out is not False:
        if st.long_rows is not None:
            raise ValueError("assemble: patterns with long rows keep their values in CSR order (SELL side path)")
        sv = sell_out if isinstance(sell_out, Tensor) else torch.empty(max(st.padded, 2), dtype=torch.float64,
                                                                       device=k.device)
        if sv.dtype != torch.float64 or sv.numel() < st.padded or not sv.is_contiguous():
            raise ValueError(f"sell_out must be a contiguous float64 tensor of at least {st.padded} entries")
    if dinv_out is not None and dinv_out is not False:
        dinv = dinv_out if isinstance(dinv_out, Tensor) else torch.empty(pattern.n_dofs, dtype=torch.float64,
                                                                         device=k.device)
    L.check(L.lib.tfem_assemble_solve(pattern.n_nod, pattern.nn, pattern.dpn, L.ptr(pattern.node_ptr),
                                      L.ptr(pattern.adj), L.ptr(pattern.indptr), L.ptr(pattern.src_ptr),
                                      L.ptr(pattern.src), L.ptr(k), L.ptr(is_con),
                                      L.ptr(ubc) if lift is not None else None, L.ptr(vals), L.ptr(lift),
                                      L.ptr(st.slice_ptr), L.ptr(sv), L.ptr(dinv), L.stream()))
    return vals, sv, dinv


class _AssembleRhs(torch.autograd.Function):
    """F = sum of the element vectors at the global DOFs (kernel `tfem_assemble_rhs`: deterministic gather over the
    pattern's incidence lists); backward = the transposed map, a plain gather g[idx]."""

    @staticmethod
    def forward(ctx, f: Tensor, pattern: "Pattern"):
        ctx.pattern = pattern
        ctx.shape = f.shape
        fe = f.detach().to(torch.float64).contiguous()
        F = torch.empty(pattern.n_dofs, dtype=torch.float64, device=fe.device)
        L.check(L.lib.tfem_assemble_rhs(pattern.n_nod, pattern.dpn, L.ptr(pattern.inc_ptr), L.ptr(pattern.inc_list),
                                        L.ptr(fe), L.ptr(F), L.stream()))
        return F.to(f.dtype)

    @staticmethod
    def backward(ctx, g: Tensor):
        p = ctx.pattern
        return g.view(p.n_nod, p.dpn)[p.elements].reshape(ctx.shape), None


def assemble_rhs(pattern: "Pattern", f: Tensor) -> Tensor:
    """Global vector [n_dofs] from element vectors [n_elem, nn*dpn] (reference base.py:428-445), bitwise
    reproducible and differentiable w.r.t. f (also twice: the backward is an indexing op)."""
    L.require_cuda(f)
    if f.numel() != pattern.n_elem * pattern.nn * pattern.dpn:
        raise ValueError(f"f must hold {pattern.n_elem} x {pattern.nn * pattern.dpn} element values")
    return _AssembleRhs.apply(f, pattern)


class ElementOperator:
    """Matrix-free operator y = sum_e P_e^T k_e P_e x on stored element matrices (kernel K8) with the Dirichlet
    masking of `assemble` applied on the fly — the optional operator of the Krylov solve: no assembly, no SELL
    copy. Accepted by `krylov_solve` in place of a `CSRMatrix`. Symmetric by construction."""

    def __init__(self, pattern: Pattern, k: Tensor, is_con: Tensor | None = None):
        L.require_cuda(k)
        nd = pattern.nn * pattern.dpn
        if k.dtype != torch.float64 or tuple(k.shape) != (pattern.n_elem, nd, nd):
            raise ValueError(f"k must be float64 with shape {(pattern.n_elem, nd, nd)}")
        self.pattern, self.k = pattern, k.contiguous()
        self.is_con = None if is_con is None else is_con.to(torch.uint8).contiguous()
        self.n = pattern.n_dofs
        self.shape = torch.Size((self.n, self.n))
        self.device = k.device
        self.symmetric = True
        self.struct = L.EbeStruct()
        self.struct.n_nod, self.struct.nn, self.struct.dpn = pattern.n_nod, pattern.nn, pattern.dpn
        self.struct.inc_ptr, self.struct.inc_list = L.ptr(pattern.inc_ptr), L.ptr(pattern.inc_list)
        self.struct.elements, self.struct.k = L.ptr(pattern.elements), L.ptr(self.k)
        self.struct.is_con = L.ptr(self.is_con)

    @property
    def ref(self):
        import ctypes

        return ctypes.byref(self.struct)

    @property
    def T(self):
        return self

    def matvec(self, x: Tensor, out: Tensor | None = None) -> Tensor:
        L.require_cuda(x)
        x = x.contiguous()
        y = out if out is not None else torch.empty_like(x)
        L.check(L.lib.tfem_ebe_spmv(self.ref, L.ptr(x), L.ptr(y), L.stream()))
        return y

    __matmul__ = matvec

    def diagonal(self) -> Tensor:
        d = torch.empty(self.n, dtype=torch.float64, device=self.device)
        L.check(L.lib.tfem_ebe_diag(self.ref, L.ptr(d), L.stream()))
        return d


class JacobiPreconditioner:
    """M = diag(A)^-1 (reference GPU path: `cupy_diags(1.0 / A_cp.diagonal())`, sparse.py:408-409).
    Returned by `sparse_solve` as `M` and accepted back, like the reference's preconditioner object."""

    def __init__(self, A=None, dinv: Tensor | None = None):
        if dinv is not None:   # written by the assembly (`assemble(..., dinv_out=)`)
            self.dinv = dinv
            self.shape = (dinv.shape[0], dinv.shape[0])
            return
        if isinstance(A, ElementOperator):
            self.dinv = 1.0 / A.diagonal()
        else:
            self.dinv = torch.empty(A.n, dtype=torch.float64, device=A.device)
            L.check(L.lib.tfem_jacobi_setup(A.n, L.ptr(A.values_), L.ptr(A.diag_pos), L.ptr(self.dinv),
                                            L.stream()))
        self.shape = (A.n, A.n)


_WORK_CACHE: dict = {}


def krylov_solve(A: "CSRMatrix | ElementOperator", b: Tensor, method: str = "cg", rtol: float = 1e-10, atol: float = 0.0,
                 x0: Tensor | None = None, M: JacobiPreconditioner | None = None, maxiter: int = 0,
                 check_every: int = 0):
    """Jacobi-preconditioned CG / MINRES on the device (kernels K5+K6). Returns (x, M, info dict).
    Raises RuntimeError("CG failed with exit code …") / ("minres failed …") like sparse.py:413,421."""
    L.require_cuda(b)
    if b.dtype != torch.float64:
        raise TypeError("torch-fem_b200 computes in float64 only (the reference runs in float64)")
    b = b.contiguous()
    if M is None:
        M = JacobiPreconditioner(A)
    x = torch.empty_like(b)
    nwork = int(L.lib.tfem_krylov_work_doubles(A.n))
    key = (b.device, nwork)
    work = _WORK_CACHE.get(key)
    if work is None:
        _WORK_CACHE.clear()
        work = torch.empty(nwork, dtype=torch.float64, device=b.device)
        _WORK_CACHE[key] = work
    info = np.zeros(8, dtype=np.float64)
    meth = {"cg": L.METHOD_CG, "minres": L.METHOD_MINRES}[method]
    if x0 is not None:
        x0 = x0.to(device=b.device, dtype=torch.float64).contiguous()
    if isinstance(A, ElementOperator):
        rc = L.lib.tfem_krylov_solve_ebe(meth, A.ref, L.ptr(M.dinv), L.ptr(b), L.ptr(x0), float(rtol), float(atol),
                                         int(maxiter), int(check_every), L.ptr(x), L.ptr(work),
                                         info.ctypes.data, L.stream())
    else:
        S = A.sell()
        rc = L.lib.tfem_krylov_solve(meth, S.ref, L.ptr(M.dinv), L.ptr(b), L.ptr(x0), float(rtol), float(atol),
                                     int(maxiter), int(check_every), L.ptr(x), L.ptr(work),
                                     info.ctypes.data, L.stream())
    stats = {"iterations": int(info[0]), "resnorm": float(info[1]), "bnorm": float(info[2]),
             "converged": bool(info[3]), "spmv": int(info[4]), "launches": int(info[5])}
    if rc in (L.ERR_NOT_CONVERGED, L.ERR_BREAKDOWN):
        name = "CG" if method == "cg" else "minres"
        raise RuntimeError(f"{name} failed with exit code {stats['iterations'] if rc == L.ERR_NOT_CONVERGED else -1}")
    L.check(rc)
    return x, M, stats


def adjoint_matrix_grad(A: CSRMatrix, lam: Tensor, x: Tensor) -> Tensor:
    """g[p] = -lam[row(p)] * x[col(p)] on A's pattern (kernel K7; reference sparse.py:212-216)."""
    L.require_cuda(lam, x)
    if lam.shape != (A.n,) or x.shape != (A.n,):
        raise ValueError("adjoint_matrix_grad expects two vectors of length n")
    # the kernel reads float64: a float32 solve (b.dtype float32) must not be reinterpreted
    lam = lam.detach().to(device=A.device, dtype=torch.float64).contiguous()
    x = x.detach().to(device=A.device, dtype=torch.float64).contiguous()
    g = torch.empty(A.nnz, dtype=torch.float64, device=A.device)
    L.check(L.lib.tfem_adjoint_matrix_grad(A.n, L.ptr(A.indptr), L.ptr(A.indices), L.ptr(lam), L.ptr(x), L.ptr(g),
                                           L.stream()))
    return g
