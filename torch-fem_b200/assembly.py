"""Several models coupled by kinematic constraints — the drop-in for the reference's `torchfem.assembly`
(src/torchfem/assembly.py:39-608; SURVEY §8(f) rank 4: the remaining caller of `assemble_matrix` / `newton_solve`).

Same public surface: `ReferencePoint`, `ReferencePointHeat`, `Assembly(parts)`, `Assembly.coupling(...)`,
`Assembly.solve(...)` with the reference's argument lists, return layout (one list entry per part) and error messages.
Couplings follow `u_secondary = u_primary + theta_primary x (x_secondary - x_primary)` (assembly.py:118-122) and are
enforced by eliminating the secondary DOFs: `u = T q`, reduced tangent `T^T K T` (assembly.py:258-334).

What is different underneath (everything stays on the GPU, all products run in this library's kernels):
* the part tangents are the device CSR matrices of kernel K2/K3; the global block-diagonal operator is their
  concatenation (reference points contribute explicit zero diagonals, which is how the reduced matrix gets a stored
  diagonal everywhere — the reference appends one afterwards, assembly.py:319-324);
* `T^T (K T)` is two runs of the hash SpGEMM K15 (`tfem_amg_spgemm_*`, scalar blocks) instead of
  `torch.sparse.mm`; both symbolic phases depend on the meshes and couplings only and are kept, so a Newton
  iteration / load increment / later solve pays the numeric phase alone, and a linear assembly whose parts reuse
  their tangent reuses the reduced matrix too;
* `T q` and `T^T f` are the CSR SpMV K5 (`tfem_spmv`, fixed summation order: the many-to-one reduction into a
  reference point's row involves no floating-point atomics), wrapped so that autograd sees them;
* the reduced system goes to `sparse.newton_solve` as a `CSRMatrix`, i.e. to the Krylov / AMG kernels.

Not carried over: plotting (`plot`, `plot2d`, `plot3d`: SURVEY §2 marks plotting out of scope) and shells as parts
(`Shell` is not on the hot path); any `Solid` / `Planar` / `SolidHeat` / `PlanarHeat` model and both kinds of
reference point can be coupled. Like the reference, an assembly solves on the reference configuration (no `nlgeom`)
and does not cut a load increment back.
"""
from __future__ import annotations

import math
from collections.abc import Iterable, Sequence

import torch
from torch import Tensor

from . import _lib as L
from .amg import BlockOperator, spgemm
from .base import FEM, Heat
from .csr import CSRMatrix, sell_structure, spmv_plan
from .sparse import describe_method, newton_solve


def _empty_index(device=None) -> Tensor:
    return torch.empty(0, dtype=torch.int64, device=device)


# An empty constraint set: `assemble_matrix(k, EMPTY)` leaves a part's matrix unmasked (reference assembly.py:19-20).
EMPTY = torch.empty(0, dtype=torch.int64, device="cpu")


class ReferencePoint:
    """A free node with rigid-body DOFs and no stiffness (reference assembly.py:39-73): three translations + three
    rotations in 3D, two + one in 2D. `forces`, `displacements`, `constraints` have shape [1, n_dofs]."""

    n_nod = 1

    def __init__(self, position: Tensor | Sequence[float]):
        self.nodes = torch.as_tensor(position).reshape(1, -1)
        if not torch.is_floating_point(self.nodes):
            self.nodes = self.nodes.to(torch.get_default_dtype())
        dim = self.nodes.shape[1]
        self.n_dof_per_node = dim + (3 if dim == 3 else 1)
        self.n_dofs = self.n_dof_per_node
        kw = dict(dtype=self.nodes.dtype, device=self.nodes.device)
        self._neumann = torch.zeros(1, self.n_dofs, **kw)
        self._dirichlet = torch.zeros(1, self.n_dofs, **kw)
        self._constraints = torch.zeros(1, self.n_dofs, dtype=torch.bool, device=self.nodes.device)

    @property
    def forces(self) -> Tensor:
        return self._neumann

    @property
    def displacements(self) -> Tensor:
        return self._dirichlet

    @property
    def constraints(self) -> Tensor:
        return self._constraints

    def __repr__(self) -> str:
        at = ", ".join(f"{x:g}" for x in self.nodes[0].tolist())
        return f"<torch-fem reference point at ({at})>"


class ReferencePointHeat:
    """A free node with one temperature and no heat capacity (reference assembly.py:76-106). `heat_flux`,
    `temperatures`, `constraints` have shape [1, 1]."""

    n_nod = 1
    n_dof_per_node = 1
    n_dofs = 1

    def __init__(self, position: Tensor | Sequence[float]):
        self.nodes = torch.as_tensor(position).reshape(1, -1)
        if not torch.is_floating_point(self.nodes):
            self.nodes = self.nodes.to(torch.get_default_dtype())
        kw = dict(dtype=self.nodes.dtype, device=self.nodes.device)
        self._neumann = torch.zeros(1, 1, **kw)
        self._dirichlet = torch.zeros(1, 1, **kw)
        self._constraints = torch.zeros(1, 1, dtype=torch.bool, device=self.nodes.device)

    @property
    def heat_flux(self) -> Tensor:
        return self._neumann

    @property
    def temperatures(self) -> Tensor:
        return self._dirichlet

    @property
    def constraints(self) -> Tensor:
        return self._constraints

    def __repr__(self) -> str:
        at = ", ".join(f"{x:g}" for x in self.nodes[0].tolist())
        return f"<torch-fem thermal reference point at ({at})>"


Part = FEM | ReferencePoint | ReferencePointHeat


# ------------------------------------------------------------------------------------------------ device operators
class _RectCSR:
    """Rectangular CSR operator on the device (the elimination map T and its transpose): int64 `indptr`, int32
    `indices` sorted per row, float64 `values`; products with kernel K5 (`tfem_spmv`)."""

    def __init__(self, n_rows: int, n_cols: int, indptr: Tensor, indices: Tensor, values: Tensor):
        L.require_cuda(indptr, indices, values)
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.indptr = indptr.contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.values = values.to(torch.float64).contiguous()
        self.nnz = int(self.values.shape[0])
        self.chunk_rows = spmv_plan(self.indptr, self.n_rows, self.nnz)
        self.transposed: "_RectCSR | None" = None

    @staticmethod
    def from_coo(rows: Tensor, cols: Tensor, values: Tensor, n_rows: int, n_cols: int) -> "_RectCSR":
        """Entries in any order, duplicates summed (sorted by (row, col) through `coalesce`)."""
        with torch.sparse.check_sparse_tensor_invariants(False):
            A = torch.sparse_coo_tensor(torch.stack([rows, cols]), values, (n_rows, n_cols)).coalesce()
        r, c = A.indices()
        indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=values.device)
        indptr[1:] = torch.cumsum(torch.bincount(r, minlength=n_rows), 0)
        return _RectCSR(n_rows, n_cols, indptr, c, A.values())

    def matvec(self, x: Tensor) -> Tensor:
        x = x.detach().to(torch.float64).contiguous()
        if x.shape != (self.n_cols,):
            raise ValueError(f"operator with {self.n_cols} columns applied to a vector of shape {tuple(x.shape)}")
        L.require_cuda(x)
        y = torch.empty(self.n_rows, dtype=torch.float64, device=x.device)
        L.check(L.lib.tfem_spmv(self.n_rows, self.nnz, L.ptr(self.indptr), L.ptr(self.indices), L.ptr(self.values),
                                L.ptr(self.chunk_rows), L.ptr(x), L.ptr(y), L.stream()))
        return y

    def block_operator(self) -> BlockOperator:
        return BlockOperator(1, self.n_rows, self.n_cols, self.indptr, self.indices, self.values)

    def __matmul__(self, x: Tensor) -> Tensor:
        """Differentiable w.r.t. x: the adjoint re-evaluates the residual through `T q` and `T^T f`."""
        if torch.is_grad_enabled() and x.requires_grad:
            return _RectMatvec.apply(x, self)
        return self.matvec(x)


class _RectMatvec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, op: _RectCSR):
        ctx.op = op
        return op.matvec(x)

    @staticmethod
    def backward(ctx, g: Tensor):
        return ctx.op.transposed @ g, None


def _blocks_from_coo(rows: Tensor, cols: Tensor, values: Tensor, nbr: int, nbc: int, d: int) -> BlockOperator:
    """Block-CSR operator (d x d blocks, `BlockOperator` value layout) from scalar COO entries without duplicates;
    blocks are created wherever an entry falls and padded with zeros."""
    I, a = rows // d, rows % d
    J, c = cols // d, cols % d
    ukey, inv = torch.unique(I * nbc + J, return_inverse=True)          # sorted by (block row, block column)
    bptr = torch.zeros(nbr + 1, dtype=torch.int64, device=rows.device)
    bptr[1:] = torch.cumsum(torch.bincount(ukey // nbc, minlength=nbr), 0)
    m = (bptr[1:] - bptr[:-1])[I]                                      # blocks in the entry's block row
    pos = d * d * bptr[I] + (a * m + (inv - bptr[I])) * d + c
    vals = torch.zeros(d * d * int(ukey.shape[0]), dtype=torch.float64, device=rows.device)
    vals.index_add_(0, pos, values.to(torch.float64))
    return BlockOperator(d, nbr, nbc, bptr, (ukey % nbc).to(torch.int32).contiguous(), vals)


def _scalar_columns(op: BlockOperator) -> Tensor:
    """int32 scalar CSR column indices of a block operator, in its value order (block row I with m blocks: scalar
    row a holds the columns d*J_s + c for s = 0..m-1, c = 0..d-1)."""
    d, dev = op.d, op.bptr.device
    cnt = op.bptr[1:] - op.bptr[:-1]
    row_cols = (op.bcol.to(torch.int64)[:, None] * d + torch.arange(d, device=dev)).reshape(-1)   # one scalar row each
    I = torch.repeat_interleave(torch.arange(op.nbr, device=dev), cnt * (d * d))
    off = torch.arange(I.shape[0], device=dev) - (d * d) * op.bptr[I]
    return row_cols[d * op.bptr[I] + off % (cnt[I] * d)].to(torch.int32).contiguous()


class _Elimination:
    """Everything about `u = T q` that depends on the meshes and couplings only: T and T^T on the device, the
    block-diagonal pattern of the part tangents in global numbering, and the symbolic phases of K T and T^T (K T).

    When every part has d = 2 or 3 DOFs per node (a 3-D reference point counts as two nodes of three) and the couplings
    eliminate whole nodes, all operators are d x d-blocked over nodes: the products then run on node blocks (`d` below),
    and the reduced matrix keeps the node-block SELL-32 layout (4/d^2 B of index per nonzero) and a block structure the
    AMG kernels can coarsen. Otherwise (`dofs=[...]` subsets, thermal parts, a 2-D point next to 2-DOF nodes) the
    operators are scalar CSR (d = 1)."""

    def __init__(self, asm: "Assembly", allow_blocks: bool = True):
        dev = asm.device
        rows, cols, vals, retained = asm._triplets()
        n, n_ret = asm.n_dofs, int(retained.shape[0])
        self.retained = retained
        self.n_retained = n_ret
        self.T = _RectCSR.from_coo(rows, cols, vals, n, n_ret)
        self.Tt = _RectCSR.from_coo(cols, rows, vals, n_ret, n)
        self.T.transposed, self.Tt.transposed = self.Tt, self.T
        self.n = n
        self._asm, self._coo = asm, (rows, cols, vals)
        self._build(self._block_size(asm, retained) if allow_blocks else 1)

    def _build(self, d: int) -> None:
        """Block-diagonal pattern of the part tangents and the operators T, T^T with d x d blocks (d = 1: scalar)."""
        asm, dev = self._asm, self._asm.device
        rows, cols, vals = self._coo
        n, n_ret = self.n, self.n_retained
        self.d = d
        ptr_parts, col_parts, self._segments = [], [], []
        nblk = 0
        for part, offset in zip(asm.parts, asm.offsets):
            if isinstance(part, FEM):
                if d == 1:
                    ptr, col = part.pattern.indptr, part.pattern.indices
                else:
                    ptr, col = part.pattern.node_ptr, part.pattern.adj
                self._segments.append(None)
            else:   # stiffness-free point: explicit zero diagonal (blocks)
                rows_here = part.n_dofs // d
                ptr = torch.arange(rows_here + 1, dtype=torch.int64, device=dev)
                col = torch.arange(rows_here, dtype=torch.int32, device=dev)
                self._segments.append(torch.zeros(d * d * rows_here, dtype=torch.float64, device=dev))
            ptr_parts.append(ptr[:-1] + nblk)
            col_parts.append(col.to(torch.int64) + offset // d)
            nblk += int(col.shape[0])
        ptr_parts.append(torch.tensor([nblk], dtype=torch.int64, device=dev))
        self.kptr = torch.cat(ptr_parts).contiguous()
        self.kcol = torch.cat(col_parts).to(torch.int32).contiguous()
        if d == 1:
            self._Tb, self._Ttb = self.T.block_operator(), self.Tt.block_operator()
        else:
            self._Tb = _blocks_from_coo(rows, cols, vals, n // d, n_ret // d, d)
            self._Ttb = _blocks_from_coo(cols, rows, vals, n_ret // d, n // d, d)
        self._kt_struct = None
        self._a_struct = None
        self._template: CSRMatrix | None = None
        self._mask_key = None
        self._mask = None
        self._last = None   # (part matrices, con key, reduced matrix) of the latest product

    @staticmethod
    def _block_size(asm: "Assembly", retained: Tensor) -> int:
        sizes = {part.n_dof_per_node for part in asm.parts if isinstance(part, FEM)}
        if len(sizes) != 1:
            return 1
        d = sizes.pop()
        if d not in (2, 3):
            return 1
        for part in asm.parts:
            if isinstance(part, FEM):
                if part.pattern.nnz != d * d * part.pattern.nnzb:     # nodes no element references: lone diagonals
                    return 1
            elif part.n_dofs % d:
                return 1
        if retained.shape[0] % d:
            return 1
        nodes = retained.view(-1, d)           # ascending: whole nodes are runs d*k, d*k+1, ...
        whole = (nodes[:, 0] % d == 0) & (nodes[:, -1] - nodes[:, 0] == d - 1)
        return d if bool(whole.all()) else 1

    def _matrix(self, A: BlockOperator) -> CSRMatrix:
        """The first reduced matrix: scalar CSR view of the product, with the SELL-32 structure (node-block columns when
        d > 1) that every later matrix on the pattern shares through `_like`."""
        n, d = self.n_retained, self.d
        if d == 1:
            return CSRMatrix(A.bptr, A.bcol, A.vals, n, symmetric=True, sell_struct=sell_structure(A.bptr, A.bcol, n))
        indptr, indices = A.indptr.contiguous(), _scalar_columns(A)
        return CSRMatrix(indptr, indices, A.vals, n, symmetric=True,
                         sell_struct=sell_structure(indptr, indices, n, (d, A.nbr, A.bptr, A.bcol)))

    def forget(self) -> None:
        """Drop the value-level caches (latest reduced matrix, Dirichlet mask); the symbolic phases stay."""
        self._last = None
        self._mask_key = None
        self._mask = None

    def reduced(self, blocks: list, con: Tensor) -> CSRMatrix:
        """`T^T K T` with the Dirichlet rows / columns of `con` (retained numbering) zeroed and unit diagonal
        (reference assembly.py:295-334), as a device CSR matrix."""
        key = (con.data_ptr(), int(con.shape[0]))
        if self._last is not None and self._last[1] == key and len(self._last[0]) == len(blocks) and all(
                a is b for a, b in zip(self._last[0], blocks)):
            return self._last[2]
        vals = torch.cat([seg if seg is not None else K._values() for seg, K in zip(self._segments, blocks)])
        d = self.d
        Kop = BlockOperator(d, self.n // d, self.n // d, self.kptr, self.kcol, vals.contiguous())
        try:
            KT, self._kt_struct = spgemm(d, Kop, self._Tb, structure=self._kt_struct)
            A, self._a_struct = spgemm(d, self._Ttb, KT, structure=self._a_struct)
        except L.TfemError as exc:
            # a reference point coupled to thousands of nodes: its product row (3,723 blocks of 3 x 3 for a face of 3,721
            # nodes) does not fit the SpGEMM's shared-memory accumulators with node blocks — scalar operators hold rows
            # of up to ~12,000 entries
            if exc.rc != L.ERR_CAPACITY or d == 1:
                raise
            self._build(1)
            return self.reduced(blocks, con)
        if self._template is None:
            self._template = self._matrix(A)
        if self._mask_key != key:
            idx = self._template._indices()
            is_con = torch.zeros(self.n_retained, dtype=torch.bool, device=vals.device)
            is_con[con] = True
            self._mask = (is_con[idx[0]] | is_con[idx[1]], self._template.diag_pos[con])
            self._mask_key = key
        kill, unit = self._mask
        if bool((unit < 0).any()):
            raise RuntimeError("the reduced tangent has no stored diagonal at a constrained DOF")
        values = A.vals.masked_fill(kill, 0.0)
        values[unit] = 1.0
        K = self._template._like(values)
        self._last = (list(blocks), key, K)
        return K


# ------------------------------------------------------------------------------------------------------- assembly
class Assembly:
    """Parts coupled by kinematic constraints (reference assembly.py:113-608). The parts keep their own nodes,
    elements, materials and boundary conditions; their DOFs are stacked in the order given.

    Attributes: `parts`, `n_dofs` (before elimination), `offsets` (first global DOF of every part), `dim`."""

    def __init__(self, parts: Sequence[Part]):
        if len({id(part) for part in parts}) != len(parts):
            raise ValueError("A part may appear only once in an assembly.")
        dims = {part.nodes.shape[1] for part in parts}
        if len(dims) != 1:
            raise ValueError("All parts of an assembly share one spatial dimension.")
        self.dim = dims.pop()
        thermal = [isinstance(part, (Heat, ReferencePointHeat)) for part in parts]
        if any(thermal) and not all(thermal):
            raise ValueError("An assembly is either mechanical or thermal, not both.")

        self.parts = list(parts)
        self.offsets, total = [], 0
        for part in self.parts:
            self.offsets.append(total)
            total += part.n_dofs
        self.n_dofs = total
        self._at = {id(part): offset for part, offset in zip(self.parts, self.offsets)}

        # models live on the GPU; points are created wherever torch's default device is and are read from there
        models = [part for part in self.parts if isinstance(part, FEM)]
        lead = models[0] if models else self.parts[0]
        self.device, self.dtype = lead.nodes.device, lead.nodes.dtype

        self._eliminated: list[Tensor] = []
        self._rows: list[tuple[Tensor, Tensor, Tensor]] = []   # (secondary DOF, primary DOF, coefficient)
        self._links: list[tuple[tuple[Part, Tensor], tuple[Part, Tensor]]] = []
        self._elimination: _Elimination | None = None
        self.node_blocks = True   # run the reduction on d x d node blocks where the couplings allow (see _Elimination)

    def __repr__(self) -> str:
        return f"<torch-fem assembly ({len(self.parts)} parts, {self.n_dofs} dofs)>"

    def _here(self, t: Tensor) -> Tensor:
        return t.to(device=self.device, dtype=self.dtype if torch.is_floating_point(t) else None)

    # ---- constraints
    def coupling(self, secondary: Part, mask: Tensor, primary: Part, primary_mask: Tensor | None = None,
                 dofs: Iterable[int] | None = None):
        """Couple the masked nodes of `secondary` to the nearest (masked) nodes of `primary` (reference
        assembly.py:180-246). `dofs` selects the DOFs per node of `secondary` that are eliminated (default: all)."""
        if id(secondary) not in self._at or id(primary) not in self._at:
            raise ValueError("Both parts must belong to this assembly.")
        n_sec, n_pri = secondary.n_dof_per_node, primary.n_dof_per_node
        dofs = list(range(n_sec)) if dofs is None else list(dofs)
        if any(dof < 0 or dof >= n_sec for dof in dofs):
            raise ValueError(f"dofs must be indices in [0, {n_sec}) for this part.")
        if any(dof >= self.dim for dof in dofs) and n_pri <= self.dim:
            raise ValueError("Rotational DOFs cannot be coupled to a primary part that has none.")

        dev = self.device
        x_sec_all, x_pri_all = self._here(secondary.nodes), self._here(primary.nodes)
        sec_nodes = torch.nonzero(mask.to(dev)).ravel()
        if primary_mask is None:
            candidates = torch.arange(primary.n_nod, device=dev)
        else:
            candidates = torch.nonzero(primary_mask.to(dev)).ravel()
        x_sec = x_sec_all[sec_nodes]
        nearest = torch.cdist(x_sec, x_pri_all[candidates]).argmin(dim=1)
        pri_nodes = candidates[nearest]
        self._links.append(((secondary, sec_nodes), (primary, pri_nodes)))

        sec_base = self._at[id(secondary)] + sec_nodes * n_sec
        pri_base = self._at[id(primary)] + pri_nodes * n_pri
        one = torch.ones(len(sec_nodes), dtype=self.dtype, device=dev)
        lever = self._skew(x_sec - x_pri_all[pri_nodes])
        axes = (0, 1, 2) if self.dim == 3 else (2,)
        for dof in dofs:
            self._eliminated.append(sec_base + dof)
            self._rows.append((sec_base + dof, pri_base + dof, one))
            if dof < self.dim and n_pri > self.dim:   # translation of a node driven by the primary's rotation(s)
                for b, axis in enumerate(axes):
                    self._rows.append((sec_base + dof, pri_base + self.dim + b, lever[:, dof, axis]))
        self._elimination = None

    def _skew(self, r: Tensor) -> Tensor:
        """[n, 3, 3] with column b = e_b x r, i.e. the matrix of `theta -> theta x r` (assembly.py:248-256); 2-D
        vectors are padded with z = 0 so that only the rotation about z (column 2) matters there."""
        n = len(r)
        r3 = torch.zeros(n, 3, dtype=r.dtype, device=r.device)
        r3[:, : self.dim] = r
        x, y, z = r3.unbind(1)
        o = torch.zeros_like(x)
        return torch.stack([torch.stack([o, z, -y], dim=1),
                            torch.stack([-z, o, x], dim=1),
                            torch.stack([y, -x, o], dim=1)], dim=1)

    def _triplets(self) -> tuple[Tensor, Tensor, Tensor, Tensor]:
        """COO entries (row, column, value) of the map T from retained to all DOFs, and the retained global DOFs
        (reference assembly.py:258-288, same checks and messages)."""
        dev = self.device
        eliminated = torch.cat(self._eliminated) if self._eliminated else _empty_index(dev)
        if len(torch.unique(eliminated)) != len(eliminated):
            raise ValueError("A DOF is eliminated by more than one constraint.")
        keep = torch.ones(self.n_dofs, dtype=torch.bool, device=dev)
        keep[eliminated] = False
        retained = torch.nonzero(keep).ravel()
        column = torch.full((self.n_dofs,), -1, dtype=torch.int64, device=dev)
        column[retained] = torch.arange(len(retained), device=dev)
        if self._rows:
            secondary, primary, coeffs = (torch.cat(x) for x in zip(*self._rows))
        else:
            secondary, primary, coeffs = _empty_index(dev), _empty_index(dev), torch.empty(0, dtype=self.dtype, device=dev)
        if bool((column[primary] < 0).any()):
            raise ValueError("A DOF is both eliminated and used as a primary. Chain the "
                             "constraints to an independent part instead.")
        rows = torch.cat([retained, secondary])
        cols = torch.cat([column[retained], column[primary]])
        values = torch.cat([torch.ones(len(retained), dtype=self.dtype, device=dev), coeffs])
        return rows, cols, values, retained

    def _build_T(self) -> tuple[Tensor, Tensor]:
        """The sparse map T [n_dofs, n_retained] with `u = T q` as a coalesced torch COO tensor, and the retained
        DOF indices — the reference's private helper of the same name (assembly.py:258-293); the solve itself
        uses the device CSR copies held by `_Elimination`."""
        rows, cols, values, retained = self._triplets()
        with torch.sparse.check_sparse_tensor_invariants(False):
            T = torch.sparse_coo_tensor(torch.stack([rows, cols]), values, (self.n_dofs, len(retained))).coalesce()
        return T, retained

    def _rigid_modes(self) -> Tensor:
        """Near-null space over all DOFs: rigid-body motions (mechanical) or a constant (thermal); reference
        assembly.py:336-358."""
        kw = dict(dtype=self.dtype, device=self.device)
        if self.parts[0].n_dof_per_node < self.dim:
            return torch.ones(self.n_dofs, 1, **kw)
        axes = (0, 1, 2) if self.dim == 3 else (2,)
        modes = torch.zeros(self.n_dofs, self.dim + len(axes), **kw)
        for part, offset in zip(self.parts, self.offsets):
            base = offset + torch.arange(part.n_nod, device=self.device) * part.n_dof_per_node
            lever = self._skew(self._here(part.nodes).detach())
            for a in range(self.dim):
                modes[base + a, a] = 1.0
                for b, axis in enumerate(axes):
                    modes[base + a, self.dim + b] = lever[:, a, axis]
            if part.n_dof_per_node > self.dim:
                for b in range(len(axes)):
                    modes[base + self.dim + b, self.dim + b] = 1.0
        return modes

    # ---- solve
    def solve(self, increments: Tensor | None = None, max_iter: int = 10, rtol: float = 1e-8, atol: float = 1e-6,
              stol: float = 1e-10, verbose: bool = False, method: str | None = None, device: str | None = None,
              return_intermediate: bool = False, aggregate_integration_points: bool = True,
              differentiable_parameters: Tensor | Iterable[Tensor] | None = None):
        """Constrained quasi-static solve by load increments (reference assembly.py:378-608). Returns
        `(u, f, flux, grad, state)`, each a list with one tensor per part; a reference point contributes empty
        flux / gradient / state. A retained DOF's force includes what the constraints transmit into it (the
        reaction where it is constrained, the coupling load at a reference point); an eliminated DOF carries the
        part's own internal force. An increment that does not converge raises (no cutback)."""
        L.require_cuda()
        dev, kw = self.device, dict(dtype=self.dtype, device=self.device)
        levels = [0.0, 1.0] if increments is None else [float(v) for v in increments]
        N = len(levels)
        if differentiable_parameters is None:
            params: tuple = ()
        elif isinstance(differentiable_parameters, Tensor):
            params = (differentiable_parameters,)
        else:
            params = tuple(differentiable_parameters)
        track = any(p.requires_grad for p in params)

        if self._elimination is None:
            self._elimination = _Elimination(self, self.node_blocks)
        elim = self._elimination
        elim.forget()   # constraints and materials may have changed since the last solve; the patterns have not
        T, Tt, retained = elim.T, elim.Tt, elim.retained

        neumann = torch.cat([self._here(p._neumann).ravel() for p in self.parts])
        dirichlet = torch.cat([self._here(p._dirichlet).ravel() for p in self.parts])
        constrained = torch.cat([p._constraints.to(dev).ravel() for p in self.parts])
        con = torch.nonzero(constrained[retained]).ravel()
        if len(con) != int(constrained.sum()):
            raise ValueError("A constrained DOF is eliminated by a constraint. Constrain the "
                             "primary part instead.")

        # per-part result arrays and the shapes that cut the flat carried state back into parts
        u = [torch.zeros(N, p.n_nod, p.n_dof_per_node, **kw) for p in self.parts]
        f = [torch.zeros(N, p.n_nod, p.n_dof_per_node, **kw) for p in self.parts]
        grad, flux, state = [], [], []
        for p in self.parts:
            if isinstance(p, FEM):
                field = (N, p.n_int, p.n_elem, *p.n_flux)
                g0 = torch.zeros(field, **kw)
                g0[:] = p.initial_grad.to(**kw)
                grad.append(g0)
                flux.append(torch.zeros(field, **kw))
                state.append(torch.zeros(N, p.n_int, p.n_elem, p.n_state, **kw))
                p.K = torch.empty(0, device=dev)   # every part caches its own tangent block
            else:
                grad.append(torch.zeros(N, 0, **kw))
                flux.append(torch.zeros(N, 0, **kw))
                state.append(torch.zeros(N, 0, **kw))
        shapes = [[x[0].shape for x in q] for q in (u, grad, flux, state)]

        def cut(flat: Tensor, part_shapes) -> list[Tensor]:
            out, start = [], 0
            for shape in part_shapes:
                size = math.prod(shape)
                out.append(flat[start:start + size].reshape(shape))
                start += size
            return out

        def with_bc(dq: Tensor, DU: Tensor) -> Tensor:
            dq = dq.clone()
            dq[con] = DU[con]
            return dq

        def integrate(prev, du, step, iteration, tangent=True):
            """All parts over the global increment du: tangent blocks, global internal force, new fields."""
            u_p, grad_p, flux_p, state_p = (cut(v, s) for v, s in zip(prev, shapes))
            blocks, F_int, fields = [], [], []
            for j, (p, offset) in enumerate(zip(self.parts, self.offsets)):
                if not isinstance(p, FEM):
                    blocks.append(None)
                    F_int.append(torch.zeros(p.n_dofs, **kw))
                    fields.append((grad_p[j], flux_p[j], state_p[j]))
                    continue
                k, f_e, *new = p.integrate_material(u_p[j], grad_p[j], flux_p[j], state_p[j],
                                                    du[offset:offset + p.n_dofs], step * p._external_gradient,
                                                    iteration, False, compute_stiffness=tangent)
                if k is not None:
                    p.K = p.assemble_matrix(k, EMPTY)
                blocks.append(p.K)
                F_int.append(p.assemble_rhs(f_e))
                fields.append(tuple(new))
            return blocks, torch.cat(F_int), list(zip(*fields))

        def make_eval_residual(F_ext, DU, step):
            # the loads of this increment are bound here, so that the adjoint replays the right ones
            def eval_residual(dq, iteration, *prev):
                du = T @ with_bc(dq, DU)
                blocks, F_int, _ = integrate(prev, du, step, iteration)
                res = Tt @ (F_int - F_ext)
                res[con] = 0.0
                return res, elim.reduced(blocks, con)

            return eval_residual

        # `T` is the identity on the retained rows: the modes restricted to them are the q that reproduces each one
        B = self._rigid_modes()[retained]

        if verbose:
            n_elem = sum(getattr(p, "n_elem", 0) for p in self.parts)
            print(f"torch-fem_b200 | solve | Assembly | {len(self.parts)} parts | {n_elem:,} elem | {self.n_dofs:,} dof"
                  f" ({elim.n_retained:,} retained) | {describe_method(elim.n_retained, 'cuda', method)}")

        dq = torch.zeros(elim.n_retained, **kw)
        carry = (torch.zeros(self.n_dofs, **kw),
                 *(torch.cat([x[0].ravel() for x in q]) for q in (grad, flux, state)))
        for n in range(1, N):
            level, step = levels[n], levels[n] - levels[n - 1]
            F_ext = level * neumann
            DU = (step * dirichlet)[retained]
            prev = tuple(x.clone() if track else x.detach() for x in carry)
            dq = newton_solve(make_eval_residual(F_ext, DU, step), dq.detach(), B, max_iter, rtol, atol, stol, None,
                              method, device, None, False, *prev, *params)

            # converged state (no tangent needed)
            dq = with_bc(dq, DU)
            du = T @ dq
            _, F_int, fields = integrate(prev, du, step, max_iter, tangent=False)
            f_cur = F_int.index_put((retained,), Tt @ F_int)
            carry = (prev[0] + du, *(torch.cat([x.ravel() for x in q]) for q in fields))
            for j, p in enumerate(self.parts):
                block = slice(self.offsets[j], self.offsets[j] + p.n_dofs)
                u[j][n] = carry[0][block].reshape(shapes[0][j])
                f[j][n] = f_cur[block].reshape(shapes[0][j])
                grad[j][n], flux[j][n], state[j][n] = (q[j] for q in fields)
            if verbose:
                print(f"  increment {n}/{N - 1} done (load factor {level:g})")

        if aggregate_integration_points:   # a reference point has no integration points to average over
            grad = [x.mean(dim=1) if x.dim() > 2 else x for x in grad]
            flux = [x.mean(dim=1) if x.dim() > 2 else x for x in flux]
            state = [x.mean(dim=1) if x.dim() > 2 else x for x in state]
        flux = [x.squeeze((-2, -1)) if x.dim() > 2 else x for x in flux]
        grad = [x.squeeze((-2, -1)) if x.dim() > 2 else x for x in grad]

        out = [u, f, flux, grad, state]
        if not track:
            out = [[x.detach() for x in q] for q in out]
        if not return_intermediate:
            out = [[x[-1] for x in q] for q in out]
        return out[0], out[1], out[2], out[3], out[4]

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is out of scope of torch-fem_b200 (SURVEY §2, row 19)")
