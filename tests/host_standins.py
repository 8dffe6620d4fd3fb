"""CPU stand-ins for the kernel-backed pieces of the models, shared by the host-logic tests (`-m "not gpu"`): a model
whose element matrices and assembly come from `oracle/fem_oracle.py` and whose residual runs the package's own torch
formulation, a host matrix that answers what the callers read of `csr.CSRMatrix`, and a dense `sparse_solve`. Test
infrastructure only: the product has no CPU path."""
import numpy as np
import torch

from oracle import fem_oracle as O


class HostMatrix:
    """What the assembly reads of `csr.CSRMatrix`."""

    def __init__(self, indptr, indices, values, n, symmetric=False, **_):
        self.indptr, self.indices, self.values_, self.n = indptr, indices, values, int(n)
        self.symmetric = symmetric

    shape = property(lambda self: torch.Size((self.n, self.n)))

    def numel(self):
        return self.n * self.n

    def _values(self):
        return self.values_

    def _indices(self):
        rows = torch.repeat_interleave(torch.arange(self.n), self.indptr[1:] - self.indptr[:-1])
        return torch.stack([rows, self.indices.to(torch.int64)])

    @property
    def diag_pos(self):
        r, c = self._indices()
        pos = torch.full((self.n,), -1, dtype=torch.int64)
        on = torch.nonzero(r == c).ravel()
        pos[r[on]] = on
        return pos

    def _like(self, values):
        return HostMatrix(self.indptr, self.indices, values, self.n, self.symmetric)

    @property
    def T(self):
        assert self.symmetric
        return self

    # the linear algebra `Heat.time_integration` does on matrices of one pattern (M + dt/2 K, M @ du)
    def __mul__(self, c):
        return self._like(self.values_ * float(c))

    __rmul__ = __mul__

    def __add__(self, other):
        assert other.indices is self.indices
        return self._like(self.values_ + other.values_)

    def __matmul__(self, x):
        if x.requires_grad:
            return _HostMatvec.apply(x, self)
        return self.matvec(x.reshape(-1)).reshape(x.shape)

    def matvec(self, x, fmt="auto"):
        r, c = self._indices()
        return torch.zeros(self.n, dtype=torch.float64).index_add_(0, r, self.values_ * x[c])

    def matmat(self, X):
        return torch.stack([self.matvec(X[:, j]) for j in range(X.shape[1])], dim=1)

    def dense(self):
        r, c = self._indices()
        out = np.zeros((self.n, self.n))
        out[r.numpy(), c.numpy()] = self.values_.numpy()
        return out


class _HostMatvec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, A):
        ctx.A = A
        return A.matvec(x.detach().reshape(-1)).reshape(x.shape)

    @staticmethod
    def backward(ctx, g):
        return ctx.A.T.matvec(g.reshape(-1)).reshape(g.shape), None


def dense_sparse_solve(A, b, B=None, stol=1e-10, device=None, method=None, M=None, x0=None):
    return torch.from_numpy(np.linalg.solve(A.dense(), b.detach().numpy())), None



def host_model(cls, nodes, elements, material, thickness=1.0):
    """An instance of the model class `cls` (Solid, Planar, SolidHeat, PlanarHeat) living on the CPU."""
    import torchfem_b200 as T
    from torchfem_b200.sparse import CachedSolve

    class Host(cls):
        def __init__(self):
            self.device = nodes.device
            self.nodes, self.elements = nodes, elements
            self.n_nod, self.n_dim = nodes.shape
            dpn = self.n_dof_per_node
            self.n_dofs, self.n_elem = dpn * self.n_nod, len(elements)
            self.n_int = len(self.etype.iweights)
            self._neumann = torch.zeros(self.n_nod, dpn)
            self._dirichlet = torch.zeros(self.n_nod, dpn)
            self._constraints = torch.zeros(self.n_nod, dpn, dtype=torch.bool)
            self._external_gradient = torch.zeros(self.n_elem, *self.n_flux)
            self.idx = torch.from_numpy(O.dof_map(elements.numpy(), dpn))
            self._glob_idx, self._k_map, self._diag_map = O.pattern(self.idx.numpy(), self.n_dofs)
            indptr, indices = O.csr_from_glob_idx(self._glob_idx, self.n_dofs)
            blocks = np.unique((self._glob_idx[0] // dpn) * self.n_nod + self._glob_idx[1] // dpn)   # node graph
            node_ptr = np.concatenate([[0], np.cumsum(np.bincount(blocks // self.n_nod, minlength=self.n_nod))])
            self.pattern = type("P", (), {
                "indptr": torch.from_numpy(indptr), "indices": torch.from_numpy(indices), "nnz": len(indices),
                "nnzb": len(blocks), "node_ptr": torch.from_numpy(node_ptr.astype(np.int64)),
                "adj": torch.from_numpy((blocks % self.n_nod).astype(np.int32))})
            self.material = material if material.is_vectorized else material.vectorize(self.n_elem)
            self.cached_solve = CachedSolve()
            self.K = torch.empty(0)
            self._shape_cache = None
            if isinstance(thickness, torch.Tensor):   # read by the planar models only
                self.thickness = thickness
            else:
                self.thickness = torch.full((self.n_elem,), float(thickness))

        def _geometry(self):
            return None   # torch formulation of the residual

        def _integrate_k_raw(self, tangent):
            bref, w = self._tables()
            fn = O.integrate_k_mech if self.KIND == T._lib.KIND_MECH else O.integrate_k_heat
            return torch.from_numpy(fn(nodes.detach().numpy(), elements.numpy(), bref.numpy(), w.numpy(),
                                       tangent.detach().numpy(), self.thickness.detach().numpy()))

        def assemble_rhs(self, f):
            F = torch.zeros(self.n_dofs, dtype=f.dtype)
            return F.index_add_(0, self.idx.ravel().long(), f.ravel())     # sequential on the CPU: deterministic

        def assemble_matrix(self, k, con):
            val = O.assemble_values(k.detach().numpy(), self._k_map, self._glob_idx, self._diag_map,
                                    con.numpy(), self.n_dofs)
            return HostMatrix(self.pattern.indptr, self.pattern.indices, torch.from_numpy(val), self.n_dofs, True)

    return Host()
