"""CPU restatement (numpy / scipy) of the aggregation-AMG preconditioner of `method="amg"` — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this; the product (torch-fem_b200/amg.py +
csrc/amg.cu) never does.

What it restates. The reference preconditions its Krylov solves with an algebraic multigrid hierarchy built by
third-party libraries that are not in /root/reference: pyamg `smoothed_aggregation_solver(A, B, smooth="jacobi")` on
the CPU (src/torchfem/sparse.py:493-512) and AmgX aggregation AMG (V cycle, one pre/post sweep, dense LU on the
coarsest level) on the GPU (src/torchfem/amgx.py:71-98, sparse.py:422-442). Neither library is installed here and
no reference test pins their numbers ("parity unpinned" for the hierarchy itself); what IS pinned is the solution of
the linear system, which any SPD preconditioner must reproduce to the Krylov tolerance (tests/test_sparse.py:49-91).
This file therefore fixes OUR algorithm — the published smoothed-aggregation method (Vanek, Mandel, Brezina 1996) with
the choices below — step by step, so that the CUDA kernels can be checked against it level by level:

  * nodes = groups of `d` DOFs (d = DOFs per node), all blocks d x d; near-null space = the d translations
    (piecewise constant per DOF component), masked at "isolated" DOFs (rows whose off-diagonal entries are all zero:
    Dirichlet rows after the reference's masking, base.py:414-419);
  * aggregation = maximal independent set of the node graph with fixed pseudo-random priorities (parallel,
    deterministic): every root collects the neighbours that prefer it (largest key). Aggregates have radius 1, which
    is what one step of prolongator smoothing can cover (distance-2 sets on Hexa1: 37 instead of 23 iterations);
    where radius 1 gives fewer than 6 nodes per aggregate (graphs of low degree: Tetra1) the distance-2 variant is
    used instead (Tetra1 20^3: operator complexity 3.2 -> 1.15, 22 -> 33 iterations);
  * prolongator smoothing P = (I - w D^-1 A) T, w = 4 / (3 rho), rho = spectral-radius estimate of D^-1 A
    (power iteration with a fixed start vector, times a safety factor);
  * Galerkin coarse operator A_c = P^T A P; zero diagonal entries (aggregates made of isolated DOFs only) become 1;
  * V(1,1) cycle with damped Jacobi (same w) — symmetric, so it is a valid CG preconditioner; dense inverse on the
    coarsest level.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

POWER_ITS = 8
RHO_SAFETY = 1.15
MIN_AGG_SIZE = 6          # nodes per aggregate below which radius-2 aggregates replace radius-1 ones
MAX_COARSE_DOFS = 1500
MAX_LEVELS = 12


def hash32(i: np.ndarray) -> np.ndarray:
    """32-bit integer mix (the finaliser of MurmurHash3) of the node index — the fixed MIS priorities."""
    h = i.astype(np.uint64) & 0xFFFFFFFF
    h = (h + 0x9E3779B9) & 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h.astype(np.uint64)


def _seg_max(vals: np.ndarray, ptr: np.ndarray) -> np.ndarray:
    """max over every CSR segment (all segments non-empty)."""
    return np.maximum.reduceat(vals, ptr[:-1])


def mis_aggregate(ptr: np.ndarray, adj: np.ndarray):
    """Aggregation by a maximal independent set of the node graph (ptr, adj) — adj rows sorted and containing the
    node itself. Parallel rounds (Luby): an undecided node whose key is the largest among its undecided neighbours
    becomes a root, its neighbours become members. Every member then joins the adjacent root with the largest key.
    Aggregates are numbered in root order. Returns (agg [n] int32, n_agg, rounds)."""
    n = len(ptr) - 1
    ids = np.arange(n, dtype=np.uint64)
    key = (hash32(ids) << np.uint64(32)) | (ids + np.uint64(1))   # unique, > 0
    state = np.zeros(n, dtype=np.int8)        # 0 undecided, 1 root, 2 member
    rounds = 0
    while (state == 0).any():
        k = np.where(state == 0, key, np.uint64(0))
        new_root = (state == 0) & (_seg_max(k[adj], ptr) == key)
        state[new_root] = 1
        covered = _seg_max(new_root[adj].astype(np.int8), ptr) > 0
        state[(state == 0) & covered] = 2
        rounds += 1
    is_root = state == 1
    agg_of_root = np.cumsum(is_root) - 1
    best = _seg_max(np.where(is_root, key, np.uint64(0))[adj], ptr)      # key of the chosen root
    root = (best & np.uint64(0xFFFFFFFF)).astype(np.int64) - 1
    return agg_of_root[root].astype(np.int32), int(is_root.sum()), rounds


def mis2_aggregate(ptr: np.ndarray, adj: np.ndarray):
    """The distance-2 variant for graphs of low degree (Tetra1: radius-1 aggregates hold ~3 nodes there and the coarse
    operators fill in): roots are more than two steps apart (largest key among the undecided nodes within distance
    2), nodes one step from a root join it, the others join the aggregate of their neighbour with the largest key."""
    n = len(ptr) - 1
    ids = np.arange(n, dtype=np.uint64)
    key = (hash32(ids) << np.uint64(32)) | (ids + np.uint64(1))
    state = np.zeros(n, dtype=np.int8)
    rounds = 0
    while (state == 0).any():
        k = np.where(state == 0, key, np.uint64(0))
        t1 = np.maximum(_seg_max(k[adj], ptr), k)
        new_root = (state == 0) & (np.maximum(_seg_max(t1[adj], ptr), t1) == key)
        state[new_root] = 1
        near1 = new_root | (_seg_max(new_root[adj].astype(np.int8), ptr) > 0)
        near2 = near1 | (_seg_max(near1[adj].astype(np.int8), ptr) > 0)
        state[(state == 0) & near2] = 2
        rounds += 1
    is_root = state == 1
    agg_of_root = np.cumsum(is_root) - 1
    best = np.maximum(_seg_max(np.where(is_root, key, np.uint64(0))[adj], ptr), np.where(is_root, key, np.uint64(0)))
    root = (best & np.uint64(0xFFFFFFFF)).astype(np.int64) - 1
    agg1 = np.where(root >= 0, agg_of_root[np.maximum(root, 0)], -1)
    best2 = _seg_max(np.where(agg1 >= 0, key, np.uint64(0))[adj], ptr)
    nb = (best2 & np.uint64(0xFFFFFFFF)).astype(np.int64) - 1
    agg = np.where(agg1 >= 0, agg1, agg1[np.maximum(nb, 0)])
    assert (agg >= 0).all()
    return agg.astype(np.int32), int(is_root.sum()), rounds


def block_graph(A: sp.csr_matrix, d: int):
    """Node graph of a matrix with d x d block structure: (ptr, adj, Gb) with Gb the nb x nb pattern as a matrix of
    ones. Explicitly stored zeros count (the kernels work on the STRUCTURAL pattern)."""
    A = A.tocsr()
    n = A.shape[0] // d
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr)) // d
    Gb = sp.csr_matrix((np.ones(len(A.indices)), (rows, A.indices // d)), shape=(n, n))
    Gb.sum_duplicates()
    Gb.sort_indices()
    Gb.data[:] = 1.0
    return Gb.indptr.astype(np.int64), Gb.indices.astype(np.int32), Gb


def with_block_pattern(A: sp.csr_matrix, Gb: sp.csr_matrix, d: int) -> sp.csr_matrix:
    """A with explicit zeros added so that every block of the pattern Gb is stored in full."""
    S = sp.kron(Gb, np.ones((d, d))).tocoo()
    A = A.tocoo()
    out = sp.coo_matrix((np.concatenate([A.data, np.zeros(len(S.data))]),
                         (np.concatenate([A.row, S.row]), np.concatenate([A.col, S.col]))), shape=A.shape).tocsr()
    out.sort_indices()
    return out


def _ones(M):
    M = M.tocsr().copy()
    M.data[:] = 1.0
    return M


def row_info(A: sp.csr_matrix):
    """diag, isolated flag (all off-diagonal entries zero); zero diagonals are set to 1 in a copy of A."""
    A = A.tocsr().copy()
    diag = A.diagonal()
    zero = diag == 0.0
    if zero.any():
        diag[zero] = 1.0
        A.setdiag(diag)            # the diagonal is stored (explicit zeros), so the structure does not change
    off = A - sp.diags(diag)
    iso = np.asarray(abs(off).sum(axis=1)).ravel() == 0.0
    return A, diag, iso


def rho_estimate(A: sp.csr_matrix, dinv: np.ndarray) -> float:
    """Power iteration on D^-1 A from a fixed start vector, times RHO_SAFETY."""
    n = A.shape[0]
    x = 1.0 + (hash32(np.arange(n)) % np.uint64(1024)).astype(np.float64) / 1024.0
    lam = 1.0
    for _ in range(POWER_ITS):
        y = dinv * (A @ x)
        lam = np.sqrt(y @ y) / np.sqrt(x @ x)
        x = y / np.sqrt(y @ y)
    return float(lam * RHO_SAFETY)


class Level:
    pass


def build_hierarchy(A: sp.csr_matrix, d: int, max_coarse=MAX_COARSE_DOFS, max_levels=MAX_LEVELS,
                    aggregation: str = "auto"):
    levels = []
    while True:
        L = Level()
        A, diag, iso = row_info(A)
        L.A, L.dinv, L.iso, L.d = A, 1.0 / diag, iso, d
        L.n = A.shape[0]
        levels.append(L)
        if L.n <= max_coarse or len(levels) >= max_levels:
            break
        L.rho = rho_estimate(A, L.dinv)
        L.omega = 4.0 / (3.0 * L.rho)
        ptr, adj, Gb = block_graph(A, d)
        n_nod = L.n // d
        L.agg_distance = 2 if aggregation == "mis2" else 1
        agg, n_agg, _ = (mis2_aggregate if aggregation == "mis2" else mis_aggregate)(ptr, adj)
        if aggregation == "auto" and n_nod < MIN_AGG_SIZE * n_agg:
            agg, n_agg, _ = mis2_aggregate(ptr, adj)
            L.agg_distance = 2
        if n_agg * d >= 0.8 * L.n:          # coarsening stalled
            break
        L.agg, L.n_agg = agg, n_agg
        n_nod = L.n // d
        rows = np.arange(L.n)
        cols = np.repeat(agg.astype(np.int64), d) * d + np.tile(np.arange(d), n_nod)
        T = sp.csr_matrix(((~iso).astype(np.float64), (rows, cols)), shape=(L.n, n_agg * d))
        P = (T - L.omega * (sp.diags(L.dinv) @ (A @ T))).tocsr()
        L.P, L.R = P, P.T.tocsr()
        # structural pattern of the Galerkin product (what the SpGEMM kernels produce), values from scipy
        Tb = sp.csr_matrix((np.ones(n_nod), (np.arange(n_nod), agg)), shape=(n_nod, n_agg))
        Pb = _ones(Gb @ Tb)
        Acb = _ones(Pb.T @ _ones(Gb @ Pb))
        A = with_block_pattern((L.R @ A @ P).tocsr(), Acb, d)
    Lc = levels[-1]
    if not hasattr(Lc, "omega"):
        Lc.rho = Lc.omega = None
    Lc.inv = np.linalg.inv(Lc.A.toarray())
    return levels


def vcycle(levels, b: np.ndarray, lvl: int = 0) -> np.ndarray:
    L = levels[lvl]
    if lvl == len(levels) - 1:
        return L.inv @ b
    x = L.omega * L.dinv * b
    r = b - L.A @ x
    x = x + L.P @ vcycle(levels, L.R @ r, lvl + 1)
    return x + L.omega * L.dinv * (b - L.A @ x)


def amg_pcg(A, b, levels, rtol=1e-10, atol=0.0, x0=None, maxiter=None):
    """Preconditioned CG with the V cycle, scipy `cg`'s stopping rule (||r|| < max(atol, rtol ||b||), tested before
    every iteration). Returns (x, info, iterations)."""
    n = len(b)
    maxiter = 10 * n if maxiter is None else maxiter
    x = np.zeros(n) if x0 is None else x0.astype(np.float64).copy()
    r = b - A @ x if x0 is not None else b.copy()
    tol = max(atol, rtol * np.linalg.norm(b))
    if np.linalg.norm(r) < tol:
        return x, 0, 0
    z = vcycle(levels, r)
    p = z.copy()
    rho = r @ z
    for it in range(1, maxiter + 1):
        q = A @ p
        alpha = rho / (p @ q)
        x += alpha * p
        r -= alpha * q
        if np.linalg.norm(r) < tol:
            return x, 0, it
        z = vcycle(levels, r)
        rho_new = r @ z
        p = z + (rho_new / rho) * p
        rho = rho_new
    return x, maxiter, maxiter
