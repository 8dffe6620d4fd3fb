"""TEST INFRASTRUCTURE ONLY — CPU (numpy/scipy) restatement of torch-fem's implicit-solve hot path.

Nothing under `oracle/` is imported by the product package (`torch-fem_b200/`); only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs use it, as the checker
and as the timed CPU baseline. Every function cites the reference lines (relative to /root/reference) it
restates. Parity pinning: `oracle/make_golden.py` runs the *unmodified* reference in the build container and
writes `tests/golden/*.npz`; `tests/test_oracle.py` checks this file against those fixtures (bit-exact for
integer structure, <=1e-13 relative for float64), so the oracle is pinned to reference outputs.

Third-party arithmetic on the path that is not in /root/reference: scipy (>=1.14 per pyproject.toml:34-42;
1.18.1 here) `cg` / `minres`; their published algorithms are restated in `jacobi_cg` / `jacobi_minres` and
checked against scipy itself in tests.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------------
# DOF map and sparsity pattern  (src/torchfem/base.py:73-119)
# --------------------------------------------------------------------------------------------------
def dof_map(elements: np.ndarray, dpn: int) -> np.ndarray:
    """idx[e, a*dpn+i] = dpn*elements[e,a] + i, int32 (base.py:73-76, 119)."""
    e = np.asarray(elements, dtype=np.int64)
    idx = (dpn * e)[:, :, None] + np.arange(dpn, dtype=np.int64)
    return idx.reshape(e.shape[0], -1).astype(np.int32)


def pattern(idx: np.ndarray, n_dofs: int, chunk_elems: int | None = None):
    """Sorted-unique packed keys (row<<32)|col of every element slot pair plus the full diagonal;
    `k_map` = position of every slot in that list, `diag_map` = position of (i,i) (base.py:78-118).

    Returns (glob_idx int64 [2,nnz], k_map int32 [n_elem*n*n], diag_map int32 [n_dofs]).
    """
    idx64 = idx.astype(np.int64)
    n_elem, n = idx64.shape
    chunk = chunk_elems or max(1, min(n_elem, (16 * 1024 * 1024) // (n * n)))
    parts = []
    for s in range(0, n_elem, chunk):
        ic = idx64[s:s + chunk]
        parts.append(np.unique(((ic[:, :, None] << 32) | ic[:, None, :]).reshape(-1)))
    diag = np.arange(n_dofs, dtype=np.int64)
    parts.append((diag << 32) | diag)
    packed = np.unique(np.concatenate(parts))
    k_parts = []
    for s in range(0, n_elem, chunk):
        ic = idx64[s:s + chunk]
        keys = ((ic[:, :, None] << 32) | ic[:, None, :]).reshape(-1)
        k_parts.append(np.searchsorted(packed, keys).astype(np.int32))
    k_map = np.concatenate(k_parts) if k_parts else np.zeros(0, np.int32)
    diag_map = np.searchsorted(packed, (diag << 32) | diag).astype(np.int32)
    glob_idx = np.stack([packed >> 32, packed & 0xFFFFFFFF])
    return glob_idx, k_map, diag_map


def csr_from_glob_idx(glob_idx: np.ndarray, n_dofs: int):
    """COO (sorted by row, then col) -> CSR indptr/indices as the GPU path of the reference does
    with bincount+cumsum (src/torchfem/sparse.py:385-391)."""
    indptr = np.zeros(n_dofs + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.bincount(glob_idx[0], minlength=n_dofs))
    return indptr, glob_idx[1].astype(np.int32)


# --------------------------------------------------------------------------------------------------
# Shape-function gradients  (src/torchfem/base.py:293-314)
# --------------------------------------------------------------------------------------------------
def shape_gradients(nodes: np.ndarray, elements: np.ndarray, bref: np.ndarray):
    """J[q,e] = bref[q] . X_e ; detJ ; B = J^-1 bref  (base.py:306-314).

    bref: [n_int, d, nn] reference-space derivatives at the integration points.
    Returns B [n_int, n_elem, d, nn], detJ [n_int, n_elem]. Raises the reference's ValueError.
    """
    X = nodes[elements]  # [n_elem, nn, d]
    J = np.einsum("qiN,ANj->qAij", bref, X)
    detJ = np.linalg.det(J)
    if np.any(detJ <= 0.0):
        raise ValueError("Negative Jacobian. Check element numbering.")
    B = np.einsum("qEij,qjN->qEiN", np.linalg.inv(J), bref)
    return B, detJ


# --------------------------------------------------------------------------------------------------
# Element matrices  (src/torchfem/base.py:1086-1090 mechanics, :1272-1278 heat; solid.py:52-54,
# planar.py:86-88 for the detJ / thickness scaling)
# --------------------------------------------------------------------------------------------------
def integrate_k_mech(nodes, elements, bref, w, C, scale=None):
    """k_e[(p,i),(r,k)] = sum_q w_q detJ_q t_e sum_{J,L} B_q[J,p] C[i,J,k,L] B_q[L,r].

    C: [n_elem,d,d,d,d] (one tangent per element) or [n_int,n_elem,d,d,d,d] (per Gauss point).
    scale: optional [n_elem] thickness (planar.py:86-88).
    """
    B, detJ = shape_gradients(nodes, elements, bref)
    n_int, n_elem, d, nn = B.shape
    k = np.zeros((n_elem, nn * d, nn * d))
    for q in range(n_int):
        Cq = C[q] if C.ndim == 6 else C
        BCB = np.einsum("eJp,eiJkL,eLr->epirk", B[q], Cq, B[q], optimize=True)
        BCB = BCB.reshape(n_elem, nn * d, nn * d)
        f = detJ[q] * w[q]
        if scale is not None:
            f = f * scale
        k += f[:, None, None] * BCB
    return k


def integrate_k_heat(nodes, elements, bref, w, kappa, scale=None):
    """k_e[N,M] = sum_q w_q detJ_q t_e sum_{ij} kappa[i,j] B_q[i,N] B_q[j,M]  (base.py:1274)."""
    B, detJ = shape_gradients(nodes, elements, bref)
    n_int, n_elem, d, nn = B.shape
    k = np.zeros((n_elem, nn, nn))
    for q in range(n_int):
        kq = kappa[q] if kappa.ndim == 4 else kappa
        BCB = np.einsum("eij,eiN,ejM->eNM", kq, B[q], B[q], optimize=True)
        f = detJ[q] * w[q]
        if scale is not None:
            f = f * scale
        k += f[:, None, None] * BCB
    return k


# --------------------------------------------------------------------------------------------------
# Assembly  (src/torchfem/base.py:398-445)
# --------------------------------------------------------------------------------------------------
def assemble_values(k, k_map, glob_idx, diag_map, con, n_dofs):
    """val = index_add(k_map, k.ravel()); zero constrained rows/cols; unit diagonal on them
    (base.py:410-419). np.add.at accumulates sequentially in slot order like CPU index_add_."""
    nnz = glob_idx.shape[1]
    val = np.zeros(nnz)
    np.add.at(val, k_map, k.ravel())
    is_con = np.zeros(n_dofs, dtype=bool)
    is_con[con] = True
    val[is_con[glob_idx[0]] | is_con[glob_idx[1]]] = 0.0
    val[diag_map[con]] = 1.0
    return val


def assemble_values_fast(k, k_map, glob_idx, diag_map, con, n_dofs):
    """Same result to round-off via bincount (used only for the timed CPU baseline)."""
    nnz = glob_idx.shape[1]
    val = np.bincount(k_map, weights=k.ravel(), minlength=nnz)
    is_con = np.zeros(n_dofs, dtype=bool)
    is_con[con] = True
    val[is_con[glob_idx[0]] | is_con[glob_idx[1]]] = 0.0
    val[diag_map[con]] = 1.0
    return val


def assemble_rhs(f, idx, n_dofs):
    """F.index_add_(0, idx.ravel(), f.ravel())  (base.py:428-445)."""
    F = np.zeros(n_dofs)
    np.add.at(F, idx.ravel(), f.ravel())
    return F


def to_csr(val, glob_idx, n_dofs):
    indptr, indices = csr_from_glob_idx(glob_idx, n_dofs)
    return sp.csr_matrix((val, indices, indptr), shape=(n_dofs, n_dofs))


# --------------------------------------------------------------------------------------------------
# Jacobi-preconditioned Krylov solvers (reference call sites src/torchfem/sparse.py:406-421 GPU,
# :493-512 CPU; algorithm = scipy.sparse.linalg.cg / minres, scipy/sparse/linalg/_isolve)
# --------------------------------------------------------------------------------------------------
def sell32_values(indptr: np.ndarray, vals: np.ndarray):
    """The solver-internal SELL-32 layout of CSR values (no reference counterpart: the reference hands CSR to cusparse,
    sparse.py:411; this restates `torch-fem_b200/csrc/sell.cuh` so that the layout written by the assembly kernel and by
    the CSR -> SELL copy is pinned by something that is not one of them). Slice t = rows 32 t .. 32 t + 31, width W_t =
    the longest row of the slice rounded up to even; entry k of the row with lane l sits at
    slice_ptr[t] + (k // 2) * 64 + 2 l + (k % 2); everything else is 0.0. Returns (slice_ptr, values)."""
    n = len(indptr) - 1
    lens = np.diff(indptr)
    n_slices = (n + 31) // 32
    padded = np.zeros(n_slices * 32, dtype=np.int64)
    padded[:n] = lens
    width = padded.reshape(n_slices, 32).max(axis=1)
    width = (width + 1) & ~1
    slice_ptr = np.zeros(n_slices + 1, dtype=np.int64)
    np.cumsum(width * 32, out=slice_ptr[1:])
    out = np.zeros(int(slice_ptr[-1]), dtype=np.float64)
    rows = np.repeat(np.arange(n), lens)
    k = np.arange(len(vals)) - np.repeat(indptr[:-1], lens)
    pos = slice_ptr[rows // 32] + (k // 2) * 64 + 2 * (rows % 32) + (k % 2)
    out[pos] = vals
    return slice_ptr, out


def jacobi_cg(A, b, rtol=1e-10, atol=0.0, x0=None, maxiter=None, dinv=None):
    """scipy `cg` restated (scipy/sparse/linalg/_isolve/iterative.py `cg`): stop when
    ||r|| < max(atol, rtol*||b||), tested at the top of every iteration; maxiter = 10 n;
    M = diag(A)^-1 (sparse.py:408-409). Returns (x, info, iterations)."""
    n = b.shape[0]
    if dinv is None:
        dinv = 1.0 / A.diagonal()
    bnrm2 = np.linalg.norm(b)
    tol = max(float(atol), float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b.copy(), 0, 0
    if maxiter is None:
        maxiter = n * 10
    x = np.zeros(n) if x0 is None else x0.astype(np.float64).copy()
    r = b - A @ x if x.any() else b.copy()
    rho_prev, p = None, None
    for it in range(maxiter):
        if np.linalg.norm(r) < tol:
            return x, 0, it
        z = dinv * r
        rho_cur = np.dot(r, z)
        if it > 0:
            beta = rho_cur / rho_prev
            p *= beta
            p += z
        else:
            p = z.copy()
        q = A @ p
        alpha = rho_cur / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho_cur
    return x, maxiter, maxiter


def jacobi_minres(A, b, rtol=1e-10, x0=None, maxiter=None, dinv=None, shift=0.0):
    """scipy `minres` (Paige-Saunders; scipy/sparse/linalg/_isolve/minres.py) restated with
    M = diag(A)^-1. Returns (x, info, iterations); info 0 on success like the reference expects
    (sparse.py:411-413)."""
    n = b.shape[0]
    if dinv is None:
        dinv = 1.0 / A.diagonal()
    if maxiter is None:
        maxiter = 5 * n
    eps = np.finfo(np.float64).eps
    x = np.zeros(n) if x0 is None else x0.astype(np.float64).copy()
    r1 = b.copy() if x0 is None else b - A @ x
    y = dinv * r1
    beta1 = np.dot(r1, y)
    if beta1 < 0:
        raise ValueError("indefinite preconditioner")
    if beta1 == 0:
        return x, 0, 0
    bnorm = np.linalg.norm(b)
    if bnorm == 0:
        return b.copy(), 0, 0
    beta1 = np.sqrt(beta1)
    oldb = 0.0
    beta = beta1
    dbar = 0.0
    epsln = 0.0
    qrnorm = beta1
    phibar = beta1
    rhs1 = beta1
    rhs2 = 0.0
    tnorm2 = 0.0
    gmax = 0.0
    gmin = np.finfo(np.float64).max
    cs = -1.0
    sn = 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    r2 = r1
    istop = 0
    itn = 0
    Anorm = 0.0
    Acond = 0.0
    rnorm = 0.0
    ynorm = 0.0
    while itn < maxiter:
        itn += 1
        s = 1.0 / beta
        v = s * y
        y = A @ v
        y = y - shift * v
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = np.dot(v, y)
        y = y - (alfa / beta) * r2
        r1 = r2
        r2 = y
        y = dinv * r2
        oldb = beta
        beta = np.dot(r2, y)
        if beta < 0:
            raise ValueError("non-symmetric matrix")
        beta = np.sqrt(beta)
        tnorm2 += alfa ** 2 + oldb ** 2 + beta ** 2
        if itn == 1:
            if beta / beta1 <= 10 * eps:
                istop = -1
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        root = np.linalg.norm([gbar, dbar])
        Arnorm = phibar * root  # noqa: F841
        gamma = np.linalg.norm([gbar, beta])
        gamma = max(gamma, eps)
        cs = gbar / gamma
        sn = beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        denom = 1.0 / gamma
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) * denom
        x = x + phi * w
        gmax = max(gmax, gamma)
        gmin = min(gmin, gamma)
        z = rhs1 / gamma
        rhs1 = rhs2 - delta * z
        rhs2 = -epsln * z
        Anorm = np.sqrt(tnorm2)
        ynorm = np.linalg.norm(x)
        epsa = Anorm * eps
        epsx = Anorm * ynorm * eps
        epsr = Anorm * ynorm * rtol
        diag = gbar
        if diag == 0:
            diag = epsa
        qrnorm = phibar
        rnorm = qrnorm
        if ynorm == 0 or Anorm == 0:
            test1 = np.inf
        else:
            test1 = rnorm / (Anorm * ynorm)
        if Anorm == 0:
            test2 = np.inf
        else:
            test2 = root / Anorm
        Acond = gmax / gmin
        if istop == 0:
            t1 = 1 + test1
            t2 = 1 + test2
            if t2 <= 1:
                istop = 2
            if t1 <= 1:
                istop = 1
            if itn >= maxiter:
                istop = 6
            if Acond >= 0.1 / eps:
                istop = 4
            if epsx >= beta1:
                istop = 3
            if test2 <= rtol:
                istop = 2
            if test1 <= rtol:
                istop = 1
        if istop != 0:
            break
    info = maxiter if istop == 6 else 0
    return x, info, itn


# --------------------------------------------------------------------------------------------------
# Synthetic inputs of the benchmark (reference generators restated: src/torchfem/mesh.py:8-46,
# benchmarks/cubes.py:9-24, materials/elasticity.py:58-70)
# --------------------------------------------------------------------------------------------------
def cube_hexa(Nx, Ny, Nz, Lx=1.0, Ly=1.0, Lz=1.0):
    """mesh.py:8-46 — node id = i*Ny*Nz + j*Nz + k, connectivity [n0,n1,n3,n2,n4,n5,n7,n6]."""
    X = np.linspace(0, Lx, Nx)
    Y = np.linspace(0, Ly, Ny)
    Z = np.linspace(0, Lz, Nz)
    x, y, z = np.meshgrid(X, Y, Z, indexing="ij")
    nodes = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    ind = np.arange(Nx * Ny * Nz, dtype=np.int64).reshape(Nx, Ny, Nz)
    n0 = ind[:-1, :-1, :-1].ravel()
    n1 = ind[1:, :-1, :-1].ravel()
    n2 = ind[:-1, 1:, :-1].ravel()
    n3 = ind[1:, 1:, :-1].ravel()
    n4 = ind[:-1, :-1, 1:].ravel()
    n5 = ind[1:, :-1, 1:].ravel()
    n6 = ind[:-1, 1:, 1:].ravel()
    n7 = ind[1:, 1:, 1:].ravel()
    return nodes, np.stack([n0, n1, n3, n2, n4, n5, n7, n6], axis=1)


def isotropic_C3d(E, nu, n_elem):
    """C_ijkl = lbd d_ij d_kl + G (d_ik d_jl + d_il d_jk)  (materials/elasticity.py:58-70)."""
    lbd = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    G = E / (2.0 * (1.0 + nu))
    I2 = np.eye(3)
    C = lbd * np.einsum("ij,kl->ijkl", I2, I2) + G * (
        np.einsum("ik,jl->ijkl", I2, I2) + np.einsum("il,jk->ijkl", I2, I2))
    return np.broadcast_to(C, (n_elem, 3, 3, 3, 3)).copy()


def cube_extension_bcs(nodes, Lx=1.0):
    """benchmarks/cubes.py:19-22: x==0 fully fixed, u_x = 0.1 prescribed at x==Lx."""
    n_nod = nodes.shape[0]
    con = np.zeros((n_nod, 3), dtype=bool)
    disp = np.zeros((n_nod, 3))
    con[nodes[:, 0] == 0.0, :] = True
    con[nodes[:, 0] == Lx, 0] = True
    disp[nodes[:, 0] == Lx, 0] = 0.1
    return con, disp


def hexa1_tables():
    """Hexa1 reference gradients at the 2x2x2 Gauss points (elements.py:1003-1075), restated in
    closed form: dN_a/dxi_c = 1/8 * s_ac * prod_{m != c} (1 + s_am xi_m)."""
    s = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                  [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
    g = 1.0 / np.sqrt(3.0)
    ip = np.array([[x1 * g, x2 * g, x3 * g] for x3 in (-1.0, 1.0) for x2 in (-1.0, 1.0)
                   for x1 in (-1.0, 1.0)])
    bref = np.zeros((8, 3, 8))
    for q in range(8):
        t = 1.0 + s * ip[q]  # [8,3]
        for c in range(3):
            others = [m for m in range(3) if m != c]
            bref[q, c] = 0.125 * s[:, c] * t[:, others[0]] * t[:, others[1]]
    return bref, np.ones(8)


def linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-8,
                                method="cg"):
    """One linear `FEM.solve()` forward restated end to end for the timed CPU baseline:
    pattern (base.py:78-118) -> k (base.py:1086-1090) -> assemble (base.py:398-426) ->
    residual for du=0 with Dirichlet increment (base.py:708-741) -> Jacobi-CG (sparse.py:414-421)
    -> du = -du_i (sparse.py:612). Returns dict with u and timings."""
    import time

    n_nod, d = nodes.shape
    n_dofs = n_nod * d
    t0 = time.perf_counter()
    idx = dof_map(elements, d)
    glob_idx, k_map, diag_map = pattern(idx, n_dofs)
    t1 = time.perf_counter()
    k = integrate_k_mech(nodes, elements, bref, w, C)
    t2 = time.perf_counter()
    con = np.nonzero(con_mask.ravel())[0]
    val = assemble_values_fast(k, k_map, glob_idx, diag_map, con, n_dofs)
    A = to_csr(val, glob_idx, n_dofs)
    t3 = time.perf_counter()
    # residual at du=0: f_int = k_e (du_bc)_e, F_ext = 0, res[con] = 0
    du_bc = np.zeros(n_dofs)
    du_bc[con] = disp.ravel()[con]
    f_e = np.einsum("eij,ej->ei", k, du_bc[idx])
    res = assemble_rhs(f_e, idx, n_dofs)
    res[con] = 0.0
    t4 = time.perf_counter()
    if method == "cg":
        x, info, its = jacobi_cg(A, res, rtol=rtol)
    else:
        x, info, its = jacobi_minres(A, res, rtol=rtol)
    t5 = time.perf_counter()
    u = -x
    u[con] = disp.ravel()[con]
    return {
        "u": u.reshape(n_nod, d), "iterations": its, "info": info, "nnz": glob_idx.shape[1],
        "t_setup": t1 - t0, "t_integrate": t2 - t1, "t_assemble": t3 - t2, "t_rhs": t4 - t3,
        "t_solve": t5 - t4, "A": A, "res": res,
    }
