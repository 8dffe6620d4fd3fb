"""The CPU oracle (oracle/fem_oracle.py) against fixtures generated from the unmodified reference
(oracle/make_golden.py). Bit-exact for integer structure, <=1e-13 relative (norm-wise) for fp64."""
import hashlib

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import HEAT_CASES, MECH_CASES, load_case
from oracle import fem_oracle as O


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_pattern_bit_exact(tag):
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    idx = O.dof_map(c["elements"], dpn)
    assert idx.dtype == np.int32 and np.array_equal(idx, c["idx"])
    glob_idx, k_map, diag_map = O.pattern(idx, dpn * c["nodes"].shape[0])
    assert np.array_equal(glob_idx, c["glob_idx"])
    assert np.array_equal(k_map, c["k_map"]) and k_map.dtype == np.int32
    assert np.array_equal(diag_map, c["diag_map"])


@pytest.mark.parametrize("tag", MECH_CASES)
def test_k_mech(tag, tables):
    c = load_case(f"case_{tag}.npz")
    et = str(c["etype"])
    k = O.integrate_k_mech(c["nodes"], c["elements"], tables[f"{et}.B_ip"],
                           tables[f"{et}.iweights"].astype(np.float64), c["C"],
                           c.get("thickness"))
    assert rel(k, c["k"]) <= 1e-13


@pytest.mark.parametrize("tag", HEAT_CASES)
def test_k_heat(tag, tables):
    c = load_case(f"case_{tag}.npz")
    et = str(c["etype"])
    k = O.integrate_k_heat(c["nodes"], c["elements"], tables[f"{et}.B_ip"],
                           tables[f"{et}.iweights"].astype(np.float64), c["kappa"],
                           c.get("thickness"))
    assert rel(k, c["k"]) <= 1e-13


def test_k_per_gauss_point_tangent(tables):
    c = load_case("case_hyper_hexa1.npz")
    k = O.integrate_k_mech(c["nodes"], c["elements"], tables["Hexa1.B_ip"],
                           tables["Hexa1.iweights"], c["C"])
    assert c["C"].ndim == 6
    assert rel(k, c["k"]) <= 1e-13


@pytest.mark.parametrize("tag", MECH_CASES + HEAT_CASES)
def test_assemble(tag):
    c = load_case(f"case_{tag}.npz")
    dpn = 1 if tag.startswith("heat") else c["nodes"].shape[1]
    n_dofs = dpn * c["nodes"].shape[0]
    for fn in (O.assemble_values, O.assemble_values_fast):
        val = fn(c["k"], c["k_map"], c["glob_idx"], c["diag_map"], c["con"], n_dofs)
        assert rel(val, c["K_val"]) <= 1e-13
        assert np.array_equal(val == 1.0, c["K_val"] == 1.0)


def test_negative_jacobian(tables):
    c = load_case("case_hexa1.npz")
    el = c["elements"].copy()
    el[0] = el[0][[1, 0, 3, 2, 5, 4, 7, 6]]  # mirrored numbering -> det < 0
    with pytest.raises(ValueError, match="Negative Jacobian"):
        O.shape_gradients(c["nodes"], el, tables["Hexa1.B_ip"])


def test_hexa1_tables_closed_form(tables):
    bref, w = O.hexa1_tables()
    assert np.array_equal(w, tables["Hexa1.iweights"])
    assert np.abs(bref - tables["Hexa1.B_ip"]).max() <= 1e-16


def test_config_a_end_to_end(tables):
    """BASELINE config[0] through the oracle vs the reference's golden vectors."""
    g = load_case("config_a.npz")
    nodes, elements = O.cube_hexa(11, 11, 11)
    idx = O.dof_map(elements, 3)
    assert sha(idx) == str(g["sha_idx"])
    glob_idx, k_map, diag_map = O.pattern(idx, 3993)
    assert sha(glob_idx) == str(g["sha_glob_idx"]) == "29c4e23ca90b6ab5"
    assert sha(k_map) == str(g["sha_k_map"]) == "f65f0b465b5b2bbf"
    assert sha(diag_map) == str(g["sha_diag_map"]) == "620cd37e28ff90bc"
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    k = O.integrate_k_mech(nodes, elements, bref, w, C)
    assert rel(k[0], g["k_e0"]) <= 1e-13 and rel(k[777], g["k_e777"]) <= 1e-13
    assert abs(np.linalg.norm(k) - g["k_fro"]) <= 1e-12 * g["k_fro"]
    con_mask, disp = O.cube_extension_bcs(nodes)
    con = np.nonzero(con_mask.ravel())[0]
    assert np.array_equal(con, g["con"])
    val = O.assemble_values(k, k_map, glob_idx, diag_map, con, 3993)
    assert abs(np.linalg.norm(val) - g["val_norm"]) <= 1e-12 * g["val_norm"]
    # exact zeros beyond the masked entries are round-off accidents: pin only the unit diagonals
    assert int((val == 1).sum()) == int(g["val_n_one"]) == 484
    assert np.abs(val[:4096] - g["val_head"]).max() <= 1e-12 * g["val_absmax"]
    out = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10)
    u_ref = g["u"]
    assert np.linalg.norm(out["u"] - u_ref) / np.linalg.norm(u_ref) <= 1e-8
    # equal-tolerance comparison against the reference's own Jacobi-CG run
    assert np.linalg.norm(out["u"] - g["u_cg"]) / np.linalg.norm(u_ref) <= 1e-8
    out2 = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-10,
                                         method="minres")
    assert np.linalg.norm(out2["u"] - g["u_minres"]) / np.linalg.norm(u_ref) <= 1e-8


def test_krylov_restatements_match_scipy():
    """jacobi_cg / jacobi_minres follow scipy's iterates step for step (same iteration counts)."""
    nodes, elements = O.cube_hexa(6, 6, 6)
    bref, w = O.hexa1_tables()
    C = O.isotropic_C3d(1000.0, 0.3, len(elements))
    con_mask, disp = O.cube_extension_bcs(nodes)
    out = O.linear_solve_reference_flow(nodes, elements, bref, w, C, con_mask, disp, rtol=1e-8)
    A, b = out["A"], out["res"]
    dinv = 1.0 / A.diagonal()
    Mop = spla.LinearOperator(A.shape, matvec=lambda x: dinv * x)
    for rtol in (1e-8, 1e-10):
        cnt = [0]
        xs, info = spla.cg(A, b, M=Mop, rtol=rtol, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
        xo, info_o, its = O.jacobi_cg(A, b, rtol=rtol)
        assert info == 0 and info_o == 0 and its == cnt[0]
        assert np.abs(xs - xo).max() <= 1e-12 * np.abs(xs).max()
        cnt = [0]
        xs, info = spla.minres(A, b, M=Mop, rtol=rtol, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
        xo, info_o, its = O.jacobi_minres(A, b, rtol=rtol)
        assert info == 0 and info_o == 0 and its == cnt[0]
        assert np.abs(xs - xo).max() <= 1e-11 * np.abs(xs).max()


def test_sell32_layout_restatement_small_case():
    """Hand-checked case of the SELL-32 restatement: 34 rows (two slices), ragged lengths, an empty row."""
    lens = np.array([3, 0, 5] + [1] * 29 + [2, 4])
    indptr = np.concatenate([[0], np.cumsum(lens)])
    vals = np.arange(1, indptr[-1] + 1, dtype=np.float64)
    slice_ptr, sv = O.sell32_values(indptr, vals)
    assert slice_ptr.tolist() == [0, 6 * 32, 6 * 32 + 4 * 32]          # widths 5 -> 6 and 4
    assert sv[0] == 1.0 and sv[1] == 2.0 and sv[64] == 3.0 and sv[65] == 0.0      # row 0: entries 0, 1 | 2, pad
    assert sv[2] == 0.0 and sv[3] == 0.0                                           # row 1 is empty
    assert sv[4] == 4.0 and sv[5] == 5.0 and sv[68] == 6.0 and sv[69] == 7.0 and sv[132] == 8.0 and sv[133] == 0.0
    base = slice_ptr[1]
    r32 = vals[indptr[32]:indptr[33]]
    assert sv[base] == r32[0] and sv[base + 1] == r32[1] and sv[base + 64] == 0.0
    r33 = vals[indptr[33]:indptr[34]]
    assert [sv[base + 2], sv[base + 3], sv[base + 66], sv[base + 67]] == r33.tolist()
    assert np.count_nonzero(sv) == len(vals)
