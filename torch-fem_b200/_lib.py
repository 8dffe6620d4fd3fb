"""ctypes binding of libtfem_b200.so (C ABI declared in include/tfem_b200.h).

Mirrors how the reference binds its only native library, AmgX (/root/reference/src/torchfem/amgx.py):
library path from an environment variable with a default next to the package (amgx.py:148-163),
`ImportError` when it cannot be loaded, every entry point returns `int rc` which `_check` turns into a
`RuntimeError` carrying `tfem_get_error_string` (amgx.py:195-201), raw device pointers from
`tensor.data_ptr()` and the caller's current CUDA stream.

There is NO fallback: if the shared library is missing or a CUDA device is absent, the ops raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ENV = "TFEM_B200_LIB"
DEFAULT_LIB = os.path.join(_HERE, "libtfem_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NOT_CONVERGED, ERR_BREAKDOWN, ERR_NCCL, ERR_COMM = range(8)
IPC_HANDLE_BYTES = 64
MAX_NEIGHBOURS = 8
TRACE_SLOTS = 16
KIND_MECH, KIND_HEAT = 0, 1
METHOD_CG, METHOD_MINRES = 0, 1
SPMV_CHUNK = 512
SELL_LONG_ROW = 1024

class SellStruct(ctypes.Structure):
    """`tfem_sell_t` of include/tfem_b200.h."""
    _fields_ = [("n_rows", c_int64), ("slice_ptr", c_void_p), ("cols", c_void_p), ("vals", c_void_p),
                ("bslice_ptr", c_void_p), ("bcols", c_void_p), ("dpn", ctypes.c_int32), ("n_long", ctypes.c_int32),
                ("long_rows", c_void_p), ("csr_indptr", c_void_p), ("csr_cols", c_void_p), ("csr_vals", c_void_p)]


_SELL_P = ctypes.POINTER(SellStruct)


class EbeStruct(ctypes.Structure):
    """`tfem_ebe_t` of include/tfem_b200.h."""
    _fields_ = [("n_nod", c_int64), ("nn", ctypes.c_int32), ("dpn", ctypes.c_int32), ("inc_ptr", c_void_p),
                ("inc_list", c_void_p), ("elements", c_void_p), ("k", c_void_p), ("is_con", c_void_p)]


_EBE_P = ctypes.POINTER(EbeStruct)


class BcsrStruct(ctypes.Structure):
    """`tfem_bcsr_t` of include/tfem_b200.h."""
    _fields_ = [("nb_rows", c_int64), ("n_blocks", c_int64), ("bptr", c_void_p), ("bcol", c_void_p),
                ("vals", c_void_p), ("d", ctypes.c_int32)]


class AmgOperatorStruct(ctypes.Structure):
    """`tfem_amg_operator_t` of include/tfem_b200.h."""
    _fields_ = [("sell", SellStruct), ("bcsr", BcsrStruct)]


class AmgLevelStruct(ctypes.Structure):
    """`tfem_amg_level_t` of include/tfem_b200.h."""
    _fields_ = [("A", AmgOperatorStruct), ("P", AmgOperatorStruct), ("R", AmgOperatorStruct), ("dinv", c_void_p),
                ("omega", c_double), ("x", c_void_p), ("b", c_void_p), ("t", c_void_p)]


_AMG_P = ctypes.POINTER(AmgLevelStruct)


class HaloSendStruct(ctypes.Structure):
    """`tfem_halo_send_t` of include/tfem_b200.h."""
    _fields_ = [("peer", ctypes.c_int32), ("count", c_int64), ("src_idx", c_void_p), ("dst_idx", c_void_p),
                ("src_start", c_int64), ("dst_start", c_int64)]


class DamgLevelStruct(ctypes.Structure):
    """`tfem_damg_level_t` of include/tfem_b200.h."""
    _fields_ = [("lv", AmgLevelStruct), ("own_lo", c_int64), ("own_hi", c_int64), ("c_own_lo", c_int64),
                ("c_own_hi", c_int64), ("n_sends", ctypes.c_int32), ("sends", c_void_p), ("n_recv", ctypes.c_int32),
                ("recv_peers", c_void_p)]


_SIGNATURES = {
    "tfem_version": (c_int, []),
    "tfem_get_error_string": (c_int, [c_int, c_char_p, c_int]),
    "tfem_pattern_phase1": (c_int, [c_int64, c_int64, c_int, c_int] + [c_void_p] * 6),
    "tfem_pattern_phase2": (c_int, [c_int64, c_int64, c_int, c_int] + [c_void_p] * 12),
    "tfem_pattern_k_map": (c_int, [c_int64, c_int64, c_int, c_int] + [c_void_p] * 6),
    "tfem_pattern_coo_rows": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "tfem_integrate_k": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfem_elem_grad": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "tfem_elem_force": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "tfem_ddot": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "tfem_ddot_outer": (c_int, [c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfem_assemble_rhs": (c_int, [c_int64, c_int] + [c_void_p] * 5),
    "tfem_assemble": (c_int, [c_int64, c_int, c_int] + [c_void_p] * 9),
    "tfem_assemble_bc": (c_int, [c_int64, c_int, c_int] + [c_void_p] * 11),
    "tfem_assemble_solve": (c_int, [c_int64, c_int, c_int] + [c_void_p] * 14),
    "tfem_spmv_num_chunks": (c_int64, [c_int64]),
    "tfem_spmv_plan": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "tfem_spmv": (c_int, [c_int64, c_int64] + [c_void_p] * 7),
    "tfem_csr_transpose": (c_int, [c_int64, c_int64, c_int64] + [c_void_p] * 7),
    "tfem_csr_diag_positions": (c_int, [c_int64] + [c_void_p] * 4),
    "tfem_jacobi_setup": (c_int, [c_int64] + [c_void_p] * 4),
    "tfem_krylov_work_doubles": (c_int64, [c_int64]),
    "tfem_krylov_solve": (c_int, [c_int, _SELL_P] + [c_void_p] * 3 + [c_double, c_double, c_int64, c_int]
                          + [c_void_p] * 4),
    "tfem_ebe_spmv": (c_int, [_EBE_P, c_void_p, c_void_p, c_void_p]),
    "tfem_ebe_diag": (c_int, [_EBE_P, c_void_p, c_void_p]),
    "tfem_krylov_solve_ebe": (c_int, [c_int, _EBE_P] + [c_void_p] * 3 + [c_double, c_double, c_int64, c_int]
                              + [c_void_p] * 4),
    "tfem_bsell_slice_ptr": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "tfem_bsell_fill": (c_int, [c_int64, c_int, c_int64] + [c_void_p] * 5),
    "tfem_sell_slice_ptr": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "tfem_sell_fill": (c_int, [c_int64] + [c_void_p] * 7),
    "tfem_sell_slice_ptr_capped": (c_int, [c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "tfem_sell_fill_capped": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64] + [c_void_p] * 4),
    "tfem_sell_fill_rect": (c_int, [c_int64, c_int64] + [c_void_p] * 7),
    "tfem_sell_spmv": (c_int, [_SELL_P, c_void_p, c_void_p, c_void_p]),
    "tfem_sell_spmm": (c_int, [_SELL_P, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "tfem_amg_row_info": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "tfem_amg_work_doubles": (c_int64, [c_int64]),
    "tfem_amg_rho": (c_int, [ctypes.POINTER(AmgOperatorStruct), c_void_p, c_int, c_void_p, ctypes.POINTER(c_double),
                             c_void_p]),
    "tfem_amg_aggregate": (c_int, [c_int64, c_void_p, c_void_p, c_int] + [c_void_p] * 4
                           + [ctypes.POINTER(c_int64), ctypes.POINTER(ctypes.c_int32), c_void_p]),
    "tfem_amg_aggregate_masked": (c_int, [c_int64, c_void_p, c_void_p, c_int] + [c_void_p] * 5
                                  + [ctypes.POINTER(c_int64), ctypes.POINTER(ctypes.c_int32), c_void_p]),
    "tfem_amg_prolongator_count_rows": (c_int, [c_int, c_int64, c_int64] + [c_void_p] * 5),
    "tfem_amg_prolongator_fill_rows": (c_int, [c_int, c_int64, c_int64] + [c_void_p] * 6 + [c_double] + [c_void_p] * 3
                                       + [c_int, c_void_p]),
    "tfem_damg_pcg_solve": (c_int, [c_void_p, ctypes.POINTER(DamgLevelStruct), c_int, _AMG_P, c_int, c_void_p, c_int64]
                            + [c_void_p] * 6 + [c_double, c_double, c_int64, c_double, c_void_p, c_void_p]),
    "tfem_amg_prolongator_count": (c_int, [c_int, c_int64] + [c_void_p] * 5),
    "tfem_amg_prolongator_fill": (c_int, [c_int, c_int64] + [c_void_p] * 6 + [c_double] + [c_void_p] * 3
                                  + [c_int, c_void_p]),
    "tfem_amg_transpose_structure": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int64] + [c_void_p] * 4),
    "tfem_amg_transpose_values": (c_int, [c_int, c_int64] + [c_void_p] * 7),
    "tfem_amg_spgemm_count": (c_int, [c_int64] + [c_void_p] * 6),
    "tfem_amg_spgemm_fill": (c_int, [c_int64] + [c_void_p] * 7),
    "tfem_amg_spgemm_numeric": (c_int, [c_int, c_int64] + [c_void_p] * 9 + [c_int, c_int, c_void_p]),
    "tfem_amg_vcycle": (c_int, [_AMG_P, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfem_amg_vcycle_block": (c_int, [_AMG_P, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfem_amg_pcg_solve": (c_int, [_AMG_P, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "tfem_adjoint_matrix_grad": (c_int, [c_int64] + [c_void_p] * 6),
    "tfem_cg_stage": (c_int, [c_int, _SELL_P, c_int64, c_int64] + [c_void_p] * 5 + [c_double, c_double, c_void_p]),
    "tfem_comm_create": (c_int, [c_int, c_int, c_int64, ctypes.POINTER(c_void_p), c_void_p]),
    "tfem_comm_connect": (c_int, [c_void_p, c_void_p]),
    "tfem_comm_destroy": (c_int, [c_void_p]),
    "tfem_comm_heap": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64)]),
    "tfem_comm_set_trace": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int]),
    "tfem_dcg_solve": (c_int, [c_void_p, _SELL_P, c_int64, c_int64, c_int64, c_int64, c_int,
                               ctypes.POINTER(HaloSendStruct), c_int, c_void_p] + [c_void_p] * 4
                       + [c_double, c_double, c_int64, c_int, c_double, c_void_p, c_void_p]),
    "tfem_krylov_work_offset": (c_int64, [c_int64, c_int]),
    "tfem_krylov_state": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
}

EXPORTED = tuple(_SIGNATURES)


def _load() -> ctypes.CDLL:
    path = os.environ.get(LIB_ENV, DEFAULT_LIB)
    try:
        lib = ctypes.CDLL(path)
    except OSError as exc:  # same contract as amgx.py:153-157
        raise ImportError(
            f"libtfem_b200.so could not be loaded from {path!r} ({exc}). Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or `make -C torch-fem_b200/csrc`, "
            f"or point {LIB_ENV} at the shared library."
        ) from exc
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
LIB_PATH = os.environ.get(LIB_ENV, DEFAULT_LIB)


class TfemError(RuntimeError):
    """A C-ABI call failed. Subclass of RuntimeError on purpose: `FEM.solve` cuts the load step back on
    exactly that exception type when the Krylov solver fails (reference base.py:831)."""

    def __init__(self, rc: int, msg: str):
        super().__init__(f"tfem_b200 error {rc}: {msg}")
        self.rc = rc


def error_string(rc: int) -> str:
    buf = ctypes.create_string_buffer(1024)
    lib.tfem_get_error_string(rc, buf, len(buf))
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != OK:
        raise TfemError(rc, error_string(rc))


def require_cuda(*tensors: torch.Tensor) -> None:
    """No CPU fallback: every tensor handed to a kernel must live on a CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "torch-fem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback. "
            "The CPU oracle under oracle/ is test infrastructure and is never used by the product."
        )
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(f"expected a CUDA tensor, got device {t.device}")


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    assert t.is_contiguous(), "kernel arguments must be contiguous"
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream
