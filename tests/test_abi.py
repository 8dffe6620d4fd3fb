"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/tfem_b200.h declares; the product refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tfem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import torchfem_b200 as T

    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(T._lib.lib, s), f"{s} declared in include/tfem_b200.h but not exported"
    assert set(T._lib.EXPORTED) == set(syms), set(T._lib.EXPORTED) ^ set(syms)
    assert T._lib.lib.tfem_version() >= 100
    assert T._lib.error_string(0) == "success"
    assert "invalid" in T._lib.error_string(1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    import torchfem_b200 as T

    el = torch.zeros(1, 8, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.csr.Pattern(el, 8, 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torch-fem_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "fem_oracle" not in text, fn


def test_solver_policy_and_amg_refuses_cpu():
    """Host logic without a device: the size policy of `resolve_method` (reference sparse.py:74-84 + the AMG threshold),
    the reference's method names / error messages, and the AMG preconditioner's refusal to run on CPU tensors."""
    import torchfem_b200 as T
    from torchfem_b200.amg import AMGPreconditioner

    S = T.sparse
    assert S.resolve_method(S.DIRECT_LIMIT - 1, "cuda", None) == "spsolve"
    assert S.resolve_method(S.DIRECT_LIMIT, "cuda", None) == "minres"
    assert S.resolve_method(S.AMG_MIN_DOFS, "cuda", None) == "amgx"
    assert S.resolve_method(10, "cuda", "cg") == "cg"
    assert S.describe_method(5 * S.AMG_MIN_DOFS, "cuda", None) == "amgx | iterative | amg | tfem_b200 | cuda"
    assert S.METHODS == ["spsolve", "minres", "cg", "pardiso", "amgx"] and "amgx" in S.available_backends
    with pytest.raises(TypeError, match="assembled CSRMatrix"):
        AMGPreconditioner(torch.eye(3))
    A = torch.eye(4).to_sparse_coo()
    with pytest.raises(ValueError, match="is not supported"):
        S.sparse_solve(A, torch.ones(4), method="multigrid")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            S.sparse_solve(A, torch.ones(4), method="amgx")
