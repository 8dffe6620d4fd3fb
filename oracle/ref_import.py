"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference torch-fem from /root/reference.

Used by `oracle/make_golden.py` (build container) to generate the committed fixtures under `tests/golden/`, by
`tests/test_reference_parity.py` and by `bench.py`'s `--impl reference` / `cpu_baseline` legs. The package is looked
for at $TFEM_REFERENCE_SRC, then `/root/reference/src` (build container), then `oracle/_ref` (the untracked offline
install `oracle/build_ref.py` stages so that the reference travels to the GPU box with the snapshot).

The reference imports matplotlib / pyvista / pyamg / meshio at module import time; none of them is
installed offline, so permissive stub modules are registered first (recipe: SURVEY.md Appendix A).
`pyamg.smoothed_aggregation_solver` is replaced by a Jacobi stand-in, which is exactly the
preconditioner of the reference's GPU path (src/torchfem/sparse.py:408-409, 416-417).
"""
import importlib.machinery
import os
import sys
import types

def _find_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("TFEM_REFERENCE_SRC"), "/root/reference/src", os.path.join(here, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "torchfem")):
            return cand
    return os.path.join(here, "_ref")


REFERENCE_SRC = _find_reference()


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        m = sys.modules.get(full)
        if m is None:
            m = _Stub(full)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(full, None)
            sys.modules[full] = m
        return m

    def __call__(self, *a, **k):
        return self

    def __or__(self, o):
        return self

    __ror__ = __or__

    def __getitem__(self, k):
        return self

    def __add__(self, o):
        return self

    __radd__ = __add__

    def __mul__(self, o):
        return self

    __rmul__ = __mul__

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


_STUBBED = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.axes",
    "matplotlib.collections", "matplotlib.colors", "matplotlib.tri", "matplotlib.transforms",
    "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d", "pyvista", "pyvista.plotting",
    "pyamg", "meshio",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "torchfem"))


def load():
    """Return the reference `torchfem` package (float64 default dtype is set as its tests do)."""
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_SRC}")
    for name in _STUBBED:
        if name not in sys.modules:
            m = _Stub(name)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import torch

    torch.set_default_dtype(torch.float64)
    import torchfem
    import torchfem.sparse as S
    from scipy.sparse.linalg import LinearOperator

    class _Jacobi:
        def __init__(self, A, B=None, smooth=None):
            self.d = 1.0 / A.diagonal()
            self.shape = A.shape

        def aspreconditioner(self):
            return LinearOperator(self.shape, matvec=lambda x: self.d * x)

    S.pyamg.smoothed_aggregation_solver = _Jacobi
    return torchfem
