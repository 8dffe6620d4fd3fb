"""torch-fem_b200 — B200-native implicit-solve hot path of torch-fem (integration -> assembly -> Krylov).

Drop-in surface (same names as the reference's `torchfem/__init__.py:1-6` for the models on the path):
`Solid`, `SolidHeat`, `Planar`, `PlanarHeat`, `Assembly`, `ReferencePoint`, `ReferencePointHeat`, plus the modules `sparse`, `materials`, `mesh`, `elements`.
All numerics run in hand-written sm_100a kernels behind the C ABI of `libtfem_b200.so`
(include/tfem_b200.h); there is no CPU fallback.
"""
from . import _lib, csr  # noqa: F401  (loads libtfem_b200.so; ImportError if it is missing)
from . import elements, materials, mesh, sparse  # noqa: F401
from .assembly import Assembly, ReferencePoint, ReferencePointHeat
from .planar import Planar, PlanarHeat
from .solid import Solid, SolidHeat

__version__ = "0.1.0"
__all__ = ["Solid", "SolidHeat", "Planar", "PlanarHeat", "Assembly", "ReferencePoint", "ReferencePointHeat", "sparse", "materials", "mesh", "elements", "csr"]
