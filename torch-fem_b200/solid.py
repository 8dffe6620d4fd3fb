"""3-D continuum models `Solid` / `SolidHeat` (drop-in for reference src/torchfem/solid.py:15-62, 236-271;
plotting is out of scope)."""
from __future__ import annotations

from functools import cached_property

import torch
from torch import Tensor

from .base import FEM, Heat, Mechanics
from .elements import Element, Hexa1, Hexa2, Tetra1, Tetra2


class SolidGeometry(FEM):
    """Element choice and integration factors shared by the solid models."""

    _ETYPES = {4: Tetra1, 8: Hexa1, 10: Tetra2, 20: Hexa2}

    def __repr__(self) -> str:
        return f"<torch-fem solid ({self.n_nod} nodes, {self.n_elem} {self.etype.__name__} elements)>"

    @property
    def etype(self) -> type[Element]:
        """Element type from the connectivity width (reference solid.py:32-44)."""
        try:
            return self._ETYPES[int(self.elements.shape[1])]
        except KeyError:
            raise ValueError("Element type not supported.") from None

    @cached_property
    def char_lengths(self) -> Tensor:
        return self.integrate_field() ** (1 / 3)

    def compute_k(self, detJ: Tensor, BCB: Tensor) -> Tensor:
        """detJ-weighted Gauss-point tangent (reference solid.py:52-54); kernel K1 applies the same factor."""
        return detJ[:, None, None] * BCB

    def compute_f(self, detJ: Tensor, B: Tensor, S: Tensor) -> Tensor:
        """Internal nodal forces detJ * B^T S, [n_elem, nn, dpn] (reference solid.py:56-58)."""
        return torch.einsum("...,...iI,...Ai->...IA", detJ, B, S)

    def compute_m(self, detJ: Tensor, rho: Tensor) -> Tensor:
        return rho * detJ


class Solid(SolidGeometry, Mechanics):
    """Solid mechanics, 3 DOFs per node."""

    @property
    def n_flux(self) -> list[int]:
        return [3, 3]


class SolidHeat(SolidGeometry, Heat):
    """Heat conduction in a solid, one temperature DOF per node."""
