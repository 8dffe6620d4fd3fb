"""2-D models `Planar` / `PlanarHeat` with element thickness (drop-in for reference
src/torchfem/planar.py:17-96, 306-339; plotting is out of scope)."""
from __future__ import annotations

from functools import cached_property

import torch
from torch import Tensor

from .base import FEM, Heat, Mechanics
from .elements import Element, Quad1, Quad2, Tria1, Tria2
from .materials import Material


class PlanarGeometry(FEM):
    _ETYPES = {3: Tria1, 4: Quad1, 6: Tria2, 8: Quad2}

    def __init__(self, nodes: Tensor, elements: Tensor, material: Material, thickness: Tensor | float = 1.0):
        super().__init__(nodes, elements, material)
        if isinstance(thickness, float):
            self.thickness = torch.full((self.n_elem,), thickness, dtype=self.nodes.dtype, device=self.device)
        else:
            self.thickness = torch.as_tensor(thickness).to(self.device)

    def __repr__(self) -> str:
        return f"<torch-fem planar ({self.n_nod} nodes, {self.n_elem} {self.etype.__name__} elements)>"

    @property
    def etype(self) -> type[Element]:
        """Element type from the connectivity width (reference planar.py:62-74)."""
        try:
            return self._ETYPES[int(self.elements.shape[1])]
        except KeyError:
            raise ValueError("Element type not supported.") from None

    @cached_property
    def char_lengths(self) -> Tensor:
        return self.integrate_field() ** (1 / 2)

    @property
    def volume_scale(self) -> Tensor:
        return self.thickness

    @property
    def _k_scale(self) -> Tensor:
        return self.thickness.detach()

    @property
    def _f_scale(self) -> Tensor:
        return self.thickness

    def compute_k(self, detJ: Tensor, BCB: Tensor) -> Tensor:
        """thickness * detJ * BCB (reference planar.py:86-88); kernel K1 applies the same factors."""
        return (self.thickness * detJ)[:, None, None] * BCB

    def compute_f(self, detJ: Tensor, B: Tensor, S: Tensor) -> Tensor:
        """thickness * detJ * B^T S (reference planar.py:90-92)."""
        return torch.einsum("...,...,...iI,...Ai->...IA", self.thickness, detJ, B, S)

    def compute_m(self, detJ: Tensor, rho: Tensor) -> Tensor:
        return rho * self.thickness * detJ


class Planar(PlanarGeometry, Mechanics):
    """Plane-stress / plane-strain mechanics, 2 DOFs per node."""

    @property
    def n_flux(self) -> list[int]:
        return [2, 2]


class PlanarHeat(PlanarGeometry, Heat):
    """Planar heat conduction, one temperature DOF per node."""
