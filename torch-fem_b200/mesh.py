"""Structured mesh generators — the synthetic-input definition of the benchmarks.

Same numbering as the reference's `torchfem.mesh` (src/torchfem/mesh.py:8-215): node id of grid point
(i, j, k) is i*Ny*Nz + j*Nz + k, elements are enumerated in raveled (i, j, k) order with the
counter-clockwise bottom/top face ordering of Hexa1 / Quad1. Written as index arithmetic on one base-corner
vector instead of eight sliced views; `tests/test_mesh.py` pins the result against the reference's output.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _grid_nodes(counts, lengths) -> Tensor:
    axes = [torch.linspace(0, L, n) for n, L in zip(counts, lengths)]
    grids = torch.meshgrid(*axes, indexing="ij")
    return torch.stack([g.reshape(-1) for g in grids], dim=1)


def cube_hexa(Nx: int, Ny: int, Nz: int, Lx: float = 1.0, Ly: float = 1.0, Lz: float = 1.0):
    """Hexa1 mesh of a box with Nx x Ny x Nz grid points (reference mesh.py:8-46)."""
    nodes = _grid_nodes((Nx, Ny, Nz), (Lx, Ly, Lz))
    i, j, k = torch.meshgrid(torch.arange(Nx - 1), torch.arange(Ny - 1), torch.arange(Nz - 1), indexing="ij")
    base = (i * (Ny * Nz) + j * Nz + k).reshape(-1)
    sx, sy, sz = Ny * Nz, Nz, 1
    # local corners (x,y,z offsets) in Hexa1 order: bottom face ccw, then top face ccw
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    elements = torch.stack([base + a * sx + b * sy + c * sz for a, b, c in offs], dim=1)
    return nodes, elements


def rect_quad(Nx: int, Ny: int, Lx: float = 1.0, Ly: float = 1.0):
    """Quad1 mesh of a rectangle with Nx x Ny grid points (reference mesh.py:103-135)."""
    nodes = _grid_nodes((Nx, Ny), (Lx, Ly))
    i, j = torch.meshgrid(torch.arange(Nx - 1), torch.arange(Ny - 1), indexing="ij")
    base = (i * Ny + j).reshape(-1)
    offs = [(0, 0), (1, 0), (1, 1), (0, 1)]
    elements = torch.stack([base + a * Ny + b for a, b in offs], dim=1)
    return nodes, elements


def cube_tetra(Nx: int, Ny: int, Nz: int, Lx: float = 1.0, Ly: float = 1.0, Lz: float = 1.0):
    """Tetra1 mesh: every hexahedron split into five tetrahedra, the two mirror-image splits alternating
    in a 3-D checkerboard; all "even" cells first, then all "odd" ones (reference mesh.py:49-100)."""
    nodes, hexes = cube_hexa(Nx, Ny, Nz, Lx, Ly, Lz)
    i, j, k = torch.meshgrid(torch.arange(Nx - 1), torch.arange(Ny - 1), torch.arange(Nz - 1), indexing="ij")
    even = ((i + j + k) % 2 == 0).reshape(-1)
    split_even = torch.tensor([[0, 1, 3, 4], [1, 2, 3, 6], [1, 3, 4, 6], [1, 4, 5, 6], [3, 4, 6, 7]])
    split_odd = torch.tensor([[4, 5, 0, 7], [5, 6, 2, 7], [5, 7, 2, 0], [5, 0, 2, 1], [7, 0, 3, 2]])
    tets = torch.cat([hexes[even][:, split_even].reshape(-1, 4), hexes[~even][:, split_odd].reshape(-1, 4)])
    return nodes, tets


def rect_tri(Nx: int, Ny: int, Lx: float = 1.0, Ly: float = 1.0, variant: str = "zigzag"):
    """Tria1 mesh from a quad grid; variants 'up', 'down', 'zigzag', 'center' (reference mesh.py:138-215)."""
    nodes, quads = rect_quad(Nx, Ny, Lx, Ly)
    a, b, c, d = quads.unbind(dim=1)  # ccw corners: a=(i,j) b=(i+1,j) c=(i+1,j+1) d=(i,j+1)
    up = (torch.stack([a, b, c], 1), torch.stack([a, c, d], 1))
    down = (torch.stack([b, c, d], 1), torch.stack([b, d, a], 1))
    if variant == "up":
        return nodes, torch.vstack(up)
    if variant == "down":
        return nodes, torch.vstack(down)
    if variant == "zigzag":
        parity = ((torch.div(a, Ny, rounding_mode="floor") + a % Ny) % 2 == 0).unsqueeze(1)
        return nodes, torch.vstack([torch.where(parity, up[0], down[0]), torch.where(parity, up[1], down[1])])
    if variant == "center":
        centers = nodes[quads].mean(dim=1)
        m = torch.arange(nodes.shape[0], nodes.shape[0] + centers.shape[0])
        tris = [torch.stack([p, q, m], 1) for p, q in ((a, b), (b, c), (c, d), (d, a))]
        return torch.vstack([nodes, centers]), torch.vstack(tris).long()
    raise ValueError(f"Unknown variant: {variant}")
