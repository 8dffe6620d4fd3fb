"""Isoparametric element tables of the hot path: Tria1/2, Quad1/2, Tetra1/2, Hexa1/2.

Same public surface as the reference's `torchfem.elements` for these types (`N`, `B`, `iso_coords`,
`ipoints`, `iweights`, `edges`, `nodes`, `iso_dim`, `iso_volume`; src/torchfem/elements.py:63-171,
309-1406) and the same numbers — node ordering, integration points and weights (including the
low-precision Tetra2 literals, elements.py:920-933, the integer Quad weights, :555-557, and the 2x2x2
reduced rule of Hexa2, :1382-1406) — but generated from closed forms over the node-coordinate table
instead of per-node expressions:

  linear box elements      N_a = 2^-d  prod_c (1 + s_ac xi_c)
  serendipity, corner a    N_a = 2^-d  prod_c (1 + s_ac xi_c) * (sum_c s_ac xi_c - (d-1))
  serendipity, mid-side a  N_a = 2^-(d-1) (1 - xi_m^2) prod_{c != m} (1 + s_ac xi_c)    (s_am = 0)
  simplices                barycentric: linear N = lambda, quadratic lambda(2 lambda - 1) and 4 lambda_a lambda_b

`tests/test_elements.py` checks every table and N/B at random points against fixtures produced by the
reference itself. The CUDA integration kernel receives `B(ipoints)` and `iweights` from here.
"""
from __future__ import annotations

from math import sqrt

import numpy as np
import torch
from torch import Tensor


class _classproperty:
    def __init__(self, fget):
        self.fget = fget

    def __get__(self, obj, cls):
        return self.fget(cls)


ELEMENT_REGISTRY: list[type["Element"]] = []


class Element:
    """Base of all element types; subclasses fill `_ISO` (node coordinates in reference space)."""

    iso_volume: float
    iso_dim: int
    nodes: int
    _ISO: list[list[float]]

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "_ISO" in cls.__dict__:
            ELEMENT_REGISTRY.append(cls)

    @_classproperty
    def iso_coords(cls) -> Tensor:
        return torch.tensor(cls._ISO)

    @_classproperty
    def edges(cls) -> Tensor:
        """Corner pairs of the element edges, ordered like the mid-side nodes of the quadratic
        element of the same shape (what `linear_to_quadratic` relies on)."""
        quad = cls._quadratic()
        iso = np.asarray(quad._ISO, dtype=float)
        nc = cls._n_corner()
        pairs = []
        for m in range(nc, len(iso)):
            hit = [(a, b) for a in range(nc) for b in range(a + 1, nc)
                   if np.allclose(0.5 * (iso[a] + iso[b]), iso[m])]
            a, b = hit[0]
            pair = quad._orient_edge(m - nc, a, b)
            pairs.append(pair + [m] if cls is quad else pair)  # quadratic types also list the mid node
        return torch.tensor(pairs)

    @classmethod
    def _orient_edge(cls, k, a, b):
        return [a, b]

    @classmethod
    def _quadratic(cls):
        return cls

    @classmethod
    def _corner_facets(cls) -> list[list[int]]:
        """Corner nodes of every codimension-1 facet, wound so that the normal points out of the element."""
        raise NotImplementedError

    @_classproperty
    def facets(cls) -> Tensor:
        """Facet connectivity (edges of a planar element, faces of a solid one), corner nodes first in outward
        winding, then — on quadratic elements — the mid-side node after each corner (reference elements.py
        `facets`, e.g. :954-966 for Hexa1); derived here from the reference coordinates."""
        iso = np.asarray(cls._ISO, dtype=float)
        nc = cls._n_corner()
        rows = []
        for corners in cls._corner_facets():
            row = list(corners)
            if cls.nodes > nc:
                ring = corners if len(corners) > 2 else corners[:1]     # an edge has one mid node, a face one per side
                for k, a in enumerate(ring):
                    b = corners[(k + 1) % len(corners)]
                    mid = 0.5 * (iso[a] + iso[b])
                    row.append(next(m for m in range(nc, cls.nodes) if np.allclose(iso[m], mid)))
            rows.append(row)
        return torch.tensor(rows)

    @_classproperty
    def facet_type(cls) -> type["Element"]:
        """Element type of a facet: Bar for planar elements, Tria / Quad for tetrahedra / hexahedra, of the same order."""
        quadratic = cls.nodes > cls._n_corner()
        if cls.iso_dim == 2:
            return Bar2 if quadratic else Bar1
        if issubclass(cls, _Simplex):
            return Tria2 if quadratic else Tria1
        return Quad2 if quadratic else Quad1

    @classmethod
    def N(cls, xi: Tensor) -> Tensor:
        raise NotImplementedError

    @classmethod
    def B(cls, xi: Tensor) -> Tensor:
        raise NotImplementedError


# ------------------------------------------------------------------------------------------ box family
class _Box(Element):
    """Linear (2^d nodes) and serendipity (2^d + mid-side nodes) quadrilaterals / hexahedra."""

    @classmethod
    def _signs(cls, xi: Tensor) -> Tensor:
        return torch.tensor(cls._ISO, dtype=xi.dtype, device=xi.device)

    @classmethod
    def _n_corner(cls):
        return 2 ** cls.iso_dim

    @classmethod
    def _corner_facets(cls):
        if cls.iso_dim == 2:     # the four sides in node order (counter-clockwise numbering: outward normals)
            return [[k, (k + 1) % 4] for k in range(4)]
        # bottom seen from below, top seen from above, then the four sides around the axis
        return [[0, 3, 2, 1], [4, 5, 6, 7]] + [[k, (k + 1) % 4, (k + 1) % 4 + 4, k + 4] for k in range(4)]

    @classmethod
    def N(cls, xi: Tensor) -> Tensor:
        s = cls._signs(xi)  # [nn, d]
        d = cls.iso_dim
        x = xi.unsqueeze(-2)  # [..., 1, d]
        t = 1.0 + s * x  # [..., nn, d]
        prod = t.prod(-1)
        nc = cls._n_corner()
        if cls.nodes == nc:
            return prod / 2 ** d
        corner = prod[..., :nc] * ((s[:nc] * x).sum(-1) - (d - 1)) / 2 ** d
        # mid-side: the zero sign marks the edge direction m; its factor (1 + 0*xi) = 1 is replaced by (1 - xi_m^2)
        sm = s[nc:]
        bubble = ((sm == 0) * (1.0 - x * x) + (sm != 0) * 1.0).prod(-1)
        mid = prod[..., nc:] * bubble / 2 ** (d - 1)
        return torch.cat([corner, mid], dim=-1)

    @classmethod
    def B(cls, xi: Tensor) -> Tensor:
        s = cls._signs(xi)
        d = cls.iso_dim
        nc = cls._n_corner()
        x = xi.unsqueeze(-2)
        t = 1.0 + s * x  # [..., nn, d]
        rows = []
        for c in range(d):
            others = [m for m in range(d) if m != c]
            po = t[..., others].prod(-1)  # prod over the other directions, [..., nn]
            if cls.nodes == nc:
                rows.append(s[:, c] * po / 2 ** d)
                continue
            # corners: d/dxi_c [ P * (S - (d-1)) ],  P = prod_c t_c,  S = sum_c s_c xi_c
            P = t[..., :nc, :].prod(-1)
            S = (s[:nc] * x).sum(-1)
            dc = s[:nc, c] * po[..., :nc] * (S - (d - 1)) + P * s[:nc, c]
            corner = dc / 2 ** d
            # mid-side nodes with edge direction m:
            #   c == m : -2 xi_c * prod_{o != m} t_o ;   c != m : (1 - xi_m^2) * s_c * prod_{o != m, c} t_o
            sm = s[nc:]
            tm = t[..., nc:, :]
            xm = x.expand(tm.shape)
            is_m = sm == 0
            along = is_m[:, c]  # this node's edge runs along c
            prod_not_m = torch.where(is_m, torch.ones_like(tm), tm).prod(-1)
            val_along = -2.0 * xi[..., c].unsqueeze(-1) * prod_not_m
            bubble = torch.where(is_m, 1.0 - xm * xm, torch.ones_like(tm)).prod(-1)
            keep = torch.ones(d, dtype=torch.bool, device=xi.device)
            keep[c] = False
            prod_rest = torch.where(is_m | ~keep, torch.ones_like(tm), tm).prod(-1)
            val_across = bubble * sm[:, c] * prod_rest
            mid = torch.where(along, val_along, val_across) / 2 ** (d - 1)
            rows.append(torch.cat([corner, mid], dim=-1))
        return torch.stack(rows, dim=-2)


def _gauss_box(d: int) -> list[list[float]]:
    g = 1.0 / sqrt(3.0)
    if d == 2:
        return [[x1 * g, x2 * g] for x2 in (-1.0, 1.0) for x1 in (-1.0, 1.0)]
    return [[x1 * g, x2 * g, x3 * g] for x3 in (-1.0, 1.0) for x2 in (-1.0, 1.0) for x1 in (-1.0, 1.0)]


class Quad1(_Box):
    iso_volume, iso_dim, nodes = 4.0, 2, 4
    _ISO = [[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]]

    @classmethod
    def _quadratic(cls):
        return Quad2

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor(_gauss_box(2))

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([1, 1, 1, 1])  # integer tensor, as in the reference (elements.py:555-557)


class Quad2(Quad1):
    nodes = 8
    _ISO = Quad1._ISO + [[0.0, -1.0], [1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]]

    @classmethod
    def _orient_edge(cls, k, a, b):
        return [3, 0] if (a, b) == (0, 3) else [a, b]


class Hexa1(_Box):
    iso_volume, iso_dim, nodes = 8.0, 3, 8
    _ISO = [[-1.0, -1.0, -1.0], [1.0, -1.0, -1.0], [1.0, 1.0, -1.0], [-1.0, 1.0, -1.0],
            [-1.0, -1.0, 1.0], [1.0, -1.0, 1.0], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.0]]

    @classmethod
    def _quadratic(cls):
        return Hexa2

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor(_gauss_box(3))

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([1.0] * 8)


class Hexa2(Hexa1):
    nodes = 20
    _ISO = Hexa1._ISO + [
        [0.0, -1.0, -1.0], [1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [-1.0, 0.0, -1.0],
        [0.0, -1.0, 1.0], [1.0, 0.0, 1.0], [0.0, 1.0, 1.0], [-1.0, 0.0, 1.0],
        [-1.0, -1.0, 0.0], [1.0, -1.0, 0.0], [1.0, 1.0, 0.0], [-1.0, 1.0, 0.0]]

    @classmethod
    def _orient_edge(cls, k, a, b):
        return {(0, 3): [3, 0], (4, 7): [7, 4]}.get((a, b), [a, b])


# ------------------------------------------------------------------------------------------ simplices
class _Simplex(Element):
    @classmethod
    def _n_corner(cls):
        return cls.iso_dim + 1

    @classmethod
    def _corner_facets(cls):
        if cls.iso_dim == 2:
            return [[0, 1], [1, 2], [2, 0]]
        iso = np.asarray(cls._ISO[:4], dtype=float)
        faces = []
        for opposite in (3, 2, 0, 1):      # the face opposite each vertex, lowest node first, normal pointing away from it
            a, b, c = [v for v in range(4) if v != opposite]
            normal = np.cross(iso[b] - iso[a], iso[c] - iso[a])
            faces.append([a, b, c] if normal @ (iso[a] - iso[opposite]) > 0 else [a, c, b])
        return faces

    @classmethod
    def _lambda(cls, xi: Tensor) -> Tensor:
        return torch.cat([1.0 - xi.sum(-1, keepdim=True), xi], dim=-1)  # [..., d+1]

    @classmethod
    def _dlambda(cls, xi: Tensor) -> Tensor:
        d = cls.iso_dim
        g = torch.cat([-torch.ones(d, 1), torch.eye(d)], dim=1).to(dtype=xi.dtype, device=xi.device)
        return g.expand(*xi.shape[:-1], d, d + 1)  # [..., d, d+1] : d lambda_a / d xi_c

    @classmethod
    def _pairs(cls):
        return cls.edges.tolist()

    @classmethod
    def N(cls, xi: Tensor) -> Tensor:
        lam = cls._lambda(xi)
        if cls.nodes == cls.iso_dim + 1:
            return lam
        corner = lam * (2.0 * lam - 1.0)
        mid = torch.stack([4.0 * lam[..., a] * lam[..., b] for a, b in cls._PAIRS], dim=-1)
        return torch.cat([corner, mid], dim=-1)

    @classmethod
    def B(cls, xi: Tensor) -> Tensor:
        dl = cls._dlambda(xi)
        if cls.nodes == cls.iso_dim + 1:
            return dl + 0.0 * xi.sum(-1)[..., None, None]
        lam = cls._lambda(xi).unsqueeze(-2)  # [..., 1, d+1]
        corner = (4.0 * lam - 1.0) * dl
        mid = torch.stack([4.0 * (dl[..., a] * lam[..., b] + lam[..., a] * dl[..., b])
                           for a, b in cls._PAIRS], dim=-1)
        return torch.cat([corner, mid], dim=-1)


class Tria1(_Simplex):
    iso_volume, iso_dim, nodes = 0.5, 2, 3
    _ISO = [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]

    @classmethod
    def _quadratic(cls):
        return Tria2

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor([[1.0 / 3.0, 1.0 / 3.0]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([0.5])


class Tria2(Tria1):
    nodes = 6
    _ISO = Tria1._ISO + [[0.5, 0.0], [0.5, 0.5], [0.0, 0.5]]
    _PAIRS = [(0, 1), (1, 2), (2, 0)]

    @classmethod
    def _orient_edge(cls, k, a, b):
        return [2, 0] if (a, b) == (0, 2) else [a, b]

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor([[0.5, 0.5], [0.5, 0.0], [0.0, 0.5]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0])


class Tetra1(_Simplex):
    iso_volume, iso_dim, nodes = 1.0 / 6.0, 3, 4
    _ISO = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]

    @classmethod
    def _quadratic(cls):
        return Tetra2

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor([[0.25, 0.25, 0.25]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([1.0 / 6.0])


class Tetra2(Tetra1):
    nodes = 10
    _ISO = Tetra1._ISO + [[0.5, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5],
                          [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]
    _PAIRS = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]

    @classmethod
    def _orient_edge(cls, k, a, b):
        return [3, 0] if (a, b) == (0, 3) else [a, b]

    @_classproperty
    def ipoints(cls) -> Tensor:
        a, b = 0.58541020, 0.13819660  # the reference's 8-digit literals (elements.py:925-933)
        return torch.tensor([[a, b, b], [b, a, b], [b, b, a], [b, b, b]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([0.041666667] * 4)  # sic (elements.py:921-922)


# ------------------------------------------------------------------------------------------ line facets
class Bar1(Element):
    """Two-node line: the facet of linear planar elements (reference elements.py:174-235). Facet use only — bar
    *models* (trusses) are outside this package."""

    iso_volume, iso_dim, nodes = 2.0, 1, 2
    _ISO = [[-1.0], [1.0]]

    @classmethod
    def _n_corner(cls):
        return 2

    @classmethod
    def N(cls, xi: Tensor) -> Tensor:
        x = xi[..., 0]
        return torch.stack([0.5 * (1.0 - x), 0.5 * (1.0 + x)], dim=-1)

    @classmethod
    def B(cls, xi: Tensor) -> Tensor:
        x = xi[..., 0]
        return torch.stack([-0.5 + 0.0 * x, 0.5 + 0.0 * x], dim=-1).unsqueeze(-2)

    @_classproperty
    def ipoints(cls) -> Tensor:
        return torch.tensor([[0.0]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([2.0])


class Bar2(Bar1):
    """Three-node line (end nodes, then the mid node): the facet of quadratic planar elements (elements.py:238-300)."""

    nodes = 3
    _ISO = [[-1.0], [1.0], [0.0]]

    @classmethod
    def N(cls, xi: Tensor) -> Tensor:
        x = xi[..., 0]
        return torch.stack([0.5 * x * (x - 1.0), 0.5 * x * (x + 1.0), 1.0 - x * x], dim=-1)

    @classmethod
    def B(cls, xi: Tensor) -> Tensor:
        x = xi[..., 0]
        return torch.stack([x - 0.5, x + 0.5, -2.0 * x], dim=-1).unsqueeze(-2)

    @_classproperty
    def ipoints(cls) -> Tensor:
        g = 1.0 / sqrt(3.0)
        return torch.tensor([[-g], [g]])

    @_classproperty
    def iweights(cls) -> Tensor:
        return torch.tensor([1.0, 1.0])


def linear_etype(nodes: Tensor, elements: Tensor) -> type[Element]:
    """Linear element type of a mesh from (nodes per element, dimension) (reference elements.py:1409-1435)."""
    table = {(3, 2): Tria1, (4, 2): Quad1, (4, 3): Tetra1, (8, 3): Hexa1}
    key = (int(elements.shape[1]), int(nodes.shape[1]))
    if key not in table:
        raise Exception("The element type is not supported. Maybe the element is already quadratic?")
    return table[key]


def linear_to_quadratic(nodes: Tensor, elements: Tensor) -> tuple[Tensor, Tensor]:
    """Insert mid-side nodes (reference elements.py:1438-1494).

    The new nodes are appended in the order of the reference's CPU branch — `np.unique` over the
    raw bytes of the (min,max) int64 edge pairs (memcmp order of little-endian words,
    elements.py:1475-1484) — so a mesh converted here is node-for-node the mesh the reference's tests
    and benchmarks build. Runs on the host (mesh generation is input preparation, not the hot path).
    """
    dev = elements.device
    el = elements.detach().cpu()
    nd = nodes.detach().cpu()
    edges = linear_etype(nodes, elements).edges.cpu()
    ev = torch.sort(el[:, edges].reshape(-1, 2), dim=1).values.numpy()
    ev = np.ascontiguousarray(ev)
    raw = ev.view(np.dtype((np.void, ev.dtype.itemsize * 2)))
    _, first, inverse = np.unique(raw, return_index=True, return_inverse=True)
    uniq = torch.from_numpy(np.ascontiguousarray(ev[first]))  # from_numpy: always a host tensor
    mids = (nd[uniq[:, 0]] + nd[uniq[:, 1]]) / 2.0
    new_nodes = torch.cat([nd, mids], dim=0)
    mid_ids = torch.from_numpy(np.ascontiguousarray(inverse.reshape(el.shape[0], -1))) + nd.shape[0]
    new_elements = torch.cat([el, mid_ids], dim=1)
    return new_nodes.to(nodes.device), new_elements.to(dev)
