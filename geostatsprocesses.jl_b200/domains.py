"""Minimal Meshes/GeoTables mirror for the boundary: CartesianGrid, PointSet, grid views, georef.

Conventions pinned by the reference: centroid = origin + (ijk - 1/2) * spacing and column-major
linear indices with x fastest (test/initialization.jl:16-21).  Indices exposed to users are 1-based
like Julia's (`view(grid, 1:5000)`, `ExplicitInit(991:1000)`).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np


class CartesianGrid:
    """CartesianGrid(nx, ny, ...) -> origin 0, spacing 1;  CartesianGrid(start, finish, dims=(...))."""

    def __init__(self, *args, dims: Optional[Sequence[int]] = None, origin=None, spacing=None):
        if dims is not None and len(args) == 2:
            start, finish = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in args)
            self.dims = tuple(int(d) for d in dims)
            self.origin = tuple(start)
            self.spacing = tuple((finish - start) / np.asarray(self.dims, dtype=np.float64))
        else:
            if dims is None:
                dims = args[0] if len(args) == 1 and not np.isscalar(args[0]) else args
            self.dims = tuple(int(d) for d in dims)
            self.origin = tuple(float(o) for o in (origin if origin is not None else [0.0] * len(self.dims)))
            self.spacing = tuple(float(s) for s in (spacing if spacing is not None else [1.0] * len(self.dims)))
        if not 1 <= len(self.dims) <= 3:
            raise ValueError("grids of dimension 1..3 are supported")

    @property
    def ndim(self) -> int:
        return len(self.dims)

    def nelements(self) -> int:
        return int(np.prod(self.dims))

    def parent(self):
        return self

    def parentindices(self) -> Optional[np.ndarray]:
        return None

    def centroids(self) -> np.ndarray:
        axes = [self.origin[a] + (np.arange(self.dims[a]) + 0.5) * self.spacing[a] for a in range(self.ndim)]
        mesh = np.meshgrid(*axes, indexing="ij")
        return np.stack([m.reshape(-1, order="F") for m in mesh], axis=1)

    def sides(self):
        return tuple(d * s for d, s in zip(self.dims, self.spacing))

    def nearest(self, p: np.ndarray) -> int:
        """0-based linear index of the element whose centroid is nearest to point p (KNearestSearch(dom, 1))."""
        lin, stride = 0, 1
        for a in range(self.ndim):
            i = int(np.floor((float(p[a]) - self.origin[a]) / self.spacing[a]))
            i = min(max(i, 0), self.dims[a] - 1)
            lin += i * stride
            stride *= self.dims[a]
        return lin

    def view(self, inds1) -> "GridView":
        return GridView(self, inds1)

    def __eq__(self, o):
        return isinstance(o, CartesianGrid) and (self.dims, self.origin, self.spacing) == (o.dims, o.origin, o.spacing)

    def __repr__(self):
        return f"CartesianGrid(dims={self.dims}, origin={self.origin}, spacing={self.spacing})"


class GridView:
    """view(grid, inds) with 1-based parent indices (test/field.jl:126-132)."""

    def __init__(self, grid: CartesianGrid, inds1):
        self.grid = grid
        self.inds1 = np.asarray(list(inds1) if not isinstance(inds1, np.ndarray) else inds1, dtype=np.int64)
        if self.inds1.min() < 1 or self.inds1.max() > grid.nelements():
            raise IndexError("view indices out of range")

    @property
    def ndim(self):
        return self.grid.ndim

    def nelements(self) -> int:
        return len(self.inds1)

    def parent(self):
        return self.grid

    def parentindices(self):
        return self.inds1

    def centroids(self) -> np.ndarray:
        return self.grid.centroids()[self.inds1 - 1]

    def nearest(self, p) -> int:
        c = self.centroids()
        return int(np.argmin(np.sum((c - np.asarray(p, dtype=np.float64)[None, :]) ** 2, axis=1)))

    def __eq__(self, o):
        return isinstance(o, GridView) and self.grid == o.grid and np.array_equal(self.inds1, o.inds1)

    def __repr__(self):
        return f"GridView({self.grid}, {len(self.inds1)} elements)"


class PointSet:
    def __init__(self, coords):
        self.coords = np.atleast_2d(np.asarray(coords, dtype=np.float64))

    @property
    def ndim(self):
        return self.coords.shape[1]

    def nelements(self) -> int:
        return self.coords.shape[0]

    def parent(self):
        return self

    def parentindices(self):
        return None

    def centroids(self) -> np.ndarray:
        return self.coords

    def nearest(self, p) -> int:
        return int(np.argmin(np.sum((self.coords - np.asarray(p, dtype=np.float64)[None, :]) ** 2, axis=1)))

    def __eq__(self, o):
        return isinstance(o, PointSet) and np.array_equal(self.coords, o.coords)

    def __repr__(self):
        return f"PointSet({self.nelements()} points)"


class GeoTable:
    """georef(table, domain): named columns over a domain."""

    def __init__(self, table: Dict[str, np.ndarray], domain):
        self.table = {k: np.asarray(v) for k, v in table.items()}
        self.domain = domain
        for k, v in self.table.items():
            if len(v) != domain.nelements():
                raise ValueError(f"column {k} has {len(v)} rows, domain has {domain.nelements()} elements")

    def __getattr__(self, name):
        t = self.__dict__.get("table", {})
        if name in t:
            return t[name]
        raise AttributeError(name)

    def __getitem__(self, name):
        return self.table[name]

    def names(self):
        return tuple(self.table.keys())

    @property
    def nrow(self) -> int:
        return self.domain.nelements()

    def __repr__(self):
        return f"GeoTable({self.nrow} rows, columns={list(self.table)})"


def georef(table: Dict[str, Sequence[float]], domain_or_coords) -> GeoTable:
    dom = domain_or_coords
    if not hasattr(dom, "nelements"):
        dom = PointSet(np.asarray(domain_or_coords, dtype=np.float64))
    return GeoTable(dict(table), dom)
