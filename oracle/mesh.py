"""Cartesian refine_global mesh hierarchy, deal.II cell order and DG numbering (oracle).

Restates what the reference obtains from deal.II for a uniformly refined
Cartesian triangulation (reference call sites
navier_stokes_TRBDF2_DG.cc:2261-2270 create_triangulation / refine_global,
:2282-2283 distribute_dofs, :2348-2349 distribute_mg_dofs):

* coarse cells of ``subdivided_hyper_rectangle`` are numbered lexicographically
  (x fastest); ``refine_global`` creates the 2^dim children of each cell
  consecutively in child order c = i + 2j + 4k, so level-l cells are ordered
  coarse cell by coarse cell, Morton (Z curve) inside each coarse cell;
* pure DG numbering: global dof = cell_rank * dofs_per_cell + comp*n^dim + lex;
* faces: face 2*d is the lower, 2*d+1 the upper face in direction d; with
  ``colorize = true`` the boundary id equals the face number.

Test infrastructure only.
"""
from __future__ import annotations

import numpy as np


def morton_encode(coords: np.ndarray, nbits: int) -> np.ndarray:
    """coords[..., dim] (each < 2**nbits) -> Morton index, x in the lowest bit of each group."""
    dim = coords.shape[-1]
    out = np.zeros(coords.shape[:-1], dtype=np.int64)
    for b in range(nbits):
        for d in range(dim):
            out |= ((coords[..., d].astype(np.int64) >> b) & 1) << (dim * b + d)
    return out


def morton_decode(idx: np.ndarray, nbits: int, dim: int) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.int64)
    out = np.zeros(idx.shape + (dim,), dtype=np.int64)
    for b in range(nbits):
        for d in range(dim):
            out[..., d] |= ((idx >> (dim * b + d)) & 1) << b
    return out


class CartesianMesh:
    """One level of a globally refined Cartesian mesh.

    Parameters mirror the C-ABI ``dgx_mesh_create_cartesian``:
    dim, coarse subdivisions ``cells0[dim]``, ``level`` refinements of every
    coarse cell, box ``lo``/``hi``, ``periodic[dim]`` flags, ``boundary_ids[2*dim]``.
    """

    def __init__(self, dim, cells0, level, lo, hi, periodic, boundary_ids=None):
        self.dim = dim
        self.cells0 = np.asarray(cells0, dtype=np.int64)[:dim]
        self.level = level
        self.lo = np.asarray(lo, dtype=np.float64)[:dim]
        self.hi = np.asarray(hi, dtype=np.float64)[:dim]
        self.periodic = np.asarray(periodic, dtype=bool)[:dim]
        if boundary_ids is None:
            boundary_ids = list(range(2 * dim))
        self.boundary_ids = np.asarray(boundary_ids, dtype=np.int64)[: 2 * dim]
        self.ncells_dir = self.cells0 << level
        self.n_cells = int(np.prod(self.ncells_dir))
        self.h = (self.hi - self.lo) / self.ncells_dir
        self._build()

    # global integer coordinates <-> cell index ------------------------------------------------
    def cell_index(self, coords):
        coords = np.asarray(coords, dtype=np.int64)
        L = self.level
        coarse = coords >> L
        local = coords & ((1 << L) - 1)
        cidx = np.zeros(coords.shape[:-1], dtype=np.int64)
        stride = 1
        for d in range(self.dim):
            cidx += coarse[..., d] * stride
            stride *= int(self.cells0[d])
        return cidx * (1 << (self.dim * L)) + morton_encode(local, L)

    def cell_coords(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        L = self.level
        per = 1 << (self.dim * L)
        cidx = idx // per
        local = morton_decode(idx % per, L, self.dim)
        out = np.zeros(idx.shape + (self.dim,), dtype=np.int64)
        for d in range(self.dim):
            out[..., d] = ((cidx % int(self.cells0[d])) << L) + local[..., d]
            cidx = cidx // int(self.cells0[d])
        return out

    def _build(self):
        idx = np.arange(self.n_cells, dtype=np.int64)
        self.coords = self.cell_coords(idx)
        nbr = np.empty((self.n_cells, 2 * self.dim), dtype=np.int64)
        for d in range(self.dim):
            for s in (0, 1):
                c = self.coords.copy()
                c[:, d] += 2 * s - 1
                outside = (c[:, d] < 0) | (c[:, d] >= self.ncells_dir[d])
                if self.periodic[d]:
                    c[:, d] %= self.ncells_dir[d]
                    nbr[:, 2 * d + s] = self.cell_index(c)
                else:
                    c[:, d] = np.clip(c[:, d], 0, self.ncells_dir[d] - 1)
                    n = self.cell_index(c)
                    n[outside] = -1 - self.boundary_ids[2 * d + s]
                    nbr[:, 2 * d + s] = n
        self.nbr = nbr

    # geometry ---------------------------------------------------------------------------------
    def cell_origin(self):
        return self.lo + self.coords * self.h

    def dof_points(self, nodes1d):
        """Support points of all DoFs of a scalar FE_DGQ field: [n_cells, n,..,n, dim] (x fastest)."""
        n = len(nodes1d)
        nodes1d = np.asarray(nodes1d, dtype=np.float64)
        grids = np.meshgrid(*([nodes1d] * self.dim), indexing="ij")   # axis order: slowest..fastest
        # axes (k, j, i): last axis is x
        ref = np.stack([grids[self.dim - 1 - d] for d in range(self.dim)], axis=-1)
        org = self.cell_origin().reshape((self.n_cells,) + (1,) * self.dim + (self.dim,))
        return org + ref * self.h

    def coarser(self):
        assert self.level > 0
        return CartesianMesh(self.dim, self.cells0, self.level - 1, self.lo, self.hi,
                             self.periodic, self.boundary_ids)

    # partition --------------------------------------------------------------------------------
    def partition(self, nranks):
        """Contiguous equal chunks of the cell order (p4est-style SFC partition)."""
        bounds = [(self.n_cells * r) // nranks for r in range(nranks + 1)]
        return np.asarray(bounds, dtype=np.int64)
