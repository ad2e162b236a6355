"""Synthetic structured meshes in the numbering the reference's examples use.

Elfel takes its meshes from MeshSteward 1.1.3 (not vendored under the reference tree;
call sites: examples/heat/poisson/t3.jl:34, q4.jl:24, examples/elasticity/stretch/t6.jl:34,
examples/stokes/colliding_flow/ht_p2_p1_gen.jl:35,42).  These are restatements of the
generators' *numbering conventions* as pinned by the reference's fixtures:

* nodes x-fastest, ``id = 1 + i + j*(nL+1)``                       (test/qmesh-xyz.dat)
* T3block ``:a``: ``for i in 1:nL, for j in 1:nW`` -> ``[f,f+1,f+nL+2], [f,f+nL+2,f+nL+1]``
  (test/qmesh-conn.dat); ``:b``: ``[f,f+1,f+nL+1], [f+1,f+nL+2,f+nL+1]`` (test/mt3gen3-conn.dat)
* T6block = T3block + mid-side nodes; corner nodes are numbered before mid-side nodes
  (test/test_stokes.jl:130 pins the counts).  Mid-side numbering order is NOT pinned by
  the reference (first-encounter order over elements and edges 1-2, 2-3, 3-1 is used);
  it does not affect parity because the engine and the oracle receive the same arrays.
* T6toT3 keeps the first three nodes of every T6 element; the pressure mesh has its own
  vertex collection made of the corner nodes (ids 1..(nL+1)(nW+1)).

All arrays are in the reference's memory layout: ``conn`` is (nel, nen) int64, 1-based
(== Julia's nen x nel column-major), ``xy`` is (nnodes, 2) float64.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

T3, Q4, T6 = 3, 4, 6
T4 = 40      # linear tetrahedron: 4 nodes, 3-D (the C ABI's EFG_T4; every other kind code equals its node count)


@dataclass
class Mesh:
    """Base incidence relation (element -> nodes) + the "geom" attribute of its vertices.

    Mirrors what FEIterator caches: ``_bir`` and ``_geom`` (src/FEIterators.jl:55-56)."""

    kind: int            # T3 / Q4 / T6 (= nodes per element) or T4
    conn: np.ndarray     # (nel, nen) int64, 1-based
    xy: np.ndarray       # (nnodes, 2) float64; (nnodes, 3) for T4

    @property
    def nel(self) -> int:
        return int(self.conn.shape[0])

    @property
    def nnodes(self) -> int:
        return int(self.xy.shape[0])


def _grid_xy(Length, Width, nL, nW):
    xs = np.arange(nL + 1, dtype=np.float64) * float(Length) / nL
    ys = np.arange(nW + 1, dtype=np.float64) * float(Width) / nW
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # row j, column i -> x fastest
    return np.stack([X.ravel(), Y.ravel()], axis=1)


def _cell_first_nodes(nL, nW):
    # element loop order: i outer (x), j inner (y); f = 0-based lower-left node
    i = np.repeat(np.arange(nL, dtype=np.int64), nW)
    j = np.tile(np.arange(nW, dtype=np.int64), nL)
    return j * (nL + 1) + i


def T3block(Length, Width, nL, nW, orientation="a") -> Mesh:
    f = _cell_first_nodes(nL, nW)
    if orientation == "a":
        t1 = np.stack([f, f + 1, f + nL + 2], axis=1)
        t2 = np.stack([f, f + nL + 2, f + nL + 1], axis=1)
    elif orientation == "b":
        t1 = np.stack([f, f + 1, f + nL + 1], axis=1)
        t2 = np.stack([f + 1, f + nL + 2, f + nL + 1], axis=1)
    else:
        raise ValueError("orientation must be 'a' or 'b'")
    conn = np.empty((2 * nL * nW, 3), dtype=np.int64)
    conn[0::2] = t1
    conn[1::2] = t2
    return Mesh(T3, conn + 1, _grid_xy(Length, Width, nL, nW))


def T4block(Length, Width, Height, nL, nW, nH) -> Mesh:
    """Tetrahedral block (examples/heat/poisson/t4.jl:23: T4block(A, A, A, N, N, N)).  Nodes x-fastest, then y, then z, like
    the 2-D blocks.  MeshSteward's own split of a hexahedral cell into tetrahedra is not vendored ("parity unpinned"; the
    engine takes the mesh as data): every cell is cut into the six tetrahedra of the Kuhn triangulation around its main
    diagonal, all positively oriented, cells in i-outer / j / k-inner order."""
    xs = np.arange(nL + 1, dtype=np.float64) * float(Length) / nL
    ys = np.arange(nW + 1, dtype=np.float64) * float(Width) / nW
    zs = np.arange(nH + 1, dtype=np.float64) * float(Height) / nH
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    i, j, k = np.meshgrid(np.arange(nL, dtype=np.int64), np.arange(nW, dtype=np.int64), np.arange(nH, dtype=np.int64), indexing="ij")
    f = (k.ravel() * (nW + 1) + j.ravel()) * (nL + 1) + i.ravel()
    sx, sy, sz = 1, nL + 1, (nL + 1) * (nW + 1)
    tets = []
    import itertools
    for perm in itertools.permutations((0, 1, 2)):        # walk from the cell's first node to the opposite one, one axis at a time
        steps = [(sx, sy, sz)[a] for a in perm]
        v = [f, f + steps[0], f + steps[0] + steps[1], f + steps[0] + steps[1] + steps[2]]
        odd = perm in ((0, 2, 1), (2, 1, 0), (1, 0, 2))   # odd permutations are negatively oriented: swap two nodes
        if odd:
            v[1], v[2] = v[2], v[1]
        tets.append(np.stack(v, axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4)
    return Mesh(T4, conn + 1, xyz)


def Q4block(Length, Width, nL, nW) -> Mesh:
    f = _cell_first_nodes(nL, nW)
    conn = np.stack([f, f + 1, f + nL + 2, f + nL + 1], axis=1)
    return Mesh(Q4, conn + 1, _grid_xy(Length, Width, nL, nW))


def T3toT6(mesh: Mesh) -> Mesh:
    """Insert mid-side nodes (generic; first-encounter numbering)."""
    assert mesh.kind == T3
    c = mesh.conn - 1
    nv = mesh.nnodes
    a = np.stack([c[:, 0], c[:, 1], c[:, 2]], axis=1).ravel()
    b = np.stack([c[:, 1], c[:, 2], c[:, 0]], axis=1).ravel()
    key = np.minimum(a, b) * np.int64(nv) + np.maximum(a, b)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq), dtype=np.int64)
    mid = (nv + rank[inv]).reshape(-1, 3)
    conn = np.concatenate([c, mid], axis=1) + 1
    ea = uniq // nv
    eb = uniq % nv
    mxy = np.empty((len(uniq), 2))
    mxy[rank] = 0.5 * (mesh.xy[ea] + mesh.xy[eb])
    return Mesh(T6, conn, np.concatenate([mesh.xy, mxy], axis=0))


def T6block(Length, Width, nL, nW, orientation="a") -> Mesh:
    return T3toT6(T3block(Length, Width, nL, nW, orientation))


def T6block_fast(Length, Width, nL, nW) -> Mesh:
    """Closed-form T6block (orientation :a), identical output to :func:`T6block` without the
    edge-dictionary pass, for benchmark-size meshes (tests/test_meshes.py checks equality).

    Walking cells i-outer/j-inner and the two triangles' edges (1-2, 2-3, 3-1), the new
    edges met in cell (i, j) are: bottom (only j == 0), right, diagonal, top, left (only
    i == 0), in that order."""
    nv = (nL + 1) * (nW + 1)
    i = np.repeat(np.arange(nL, dtype=np.int64), nW)
    j = np.tile(np.arange(nW, dtype=np.int64), nL)
    # number of new edges in the cells before (i, j) in walk order
    # column i=0 has 5 + 4*(nW-1) new edges, later columns 4 + 3*(nW-1)
    col0 = 4 * nW + 1
    coln = 3 * nW + 1
    base = np.where(i == 0, 0, col0 + (i - 1) * coln)
    per = np.where(i == 0, 4, 3)
    first_extra = 1  # the bottom edge exists only in j == 0 cells
    start = base + np.where(j == 0, 0, first_extra + per * j)
    has_bottom = (j == 0)
    has_left = (i == 0)
    k = start.copy()
    bottom = np.where(has_bottom, k, -1)
    k = k + has_bottom
    right = k
    diag = k + 1
    top = k + 2
    left_own = k + 3
    # ids of edges owned by neighbours
    # bottom of (i, j>0) = top of (i, j-1); left of (i>0, j) = right of (i-1, j)
    top2 = top.reshape(nL, nW)
    right2 = right.reshape(nL, nW)
    bottom2 = bottom.reshape(nL, nW).copy()
    bottom2[:, 1:] = top2[:, :-1]
    left2 = left_own.reshape(nL, nW).copy()
    left2[1:, :] = right2[:-1, :]
    bottom = bottom2.ravel() + nv
    left = left2.ravel() + nv
    right = right + nv
    diag = diag + nv
    top = top + nv
    f = j * (nL + 1) + i
    conn = np.empty((2 * nL * nW, 6), dtype=np.int64)
    conn[0::2] = np.stack([f, f + 1, f + nL + 2, bottom, right, diag], axis=1)
    conn[1::2] = np.stack([f, f + nL + 2, f + nL + 1, diag, top, left], axis=1)
    nedges = col0 + (nL - 1) * coln
    vxy = _grid_xy(Length, Width, nL, nW)
    mxy = np.empty((nedges, 2))
    c0 = conn[:, :3]
    for (a, b, m) in ((0, 1, 3), (1, 2, 4), (2, 0, 5)):
        mxy[conn[:, m] - nv] = 0.5 * (vxy[c0[:, a]] + vxy[c0[:, b]])
    return Mesh(T6, conn + 1, np.concatenate([vxy, mxy], axis=0))


def T6toT3(mesh: Mesh) -> Mesh:
    """Pressure mesh of a Taylor-Hood pair: corner nodes only, own vertex collection."""
    assert mesh.kind == T6
    conn = np.ascontiguousarray(mesh.conn[:, :3])
    nv = int(conn.max())
    return Mesh(T3, conn, np.ascontiguousarray(mesh.xy[:nv]))


def transform(mesh: Mesh, fn) -> Mesh:
    """``transform(ir, x -> ...)`` of MeshSteward: map the coordinates in place."""
    mesh.xy[:] = fn(mesh.xy)
    return mesh


def boundary_nodes(mesh: Mesh) -> np.ndarray:
    """``connectedv(boundary(mesh))``: 1-based ids of the nodes on boundary edges (sorted)."""
    c = mesh.conn - 1
    nc = 4 if mesh.kind == Q4 else 3
    a = c[:, :nc].ravel()
    b = np.roll(c[:, :nc], -1, axis=1).ravel()
    nn = np.int64(mesh.nnodes)
    key = np.minimum(a, b) * nn + np.maximum(a, b)
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    on = cnt[inv] == 1
    nodes = [a[on], b[on]]
    if mesh.kind == T6:
        nodes.append(c[:, 3:6].ravel()[on])
    return np.unique(np.concatenate(nodes)) + 1


def vselect(xy: np.ndarray, box=None, inflate=0.0, nearestto=None) -> np.ndarray:
    """``vselect(locs; box=..., inflate=...)`` / ``vselect(locs; nearestto=...)`` (1-based ids)."""
    if nearestto is not None:
        d = ((xy - np.asarray(nearestto, dtype=np.float64)) ** 2).sum(axis=1)
        return np.array([int(np.argmin(d)) + 1], dtype=np.int64)
    x0, x1, y0, y1 = box
    m = ((xy[:, 0] >= x0 - inflate) & (xy[:, 0] <= x1 + inflate)
         & (xy[:, 1] >= y0 - inflate) & (xy[:, 1] <= y1 + inflate))
    return np.nonzero(m)[0].astype(np.int64) + 1


def jitter(mesh: Mesh, seed=20260101, frac=0.2, midfrac=0.05) -> Mesh:
    """Parity-only variant: move interior corner nodes by U(-frac*h, frac*h) and
    T6 mid-side nodes additionally by U(-midfrac*h, midfrac*h) so the Jacobian differs per qp."""
    rng = np.random.default_rng(seed)
    xy = mesh.xy.copy()
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    ncorner = int(mesh.conn[:, : (4 if mesh.kind == Q4 else 3)].max())
    h = np.sqrt((hi - lo).prod() / max(ncorner, 1))
    tol = 1e-9 * (hi - lo).max()
    interior = np.all((xy > lo + tol) & (xy < hi - tol), axis=1)
    d = rng.uniform(-frac * h, frac * h, size=(ncorner, 2))
    xy[:ncorner] += d * interior[:ncorner, None]
    if mesh.kind == T6:
        c = mesh.conn - 1
        for (a, b, m) in ((0, 1, 3), (1, 2, 4), (2, 0, 5)):
            xy[c[:, m]] = 0.5 * (xy[c[:, a]] + xy[c[:, b]])
        dm = rng.uniform(-midfrac * h, midfrac * h, size=(mesh.nnodes - ncorner, 2))
        xy[ncorner:] += dm * interior[ncorner:, None]
    return Mesh(mesh.kind, mesh.conn.copy(), xy)
