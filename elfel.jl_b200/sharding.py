"""Owner-computes sharding of the assembly over the GPUs of one node.

The matrix columns (= dof numbers) are split between the ranks; rank r assembles every element that
touches one of its columns (halo elements are replicated) and keeps only its columns, so the numeric
path needs no communication.  A rank owns the dofs of the nodes of one horizontal band of the mesh.
With the reference's numbering (nodes x-fastest; T6 mid-side nodes numbered after the vertices; data
dofs after the free dofs) a band is a *set of column ranges*, handed to efg_set_column_ranges(); the
global matrix is the column-wise interleave of the ranks' blocks -- still no reduction.

Everything here is host/driver logic written with torch tensor ops so that it runs on the GPU for
benchmark-size meshes (a 4000 x 32000 T6 mesh is built in well under a second) and on the CPU in tests.
The mesh generators restate meshes.T6block_fast / T3block / Q4block (tests check equality).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .meshes import Q4, T3, T6


def _grid_xy(Length, Width, nL, nW, dev):
    xs = torch.arange(nL + 1, dtype=torch.float64, device=dev) * float(Length) / nL
    ys = torch.arange(nW + 1, dtype=torch.float64, device=dev) * float(Width) / nW
    return torch.stack([xs.repeat(nW + 1), ys.repeat_interleave(nL + 1)], dim=1)


def _cells(nL, nW, dev):
    i = torch.arange(nL, dtype=torch.int64, device=dev).repeat_interleave(nW)
    j = torch.arange(nW, dtype=torch.int64, device=dev).repeat(nL)
    return i, j


def t3block_torch(Length, Width, nL, nW, dev="cpu"):
    i, j = _cells(nL, nW, dev)
    f = j * (nL + 1) + i
    conn = torch.empty((2 * nL * nW, 3), dtype=torch.int64, device=dev)
    conn[0::2] = torch.stack([f, f + 1, f + nL + 2], dim=1)
    conn[1::2] = torch.stack([f, f + nL + 2, f + nL + 1], dim=1)
    return T3, conn + 1, _grid_xy(Length, Width, nL, nW, dev)


def q4block_torch(Length, Width, nL, nW, dev="cpu"):
    i, j = _cells(nL, nW, dev)
    f = j * (nL + 1) + i
    conn = torch.stack([f, f + 1, f + nL + 2, f + nL + 1], dim=1)
    return Q4, conn + 1, _grid_xy(Length, Width, nL, nW, dev)


def t6block_torch(Length, Width, nL, nW, dev="cpu"):
    """Same numbering as meshes.T6block_fast (mid-side nodes in first-encounter order)."""
    nv = (nL + 1) * (nW + 1)
    i, j = _cells(nL, nW, dev)
    col0, coln = 4 * nW + 1, 3 * nW + 1
    base = torch.where(i == 0, torch.zeros_like(i), col0 + (i - 1) * coln)
    per = torch.where(i == 0, torch.full_like(i, 4), torch.full_like(i, 3))
    start = base + torch.where(j == 0, torch.zeros_like(j), 1 + per * j)
    k = start + (j == 0).to(torch.int64)
    bottom = torch.where(j == 0, start, torch.full_like(start, -1))
    right, diag, top, left_own = k, k + 1, k + 2, k + 3
    top2, right2 = top.view(nL, nW), right.view(nL, nW)
    bottom2 = bottom.view(nL, nW).clone()
    bottom2[:, 1:] = top2[:, :-1]
    left2 = left_own.view(nL, nW).clone()
    left2[1:, :] = right2[:-1, :]
    bottom, left = bottom2.reshape(-1) + nv, left2.reshape(-1) + nv
    right, diag, top = right + nv, diag + nv, top + nv
    f = j * (nL + 1) + i
    conn = torch.empty((2 * nL * nW, 6), dtype=torch.int64, device=dev)
    conn[0::2] = torch.stack([f, f + 1, f + nL + 2, bottom, right, diag], dim=1)
    conn[1::2] = torch.stack([f, f + nL + 2, f + nL + 1, diag, top, left], dim=1)
    nedges = col0 + (nL - 1) * coln
    vxy = _grid_xy(Length, Width, nL, nW, dev)
    mxy = torch.empty((nedges, 2), dtype=torch.float64, device=dev)
    for (a, b, m) in ((0, 1, 3), (1, 2, 4), (2, 0, 5)):
        mxy[conn[:, m] - nv] = 0.5 * (vxy[conn[:, a]] + vxy[conn[:, b]])
    return T6, conn + 1, torch.cat([vxy, mxy], dim=0)


def boundary_mask(xy, Length, Width):
    tol = 1e-9 * max(Length, Width)
    return ((xy[:, 0] < tol) | (xy[:, 0] > Length - tol) | (xy[:, 1] < tol) | (xy[:, 1] > Width - tol))


def number_dofs(isdatum):
    """numberfreedofs! then numberdatadofs! (src/FEFields.jl:137-177): free dofs first in term order."""
    flat = isdatum.reshape(-1)
    free = ~flat
    nfree = int(free.sum().item())
    nums = torch.empty(flat.numel(), dtype=torch.int64, device=flat.device)
    nums[free] = torch.arange(1, nfree + 1, dtype=torch.int64, device=flat.device)
    nums[flat] = torch.arange(nfree + 1, flat.numel() + 1, dtype=torch.int64, device=flat.device)
    return nums.view(isdatum.shape), nfree


def ranges_of(mask):
    """Maximal runs of True in a 1-D bool tensor -> (firsts, lasts), 1-based inclusive (numpy int64)."""
    m = mask.to(torch.int8)
    d = torch.diff(m, prepend=m.new_zeros(1), append=m.new_zeros(1))
    firsts = torch.nonzero(d == 1).reshape(-1) + 1
    lasts = torch.nonzero(d == -1).reshape(-1)
    return firsts.cpu().numpy().astype(np.int64), lasts.cpu().numpy().astype(np.int64)


@dataclass
class _Mesh:
    kind: int
    conn: object
    xy: object

    @property
    def nel(self):
        return int(self.conn.shape[0])

    @property
    def nnodes(self):
        return int(self.xy.shape[0])


class _Field:
    def __init__(self, dofnums):
        self.dofnums = dofnums


class _Space:
    def __init__(self, dofnums):
        self.field = _Field(dofnums)


@dataclass
class Shard:
    name: str
    form: object
    quad: int
    meshes: list
    spaces: list
    space_mesh: list
    ndofs: int                 # GLOBAL matrix dimension
    ranges: tuple              # (firsts, lasts) owned column ranges, 1-based
    nel_global: int

    @property
    def nel(self):
        return self.meshes[0].nel

    def dofs(self):
        return [s.field.dofnums for s in self.spaces]


def build_global(workload, n, world, dev):
    """Global synthetic problem of the weak-scaling run: unit-width, `world` units tall."""
    from .assemblers import ElasticityForm, HeatForm
    from .problems import plane_stress_D
    L, W, nL, nW = 1.0, float(world), n, n * world
    if workload == "heat_t6":
        kind, conn, xy = t6block_torch(L, W, nL, nW, dev); ncomp, form, quad = 1, HeatForm(1.0), 3
    elif workload == "heat_t3":
        kind, conn, xy = t3block_torch(L, W, nL, nW, dev); ncomp, form, quad = 1, HeatForm(1.0), 1
    elif workload == "heat_q4":
        kind, conn, xy = q4block_torch(L, W, nL, nW, dev); ncomp, form, quad = 1, HeatForm(1.0), 2
    elif workload == "elasticity_t6":
        kind, conn, xy = t6block_torch(L, W, nL, nW, dev); ncomp, form, quad = 2, ElasticityForm(plane_stress_D()), 3
    else:
        raise ValueError(f"workload {workload} has no sharded generator")
    if ncomp == 1:
        isdatum = boundary_mask(xy, L, W).view(-1, 1)
    else:   # examples/elasticity/stretch/t6.jl: edges x=0 and x=A, both components
        tol = 1e-9
        on = (xy[:, 0] < tol) | (xy[:, 0] > L - tol)
        isdatum = torch.stack([on, on], dim=1)
    dofnums, _ = number_dofs(isdatum)
    # band of every node: by cell row (half-step integer y index avoids floating point at the cuts)
    yi = torch.round(xy[:, 1] * (2.0 * nW / W)).to(torch.int64)
    band = torch.clamp((yi // 2) // n, max=world - 1)
    return kind, conn, xy, dofnums, band, form, quad


def shard_from_global(kind, conn, xy, dofnums, band, rank):
    """Rank's sub-mesh (elements touching an owned dof, nodes renumbered locally, GLOBAL dof numbers)."""
    ndofs = dofnums.numel()
    owner = torch.empty(ndofs, dtype=torch.int64, device=conn.device)
    owner[dofnums.reshape(-1) - 1] = band.repeat_interleave(dofnums.shape[1])
    firsts, lasts = ranges_of(owner == rank)
    touches = (band[conn - 1] == rank).any(dim=1)
    sconn = conn[touches]
    uniq, inv = torch.unique(sconn, return_inverse=True)
    lconn = (inv + 1).contiguous()
    return lconn, xy[uniq - 1].contiguous(), dofnums[uniq - 1].contiguous(), (firsts, lasts)


def shard_problem(efg, workload, n, rank, world, dev=None):
    """-> (Shard, (firsts, lasts), nel_global).  The global arrays are dropped before returning."""
    dev = dev if dev is not None else (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")
    kind, conn, xy, dofnums, band, form, quad = build_global(workload, n, world, dev)
    nel_global, ndofs = int(conn.shape[0]), int(dofnums.numel())
    lconn, lxy, ldof, ranges = shard_from_global(kind, conn, xy, dofnums, band, rank)
    del conn, xy, dofnums, band
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    sh = Shard(f"{workload}_N{n}_rank{rank}of{world}", form, quad, [_Mesh(kind, lconn, lxy)], [_Space(ldof)], [0],
               ndofs, ranges, nel_global)
    return sh, ranges, nel_global


def merge_blocks(ncol, blocks):
    """Interleave per-rank CSC blocks [(firsts, lasts, colptr, rowval, nzval), ...] into the global CSC."""
    counts = np.zeros(ncol, dtype=np.int64)
    for firsts, lasts, colptr, _, _ in blocks:
        cols = np.concatenate([np.arange(f, l + 1) for f, l in zip(firsts, lasts)]) if len(firsts) else np.zeros(0, np.int64)
        counts[cols - 1] = np.diff(colptr)
    gcolptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    nnz = int(gcolptr[-1] - 1)
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz, dtype=np.float64)
    for firsts, lasts, colptr, rv, nz in blocks:
        lc = 0
        for f, l in zip(firsts, lasts):
            k = l - f + 1
            a, b = colptr[lc] - 1, colptr[lc + k] - 1
            ga = gcolptr[f - 1] - 1
            rowval[ga: ga + (b - a)] = rv[a:b]
            nzval[ga: ga + (b - a)] = nz[a:b]
            lc += k
    return gcolptr, rowval, nzval


def merge_vectors(n, blocks):
    """Scatter per-rank system-vector blocks [(firsts, lasts, values of the owned rows, ascending), ...] into the global
    vector of length n (owner-computes: every row is owned by exactly one rank, so no reduction is needed)."""
    out = np.empty(n, dtype=np.float64)
    for firsts, lasts, vals in blocks:
        o = 0
        for f, l in zip(firsts, lasts):
            k = l - f + 1
            out[f - 1: l] = vals[o: o + k]
            o += k
    return out
