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


@dataclass
class Band:
    """One rank's share of a structured block mesh generated WITHOUT the global arrays: local connectivity, the
    coordinates of its nodes, their GLOBAL dof numbers and the owned column ranges (1-based, inclusive)."""
    kind: int
    conn: object
    xy: object
    dofnums: object
    firsts: np.ndarray
    lasts: np.ndarray
    ndofs: int
    nel_global: int
    nnz_global: int
    rows: tuple      # owned node rows [j0, j1)


def band_rows(N, rank, world):
    """Node rows [j0, j1) owned by `rank`: cell rows are dealt out in `world` near-equal contiguous blocks, a node row
    goes with the cell row above it, the last rank also takes the top node row."""
    base, rem = divmod(N, world)
    j0 = rank * base + min(rank, rem)
    j1 = j0 + base + (1 if rank < rem else 0)
    return j0, (N + 1 if rank == world - 1 else j1)


def heat_block_numbering(N, j, i):
    """Closed form of setebc!(all boundary nodes) + numberfreedofs! + numberdatadofs! (src/FEFields.jl:137-177) on the
    (N+1) x (N+1) vertex grid of T3block/Q4block (nodes x-fastest): free dofs 1..(N-1)^2 in node order, then the data
    dofs in node order.  j, i: integer tensors (node row, node column)."""
    nfree = (N - 1) * (N - 1)
    free = (j - 1) * (N - 1) + i
    d_bottom = nfree + 1 + i
    d_side = nfree + 1 + (N + 1) + 2 * (j - 1) + (i == N).to(j.dtype)
    d_top = nfree + 1 + (N + 1) + 2 * (N - 1) + i
    interior = (i > 0) & (i < N) & (j > 0) & (j < N)
    return torch.where(interior, free, torch.where(j == 0, d_bottom, torch.where(j == N, d_top, d_side)))


def heat_block_owned_ranges(N, j0, j1):
    """Column ranges (1-based inclusive) of the dofs of node rows [j0, j1) under heat_block_numbering."""
    nfree = (N - 1) * (N - 1)
    lo, hi = max(j0, 1), min(j1 - 1, N - 1)          # interior node rows owned
    r = []
    if lo <= hi:
        r.append(((lo - 1) * (N - 1) + 1, hi * (N - 1)))
    if j0 == 0:
        r.append((nfree + 1, nfree + N + 1))
    if lo <= hi:
        r.append((nfree + N + 2 + 2 * (lo - 1), nfree + N + 1 + 2 * hi))
    if j1 == N + 1:
        r.append((nfree + N + 2 + 2 * (N - 1), nfree + 2 * N + 2 + 2 * (N - 1)))
    r = [x for x in r if x[1] >= x[0]]
    merged = []
    for f, l in sorted(r):
        if merged and f == merged[-1][1] + 1:
            merged[-1] = (merged[-1][0], l)
        else:
            merged.append((f, l))
    return (np.array([m[0] for m in merged], dtype=np.int64), np.array([m[1] for m in merged], dtype=np.int64))


def block_band(kind, N, rank, world, dev="cpu"):
    """Rank's share of the unit-square N x N heat problem on T3block / Q4block (BASELINE configs 1 and 5), strong
    scaling: the cells of the node rows it owns plus one halo cell row per cut.  Element order, node order and
    coordinates are those of the global generators restricted to the band, so per-nonzero sums see their contributions in
    the same order as the unsharded assembly (bit-identical blocks)."""
    if kind not in (Q4, T3):
        raise ValueError("block_band: T3 / Q4 block meshes only")
    j0, j1 = band_rows(N, rank, world)
    ja, jb = max(j0 - 1, 0), min(j1, N)               # cell rows [ja, jb); node rows [ja, jb]
    nr = jb - ja
    i = torch.arange(N, dtype=torch.int64, device=dev).repeat_interleave(nr)
    j = torch.arange(nr, dtype=torch.int64, device=dev).repeat(N)
    f = j * (N + 1) + i                                # local 0-based id of the cell's lower-left node
    if kind == Q4:
        conn = torch.stack([f, f + 1, f + N + 2, f + N + 1], dim=1) + 1
    else:
        conn = torch.empty((2 * N * nr, 3), dtype=torch.int64, device=dev)
        conn[0::2] = torch.stack([f, f + 1, f + N + 2], dim=1) + 1
        conn[1::2] = torch.stack([f, f + N + 2, f + N + 1], dim=1) + 1
    del i, j, f
    xs = torch.arange(N + 1, dtype=torch.float64, device=dev) * 1.0 / N
    ys = torch.arange(ja, jb + 1, dtype=torch.float64, device=dev) * 1.0 / N
    xy = torch.stack([xs.repeat(nr + 1), ys.repeat_interleave(N + 1)], dim=1)
    ni = torch.arange(N + 1, dtype=torch.int64, device=dev).repeat(nr + 1)
    nj = torch.arange(ja, jb + 1, dtype=torch.int64, device=dev).repeat_interleave(N + 1)
    dofnums = heat_block_numbering(N, nj, ni).view(-1, 1).contiguous()
    firsts, lasts = heat_block_owned_ranges(N, j0, j1)
    nelg = N * N * (1 if kind == Q4 else 2)
    nnzg = (3 * N + 1) ** 2 if kind == Q4 else 7 * N * N + 6 * N + 1
    return Band(kind, conn.contiguous(), xy.contiguous(), dofnums, firsts, lasts, (N + 1) * (N + 1), nelg, nnzg, (j0, j1))


def q4_band(N, rank, world, dev="cpu"):
    return block_band(Q4, N, rank, world, dev)


def pattern_checksum(colptr, rowval, firsts, lasts):
    """Order-independent checksum of a CSC column block: sum over stored entries of mix(row, col) mod 2^63 (torch int64
    wrap-around arithmetic).  colptr (ncl+1) / rowval 1-based; firsts/lasts: the global column ranges of the block."""
    dev = rowval.device
    cols = torch.cat([torch.arange(int(f), int(l) + 1, dtype=torch.int64, device=dev) for f, l in zip(firsts, lasts)])
    cnt = (colptr[1:] - colptr[:-1])
    col_of = torch.repeat_interleave(cols, cnt)
    return int(_mix(rowval, col_of).sum().item())


def _mix(r, c):
    return (r * 0x9E3779B1 + c * 0x85EBCA77) ^ ((r + 0x27D4EB2F) * (c + 0x165667B1))


def q4_expected_checksums(N, j0, j1, dev="cpu", chunk_rows=256):
    """What the heat Q4 assembly on the N x N block must produce for the columns of node rows [j0, j1), derived from the
    grid alone (9-point node adjacency, uniform square cells: element matrix 1/6 [4 -1 -2 -1; ...]): (nnz, pattern
    checksum, sum of nzval^2).  Independent of the library: used by bench.py to validate the sharded result."""
    nnz, chk = 0, 0
    sq = 0.0
    for a in range(j0, j1, chunk_rows):
        b = min(a + chunk_rows, j1)
        cj = torch.arange(a, b, dtype=torch.int64, device=dev).repeat_interleave(N + 1)
        ci = torch.arange(N + 1, dtype=torch.int64, device=dev).repeat(b - a)
        cd = heat_block_numbering(N, cj, ci)
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                rj, ri = cj + dj, ci + di
                ok = (rj >= 0) & (rj <= N) & (ri >= 0) & (ri <= N)
                rd = heat_block_numbering(N, rj[ok], ri[ok])
                nnz += int(ok.sum().item())
                chk = (chk + int(_mix(rd, cd[ok]).sum().item())) % (1 << 64)
                # number of cells shared by the two nodes -> value: diagonal 4/6 per cell, edge neighbour -1/6 per cell, corner neighbour -2/6
                cjo, cio, rjo, rio = cj[ok], ci[ok], rj[ok], ri[ok]
                nx = (torch.minimum(cio, rio) < N).to(torch.float64) * (di != 0) + (di == 0) * ((cio > 0).to(torch.float64) + (cio < N).to(torch.float64))
                ny = (torch.minimum(cjo, rjo) < N).to(torch.float64) * (dj != 0) + (dj == 0) * ((cjo > 0).to(torch.float64) + (cjo < N).to(torch.float64))
                ncell = nx * ny
                per = 4.0 / 6 if (di == 0 and dj == 0) else (-2.0 / 6 if (di != 0 and dj != 0) else -1.0 / 6)
                sq += float(((ncell * per) ** 2).sum().item())
    return nnz, chk, sq


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
