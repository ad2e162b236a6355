"""Host-side mirror of the slice of Elfel's FESpace / FEField API that *defines the inputs* of
the assembly path (dof numbering, EBC flags, element-dof order).  It stays host code in the
reference too (src/FESpaces.jl, src/FEFields.jl).  The H1 T3/T6/Q4 elements carry vertex (dim-0)
dofs only (src/FElements.jl:237,262,304: ndofperfeat=[1,0,0,0]); FEH1_T3_BUBBLE adds one dof on the
cell (:339, [1,0,1,0]) and FEL2_T3 / FEL2_Q4 have only that one (:410,:439, [0,0,1,0]) -- SURVEY 8f row f5.

Names follow the reference (``!`` dropped): FESpace, setebc, numberfreedofs, numberdatadofs,
numberdofs, ndofs, nunknowns, edofcompnt, edofbfnum, ndofsperel.
"""
from __future__ import annotations

import numpy as np

from .meshes import Mesh, T3, Q4, T6, T4


class FE:
    """Finite element type tag; ``FEH1_T3()`` etc. (src/FElements.jl:225-320, 324-355, 394-448).

    kind: element kind of the mesh it lives on; ndofperfeat: dofs per vertex / edge / cell (FEData);
    nbf: number of scalar basis functions; fe_id: the C ABI's EFG_FE_* code."""

    def __init__(self, kind: int, name: str, ndofperfeat=(1, 0, 0), fe_id: int = 0):
        self.kind, self.name, self.ndofperfeat, self.fe_id = kind, name, tuple(ndofperfeat), fe_id
        self.nen = 4 if kind == T4 else kind            # nodes per element
        self.nbf = self.nen * self.ndofperfeat[0] + self.ndofperfeat[2]

    def __repr__(self):
        return self.name


def FEH1_T3():
    return FE(T3, "FEH1_T3")


def FEH1_T6():
    return FE(T6, "FEH1_T6")


def FEH1_Q4():
    return FE(Q4, "FEH1_Q4")


def FEH1_T4():
    """src/FElements.jl:359-386 (3-D; examples/heat/poisson/t4.jl)"""
    return FE(T4, "FEH1_T4")


def FEH1_T3_BUBBLE():
    return FE(T3, "FEH1_T3_BUBBLE", (1, 0, 1), 7)


def FEL2_T3():
    return FE(T3, "FEL2_T3", (0, 0, 1), 1)


def FEL2_Q4():
    return FE(Q4, "FEL2_Q4", (0, 0, 1), 1)


def FEL2_T4():
    """src/FElements.jl:454-477: one constant basis function on the cell of a tetrahedron (ndofperfeat [0,0,0,1]; here the
    third entry of ndofperfeat is the cell whatever its dimension).  The reference has no example that assembles with it
    (test/test_felements.jl:117-135 checks its tables only); the space machinery (cell field, numbering) is the generic one."""
    return FE(T4, "FEL2_T4", (0, 0, 1), 1)


def bfun(fe: FE, pc):
    """Scalar basis functions at parametric point pc (src/FElements.jl:239-246, 264-288, 306-320, 341-347, 412-414, 470-472)."""
    if fe.ndofperfeat[0] == 0:
        return np.array([1.0])
    if fe.kind == T4:
        return np.array([1 - pc[0] - pc[1] - pc[2], pc[0], pc[1], pc[2]], dtype=np.float64)
    r, s = float(pc[0]), float(pc[1])
    if fe.ndofperfeat[0] == 0:
        return np.array([1.0])
    if fe.kind == T3:
        N = [(1 - r - s), r, s]
        if fe.ndofperfeat[2]:
            N.append((1 - r - s) * r * s)
        return np.array(N)
    if fe.kind == T6:
        t = 1. - r - s
        return np.array([t * (t + t - 1), r * (r + r - 1), s * (s + s - 1), 4 * r * t, 4 * r * s, 4 * s * t])
    return np.array([0.25 * (1. - r) * (1. - s), 0.25 * (1. + r) * (1. - s), 0.25 * (1. + r) * (1. + s), 0.25 * (1. - r) * (1. + s)])


class FEField:
    """FEField{N,T,IT} (src/FEFields.jl:14-33): per-term dof numbers, datum flags, values."""

    def __init__(self, ncomp: int, nterms: int):
        self.dofnums = np.zeros((nterms, ncomp), dtype=np.int64)   # Vector{SVector{N,IT}} layout
        self.isdatum = np.zeros((nterms, ncomp), dtype=bool)
        self.dofvals = np.zeros((nterms, ncomp), dtype=np.float64)

    @property
    def nterms(self):
        return self.dofnums.shape[0]

    @property
    def ndofsperterm(self):
        return self.dofnums.shape[1]

    def setebc(self, tid, comp, val):  # src/FEFields.jl:124-128 (1-based tid, comp)
        self.isdatum[tid - 1, comp - 1] = True
        self.dofvals[tid - 1, comp - 1] = val

    def numberfreedofs(self, firstnum=1):  # src/FEFields.jl:137-155
        free = ~self.isdatum.ravel()
        nums = np.zeros(free.size, dtype=np.int64)
        nums[free] = firstnum + np.arange(int(free.sum()), dtype=np.int64)
        self.dofnums = nums.reshape(self.dofnums.shape)

    def numberdatadofs(self, firstnum=1):  # src/FEFields.jl:164-177
        dat = self.isdatum.ravel()
        nums = self.dofnums.ravel().copy()
        nums[dat] = firstnum + np.arange(int(dat.sum()), dtype=np.int64)
        self.dofnums = nums.reshape(self.dofnums.shape)

    def freedofnums(self):  # src/FEFields.jl:187-203
        v = self.dofnums[~self.isdatum]
        if v.size == 0:
            return (np.iinfo(np.int64).max, 0, 0)
        return (int(v.min()), int(v.max()), int(v.size))

    def datadofnums(self):  # src/FEFields.jl:217-233
        v = self.dofnums[self.isdatum]
        if v.size == 0:
            return (np.iinfo(np.int64).max, 0, 0)
        return (int(v.min()), int(v.max()), int(v.size))


class FESpace:
    """FESpace{FET,T} (src/FESpaces.jl:25-40): one FEField per entity dimension that carries dofs (_makefields,
    src/FESpaces.jl:75-85).  ``field`` = the vertex field (_irsfields[0]) or None, ``cellfield`` = the cell field
    (_irsfields[2]) or None."""

    def __init__(self, mesh: Mesh, fe: FE, nfecopies: int = 1):
        assert mesh.kind == fe.kind, "finite element type does not match the mesh"
        self.mesh, self.fe, self.nfecopies = mesh, fe, nfecopies
        self.field = FEField(nfecopies, mesh.nnodes) if fe.ndofperfeat[0] else None   # _irsfields[0][2]
        self.cellfield = FEField(nfecopies, mesh.nel) if fe.ndofperfeat[2] else None  # _irsfields[2][2]
        # _number_edofs (src/FESpaces.jl:87-105): entity dimension, entity, copy-minor
        nbf = fe.nbf
        self._edofbfnum = np.repeat(np.arange(1, nbf + 1), nfecopies)
        self._edofcompnt = np.tile(np.arange(1, nfecopies + 1), nbf)
        nv = fe.nen * fe.ndofperfeat[0] * nfecopies
        self._edofmdim = np.concatenate([np.zeros(nv, dtype=np.int64), np.full(nbf * nfecopies - nv, 2, dtype=np.int64)])

    def fields(self):
        """The fields in ascending entity dimension.  The reference iterates ``keys(fesp._irsfields)`` of a Dict{Any,Any}
        with keys 0 and 2 (src/FESpaces.jl:142,164); Julia's Dict order for these two keys cannot be checked here (no
        Julia), so the numbering ORDER between the vertex and the cell field of one bubble space is 'unpinned'.  Ascending
        is the only order in which numberdatadofs! works for the examples (a cell field without data dofs reports
        lnum = 0 and would restart the vertex field's data numbers at 1).  The order only permutes the global numbers:
        the engine takes the dof numbers as data."""
        return [f for f in (self.field, self.cellfield) if f is not None]


def edofbfnum(fesp):
    return fesp._edofbfnum


def edofcompnt(fesp):
    return fesp._edofcompnt


def edofmdim(fesp):
    return fesp._edofmdim


def ndofsperel(fesp):  # src/FESpaces.jl:115
    return fesp.fe.nbf * fesp.nfecopies


def setebc(fesp, m, eid, comp, val):  # src/FESpaces.jl:287-291
    f = fesp.field if m == 0 else (fesp.cellfield if m == 2 else None)
    assert f is not None, f"{fesp.fe} has no dofs on entities of dimension {m}"
    f.setebc(int(eid), comp, val)
    return fesp


def numberfreedofs(fesp, firstnum=1):  # src/FESpaces.jl:141-151
    for f in fesp.fields():
        f.numberfreedofs(firstnum)
        firstnum = f.freedofnums()[1] + 1       # literally :147-148 (lnum = 0 for a field without free dofs)
    return fesp


def nunknowns(fesp):  # src/FESpaces.jl:192-201
    return sum(f.freedofnums()[2] for f in fesp.fields())


def ndofs(fesp):  # src/FESpaces.jl:180-185
    return sum(f.dofnums.size for f in fesp.fields())


def highestfreedofnum(fesp):
    return max([f.freedofnums()[1] for f in fesp.fields()] + [0])


def highestdatadofnum(fesp):
    return max([f.datadofnums()[1] for f in fesp.fields()] + [0])


def numberdatadofs(fesp, firstnum=0):  # src/FESpaces.jl:162-173
    firstnum = nunknowns(fesp) + 1 if firstnum == 0 else firstnum
    for f in fesp.fields():
        f.numberdatadofs(firstnum)
        firstnum = f.datadofnums()[1] + 1       # literally :169-170
    return fesp


def numberdofs(fesps):  # src/FESpaces.jl:250-273
    if isinstance(fesps, FESpace):
        fesps = [fesps]
    numberfreedofs(fesps[0], 1)
    for i in range(1, len(fesps)):
        numberfreedofs(fesps[i], highestfreedofnum(fesps[i - 1]) + 1)
    numberdatadofs(fesps[0], highestfreedofnum(fesps[-1]) + 1)
    for i in range(1, len(fesps)):
        numberdatadofs(fesps[i], highestdatadofnum(fesps[i - 1]) + 1)


def gathersysvec(fesps, n=None):
    """gathersysvec!(U, fesps) (src/FESpaces.jl:316-333): dof values in global dof order."""
    if isinstance(fesps, FESpace):
        fesps = [fesps]
    n = sum(ndofs(f) for f in fesps) if n is None else n
    U = np.zeros(n)
    for sp in fesps:
        for f in sp.fields():
            U[f.dofnums.ravel() - 1] = f.dofvals.ravel()
    return U


def scattersysvec(fesps, U):
    """scattersysvec!(fesps, U) (src/FESpaces.jl:340-365)."""
    if isinstance(fesps, FESpace):
        fesps = [fesps]
    for sp in fesps:
        for f in sp.fields():
            f.dofvals = U[f.dofnums.ravel() - 1].reshape(f.dofvals.shape)
