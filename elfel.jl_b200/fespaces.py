"""Host-side mirror of the slice of Elfel's FESpace / FEField API that *defines the inputs* of
the assembly path (dof numbering, EBC flags, element-dof order).  It stays host code in the
reference too (src/FESpaces.jl, src/FEFields.jl); only nodal (dim-0) dofs exist for the
H1 T3/T6/Q4 elements the path covers (src/FElements.jl:237,262,304: ndofperfeat=[1,0,0,0]).

Names follow the reference (``!`` dropped): FESpace, setebc, numberfreedofs, numberdatadofs,
numberdofs, ndofs, nunknowns, edofcompnt, edofbfnum, ndofsperel.
"""
from __future__ import annotations

import numpy as np

from .meshes import Mesh, T3, Q4, T6


class FE:
    """Finite element type tag; ``FEH1_T3()`` etc. (src/FElements.jl:225-320)."""

    def __init__(self, kind: int, name: str):
        self.kind, self.name = kind, name

    def __repr__(self):
        return self.name


def FEH1_T3():
    return FE(T3, "FEH1_T3")


def FEH1_T6():
    return FE(T6, "FEH1_T6")


def FEH1_Q4():
    return FE(Q4, "FEH1_Q4")


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
    """FESpace{FET,T} (src/FESpaces.jl:25-40) restricted to nodal H1 elements."""

    def __init__(self, mesh: Mesh, fe: FE, nfecopies: int = 1):
        assert mesh.kind == fe.kind, "finite element type does not match the mesh"
        self.mesh, self.fe, self.nfecopies = mesh, fe, nfecopies
        self.field = FEField(nfecopies, mesh.nnodes)  # _irsfields[0][2]
        # _number_edofs (src/FESpaces.jl:87-105): node-major, copy-minor
        nbf = fe.kind
        self._edofbfnum = np.repeat(np.arange(1, nbf + 1), nfecopies)
        self._edofcompnt = np.tile(np.arange(1, nfecopies + 1), nbf)


def edofbfnum(fesp):
    return fesp._edofbfnum


def edofcompnt(fesp):
    return fesp._edofcompnt


def ndofsperel(fesp):  # src/FESpaces.jl:115
    return fesp.fe.kind * fesp.nfecopies


def setebc(fesp, m, eid, comp, val):  # src/FESpaces.jl:287-291
    assert m == 0, "only vertex dofs exist for H1 T3/T6/Q4"
    fesp.field.setebc(int(eid), comp, val)
    return fesp


def numberfreedofs(fesp, firstnum=1):  # src/FESpaces.jl:141-151
    fesp.field.numberfreedofs(firstnum)
    return fesp


def nunknowns(fesp):  # src/FESpaces.jl:192-201
    return fesp.field.freedofnums()[2]


def ndofs(fesp):  # src/FESpaces.jl:180-185
    return fesp.field.dofnums.size


def highestfreedofnum(fesp):
    return fesp.field.freedofnums()[1]


def highestdatadofnum(fesp):
    return fesp.field.datadofnums()[1]


def numberdatadofs(fesp, firstnum=0):  # src/FESpaces.jl:162-173
    firstnum = nunknowns(fesp) + 1 if firstnum == 0 else firstnum
    fesp.field.numberdatadofs(firstnum)
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
    for f in fesps:
        U[f.field.dofnums.ravel() - 1] = f.field.dofvals.ravel()
    return U


def scattersysvec(fesps, U):
    """scattersysvec!(fesps, U) (src/FESpaces.jl:340-365)."""
    if isinstance(fesps, FESpace):
        fesps = [fesps]
    for f in fesps:
        f.field.dofvals = U[f.field.dofnums.ravel() - 1].reshape(f.field.dofvals.shape)
