"""The reference's example set-ups (mesh, spaces, EBC, numbering) on synthetic structured unit-square
meshes -- the inputs of the five BASELINE configs.  Host logic only.

heat:       examples/heat/poisson/t3.jl:93-104, q4.jl (all boundary nodes prescribed)
elasticity: examples/elasticity/stretch/t6.jl:78-99 (edges x=0 and x=A, both components)
stokes:     examples/stokes/colliding_flow/ht_p2_p1_gen.jl:152-177 (whole velocity boundary + one
            pressure node nearest the centre; numberdofs!([Uh, Ph]))
p1b_p1:     examples/stokes/colliding_flow/p1b_p1.jl:175-200 (FEH1_T3_BUBBLE velocities, FEH1_T3 pressure, one T3 mesh)
q1_q0:      examples/stokes/colliding_flow/q1_q0.jl:131-153 (FEH1_Q4 velocities, FEL2_Q4 pressure: the CELL nearest the
            centre is pinned, setebc!(pfesp, 2, ...))
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .assemblers import (ElasticityForm, HeatForm, StokesGenForm, StokesReddyForm, StokesVeclapAltForm,
                         StokesVeclapForm)
from .fespaces import (FEH1_Q4, FEH1_T3, FEH1_T4, FEH1_T6, FEH1_T3_BUBBLE, FEL2_Q4, FEL2_T3, FESpace, ndofs, numberdatadofs, numberdofs,
                       numberfreedofs, setebc)
from .meshes import Q4, T3, T4, T6, Mesh, Q4block, T3block, T4block, T6block, T6block_fast, T6toT3, jitter, transform


@dataclass
class Problem:
    name: str
    form: object                 # a *Form object (form_id, params())
    quad: int                    # triangle npts / square Gauss order
    meshes: list                 # [mesh0] or [vmesh, pmesh]
    spaces: list                 # FESpace list in space-slot order
    space_mesh: list             # mesh slot of each space
    ndofs: int = 0

    @property
    def nel(self):
        return self.meshes[0].nel

    def dofs(self):
        """vertex-field dof numbers per space (None for an L2 space)"""
        return [None if s.field is None else s.field.dofnums for s in self.spaces]

    def cell_dofs(self):
        """cell-field dof numbers per space (None where the element has no cell dof)"""
        return [None if s.cellfield is None else s.cellfield.dofnums for s in self.spaces]


def _structured_boundary(nL, nW, kind, mesh):
    """Boundary node ids (1-based) of the structured block without the edge-dictionary pass."""
    nv = (nL + 1) * (nW + 1)
    i = np.arange(nv) % (nL + 1)
    j = np.arange(nv) // (nL + 1)
    on = (i == 0) | (i == nL) | (j == 0) | (j == nW)
    ids = [np.nonzero(on)[0]]
    if kind == T6:
        xy = mesh.xy[nv:]
        lo, hi = mesh.xy[:nv].min(axis=0), mesh.xy[:nv].max(axis=0)
        tol = 1e-9 * (hi - lo).max()
        onm = (np.abs(xy[:, 0] - lo[0]) < tol) | (np.abs(xy[:, 0] - hi[0]) < tol) | \
              (np.abs(xy[:, 1] - lo[1]) < tol) | (np.abs(xy[:, 1] - hi[1]) < tol)
        ids.append(nv + np.nonzero(onm)[0])
    return np.concatenate(ids) + 1


def _setebc_nodes(fesp, ids, comps=(1,)):
    f = fesp.field
    for c in comps:
        f.isdatum[ids - 1, c - 1] = True


def heat_problem(kind: int, N: int, perturb: bool = False, kappa: float = 1.0, quad=None) -> Problem:
    if kind == T3:
        mesh, fe, q = T3block(1.0, 1.0, N, N), FEH1_T3(), 1
    elif kind == T6:
        mesh, fe, q = T6block_fast(1.0, 1.0, N, N), FEH1_T6(), 3
    elif kind == Q4:
        mesh, fe, q = Q4block(1.0, 1.0, N, N), FEH1_Q4(), 2
    elif kind == T4:      # examples/heat/poisson/t4.jl: unit cube, every boundary node prescribed, default rule (1 point)
        mesh, fe, q = T4block(1.0, 1.0, 1.0, N, N, N), FEH1_T4(), 1
    else:
        raise ValueError(kind)
    if kind == T4:
        on = ((mesh.xy == 0.0) | (mesh.xy == 1.0)).any(axis=1)
        bnd = np.nonzero(on)[0] + 1
        if perturb:       # interior nodes moved by up to 0.2 h
            rng = np.random.default_rng(20260101)
            xyz = mesh.xy.copy()
            xyz[~on] += rng.uniform(-0.2 / N, 0.2 / N, size=(int((~on).sum()), 3))
            mesh = Mesh(T4, mesh.conn, xyz)
    else:
        bnd = _structured_boundary(N, N, kind, mesh)
        if perturb:
            mesh = jitter(mesh)
    fesp = FESpace(mesh, fe)
    _setebc_nodes(fesp, bnd)
    numberfreedofs(fesp)
    numberdatadofs(fesp)
    return Problem(f"heat_{fe.name[5:].lower()}_N{N}", HeatForm(kappa), quad or q, [mesh], [fesp], [0], ndofs(fesp))


def plane_stress_D(E=1.0, nu=1.0 / 3):
    # examples/elasticity/stretch/t6.jl:25-29
    return E / (1 - nu ** 2) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])


def elasticity_problem(N: int, kind: int = T6, perturb: bool = False, A: float = 1.0) -> Problem:
    if kind == T6:
        mesh, fe, q = T6block_fast(A, A, N, N), FEH1_T6(), 3
    elif kind == T3:
        mesh, fe, q = T3block(A, A, N, N), FEH1_T3(), 1
    else:
        mesh, fe, q = Q4block(A, A, N, N), FEH1_Q4(), 2
    inflate = A / N / 100
    left = np.nonzero(np.abs(mesh.xy[:, 0]) <= inflate)[0] + 1
    right = np.nonzero(np.abs(mesh.xy[:, 0] - A) <= inflate)[0] + 1
    if perturb:
        mesh = jitter(mesh)
    fesp = FESpace(mesh, fe, 2)
    _setebc_nodes(fesp, left, (1, 2))
    _setebc_nodes(fesp, right, (1, 2))
    numberfreedofs(fesp)
    numberdatadofs(fesp)
    return Problem(f"elasticity_{fe.name[5:].lower()}_N{N}", ElasticityForm(plane_stress_D()), q, [mesh], [fesp], [0], ndofs(fesp))


def stokes_problem(N: int, formulation: str = "gen", perturb: bool = False, A: float = 1.0) -> Problem:
    vmesh = T6block_fast(2 * A, 2 * A, N, N)
    transform(vmesh, lambda x: x - A)
    bnd = _structured_boundary(N, N, T6, vmesh)
    if perturb:
        vmesh = jitter(vmesh)
    pmesh = T6toT3(vmesh)
    Ph = FESpace(pmesh, FEH1_T3(), 1)
    d = ((pmesh.xy - np.array([0.0, 0.0])) ** 2).sum(axis=1)
    _setebc_nodes(Ph, np.array([int(np.argmin(d)) + 1]))
    if formulation in ("gen", "veclap_alt"):
        Uh = FESpace(vmesh, FEH1_T6(), 2)
        _setebc_nodes(Uh, bnd, (1, 2))
        spaces, smesh = [Uh, Ph], [0, 1]
        form = StokesGenForm(np.diag([2.0, 2.0, 1.0])) if formulation == "gen" else StokesVeclapAltForm(1.0)
    elif formulation in ("reddy", "veclap"):
        ux, uy = FESpace(vmesh, FEH1_T6(), 1), FESpace(vmesh, FEH1_T6(), 1)
        _setebc_nodes(ux, bnd); _setebc_nodes(uy, bnd)
        spaces, smesh = [ux, uy, Ph], [0, 0, 1]
        form = StokesReddyForm(1.0) if formulation == "reddy" else StokesVeclapForm(1.0)
    else:
        raise ValueError(formulation)
    numberdofs(spaces)
    return Problem(f"stokes_{formulation}_N{N}", form, 3, [vmesh, pmesh], spaces, smesh, sum(ndofs(s) for s in spaces))


def stokes_f5_problem(N: int, pair: str = "p1b_p1", formulation: str = "reddy", perturb: bool = False, A: float = 1.0) -> Problem:
    """The colliding-flow set-up on the element pairs of SURVEY 8f row f5, all three spaces on ONE mesh:
    pair = "p1b_p1" (FEH1_T3_BUBBLE / FEH1_T3, npts 3), "q1_q0" (FEH1_Q4 / FEL2_Q4, Gauss order 2) or
    "p1_p0" (FEH1_T3 / FEL2_T3, npts 3)."""
    if pair == "q1_q0":
        mesh, vfe, pfe, q = Q4block(2 * A, 2 * A, N, N), FEH1_Q4(), FEL2_Q4(), 2
    else:
        mesh = T3block(2 * A, 2 * A, N, N)
        vfe, pfe, q = (FEH1_T3_BUBBLE(), FEH1_T3(), 3) if pair == "p1b_p1" else (FEH1_T3(), FEL2_T3(), 3)
    transform(mesh, lambda x: x - A)
    bnd = _structured_boundary(N, N, mesh.kind, mesh)
    regular_xy = mesh.xy
    if perturb:
        mesh = jitter(mesh)
    ux, uy, Ph = FESpace(mesh, vfe, 1), FESpace(mesh, vfe, 1), FESpace(mesh, pfe, 1)
    _setebc_nodes(ux, bnd); _setebc_nodes(uy, bnd)
    if Ph.field is not None:      # the pressure node nearest the centre
        _setebc_nodes(Ph, np.array([int(np.argmin((regular_xy ** 2).sum(axis=1))) + 1]))
    else:                         # q1_q0.jl:147-148: setebc!(pfesp, 2, atcenter[1], 1, 0.0) -- the id of the NODE nearest the centre, used as a cell id
        Ph.cellfield.isdatum[int(np.argmin((regular_xy ** 2).sum(axis=1))), 0] = True
    spaces = [ux, uy, Ph]
    numberdofs(spaces)
    form = StokesReddyForm(1.0) if formulation == "reddy" else StokesVeclapForm(1.0)
    return Problem(f"stokes_{pair}_{formulation}_N{N}", form, q, [mesh], spaces, [0, 0, 0], sum(ndofs(s) for s in spaces))


def load_problem(engine, prob: Problem, column_range=None):
    """Push a Problem through the C ABI: efg_set_mesh / efg_set_space / efg_start (/ efg_set_column_range)."""
    for slot, m in enumerate(prob.meshes):
        engine.set_mesh(slot, m.kind, np.ascontiguousarray(m.conn, dtype=np.int64),
                        np.ascontiguousarray(m.xy, dtype=np.float64))
    for slot, (s, ms) in enumerate(zip(prob.spaces, prob.space_mesh)):
        nd = None if s.field is None else np.ascontiguousarray(s.field.dofnums, dtype=np.int64)
        cd = None if s.cellfield is None else np.ascontiguousarray(s.cellfield.dofnums, dtype=np.int64)
        engine.set_space(slot, ms, nd, s.fe.fe_id, cd)
    engine.start(prob.ndofs, prob.ndofs)
    if column_range is not None:
        engine.set_column_range(*column_range)


def oracle_args(prob: Problem):
    """(form_id, quad, vmesh, pmesh, dofs, params) for oracle.assemble -- used by tests/bench only."""
    pm = prob.meshes[1] if len(prob.meshes) > 1 else None
    if any(s.cellfield is not None for s in prob.spaces):      # row f5: one mesh, elements named per space
        dofs = {"dofs": prob.dofs(), "cell_dofs": prob.cell_dofs(), "vfe": prob.spaces[0].fe.fe_id, "pfe": prob.spaces[2].fe.fe_id}
        return prob.form.form_id, prob.quad, prob.meshes[0], prob.meshes[0], dofs, prob.form.params()
    return prob.form.form_id, prob.quad, prob.meshes[0], pm, prob.dofs(), prob.form.params()
