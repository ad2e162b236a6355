"""ctypes driver of the CPU oracle (oracle/elfel_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libelfel_oracle.so")

FORM_HEAT, FORM_ELASTICITY, FORM_STOKES_GEN, FORM_STOKES_REDDY, FORM_STOKES_VECLAP_ALT, FORM_STOKES_VECLAP = 1, 2, 3, 4, 5, 6

FE_H1, FE_L2, FE_T3_BUBBLE = 0, 1, 7      # elements with a dof on the cell


class _Ext(C.Structure):      # efo_ext
    _fields_ = [("vfe", C.c_int), ("pfe", C.c_int), ("cdof0", C.POINTER(C.c_int64)), ("cdof1", C.POINTER(C.c_int64)),
                ("cdof2", C.POINTER(C.c_int64))]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "elfel_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i64p, f64p = C.POINTER(C.c_int64), C.POINTER(C.c_double)
        L.efo_quadrature.argtypes = [C.c_int, C.c_int, f64p, f64p]
        L.efo_quadrature.restype = C.c_int
        L.efo_bfun.argtypes = [C.c_int, C.c_double, C.c_double, f64p]
        L.efo_bfungradpar.argtypes = [C.c_int, C.c_double, C.c_double, f64p]
        L.efo_triplets_per_element.argtypes = [C.c_int, C.c_int, C.c_int]
        L.efo_triplets_per_element.restype = C.c_int64
        L.efo_assemble_coo.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64,
                                       i64p, C.c_int, f64p, i64p, C.c_int, f64p,
                                       i64p, i64p, i64p, f64p, i64p, i64p, f64p, C.POINTER(_Ext)]
        L.efo_assemble_coo.restype = C.c_int64
        common = [C.c_int, C.c_int, C.c_int64, i64p, C.c_int64, i64p, C.c_int, f64p, i64p, C.c_int, f64p, i64p, i64p, i64p]
        L.efo_direct_pattern.argtypes = common + [C.c_int64, C.c_int64, C.c_int64, C.c_int64, i64p, i64p, C.POINTER(_Ext)]
        L.efo_direct_pattern.restype = C.c_int64
        L.efo_direct_values.argtypes = common + [f64p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, i64p, i64p, f64p, C.POINTER(_Ext)]
        L.efo_direct_values.restype = C.c_int64
        L.efo_sparse.argtypes = [C.c_int64, C.c_int64, C.c_int64, i64p, i64p, f64p, i64p, i64p, f64p]
        L.efo_sparse.restype = C.c_int64
        L.efo_assemble_vec_heat.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, C.c_int, f64p, i64p, C.c_double, C.c_int64, f64p]
        L.efo_assemble_vec_heat.restype = C.c_int64
        L.efo_spmv_csc.argtypes = [C.c_int64, C.c_int64, i64p, i64p, f64p, f64p, f64p]
        L.efo_spmv_csc.restype = None
        L.efo_qp_locations.argtypes = [C.c_int, C.c_int64, i64p, C.c_int, f64p, f64p]
        L.efo_qp_locations.restype = C.c_int64
        L.efo_l2_error.argtypes = [C.c_int, C.c_int64, i64p, C.c_int, f64p, C.c_int, i64p, C.c_int, C.c_int,
                                   i64p, C.c_int, C.c_int, f64p, f64p]
        L.efo_l2_error.restype = C.c_double
        L.efo_l2_error_fe.argtypes = [C.c_int, C.c_int64, i64p, C.c_int, f64p, C.c_int, C.c_int, i64p, i64p, i64p, i64p, f64p, f64p]
        L.efo_l2_error_fe.restype = C.c_double
        _lib = L
    return _lib


def _p(a, ty):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ty))


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def _spaces(dofs):
    """dofs: a list of up to three (nnodes, ncomp) dofnums arrays, or -- spaces with cell dofs, row f5 -- a dict
    {"dofs": [...], "cell_dofs": [...], "vfe": FE_*, "pfe": FE_*}.  -> (d[3], ext pointer or None, keep-alive, (vfe, pfe))"""
    if isinstance(dofs, dict):
        d = [_c(x, np.int64) for x in dofs["dofs"]] + [None] * (3 - len(dofs["dofs"]))
        cd = [_c(x, np.int64) for x in dofs["cell_dofs"]] + [None] * (3 - len(dofs["cell_dofs"]))
        ext = _Ext(int(dofs["vfe"]), int(dofs["pfe"]), _p(cd[0], C.c_int64), _p(cd[1], C.c_int64), _p(cd[2], C.c_int64))
        return d, C.byref(ext), (cd, ext), (int(dofs["vfe"]), int(dofs["pfe"]))
    d = [_c(x, np.int64) for x in dofs] + [None] * (3 - len(dofs))
    return d, None, None, (0, 0)


def quadrature(kind, rule):
    pc = np.zeros((25, 2))
    w = np.zeros(25)
    n = lib().efo_quadrature(kind, rule, _p(pc, C.c_double), _p(w, C.c_double))
    if n < 0:
        raise ValueError("quadrature rule not available")
    return pc[:n].copy(), w[:n].copy()


def bfun(kind, r, s):
    N = np.zeros(lib().efo_nbf(kind))
    lib().efo_bfun(kind, r, s, _p(N, C.c_double))
    return N


def bfungradpar(kind, r, s):
    g = np.zeros((lib().efo_nbf(kind), 2))
    lib().efo_bfungradpar(kind, r, s, _p(g, C.c_double))
    return g


def assemble_coo(form, quad, vmesh, pmesh, dofs, params, e0=0, e1=None):
    """Element loop -> COO triplets (row, col, val), in the reference's append order.

    vmesh/pmesh: objects with .kind, .conn (nel,nen) int64 1-based, .xy (nnodes,2);
    dofs: list of up to three (nnodes, ncomp) int64 dofnums arrays; params: float64 vector."""
    L = lib()
    e1 = vmesh.conn.shape[0] if e1 is None else e1
    pk = pmesh.kind if pmesh is not None else 0
    d, ext, _keep, (vfe, pfe) = _spaces(dofs)
    tpe = L.efo_triplets_per_element(form, vfe or vmesh.kind, pfe or pk)
    nt = tpe * (e1 - e0)
    row = np.empty(nt, dtype=np.int64)
    col = np.empty(nt, dtype=np.int64)
    val = np.empty(nt, dtype=np.float64)
    vconn = _c(vmesh.conn, np.int64)
    vxy = _c(vmesh.xy, np.float64)
    pconn = _c(pmesh.conn, np.int64) if pmesh is not None else None
    pxy = _c(pmesh.xy, np.float64) if pmesh is not None else None
    prm = _c(np.atleast_1d(params), np.float64)
    n = L.efo_assemble_coo(form, quad, e0, e1, _p(vconn, C.c_int64), vmesh.kind, _p(vxy, C.c_double),
                           _p(pconn, C.c_int64), pk, _p(pxy, C.c_double),
                           _p(d[0], C.c_int64), _p(d[1], C.c_int64), _p(d[2], C.c_int64),
                           _p(prm, C.c_double), _p(row, C.c_int64), _p(col, C.c_int64), _p(val, C.c_double), ext)
    if n != nt:
        raise RuntimeError(f"oracle element loop failed (rc={n})")
    return row, col, val


def sparse(row, col, val, nrow, ncol):
    """finish!: SparseArrays.sparse(I,J,V,m,n) -> (colptr, rowval, nzval), Int64 1-based."""
    L = lib()
    nt = len(row)
    colptr = np.empty(ncol + 1, dtype=np.int64)
    rowval = np.empty(max(nt, 1), dtype=np.int64)
    nzval = np.empty(max(nt, 1), dtype=np.float64)
    row = _c(row, np.int64); col = _c(col, np.int64); val = _c(val, np.float64)
    nnz = L.efo_sparse(nrow, ncol, nt, _p(row, C.c_int64), _p(col, C.c_int64), _p(val, C.c_double),
                       _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_double))
    if nnz == -1:
        raise ValueError("ArgumentError: row/column index out of range (dof number 0 or > nrow?)")
    if nnz < 0:
        raise MemoryError("oracle sparse(): allocation failed")
    return colptr, rowval[:nnz].copy(), nzval[:nnz].copy()


def assemble(form, quad, vmesh, pmesh, dofs, params, nrow, ncol, timing=None):
    """start! / element loop / finish! -> CSC, exactly what the reference's assembleK returns."""
    t0 = time.perf_counter()
    row, col, val = assemble_coo(form, quad, vmesh, pmesh, dofs, params)
    t1 = time.perf_counter()
    out = sparse(row, col, val, nrow, ncol)
    t2 = time.perf_counter()
    if timing is not None:
        timing["integrate_s"] = t1 - t0
        timing["finish_s"] = t2 - t1
    return out


def elements_touching(dofs_per_mesh, c0, c1):
    """Ascending 0-based numbers of the elements with at least one dof in the column block [c0, c1] (1-based inclusive).
    dofs_per_mesh: [(conn (nel, nen), dofnums (nnodes, ncomp)), ...] for every space of the form; conn = None for a
    cell field (dofnums (nel, ncomp): term e belongs to element e)."""
    hit = None
    for conn, dofnums in dofs_per_mesh:
        d = np.asarray(dofnums)
        inb = ((d >= c0) & (d <= c1)).any(axis=1)
        h = inb if conn is None else inb[np.asarray(conn) - 1].any(axis=1)
        hit = h if hit is None else (hit | h)
    return np.nonzero(hit)[0].astype(np.int64)


def assemble_direct(form, quad, vmesh, pmesh, dofs, params, nrow, ncol, c0=1, c1=None, elist=None, timing=None):
    """Direct-accumulate oracle: CSC of the column block [c0, c1] (1-based inclusive; default: the whole
    matrix) -- the same element traversal and left-to-right sums as assemble(), into a pre-built pattern, without the
    24 B/triplet COO list.  colptr is rebased to the block.  Bit-identical to assemble() (tested)."""
    L = lib()
    c1 = ncol if c1 is None else c1
    pk = pmesh.kind if pmesh is not None else 0
    vconn = _c(vmesh.conn, np.int64); vxy = _c(vmesh.xy, np.float64)
    pconn = _c(pmesh.conn, np.int64) if pmesh is not None else None
    pxy = _c(pmesh.xy, np.float64) if pmesh is not None else None
    d, ext, _keep, _ = _spaces(dofs)
    prm = _c(np.atleast_1d(params), np.float64)
    el = _c(elist, np.int64) if elist is not None else None
    nsel = 0 if el is None else len(el)
    nel = vconn.shape[0]
    head = (form, quad, nel, _p(el, C.c_int64), nsel, _p(vconn, C.c_int64), vmesh.kind, _p(vxy, C.c_double),
            _p(pconn, C.c_int64), pk, _p(pxy, C.c_double), _p(d[0], C.c_int64), _p(d[1], C.c_int64), _p(d[2], C.c_int64))
    t0 = time.perf_counter()
    colptr = np.empty(c1 - c0 + 2, dtype=np.int64)
    nnz = L.efo_direct_pattern(*head, nrow, ncol, c0, c1, _p(colptr, C.c_int64), None, ext)
    if nnz == -1:
        raise ValueError("ArgumentError: row/column index out of range (dof number 0 or > nrow?)")
    if nnz < 0:
        raise MemoryError("oracle direct mode: allocation failed")
    rowval = np.empty(max(nnz, 1), dtype=np.int64)
    L.efo_direct_pattern(*head, nrow, ncol, c0, c1, _p(colptr, C.c_int64), _p(rowval, C.c_int64), ext)
    t1 = time.perf_counter()
    nzval = np.empty(max(nnz, 1), dtype=np.float64)
    rc = L.efo_direct_values(*head, _p(prm, C.c_double), nrow, ncol, c0, c1, _p(colptr, C.c_int64), _p(rowval, C.c_int64),
                             _p(nzval, C.c_double), ext)
    if rc != nnz:
        raise RuntimeError(f"oracle direct accumulation failed (rc={rc})")
    if timing is not None:
        timing["pattern_s"] = t1 - t0
        timing["values_s"] = time.perf_counter() - t1
    return colptr, rowval[:nnz], nzval[:nnz]


def assemble_direct_parallel(prob_args, nrow, ncol, space_pairs, nblocks=None, threads=None, c0=1, c1=None):
    """assemble_direct over `nblocks` column blocks of [c0, c1] on a thread pool (ctypes releases the GIL), each block
    visiting only the elements that touch it; the blocks are concatenated.  Same result as assemble_direct, used to check
    matrices of the BASELINE sizes in bounded time and memory.  prob_args = (form, quad, vmesh, pmesh, dofs, params);
    space_pairs = [(conn, dofnums) per space] (see elements_touching)."""
    from concurrent.futures import ThreadPoolExecutor
    c1 = ncol if c1 is None else c1
    threads = threads or min(os.cpu_count() or 1, 16)
    nblocks = nblocks or max(threads * 2, 1)
    nblocks = max(1, min(nblocks, c1 - c0 + 1))
    cuts = np.linspace(c0 - 1, c1, nblocks + 1).astype(np.int64)
    lib()

    def one(k):
        a, b = int(cuts[k]) + 1, int(cuts[k + 1])
        el = elements_touching(space_pairs, a, b)
        return assemble_direct(*prob_args, nrow, ncol, c0=a, c1=b, elist=el)
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(one, range(nblocks)))
    colptr = [np.array([1], dtype=np.int64)]
    off = 0
    for cp, _, _ in parts:
        colptr.append(cp[1:] + off)
        off += int(cp[-1]) - 1
    return np.concatenate(colptr), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def assemble_vec_heat(quad, mesh, dofnums, Q, nrow, e0=0, e1=None):
    """start!(av, nrow) / `fe[j] += N[j]*Q*JxW` element loop / finish!(av) -> the load vector (nrow,)."""
    conn = _c(mesh.conn, np.int64); xy = _c(mesh.xy, np.float64); d = _c(dofnums, np.int64)
    e1 = conn.shape[0] if e1 is None else e1
    val = np.empty(nrow, dtype=np.float64)
    rc = lib().efo_assemble_vec_heat(quad, e0, e1, _p(conn, C.c_int64), mesh.kind, _p(xy, C.c_double),
                                     _p(d, C.c_int64), float(Q), nrow, _p(val, C.c_double))
    if rc == -2:
        raise IndexError("BoundsError: dof number outside 1..nrow")
    if rc != 0:
        raise ValueError("quadrature rule not available")
    return val


def spmv_csc(nrow, ncol, colptr, rowval, nzval, x):
    """K * x with SparseArrays' column-sweep accumulation order."""
    colptr = _c(colptr, np.int64); rowval = _c(rowval, np.int64); nzval = _c(nzval, np.float64); x = _c(x, np.float64)
    y = np.empty(nrow, dtype=np.float64)
    lib().efo_spmv_csc(nrow, ncol, _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_double),
                       _p(x, C.c_double), _p(y, C.c_double))
    return y


def qp_locations(quad, mesh):
    """location(el, qp) for every element / quadrature point -> (nel, npts, 2)."""
    conn = _c(mesh.conn, np.int64); xy = _c(mesh.xy, np.float64)
    out = np.empty((conn.shape[0], MAXQP_, 2))
    n = lib().efo_qp_locations(quad, conn.shape[0], _p(conn, C.c_int64), mesh.kind, _p(xy, C.c_double), _p(out, C.c_double))
    if n < 0:
        raise ValueError("quadrature rule not available")
    return out.reshape(-1)[:conn.shape[0] * n * 2].reshape(conn.shape[0], n, 2).copy()


MAXQP_ = 25


def l2_error(quad, mesh, comps, U, truth):
    """evaluate_*_error: comps = [(dofnums, comp0based), ...] (1 or 2 entries, all on `mesh`), U the system vector,
    truth (nel, npts, ncomp) the true solution at the quadrature-point locations -> sqrt(E)."""
    conn = _c(mesh.conn, np.int64); xy = _c(mesh.xy, np.float64)
    d = [_c(c[0], np.int64) for c in comps]
    ncs = [x.shape[1] for x in d]
    cc = [int(c[1]) for c in comps]
    if len(d) == 1:
        d.append(d[0]); ncs.append(ncs[0]); cc.append(cc[0])
    U = _c(U, np.float64); truth = _c(truth, np.float64)
    r = lib().efo_l2_error(quad, conn.shape[0], _p(conn, C.c_int64), mesh.kind, _p(xy, C.c_double), len(comps),
                           _p(d[0], C.c_int64), ncs[0], cc[0], _p(d[1], C.c_int64), ncs[1], cc[1],
                           _p(U, C.c_double), _p(truth, C.c_double))
    if r < 0:
        raise ValueError("quadrature rule not available")
    return r


def l2_error_fe(quad, mesh, fe, comps, U, truth):
    """evaluate_*_error for scalar spaces whose element has a cell dof (FEH1_T3_BUBBLE, row f5): comps =
    [(dofnums (nnodes,1), cell_dofnums (nel,1)), ...] (1 or 2 entries)."""
    conn = _c(mesh.conn, np.int64); xy = _c(mesh.xy, np.float64)
    d = [(_c(c[0], np.int64), _c(c[1], np.int64)) for c in comps]
    if len(d) == 1:
        d.append(d[0])
    U = _c(U, np.float64); truth = _c(truth, np.float64)
    r = lib().efo_l2_error_fe(quad, conn.shape[0], _p(conn, C.c_int64), mesh.kind, _p(xy, C.c_double), int(fe), len(comps),
                              _p(d[0][0], C.c_int64), _p(d[0][1], C.c_int64), _p(d[1][0], C.c_int64), _p(d[1][1], C.c_int64),
                              _p(U, C.c_double), _p(truth, C.c_double))
    if r < 0:
        raise ValueError("quadrature rule not available")
    return r
