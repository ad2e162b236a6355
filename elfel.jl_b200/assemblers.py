"""Host-side mirror of the reference's assembler interface for the accelerated path.

Reference (Julia)                                   here (Python, same names without ``!``)
--------------------------------------------------  ------------------------------------------------
FEIterator(fesp)            src/FEIterators.jl:54   FEIterator(fesp)
QPIterator(fesp, (kind=:default, npts=3))           QPIterator(fesp, kind="default", npts=3)
                            src/QPIterators.jl:79   QPIterator(fesp, kind="Gauss", order=2)
SysmatAssemblerSparse(0.0)  src/Assemblers.jl:58    SysmatAssemblerGPU(0.0)
start!(ass, nrow, ncol)     src/Assemblers.jl:67    start(ass, nrow, ncol)
for el in elit ... assemble!(ass, ke) end           assemble(ass, HeatForm(kappa), elit, qpit)
                 (user closure, examples/*)         assemble(ass, ElasticityForm(D), elit, qpit)
                                                    assemble(ass, StokesGenForm(D), (uel, pel), (uqp, pqp)) ...
finish!(ass)                src/Assemblers.jl:121   finish(ass) -> SparseMatrixCSC(m, n, colptr, rowval, nzval)
SysvecAssembler(0.0)        src/Assemblers.jl:183   SysvecAssemblerGPU(0.0, like=am)   (shares am's device context)
start!(av, nrow)            src/Assemblers.jl:196   start(av, nrow)
init!(fe, eldofs(el)); fe[j] += N[j]*Q*JxW;         assemble(av, HeatLoadForm(Q), elit, qpit)
  assemble!(av, fe)         examples/heat/.../t3.jl:44-61
finish!(av)                 src/Assemblers.jl:230   finish(av) -> numpy vector
KT = K * T                  examples/heat/.../t3.jl:78   mul(am, T)              (on the device, Julia's summation order)
K[1:nu, 1:nu]               examples/heat/.../t3.jl:79   block(am, 1, nu, 1, nu) (sliced on the device)

The element loop + COO append + sparse() of the reference collapse into one ``assemble`` call that
runs on the GPU through the C ABI (include/elfel_gpu.h).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .fespaces import FESpace
from .meshes import Q4


# ------------------------------------------------------------------------------------------------
# iterators (thin: they only carry what the engine reads from them)
# ------------------------------------------------------------------------------------------------
class FEIterator:
    """Caches the base incidence relation, geometry and the field of a space (src/FEIterators.jl:54-81)."""

    def __init__(self, fesp: FESpace):
        self.fesp = fesp
        self._bir = fesp.mesh.conn        # element -> nodes, (nel, nen) int64 1-based
        self._geom = fesp.mesh.xy         # (nnodes, 2)
        self._fld0 = fesp.field           # vertex dofs (None for an L2 space)
        self._fld2 = fesp.cellfield       # cell dofs (FEH1_T3_BUBBLE, FEL2_*), else None

    def __len__(self):
        return self._bir.shape[0]


class QPIterator:
    """Quadrature settings of a space (src/QPIterators.jl:79-84, src/RefShapes.jl:301-323,333-366)."""

    def __init__(self, fesp: FESpace, kind="default", npts=None, order=None):
        self.fesp = fesp
        if fesp.fe.kind == Q4:
            if kind not in ("default", "Gauss"):
                raise ValueError(f"Integration rule {kind} not available")
            self.rule = 1 if order is None else int(order)
        else:
            if kind != "default":
                raise ValueError(f"Integration rule {kind} not available")
            self.rule = 1 if npts is None else int(npts)


# ------------------------------------------------------------------------------------------------
# weak forms = the integrate! closures of the reference's examples
# ------------------------------------------------------------------------------------------------
@dataclass
class HeatForm:           # examples/heat/poisson/t3.jl:53-58
    kappa: float
    form_id = _lib.FORM_HEAT

    def params(self):
        return np.array([self.kappa], dtype=np.float64)


@dataclass
class HeatLoadForm:       # examples/heat/poisson/t3.jl:57  fe[j] += N[j] * Q * JxW   (vector form)
    Q: float = 0.0
    vform_id = _lib.VFORM_HEAT_LOAD

    def params(self):
        return np.array([self.Q], dtype=np.float64)


@dataclass
class ElasticityForm:     # examples/elasticity/stretch/t6.jl:42-58 ; D is the 3x3 material matrix
    D: np.ndarray
    form_id = _lib.FORM_ELASTICITY

    def params(self):
        return np.asarray(self.D, dtype=np.float64).T.ravel().copy()   # column-major like SMatrix{3,3}


@dataclass
class StokesGenForm:      # examples/stokes/colliding_flow/ht_p2_p1_gen.jl:48-79 ; spaces (Uh, Ph)
    D: np.ndarray
    form_id = _lib.FORM_STOKES_GEN

    def params(self):
        return np.asarray(self.D, dtype=np.float64).T.ravel().copy()


@dataclass
class StokesReddyForm:    # examples/stokes/colliding_flow/ht_p2_p1.jl:56-104 ; spaces (ux, uy, p)
    mu: float
    form_id = _lib.FORM_STOKES_REDDY

    def params(self):
        return np.array([self.mu], dtype=np.float64)


@dataclass
class StokesVeclapAltForm:  # examples/stokes/colliding_flow/ht_p2_p1_veclap_alt.jl:61-90 ; spaces (u, p)
    mu: float
    form_id = _lib.FORM_STOKES_VECLAP_ALT

    def params(self):
        return np.array([self.mu], dtype=np.float64)


@dataclass
class StokesVeclapForm:   # examples/stokes/colliding_flow/ht_p2_p1_veclap.jl:56-97 ; spaces (ux, uy, p)
    mu: float
    form_id = _lib.FORM_STOKES_VECLAP

    def params(self):
        return np.array([self.mu], dtype=np.float64)


@dataclass
class SparseMatrixCSC:
    """Julia's SparseMatrixCSC{Float64,Int64}: 1-based colptr / rowval."""
    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))


def _ptr(a):
    """Raw address of a numpy array or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class _DeviceArray:
    """A 1-D device array owned by an Engine, described through the CUDA array interface (version 3)."""

    def __init__(self, ptr, n, typestr, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr or 0), False), "version": 3,
                                         "strides": None, "stream": 1}


class Engine:
    """One ``efg_ctx`` (one device, one stream).  Thin, 1:1 with the C ABI."""

    def __init__(self, device: int = 0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.efg_create(device, C.byref(h))
        if rc != _lib.OK:
            raise _lib.EfgError(rc, "efg_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.ncols_local = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.efg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(self.h, rc)

    def set_option(self, opt, value):
        self._ck(self.L.efg_set_option(self.h, opt, int(value)))

    def stat(self, which) -> float:
        v = C.c_double()
        self._ck(self.L.efg_get_stat(self.h, which, C.byref(v)))
        return v.value

    def stream(self) -> int:
        s = C.c_void_p()
        self._ck(self.L.efg_get_stream(self.h, C.byref(s)))
        return s.value or 0

    def synchronize(self):
        self._ck(self.L.efg_synchronize(self.h))

    def set_mesh(self, slot, kind, conn, xy):
        """conn: (nel, nen) int64 1-based, xy: (nnodes, 2) float64 -- numpy (host) or torch (host/device)."""
        nel, nnodes = int(conn.shape[0]), int(xy.shape[0])
        if int(xy.shape[1]) == 3:       # FEH1_T4: xyz (nnodes, 3)
            self._ck(self.L.efg_set_mesh3(self.h, slot, kind, nel, nnodes, _ptr(conn), _ptr(xy)))
            return
        self._ck(self.L.efg_set_mesh(self.h, slot, kind, nel, nnodes, _ptr(conn), _ptr(xy)))

    def set_space(self, slot, mesh_slot, dofnums, fe=_lib.FE_H1, cell_dofnums=None):
        """dofnums: (nnodes, ncomp) int64 1-based (None for an L2 space); fe / cell_dofnums (nel, ncomp): spaces whose
        element carries a dof on the cell (efg_set_space_fe)."""
        if fe == _lib.FE_H1:
            nnodes, ncomp = int(dofnums.shape[0]), int(dofnums.shape[1])
            self._ck(self.L.efg_set_space(self.h, slot, mesh_slot, ncomp, nnodes, _ptr(dofnums)))
            return
        nel, ncomp = int(cell_dofnums.shape[0]), int(cell_dofnums.shape[1])
        nnodes = 0 if dofnums is None else int(dofnums.shape[0])
        self._ck(self.L.efg_set_space_fe(self.h, slot, mesh_slot, int(fe), ncomp, nnodes, _ptr(dofnums), nel, _ptr(cell_dofnums)))

    def start(self, nrow, ncol):
        self._ck(self.L.efg_start(self.h, int(nrow), int(ncol)))
        self.nrow, self.ncol, self.ncols_local = int(nrow), int(ncol), int(ncol)

    def set_column_range(self, first, last):
        self._ck(self.L.efg_set_column_range(self.h, int(first), int(last)))
        self.ncols_local = int(last) - int(first) + 1

    def set_column_ranges(self, firsts, lasts):
        f = np.ascontiguousarray(firsts, dtype=np.int64)
        l = np.ascontiguousarray(lasts, dtype=np.int64)
        self._ck(self.L.efg_set_column_ranges(self.h, len(f), f.ctypes.data_as(C.POINTER(C.c_int64)),
                                              l.ctypes.data_as(C.POINTER(C.c_int64))))
        self.ncols_local = int((l - f + 1).sum())

    def symbolic(self, form_id, quad) -> int:
        nnz = C.c_int64()
        self._ck(self.L.efg_symbolic(self.h, form_id, quad, C.byref(nnz)))
        self.nnz = nnz.value
        return nnz.value

    def pattern(self, form_id, quad) -> int:
        """First half of the symbolic phase: the CSC pattern only (-> nnz); the tiles are built by the next numeric call."""
        nnz = C.c_int64()
        self._ck(self.L.efg_pattern(self.h, form_id, quad, C.byref(nnz)))
        self.nnz = nnz.value
        return nnz.value

    def fetch_pattern_async(self, colptr, rowval):
        """Start copying colptr / rowval to the caller's arrays; completed by the next fetch_csc call."""
        self._keep_out = (colptr, rowval)
        self._ck(self.L.efg_fetch_pattern_async(self.h, _ptr(colptr), _ptr(rowval)))

    def numeric(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efg_numeric(self.h, p.ctypes.data_as(C.POINTER(C.c_double)), len(p)))

    def numeric_with_load(self, params, Q):
        """K and the heat load vector in one pass (efg_numeric_with_load; set OPT_FUSE_LOAD before the symbolic phase)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efg_numeric_with_load(self.h, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), float(Q)))
        self.nvec = self.nrow

    def assemble(self, form_id, quad, params) -> int:
        nnz = C.c_int64()
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efg_assemble(self.h, form_id, quad, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), C.byref(nnz)))
        self.nnz = nnz.value
        return nnz.value

    def fetch_csc(self, colptr=None, rowval=None, nzval=None):
        """Fill caller-allocated arrays (numpy or torch); allocates numpy arrays when all are None."""
        if colptr is None and rowval is None and nzval is None:
            colptr = np.empty(self.ncols_local + 1, dtype=np.int64)
            rowval = np.empty(self.nnz, dtype=np.int64)
            nzval = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.L.efg_fetch_csc(self.h, _ptr(colptr), _ptr(rowval), _ptr(nzval)))
        return colptr, rowval, nzval


    def device_csc(self):
        """Zero-copy views of the device-resident result (efg_device_csc): (colptr int64 1-based, rowval int32 0-BASED,
        nzval float64) as objects exposing ``__cuda_array_interface__`` (torch.as_tensor / cupy.asarray wrap them without
        a copy).  Valid until the next symbolic phase / efg_set_mesh / close."""
        cp, rv, nz = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.efg_device_csc(self.h, C.byref(cp), C.byref(rv), C.byref(nz)))
        return (_DeviceArray(cp.value, self.ncols_local + 1, "<i8", self), _DeviceArray(rv.value, self.nnz, "<i4", self),
                _DeviceArray(nz.value, self.nnz, "<f8", self))

    # ---- inputs made on the device -------------------------------------------------
    def gen_mesh(self, slot, kind, nL, nW, Length=1.0, Width=1.0, xshift=0.0, yshift=0.0):
        self._ck(self.L.efg_gen_mesh(self.h, slot, kind, int(nL), int(nW), float(Length), float(Width), float(xshift), float(yshift)))

    def gen_mesh_corners(self, slot_dst, slot_src):
        self._ck(self.L.efg_gen_mesh_corners(self.h, slot_dst, slot_src))

    def gen_space(self, slot, mesh_slot, ncomp):
        self._ck(self.L.efg_gen_space(self.h, slot, mesh_slot, ncomp))

    def setebc_box(self, space_slot, comp, x0, x1, y0, y1):
        self._ck(self.L.efg_setebc_box(self.h, space_slot, comp, float(x0), float(x1), float(y0), float(y1)))

    def setebc_nodes(self, space_slot, comp, node_ids):
        ids = np.ascontiguousarray(node_ids, dtype=np.int64)
        self._ck(self.L.efg_setebc_nodes(self.h, space_slot, comp, len(ids), ids.ctypes.data))

    def number_dofs(self, space_slots):
        arr = (C.c_int * len(space_slots))(*[int(s) for s in space_slots])
        nfree, ndofs = C.c_int64(), C.c_int64()
        self._ck(self.L.efg_number_dofs(self.h, len(space_slots), arr, C.byref(nfree), C.byref(ndofs)))
        return nfree.value, ndofs.value

    def fetch_mesh(self, slot, kind):
        nel, nn = C.c_int64(), C.c_int64()
        self._ck(self.L.efg_fetch_mesh(self.h, slot, C.byref(nel), C.byref(nn), None, None))
        conn = np.empty((nel.value, kind), dtype=np.int64)
        xy = np.empty((nn.value, 2), dtype=np.float64)
        self._ck(self.L.efg_fetch_mesh(self.h, slot, None, None, _ptr(conn), _ptr(xy)))
        return conn, xy

    def fetch_dofnums(self, space_slot, nnodes, ncomp):
        out = np.empty((int(nnodes), int(ncomp)), dtype=np.int64)
        self._ck(self.L.efg_fetch_dofnums(self.h, space_slot, _ptr(out)))
        return out

    # ---- SURVEY 8f rows f1 / f2 -------------------------------------------------------------------
    def vec_assemble(self, vform_id, quad, params, nrow):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efg_vec_assemble(self.h, vform_id, quad, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), int(nrow)))
        n = C.c_int64()
        self._ck(self.L.efg_device_vec(self.h, None, C.byref(n)))
        self.nvec = n.value

    def fetch_vec(self, out=None):
        if out is None:
            out = np.empty(self.nvec, dtype=np.float64)
        self._ck(self.L.efg_fetch_vec(self.h, _ptr(out)))
        return out

    def spmv(self, x, y=None):
        """y = K*x in SparseArrays' accumulation order; x, y numpy (host) or torch (host/device) float64."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty(self.nrow, dtype=np.float64)
        self._ck(self.L.efg_spmv(self.h, _ptr(x), _ptr(y)))
        return y

    def block(self, r0, r1, c0, c1):
        """K[r0:r1, c0:c1] (1-based inclusive) -> (colptr, rowval, nzval) of the block, sliced on the device."""
        nnz = C.c_int64()
        self._ck(self.L.efg_block_nnz(self.h, int(r0), int(r1), int(c0), int(c1), C.byref(nnz)))
        colptr = np.empty(int(c1) - int(c0) + 2, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value, dtype=np.float64)
        self._ck(self.L.efg_fetch_block(self.h, _ptr(colptr), _ptr(rowval), _ptr(nzval)))
        return colptr, rowval, nzval


    # ---- SURVEY 8f row f3 ---------------------------------------------------------------------------
    def qp_locations(self, mesh_slot, quad, nel):
        """location(el, qp) of every element / quadrature point of a mesh -> (nel, npts, 2)."""
        npts = C.c_int64()
        self._ck(self.L.efg_qp_locations(self.h, mesh_slot, quad, None, C.byref(npts)))
        out = np.empty((int(nel), npts.value, 2), dtype=np.float64)
        self._ck(self.L.efg_qp_locations(self.h, mesh_slot, quad, _ptr(out), C.byref(npts)))
        return out

    def l2_error(self, comps, quad, U, truth) -> float:
        """comps: [(space_slot, component0based), ...] (1 or 2); U: system vector; truth: (nel, npts, ncomp)."""
        slots = (C.c_int * len(comps))(*[int(c[0]) for c in comps])
        cc = (C.c_int * len(comps))(*[int(c[1]) for c in comps])
        if isinstance(U, np.ndarray):
            U = np.ascontiguousarray(U, dtype=np.float64)
        if isinstance(truth, np.ndarray):
            truth = np.ascontiguousarray(truth, dtype=np.float64)
        out = C.c_double()
        self._ck(self.L.efg_l2_error(self.h, len(comps), slots, cc, quad, _ptr(U), int(U.shape[0]), _ptr(truth), C.byref(out)))
        return out.value


class MultiEngine:
    """One ``efg_multi``: several GPUs behind one handle, global arrays in, global CSC out (include/elfel_gpu.h, efgm_*).
    The global arrays passed to set_mesh / set_space must stay alive until assemble() has returned."""

    def __init__(self, devices):
        self.L = _lib.load()
        devs = list(devices)
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        rc = self.L.efgm_create(len(devs), arr, C.byref(h))
        if rc != _lib.OK:
            raise _lib.EfgError(rc, "efgm_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.efgm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != _lib.OK:
            msg = self.L.efgm_last_error(self.h).decode()
            raise (_lib.ArgumentError if rc == _lib.ERR_INDEX else _lib.EfgError)(rc, msg)

    def set_option(self, opt, value):
        self._ck(self.L.efgm_set_option(self.h, opt, int(value)))

    def set_mesh(self, slot, kind, conn, xy):
        self._keep += [conn, xy]
        self._ck(self.L.efgm_set_mesh(self.h, slot, kind, int(conn.shape[0]), int(xy.shape[0]), _ptr(conn), _ptr(xy)))

    def set_space(self, slot, mesh_slot, dofnums):
        self._keep.append(dofnums)
        self._ck(self.L.efgm_set_space(self.h, slot, mesh_slot, int(dofnums.shape[1]), int(dofnums.shape[0]), _ptr(dofnums)))

    def start(self, nrow, ncol):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self._ck(self.L.efgm_start(self.h, int(nrow), int(ncol)))

    def assemble(self, form_id, quad, params) -> int:
        nnz = C.c_int64()
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efgm_assemble(self.h, form_id, quad, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), C.byref(nnz)))
        self.nnz = nnz.value
        self._keep = []
        return nnz.value

    def numeric(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.L.efgm_numeric(self.h, p.ctypes.data_as(C.POINTER(C.c_double)), len(p)))

    def fetch_csc(self):
        colptr = np.empty(self.ncol + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        nzval = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.L.efgm_fetch_csc(self.h, _ptr(colptr), _ptr(rowval), _ptr(nzval)))
        return colptr, rowval, nzval

    def stat(self, which, device=-1) -> float:
        v = C.c_double()
        self._ck(self.L.efgm_get_stat(self.h, which, device, C.byref(v)))
        return v.value

    def device_ranges(self, device):
        """Column ranges (firsts, lasts; 1-based inclusive) of the block device index `device` holds."""
        n, f, l = C.c_int64(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        self._ck(self.L.efgm_device_ctx(self.h, device, None, C.byref(n), C.byref(f), C.byref(l)))
        return (np.array([f[i] for i in range(n.value)], dtype=np.int64), np.array([l[i] for i in range(n.value)], dtype=np.int64))


class SysmatAssemblerGPU:
    """Selected in place of SysmatAssemblerSparse; same start / assemble / finish life cycle."""

    def __init__(self, zero: float = 0.0, device: int = 0):
        if not isinstance(zero, float):
            raise TypeError("only Float64 matrices are assembled")
        self.engine = Engine(device)
        self.nrow = self.ncol = 0
        self._started = False
        self._assembled = False

    def set_option(self, opt, value):
        self.engine.set_option(opt, value)
        return self


class SysvecAssemblerGPU:
    """Selected in place of SysvecAssembler (src/Assemblers.jl:170-232).  ``like=am`` shares the device context of a
    SysmatAssemblerGPU, so the mesh and dof maps uploaded for the matrix are reused for the vector.  Call order on a
    shared context: ``assemble(am, ...)`` first, then ``assemble(av, ...)`` (a later matrix assembly re-uploads the inputs and
    drops the vector: fetch it with ``finish(av)`` before) -- or ``assemble_both`` for K and F in one pass."""

    def __init__(self, zero: float = 0.0, device: int = 0, like: "SysmatAssemblerGPU | None" = None):
        if not isinstance(zero, float):
            raise TypeError("only Float64 vectors are assembled")
        self.engine = like.engine if like is not None else Engine(device)
        self.ndofs = 0
        self._started = False
        self._assembled = False


def _load_spaces(eng, elits, reuse=False):
    """Upload meshes + dof maps.  ``reuse``: skip the upload when this engine was last loaded from exactly these
    objects (the vector half of one integrate! call that follows the matrix half on a shared context)."""
    # the arrays themselves (strong references, compared by identity): numberfreedofs / numberdatadofs / setebc replace
    # field.dofnums by a new array, which is then seen as a change; an id() of a freed array could be recycled.  In-place
    # edits of mesh.xy / dofnums are NOT detected: call assemble(am, ...) again (it always re-uploads) after such edits.
    token = []
    for it in elits:
        token += [it.fesp.mesh.conn, it.fesp.mesh.xy, None if it._fld0 is None else it._fld0.dofnums,
                  None if it._fld2 is None else it._fld2.dofnums]
    old = getattr(eng, "_loaded", None)
    if reuse and old is not None and len(old) == len(token) and all(a is b for a, b in zip(old, token)):
        return
    meshes = []
    for it in elits:
        if not any(it.fesp.mesh is m for m in meshes):
            meshes.append(it.fesp.mesh)
    if len(meshes) > 2:
        raise ValueError("at most two meshes (velocity, pressure)")
    for slot, m in enumerate(meshes):
        eng.set_mesh(slot, m.kind, np.ascontiguousarray(m.conn, dtype=np.int64), np.ascontiguousarray(m.xy, dtype=np.float64))
    for slot, it in enumerate(elits):
        mslot = [i for i, m in enumerate(meshes) if m is it.fesp.mesh][0]
        nd = None if it._fld0 is None else np.ascontiguousarray(it._fld0.dofnums, dtype=np.int64)
        cd = None if it._fld2 is None else np.ascontiguousarray(it._fld2.dofnums, dtype=np.int64)
        eng.set_space(slot, mslot, nd, it.fesp.fe.fe_id, cd)
    eng._loaded = token
    eng._keep = [it for it in elits]


def start(ass, nrow, ncol=None):
    if isinstance(ass, SysvecAssemblerGPU):   # start!(av, nrow): src/Assemblers.jl:196-200
        ass.ndofs = int(nrow)
        ass._started, ass._assembled = True, False
        return ass
    ass.engine.start(nrow, ncol)
    ass.nrow, ass.ncol = int(nrow), int(ncol)
    ass._started, ass._assembled = True, False
    return ass


def assemble(ass: SysmatAssemblerGPU, form, elits, qpits):
    """One call replaces the reference's whole element loop for ``form``."""
    if not ass._started:
        raise _lib.EfgError(_lib.ERR_STATE, "assemble before start")
    if isinstance(elits, FEIterator):
        elits, qpits = (elits,), (qpits,)
    if len({q.rule for q in qpits}) != 1:
        raise ValueError("all spaces of a mixed form must use the same quadrature rule")
    eng = ass.engine
    _load_spaces(eng, elits, reuse=isinstance(ass, SysvecAssemblerGPU))
    if isinstance(ass, SysvecAssemblerGPU):
        eng.vec_assemble(form.vform_id, qpits[0].rule, form.params(), ass.ndofs)
        ass._assembled = True
        return ass
    eng.start(ass.nrow, ass.ncol)
    # structure first (-> nnz: the SparseMatrixCSC arrays can be allocated), and it starts travelling to the host on the
    # copy stream while the scatter maps and the values are computed; finish() waits for it and fetches the values
    nnz = eng.pattern(form.form_id, qpits[0].rule)
    ass._out = (np.empty(ass.ncol + 1, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.float64))
    eng.fetch_pattern_async(ass._out[0], ass._out[1])
    eng.numeric(form.params())
    ass._assembled = True
    return ass


def assemble_both(am: SysmatAssemblerGPU, av: "SysvecAssemblerGPU", form, vform, elit, qpit):
    """ONE integrate! pass for the matrix and the vector, like the reference's heat loops, which fill ``ke`` and ``fe`` in the
    same quadrature loop and call ``assemble!(am, ke); assemble!(av, fe)`` per element (examples/heat/poisson/t3.jl:41-64):
    ``assemble_both(am, av, HeatForm(kappa), HeatLoadForm(Q), elit, qpit)`` then ``finish(am)``, ``finish(av)``.
    ``av`` must share ``am``'s context (``SysvecAssemblerGPU(0.0, like=am)``).  The fused kernel stages the element load
    vector next to the element matrix (efg_numeric_with_load); F is bit-identical to the separate ``assemble(av, ...)``."""
    if av.engine is not am.engine:
        raise ValueError("assemble_both: the vector assembler must share the matrix assembler's context (like=am)")
    if not (am._started and av._started):
        raise _lib.EfgError(_lib.ERR_STATE, "assemble before start")
    eng = am.engine
    eng.set_option(_lib.OPT_FUSE_LOAD, 1)
    _load_spaces(eng, (elit,))
    eng.start(am.nrow, am.ncol)
    nnz = eng.pattern(form.form_id, qpit.rule)
    am._out = (np.empty(am.ncol + 1, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.float64))
    eng.fetch_pattern_async(am._out[0], am._out[1])
    eng.numeric_with_load(form.params(), vform.params()[0])
    am._assembled = av._assembled = True
    return am, av


def finish(ass):
    if not ass._assembled:
        raise _lib.EfgError(_lib.ERR_STATE, "finish before assemble")
    if isinstance(ass, SysvecAssemblerGPU):   # finish!(av): src/Assemblers.jl:230-232
        return ass.engine.fetch_vec()
    colptr, rowval, nzval = ass._out
    ass.engine.fetch_csc(None, None, nzval)
    return SparseMatrixCSC(ass.nrow, ass.ncol, colptr, rowval, nzval)


def mul(ass: SysmatAssemblerGPU, x):
    """``K * x`` (examples/heat/poisson/t3.jl:78) on the device, without fetching K."""
    if not ass._assembled:
        raise _lib.EfgError(_lib.ERR_STATE, "mul before assemble")
    return ass.engine.spmv(x)


def block(ass: SysmatAssemblerGPU, r0, r1, c0, c1) -> SparseMatrixCSC:
    """``K[r0:r1, c0:c1]`` (1-based inclusive; examples/heat/poisson/t3.jl:79) sliced on the device."""
    if not ass._assembled:
        raise _lib.EfgError(_lib.ERR_STATE, "block before assemble")
    colptr, rowval, nzval = ass.engine.block(r0, r1, c0, c1)
    return SparseMatrixCSC(int(r1) - int(r0) + 1, int(c1) - int(c0) + 1, colptr, rowval, nzval)


def evaluate_error(ass: SysmatAssemblerGPU, elits, qpit, U, truefs):
    """The evaluate_pressure_error / evaluate_velocity_error loops of the Stokes examples
    (examples/stokes/colliding_flow/ht_p2_p1.jl:120-178) on the device.  ``elits``: one FEIterator per field component
    (the same iterator twice for the two components of a vector space); ``truefs``: one vectorised callable f(x, y) per
    component, evaluated on the host at location(el, qp); ``U``: the system vector.  The spaces must be the ones the
    assembler's context holds (the last ``assemble`` call)."""
    eng = ass.engine
    if isinstance(elits, FEIterator):
        elits, truefs = (elits,), (truefs,)
    loaded = getattr(eng, "_keep", [])
    comps, seen = [], {}
    for it in elits:
        slot = [i for i, k in enumerate(loaded) if k.fesp is it.fesp]
        if not slot:
            raise _lib.EfgError(_lib.ERR_STATE, "evaluate_error: this space was not part of the last assemble call")
        comps.append((slot[0], seen.get(slot[0], 0)))
        seen[slot[0]] = seen.get(slot[0], 0) + 1
    mesh = elits[0].fesp.mesh
    mslot = 0 if loaded[0].fesp.mesh is mesh else 1
    loc = eng.qp_locations(mslot, qpit.rule, mesh.conn.shape[0])
    truth = np.stack([np.asarray(f(loc[..., 0], loc[..., 1]), dtype=np.float64) for f in truefs], axis=-1)
    return eng.l2_error(comps, qpit.rule, U, truth)
