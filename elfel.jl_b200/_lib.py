"""ctypes binding of include/elfel_gpu.h (libelfelgpu.so) -- the same C ABI a Julia shim would
``ccall``.  No torch types cross this boundary: plain pointers and sizes.

``build()`` compiles the CUDA library in-tree for sm_100a with nvcc (cross-compiles without a GPU).
There is no CPU fallback: ``load()`` raises if the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("EFG_LIB") or os.path.join(_HERE, "libelfelgpu.so")   # EFG_LIB: A/B kernel variants while tuning
HEADER = os.path.join(os.path.dirname(_HERE), "include", "elfel_gpu.h")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wno-deprecated-declarations",
              "-Wno-deprecated-declarations", "-shared"]

# error codes / constants (mirror of the header)
OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_INDEX, ERR_STATE, ERR_LIMIT = 0, -1, -2, -3, -4, -5, -6
FORM_HEAT, FORM_ELASTICITY, FORM_STOKES_GEN, FORM_STOKES_REDDY, FORM_STOKES_VECLAP_ALT, FORM_STOKES_VECLAP = 1, 2, 3, 4, 5, 6
OPT_PATH, OPT_STRICT_FP, OPT_TILE_ELEMS, OPT_SFC_ORDER, OPT_FUSE_LOAD, OPT_DEFER_XY, OPT_HOST_WIDEN = 1, 2, 3, 4, 5, 6, 7
PATH_AUTO, PATH_TWOPASS, PATH_TILED = 0, 1, 2
(STAT_SYMBOLIC_MS, STAT_NUMERIC_MS, STAT_KERNEL_LAUNCHES, STAT_NUMERIC_LAUNCHES, STAT_DEVICE_BYTES,
 STAT_NTILES, STAT_TILE_ELEMS, STAT_NUMERIC_BYTES, STAT_PATH, STAT_VEC_MS, STAT_SPMV_MS) = range(1, 12)
VFORM_HEAT_LOAD = 1
FE_H1, FE_L2, FE_T3_BUBBLE = 0, 1, 7     # efg_set_space_fe

EXPORTS = ["efg_create", "efg_destroy", "efg_last_error", "efg_set_option", "efg_get_stat", "efg_get_stream",
           "efg_synchronize", "efg_set_mesh", "efg_set_space", "efg_start", "efg_set_column_range",
           "efg_set_column_ranges", "efg_pattern", "efg_fetch_pattern_async", "efg_symbolic", "efg_numeric", "efg_numeric_with_load", "efg_assemble", "efg_fetch_csc", "efg_device_csc", "efg_version",
           "efg_vec_assemble", "efg_fetch_vec", "efg_device_vec", "efg_spmv", "efg_block_nnz", "efg_fetch_block",
           "efg_qp_locations", "efg_l2_error",
           "efg_set_space_fe", "efg_set_mesh3", "efg_gen_mesh", "efg_gen_mesh_corners", "efg_gen_space", "efg_setebc_box", "efg_setebc_nodes", "efg_number_dofs",
           "efg_fetch_mesh", "efg_fetch_dofnums",
           "efgm_create", "efgm_destroy", "efgm_last_error", "efgm_device_count", "efgm_set_option", "efgm_set_mesh", "efgm_set_space",
           "efgm_start", "efgm_assemble", "efgm_numeric", "efgm_fetch_csc", "efgm_get_stat", "efgm_device_ctx"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(s) > t for s in _sources() + [HEADER])


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> elfel.jl_b200/libelfelgpu.so"""
    if not force and not needs_build():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("EFG_NVCC_EXTRA", "").split()      # e.g. -DTL_NEUTRAL=1 while A/B-ing kernel variants
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, os.path.join(CSRC, "elfel_gpu.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO_PATH


_lib = None


def load():
    """Load libelfelgpu.so and declare the prototypes of include/elfel_gpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the assembly path)")
    L = C.CDLL(SO_PATH)
    vp, i64, i64p, f64p, ci = C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int
    L.efg_version.restype = C.c_char_p
    L.efg_create.argtypes = [ci, C.POINTER(vp)]
    L.efg_destroy.argtypes = [vp]
    L.efg_last_error.argtypes = [vp]
    L.efg_last_error.restype = C.c_char_p
    L.efg_set_option.argtypes = [vp, ci, i64]
    L.efg_get_stat.argtypes = [vp, ci, f64p]
    L.efg_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.efg_synchronize.argtypes = [vp]
    L.efg_set_mesh.argtypes = [vp, ci, ci, i64, i64, vp, vp]
    L.efg_set_space.argtypes = [vp, ci, ci, ci, i64, vp]
    L.efg_numeric_with_load.argtypes = [vp, f64p, ci, C.c_double]
    L.efg_set_mesh3.argtypes = [vp, ci, ci, i64, i64, vp, vp]
    L.efg_set_space_fe.argtypes = [vp, ci, ci, ci, ci, i64, vp, i64, vp]
    L.efg_start.argtypes = [vp, i64, i64]
    L.efg_set_column_range.argtypes = [vp, i64, i64]
    if hasattr(L, "efg_set_column_ranges"):
        L.efg_set_column_ranges.argtypes = [vp, i64, i64p, i64p]
    L.efg_symbolic.argtypes = [vp, ci, ci, i64p]
    L.efg_pattern.argtypes = [vp, ci, ci, i64p]
    L.efg_fetch_pattern_async.argtypes = [vp, vp, vp]
    L.efg_numeric.argtypes = [vp, f64p, ci]
    L.efg_assemble.argtypes = [vp, ci, ci, f64p, ci, i64p]
    L.efg_fetch_csc.argtypes = [vp, vp, vp, vp]
    L.efg_device_csc.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.efg_vec_assemble.argtypes = [vp, ci, ci, f64p, ci, i64]
    L.efg_fetch_vec.argtypes = [vp, vp]
    L.efg_device_vec.argtypes = [vp, C.POINTER(vp), i64p]
    L.efg_spmv.argtypes = [vp, vp, vp]
    L.efg_block_nnz.argtypes = [vp, i64, i64, i64, i64, i64p]
    L.efg_fetch_block.argtypes = [vp, vp, vp, vp]
    L.efg_qp_locations.argtypes = [vp, ci, ci, vp, i64p]
    L.efg_l2_error.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci), ci, vp, i64, vp, f64p]
    f64 = C.c_double
    L.efg_gen_mesh.argtypes = [vp, ci, ci, i64, i64, f64, f64, f64, f64]
    L.efg_gen_mesh_corners.argtypes = [vp, ci, ci]
    L.efg_gen_space.argtypes = [vp, ci, ci, ci]
    L.efg_setebc_box.argtypes = [vp, ci, ci, f64, f64, f64, f64]
    L.efg_setebc_nodes.argtypes = [vp, ci, ci, i64, vp]
    L.efg_number_dofs.argtypes = [vp, ci, C.POINTER(ci), i64p, i64p]
    L.efg_fetch_mesh.argtypes = [vp, ci, i64p, i64p, vp, vp]
    L.efg_fetch_dofnums.argtypes = [vp, ci, vp]
    L.efgm_create.argtypes = [ci, C.POINTER(ci), C.POINTER(vp)]
    L.efgm_destroy.argtypes = [vp]
    L.efgm_last_error.argtypes = [vp]
    L.efgm_last_error.restype = C.c_char_p
    L.efgm_device_count.argtypes = [vp]
    L.efgm_set_option.argtypes = [vp, ci, i64]
    L.efgm_set_mesh.argtypes = [vp, ci, ci, i64, i64, vp, vp]
    L.efgm_set_space.argtypes = [vp, ci, ci, ci, i64, vp]
    L.efgm_start.argtypes = [vp, i64, i64]
    L.efgm_assemble.argtypes = [vp, ci, ci, f64p, ci, i64p]
    L.efgm_numeric.argtypes = [vp, f64p, ci]
    L.efgm_fetch_csc.argtypes = [vp, vp, vp, vp]
    L.efgm_get_stat.argtypes = [vp, ci, ci, f64p]
    L.efgm_device_ctx.argtypes = [vp, ci, C.POINTER(vp), i64p, C.POINTER(i64p), C.POINTER(i64p)]
    for name in EXPORTS:
        if name not in ("efg_version", "efg_last_error", "efgm_last_error") and hasattr(L, name):
            getattr(L, name).restype = ci
    _lib = L
    return L


class EfgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libelfelgpu error {code}: {msg}")
        self.code = code


class ArgumentError(EfgError, ValueError):
    """EFG_ERR_INDEX: what Julia's sparse() raises for an index < 1 or > m/n."""


def check(ctx, rc):
    if rc != OK:
        L = load()
        msg = L.efg_last_error(ctx).decode() if ctx else "no context (no CUDA device?)"
        raise (ArgumentError if rc == ERR_INDEX else EfgError)(rc, msg)
