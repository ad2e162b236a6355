"""elfel.jl_b200 -- B200-native engine for Elfel.jl's one data-parallel hot path: the
per-element quadrature loop + global scatter into the sparse system matrix
(FEIterator/QPIterator -> LocalMatrixAssembler -> SysmatAssemblerSparse start!/assemble!/finish!).

The directory name contains a dot, so it is loaded through the root-level shim
``elfel_jl_b200.py`` (``import elfel_jl_b200``).

Contents: ``csrc/`` (hand-written sm_100a CUDA kernels + the C-ABI library ``libelfelgpu.so``),
``_lib.py`` (ctypes binding of include/elfel_gpu.h), and a thin host-side mirror of the
reference interface for this path (meshes, FESpace numbering, iterators, assembler).
There is NO CPU fallback: every assembly call goes through the CUDA library and fails
loudly if it is missing.
"""
from .meshes import (Mesh, T3, Q4, T6, T4, T4block, T3block, Q4block, T6block, T6block_fast, T3toT6, T6toT3,
                     transform, boundary_nodes, vselect, jitter)
from .fespaces import (FE, FEH1_T3, FEH1_T6, FEH1_Q4, FEH1_T3_BUBBLE, FEH1_T4, FEL2_T3, FEL2_Q4, FEL2_T4, bfun, edofmdim, FEField, FESpace, edofbfnum, edofcompnt,
                       ndofsperel, setebc, numberfreedofs, numberdatadofs, numberdofs, nunknowns,
                       ndofs, highestfreedofnum, highestdatadofnum, gathersysvec, scattersysvec)
from . import _lib
from ._lib import build, EfgError, ArgumentError
from .assemblers import (FEIterator, QPIterator, HeatForm, HeatLoadForm, SysvecAssemblerGPU, mul, block, evaluate_error, ElasticityForm, StokesGenForm, StokesReddyForm,
                         StokesVeclapAltForm, StokesVeclapForm, SparseMatrixCSC, Engine, MultiEngine, SysmatAssemblerGPU,
                         start, assemble, assemble_both, finish)
from .problems import (Problem, heat_problem, elasticity_problem, stokes_problem, stokes_f5_problem, plane_stress_D, load_problem,
                       oracle_args)
