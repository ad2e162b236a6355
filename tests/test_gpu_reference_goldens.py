"""The reference's own end-to-end golden values, reproduced with the GPU assembler driven through the mirror of the
reference API (SysmatAssemblerGPU / start / assemble / finish): K from the GPU, solve with scipy like the examples'
solve!, then the error norms / solution vector the reference tests assert (test/test_heat.jl:110,
test/test_stokes.jl:550-560,777-780)."""
import numpy as np
import pytest

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
from test_oracle_golden import _run_stokes, _csc, _solve

pytestmark = pytest.mark.gpu

FORMS = {1: efg.HeatForm, 3: efg.StokesGenForm, 4: efg.StokesReddyForm, 5: efg.StokesVeclapAltForm, 6: efg.StokesVeclapForm}


def _gpu_assemble(form_id, spaces, params, tndof):
    form = FORMS[form_id](np.asarray(params).reshape(3, 3).T) if form_id == 3 else FORMS[form_id](float(params[0]))
    elits = tuple(efg.FEIterator(s) for s in spaces)
    qpits = tuple(efg.QPIterator(s, kind="default", npts=3) for s in spaces)
    ass = efg.SysmatAssemblerGPU(0.0)
    efg.start(ass, tndof, tndof)
    efg.assemble(ass, form, elits, qpits)
    K = efg.finish(ass)
    return K.colptr, K.rowval, K.nzval


def test_stokes_reddy_goldens_gpu(oracle):
    ref = [(3.5171450671095306, 0.2968271617227661), (0.5999467323539439, 0.03781189670123018),
           (0.12350320261417459, 0.004741849976722882)]
    for N, r in zip((4, 8, 16), ref):
        ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_REDDY, N, True, assemble=_gpu_assemble)
        assert np.allclose([ep, ev], r, rtol=1e-9, atol=0)


def test_stokes_veclap_alt_and_gen_goldens_gpu(oracle):
    ep, ev, _, _, nnz = _run_stokes(oracle, oracle.FORM_STOKES_VECLAP_ALT, 4, False, assemble=_gpu_assemble)
    assert np.allclose([ep, ev], [2.596076907594511, 0.3001331486426876], rtol=1e-9, atol=0)
    assert nnz == 260 * 16 + 104 * 4 + 8
    ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_GEN, 4, False, assemble=_gpu_assemble)
    assert np.allclose([ep, ev], (3.5171450671095306, 0.2968271617227661), rtol=1e-9, atol=0)


def test_heat_t3_n4_golden_solution_gpu():
    N, kappa, Q = 4, 1.0, -6.0
    tempf = lambda x, y: 1.0 + x ** 2 + 2.0 * y ** 2
    mesh = efg.T3block(1.0, 1.0, N, N)
    fesp = efg.FESpace(mesh, efg.FEH1_T3())
    for i in efg.boundary_nodes(mesh):
        efg.setebc(fesp, 0, i, 1, tempf(*mesh.xy[i - 1]))
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    n = efg.ndofs(fesp)
    ass = efg.start(efg.SysmatAssemblerGPU(0.0), n, n)
    efg.assemble(ass, efg.HeatForm(kappa), efg.FEIterator(fesp), efg.QPIterator(fesp, kind="default"))
    K = efg.finish(ass).to_scipy()
    F = np.zeros(n)
    d = fesp.field.dofnums[:, 0]
    for e in range(mesh.nel):
        nodes = mesh.conn[e] - 1
        x = mesh.xy[nodes]
        J = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        F[d[nodes] - 1] += (1 / 3) * Q * (J * 0.5)
    T = _solve(K, efg.gathersysvec(fesp), F, efg.nunknowns(fesp))
    ref = [1.1875, 1.3749999999999998, 1.6874999999999998, 1.5624999999999998, 1.7499999999999998,
           2.0625, 2.1875, 2.375, 2.6875, 1.0, 1.0625, 1.25, 1.5625, 2.0, 1.125, 2.125, 1.5, 2.5,
           2.125, 3.125, 3.0, 3.0625, 3.25, 3.5625, 4.0]
    assert np.allclose(T, ref, rtol=0, atol=1e-13)
