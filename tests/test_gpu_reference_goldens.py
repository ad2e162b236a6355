"""The reference's own end-to-end golden values, reproduced with the GPU assembler driven through the mirror of the
reference API (SysmatAssemblerGPU / start / assemble / finish): K from the GPU, solve with scipy like the examples'
solve!, then the error norms / solution vector the reference tests assert (test/test_heat.jl:110,
test/test_stokes.jl:550-560,777-780)."""
import numpy as np
import pytest

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
from test_oracle_golden import _run_stokes, _csc, _solve, _oracle_errors, _truep, _trueux, _trueuy

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
    _LAST.update(ass=ass, elits=elits, qpits=qpits)
    return K.colptr, K.rowval, K.nzval


_LAST = {}


def _gpu_errors(oracle, vmesh, pmesh, spaces, Uv, three_spaces):
    """evaluate_pressure_error / evaluate_velocity_error on the device (efg_qp_locations + efg_l2_error), checked
    against the oracle's restatement: locations bit-identical, error norms to 1e-12 relative."""
    ass, elits, qpits = _LAST["ass"], _LAST["elits"], _LAST["qpits"]
    ep = efg.evaluate_error(ass, elits[-1], qpits[-1], Uv, _truep)
    uel = (elits[0], elits[1]) if three_spaces else (elits[0], elits[0])
    ev = efg.evaluate_error(ass, uel, qpits[0], Uv, (_trueux, _trueuy))
    oep, oev = _oracle_errors(oracle, vmesh, pmesh, spaces, Uv, three_spaces)
    assert abs(ep - oep) <= 1e-12 * abs(oep) and abs(ev - oev) <= 1e-12 * abs(oev)
    eng = ass.engine
    assert np.array_equal(eng.qp_locations(0, 3, vmesh.conn.shape[0]), oracle.qp_locations(3, vmesh))
    assert np.array_equal(eng.qp_locations(1, 3, pmesh.conn.shape[0]), oracle.qp_locations(3, pmesh))
    return ep, ev


def test_stokes_reddy_goldens_gpu(oracle):
    ref = [(3.5171450671095306, 0.2968271617227661), (0.5999467323539439, 0.03781189670123018),
           (0.12350320261417459, 0.004741849976722882)]
    for N, r in zip((4, 8, 16), ref):
        ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_REDDY, N, True, assemble=_gpu_assemble, errors=_gpu_errors)
        assert np.allclose([ep, ev], r, rtol=1e-9, atol=0)


def test_stokes_veclap_alt_and_gen_goldens_gpu(oracle):
    ep, ev, _, _, nnz = _run_stokes(oracle, oracle.FORM_STOKES_VECLAP_ALT, 4, False, assemble=_gpu_assemble, errors=_gpu_errors)
    assert np.allclose([ep, ev], [2.596076907594511, 0.3001331486426876], rtol=1e-9, atol=0)
    assert nnz == 260 * 16 + 104 * 4 + 8
    ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_GEN, 4, False, assemble=_gpu_assemble, errors=_gpu_errors)
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


def test_error_integrator_other_elements_and_errors(oracle):
    """Q4 / T3 / T6 scalar fields with their own rules, a jittered mesh, a device-resident U; argument errors."""
    import torch
    rng = np.random.default_rng(5)
    f = lambda x, y: np.sin(3 * x) * np.cos(2 * y)
    for kind, N, quad in ((efg.Q4, 23, None), (efg.T3, 31, 3), (efg.T6, 17, None), (efg.T3, 12, None)):
        prob = efg.heat_problem(kind, N, perturb=True, quad=quad)
        eng = efg.Engine(0)
        efg.load_problem(eng, prob)
        U = rng.standard_normal(prob.ndofs)
        loc = eng.qp_locations(0, prob.quad, prob.nel)
        oloc = oracle.qp_locations(prob.quad, prob.meshes[0])
        assert np.array_equal(loc, oloc)
        truth = f(loc[..., 0], loc[..., 1])[..., None]
        want = oracle.l2_error(prob.quad, prob.meshes[0], [(prob.spaces[0].field.dofnums, 0)], U, truth)
        got = eng.l2_error([(0, 0)], prob.quad, U, truth)
        assert abs(got - want) <= 1e-12 * want
        got_d = eng.l2_error([(0, 0)], prob.quad, torch.from_numpy(U).cuda(), torch.from_numpy(np.ascontiguousarray(truth)).cuda())
        assert got_d == got                                  # same bits: fixed-shape reduction
        with pytest.raises(_lib.ArgumentError):              # U shorter than the dof numbers
            eng.l2_error([(0, 0)], prob.quad, U[: prob.ndofs // 2], truth)
        with pytest.raises(_lib.EfgError):
            eng.l2_error([(0, 1)], prob.quad, U, truth)      # component 1 of a scalar space
        with pytest.raises(_lib.EfgError):
            eng.l2_error([(2, 0)], prob.quad, U, truth)      # space slot never set
        eng.close()
