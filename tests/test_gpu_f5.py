"""SURVEY 8f row f5 on the GPU: the Reddy / veclap loops on FEH1_T3_BUBBLE / FEH1_T3 (p1b_p1.jl), FEH1_Q4 / FEL2_Q4
(q1_q0.jl) and FEH1_T3 / FEL2_T3 through efg_set_space_fe -- pattern bit-exact, values 1e-12 / 1e-14, strict mode ==
the oracle; the p1b_p1 example end to end through the mirrored assembler API with the error norms on the device."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-12, 1e-14


def _gpu(prob, path, strict):
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_PATH, path)
    eng.set_option(_lib.OPT_STRICT_FP, strict)
    efg.load_problem(eng, prob)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    out = eng.fetch_csc()
    assert int(eng.stat(_lib.STAT_PATH)) == path
    eng.close()
    return out


@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
@pytest.mark.parametrize("perturb", [False, True], ids=["regular", "jittered"])
@pytest.mark.parametrize("formulation", ["reddy", "veclap"])
@pytest.mark.parametrize("pair,N", [("p1b_p1", 23), ("q1_q0", 27), ("p1_p0", 19)])
def test_f5_pairs_vs_oracle(oracle, pair, N, formulation, perturb, path):
    prob = efg.stokes_f5_problem(N, pair, formulation, perturb)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    for strict in (1, 0):
        cp, rv, nz = _gpu(prob, path, strict)
        assert np.array_equal(cp, ocp), "colptr not bit-exact"
        assert np.array_equal(rv, orv), "rowval not bit-exact"
        if strict:
            assert np.array_equal(nz, onz), f"strict mode differs from the oracle: max |d| = {np.abs(nz - onz).max()}"
        else:
            assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"max |d| = {np.abs(nz - onz).max()}"


@pytest.mark.parametrize("pair,N", [("p1b_p1", 257), ("q1_q0", 300)])
def test_f5_pairs_larger_mesh(oracle, pair, N):
    prob = efg.stokes_f5_problem(N, pair, "reddy", True)
    pairs = []
    for s in prob.spaces:
        if s.field is not None:
            pairs.append((prob.meshes[0].conn, s.field.dofnums))
        if s.cellfield is not None:
            pairs.append((None, s.cellfield.dofnums))
    ocp, orv, onz = oracle.assemble_direct_parallel(efg.oracle_args(prob), prob.ndofs, prob.ndofs, pairs)
    cp, rv, nz = _gpu(prob, _lib.PATH_TILED, 0)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
    assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"max |d| = {np.abs(nz - onz).max()}"


def test_spaces_with_cell_dofs_are_checked():
    prob = efg.stokes_f5_problem(6, "p1b_p1")
    eng = efg.Engine(0)
    efg.load_problem(eng, prob)
    m = prob.meshes[0]
    ux = prob.spaces[0]
    with pytest.raises(_lib.EfgError):          # cell field of the wrong length
        eng.set_space(0, 0, ux.field.dofnums, _lib.FE_T3_BUBBLE, ux.cellfield.dofnums[:-1])
    with pytest.raises(_lib.EfgError):          # a bubble space on a Q4 mesh
        e2 = efg.Engine(0)
        q = efg.Q4block(1.0, 1.0, 3, 3)
        e2.set_mesh(0, q.kind, q.conn, q.xy)
        e2.set_space(0, 0, np.ones((q.nnodes, 1), dtype=np.int64), _lib.FE_T3_BUBBLE, np.ones((q.nel, 1), dtype=np.int64))
    # mixing a bubble ux with a plain uy is not a known pair
    efg.load_problem(eng, prob)
    eng.set_space(1, 0, prob.spaces[1].field.dofnums)
    with pytest.raises(_lib.EfgError):
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    # the heat form refuses a space without vertex dofs
    p0 = efg.stokes_f5_problem(5, "q1_q0")
    e3 = efg.Engine(0)
    efg.load_problem(e3, p0)
    e3.set_space(0, 0, None, _lib.FE_L2, p0.spaces[2].cellfield.dofnums)
    with pytest.raises(_lib.EfgError):
        e3.assemble(_lib.FORM_HEAT, 2, [1.0])
    eng.close(); e3.close()


def test_p1b_p1_example_end_to_end_on_the_device(oracle):
    """examples/stokes/colliding_flow/p1b_p1.jl through the mirrored API: assemble on the GPU, solve on the host, both error
    norms on the GPU (the velocity norm sums the bubble function too) == the oracle's loops to 1e-12."""
    trueux = lambda x, y: 20 * x * y ** 3
    trueuy = lambda x, y: 5 * x ** 4 - 5 * y ** 4
    truep = lambda x, y: 60 * x ** 2 * y - 20 * y ** 3
    N = 16
    prob = efg.stokes_f5_problem(N, "p1b_p1")
    ux, uy, ph = prob.spaces
    xy = prob.meshes[0].xy
    for s_, f in ((ux, trueux), (uy, trueuy)):
        d = s_.field.isdatum[:, 0]
        s_.field.dofvals[d, 0] = f(xy[d, 0], xy[d, 1])
    elits = (efg.FEIterator(ux), efg.FEIterator(uy), efg.FEIterator(ph))
    qpits = tuple(efg.QPIterator(s_, npts=3) for s_ in prob.spaces)
    ass = efg.SysmatAssemblerGPU(0.0)
    efg.start(ass, prob.ndofs, prob.ndofs)
    efg.assemble(ass, efg.StokesReddyForm(1.0), elits, qpits)
    K = efg.finish(ass).to_scipy()
    U = efg.gathersysvec(prob.spaces)
    nu = sum(efg.nunknowns(s_) for s_ in prob.spaces)
    KT = K @ U
    U[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), -KT[:nu])
    ev = efg.evaluate_error(ass, (elits[0], elits[1]), qpits[0], U, (trueux, trueuy))
    ep = efg.evaluate_error(ass, elits[2], qpits[2], U, truep)
    m = prob.meshes[0]
    loc = oracle.qp_locations(3, m)
    ev0 = oracle.l2_error_fe(3, m, oracle.FE_T3_BUBBLE, [(ux.field.dofnums, ux.cellfield.dofnums), (uy.field.dofnums, uy.cellfield.dofnums)], U,
                             np.stack([trueux(loc[..., 0], loc[..., 1]), trueuy(loc[..., 0], loc[..., 1])], -1))
    ep0 = oracle.l2_error(3, m, [(ph.field.dofnums, 0)], U, truep(loc[..., 0], loc[..., 1])[..., None])
    assert abs(ev - ev0) <= 1e-12 * ev0 and abs(ep - ep0) <= 1e-12 * ep0
    assert ev < 0.25 and ep < 6.0          # N = 16: (5.33, 0.220) with the oracle matrix (tests/test_f5_elements.py)


# ---- FEH1_T4 (3-D): examples/heat/poisson/t4.jl, tiled and two-pass paths -----------------------------------
@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
@pytest.mark.parametrize("quad", [1, 4, 5])
@pytest.mark.parametrize("perturb", [False, True], ids=["regular", "jittered"])
def test_t4_heat_vs_oracle(oracle, quad, perturb, path):
    prob = efg.heat_problem(efg.T4, 9, perturb, quad=quad)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    oF = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, -6.0, prob.ndofs)
    for strict in (1, 0):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_STRICT_FP, strict)
        eng.set_option(_lib.OPT_PATH, path)
        efg.load_problem(eng, prob)
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
        cp, rv, nz = eng.fetch_csc()
        assert int(eng.stat(_lib.STAT_PATH)) == path
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], prob.ndofs)
        F = eng.fetch_vec()
        eng.close()
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
        assert F.tobytes() == oF.tobytes()                     # the load vector is bit-identical in both modes
        if strict:
            assert np.array_equal(nz, onz), f"max |d| = {np.abs(nz - onz).max()}"
        else:
            assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"max |d| = {np.abs(nz - onz).max()}"


def test_t4_example_end_to_end_through_the_mirrored_api():
    """examples/heat/poisson/t4.jl: K and F on the GPU from one shared context, solve on the host, checkcorrectness bar 1e-9."""
    tempf = lambda x, y, z: 1.0 + x ** 2 + 2.0 * y ** 2
    N = 12
    prob = efg.heat_problem(efg.T4, N)
    fesp, xyz = prob.spaces[0], prob.meshes[0].xy
    d = fesp.field.isdatum[:, 0]
    fesp.field.dofvals[d, 0] = tempf(xyz[d, 0], xyz[d, 1], xyz[d, 2])
    elit, qpit = efg.FEIterator(fesp), efg.QPIterator(fesp, kind="default")
    am = efg.SysmatAssemblerGPU(0.0)
    av = efg.SysvecAssemblerGPU(0.0, like=am)
    efg.start(am, prob.ndofs, prob.ndofs); efg.start(av, prob.ndofs)
    efg.assemble(am, efg.HeatForm(1.0), elit, qpit)
    efg.assemble(av, efg.HeatLoadForm(-6.0), elit, qpit)
    K, F = efg.finish(am).to_scipy(), efg.finish(av)
    T = efg.gathersysvec(fesp)
    nu = efg.nunknowns(fesp)
    KT = efg.mul(am, T)
    T[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), F[:nu] - KT[:nu])
    err = np.abs(T[fesp.field.dofnums[:, 0] - 1] - tempf(xyz[:, 0], xyz[:, 1], xyz[:, 2])).mean()
    assert err <= 1e-9, err


def test_t4_larger_mesh_tiled_equals_twopass(oracle):
    """40^3 cells = 384 000 tetrahedra (jittered): the tiled kernel (3-D Morton tiles, geometry blocks with a z plane) against the
    two-pass path bit for bit in strict mode, and against the direct oracle."""
    prob = efg.heat_problem(efg.T4, 40, True, quad=4)
    out = {}
    for path in (_lib.PATH_TWOPASS, _lib.PATH_TILED):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_STRICT_FP, 1)
        eng.set_option(_lib.OPT_PATH, path)
        efg.load_problem(eng, prob)
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
        out[path] = eng.fetch_csc()
        assert int(eng.stat(_lib.STAT_PATH)) == path
        eng.close()
    a, b = out[_lib.PATH_TWOPASS], out[_lib.PATH_TILED]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2].tobytes() == b[2].tobytes()
    ocp, orv, onz = oracle.assemble_direct_parallel(efg.oracle_args(prob), prob.ndofs, prob.ndofs, [(prob.meshes[0].conn, prob.spaces[0].field.dofnums)])
    assert np.array_equal(b[0], ocp) and np.array_equal(b[1], orv) and np.array_equal(b[2], onz)
