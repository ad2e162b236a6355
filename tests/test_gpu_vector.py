"""GPU parity for the callers either side of the matrix path, through the C ABI:
system-vector assembly (SysvecAssembler / LocalVectorAssembler, `fe[j] += N[j]*Q*JxW`), `K*T`, and the
`K[1:nu,1:nu]` / `K[1:nu,nu+1:end]` partition.  Bar: BIT-IDENTICAL to the CPU oracle (these kernels use
individually rounded operations in the reference's accumulation order), in both FP modes."""
import numpy as np
import pytest
import scipy.sparse as sp

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
from test_oracle_golden import _solve

pytestmark = pytest.mark.gpu

VCASES = [("t3_1pt", efg.T3, 37, None), ("t3_3pt", efg.T3, 20, 3), ("t6_3pt", efg.T6, 33, None), ("t6_1pt", efg.T6, 12, 1),
          ("q4_o2", efg.Q4, 41, None), ("q4_o3", efg.Q4, 17, 3), ("q4_o1", efg.Q4, 9, 1)]


@pytest.mark.parametrize("perturb", [False, True], ids=["regular", "jittered"])
@pytest.mark.parametrize("case", VCASES, ids=[c[0] for c in VCASES])
def test_load_vector_bit_identical(oracle, case, perturb):
    _, kind, N, quad = case
    prob = efg.heat_problem(kind, N, perturb, quad=quad)
    fesp = prob.spaces[0]
    Q = -6.0
    want = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], fesp.field.dofnums, Q, prob.ndofs)
    for strict in (0, 1):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_STRICT_FP, strict)
        efg.load_problem(eng, prob)
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [Q], prob.ndofs)
        got = eng.fetch_vec()
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [2 * Q], prob.ndofs)     # cached dof map, new parameter
        got2 = eng.fetch_vec()
        eng.close()
        assert np.array_equal(got, want), f"max |d| = {np.abs(got - want).max()}"
        assert np.array_equal(got2, oracle.assemble_vec_heat(prob.quad, prob.meshes[0], fesp.field.dofnums, 2 * Q, prob.ndofs))


def test_load_vector_errors():
    prob = efg.heat_problem(efg.T3, 6)
    eng = efg.Engine(0)
    with pytest.raises(_lib.EfgError):                       # no mesh yet
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, 1, [1.0], 10)
    efg.load_problem(eng, prob)
    with pytest.raises(_lib.ArgumentError):                  # BoundsError: dof numbers exceed nrow
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, 1, [1.0], prob.ndofs - 1)
    with pytest.raises(_lib.EfgError):                       # unknown vector form / rule
        eng.vec_assemble(99, 1, [1.0], prob.ndofs)
    with pytest.raises(_lib.EfgError):
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, 5, [1.0], prob.ndofs)     # no 5-point triangle rule
    eng.vec_assemble(_lib.VFORM_HEAT_LOAD, 1, [1.0], prob.ndofs)
    assert eng.fetch_vec().shape == (prob.ndofs,)
    eng.close()


def test_load_vector_sharded_rows_concatenate(oracle):
    """Owner-computes: with a column range set on the n x n system only the owned rows are assembled."""
    prob = efg.heat_problem(efg.Q4, 19, perturb=True)
    n = prob.ndofs
    want = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, 3.0, n)
    cuts = [0, n // 3, n // 2, n]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng = efg.Engine(0)
        efg.load_problem(eng, prob, column_range=(a + 1, b))
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [3.0], n)
        parts.append(eng.fetch_vec())
        eng.close()
    assert np.array_equal(np.concatenate(parts), want)


@pytest.mark.parametrize("name", ["heat_t6", "heat_q4", "elast_t6", "stokes_gen"])
def test_spmv_and_blocks_bit_identical(oracle, name):
    prob = {"heat_t6": lambda: efg.heat_problem(efg.T6, 21, True), "heat_q4": lambda: efg.heat_problem(efg.Q4, 33, True), "elast_t6": lambda: efg.elasticity_problem(13, efg.T6, True),
            "stokes_gen": lambda: efg.stokes_problem(9, "gen", True)}[name]()
    n = prob.ndofs
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_STRICT_FP, 1)
    efg.load_problem(eng, prob)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp, rv, nz = eng.fetch_csc()
    y = eng.spmv(x)
    assert np.array_equal(y, oracle.spmv_csc(n, n, cp, rv, nz, x))
    if name.startswith("heat"):     # the heat forms take the symmetric shortcut (columns read as rows): K must equal K' bit for bit
        Kc = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(n, n))
        Kt = Kc.T.tocsc(); Kt.sort_indices()
        assert np.array_equal(Kt.indptr, Kc.indptr) and np.array_equal(Kt.indices, Kc.indices) and np.array_equal(Kt.data, Kc.data)
        eng.set_option(_lib.OPT_STRICT_FP, 0)          # ... in the default FP mode too
        eng.numeric(prob.form.params())
        _, _, nz0 = eng.fetch_csc()
        assert np.array_equal(eng.spmv(x), oracle.spmv_csc(n, n, cp, rv, nz0, x))
        eng.set_option(_lib.OPT_STRICT_FP, 1)
        eng.numeric(prob.form.params())
    # device vectors (torch) take the zero-copy route
    import torch
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    eng.spmv(xd, yd)
    assert np.array_equal(yd.cpu().numpy(), y)
    # partition as in solve!: K[1:nu,1:nu], K[1:nu,nu+1:end]; plus an empty and an interior block
    nu = sum(efg.nunknowns(s) for s in prob.spaces)
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(n, n))
    for (r0, r1, c0, c1) in [(1, nu, 1, nu), (1, nu, nu + 1, n), (5, 4, 1, n), (n // 3, 2 * n // 3, n // 4, n // 2), (1, n, 1, n)]:
        bcp, brv, bnz = eng.block(r0, r1, c0, c1)
        ref = K[r0 - 1:r1, c0 - 1:c1].tocsc()
        ref.sort_indices()
        assert np.array_equal(bcp, ref.indptr + 1) and np.array_equal(brv, ref.indices + 1) and np.array_equal(bnz, ref.data)
    with pytest.raises(_lib.ArgumentError):
        eng.block(1, n + 1, 1, n)
    eng.close()


def test_heat_examples_end_to_end_through_the_mirrored_api():
    """examples/heat/poisson/t3.jl assembleKF + solve! with both assemblers on the GPU: K and F from one shared
    context, KT = K*T and K[1:nu,1:nu] on the device; reproduces test/test_heat.jl:110 (T3 N=4 golden) and the
    Q4 N=100 accuracy check (test/test_heat.jl:214-221)."""
    import scipy.sparse.linalg as spla
    tempf = lambda x, y: 1.0 + x ** 2 + 2.0 * y ** 2
    for kind, N, qkw in ((efg.T3, 4, dict(kind="default")), (efg.Q4, 100, dict(kind="Gauss", order=2))):
        mesh = efg.T3block(1.0, 1.0, N, N) if kind == efg.T3 else efg.Q4block(1.0, 1.0, N, N)
        fesp = efg.FESpace(mesh, efg.FEH1_T3() if kind == efg.T3 else efg.FEH1_Q4())
        for i in efg.boundary_nodes(mesh):
            efg.setebc(fesp, 0, i, 1, tempf(*mesh.xy[i - 1]))
        efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
        n, nu = efg.ndofs(fesp), efg.nunknowns(fesp)
        elit, qpit = efg.FEIterator(fesp), efg.QPIterator(fesp, **qkw)
        am = efg.start(efg.SysmatAssemblerGPU(0.0), n, n)
        av = efg.start(efg.SysvecAssemblerGPU(0.0, like=am), n)
        efg.assemble(am, efg.HeatForm(1.0), elit, qpit)
        efg.assemble(av, efg.HeatLoadForm(-6.0), elit, qpit)
        F = efg.finish(av)
        T = efg.gathersysvec(fesp)
        KT = efg.mul(am, T)                                   # KT = K * T
        Kff = efg.block(am, 1, nu, 1, nu).to_scipy()          # K[1:nu, 1:nu]
        T[:nu] = spla.spsolve(Kff.tocsc(), F[:nu] - KT[:nu])
        if kind == efg.T3:
            ref = [1.1875, 1.3749999999999998, 1.6874999999999998, 1.5624999999999998, 1.7499999999999998,
                   2.0625, 2.1875, 2.375, 2.6875, 1.0, 1.0625, 1.25, 1.5625, 2.0, 1.125, 2.125, 1.5, 2.5,
                   2.125, 3.125, 3.0, 3.0625, 3.25, 3.5625, 4.0]
            assert np.allclose(T, ref, rtol=0, atol=1e-13)
        else:
            assert nu == 9801
            efg.scattersysvec(fesp, T)
            assert np.abs(fesp.field.dofvals[:, 0] - tempf(mesh.xy[:, 0], mesh.xy[:, 1])).mean() <= 1.0e-9


def test_load_vector_and_spmv_larger_mesh(oracle):
    """Beyond one radix-sort tile and several CTAs of the matrix kernel: 180 k T6 elements, default tile size."""
    prob = efg.heat_problem(efg.T6, 300, perturb=True)
    n = prob.ndofs
    eng = efg.Engine(0)
    efg.load_problem(eng, prob)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp, rv, nz = eng.fetch_csc()
    eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], n)
    F = eng.fetch_vec()
    assert np.array_equal(F, oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, -6.0, n))
    x = np.random.default_rng(11).standard_normal(n)
    assert np.array_equal(eng.spmv(x), oracle.spmv_csc(n, n, cp, rv, nz, x))
    # the vector data survives a new numeric phase (same pattern), the row-major view too
    eng.numeric([2.0])
    _, _, nz2 = eng.fetch_csc()
    assert np.array_equal(eng.spmv(x), oracle.spmv_csc(n, n, cp, rv, nz2, x))
    assert np.array_equal(eng.fetch_vec(), F)
    eng.close()


def test_reassembly_reuses_tile_size_and_stays_bit_identical(oracle):
    """efg_set_mesh invalidates the symbolic data; the tile size found by the first (occupancy-driven) search is
    reused, and the result of the second assembly is the same bits."""
    prob = efg.heat_problem(efg.T6, 160)
    eng = efg.Engine(0)
    efg.load_problem(eng, prob)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    a = eng.fetch_csc()
    tiles1 = eng.stat(_lib.STAT_NTILES)
    efg.load_problem(eng, prob)                 # set_mesh / set_space / start again
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    b = eng.fetch_csc()
    assert eng.stat(_lib.STAT_NTILES) == tiles1
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    assert np.array_equal(a[0], ocp) and np.array_equal(a[1], orv)
    assert np.all(np.abs(a[2] - onz) <= 1e-14 + 1e-12 * np.abs(onz))
    eng.close()


# ---- K and F from ONE pass (efg_numeric_with_load): the element load vector rides in the matrix kernel's stage ----------
@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("kind,N,quad", [(efg.T3, 61, 1), (efg.T3, 33, 3), (efg.T6, 47, 3), (efg.T6, 30, 1), (efg.Q4, 53, 2), (efg.Q4, 20, 3)])
def test_fused_matrix_and_load_vector(oracle, kind, N, quad, strict):
    prob = efg.heat_problem(kind, N, True, quad=quad)
    Q = -6.0
    oF = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, Q, prob.ndofs)
    # reference result of the library's own separate paths
    e0 = efg.Engine(0)
    e0.set_option(_lib.OPT_STRICT_FP, strict)
    efg.load_problem(e0, prob)
    e0.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp0, rv0, nz0 = e0.fetch_csc()
    e0.close()
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_STRICT_FP, strict)
    eng.set_option(_lib.OPT_FUSE_LOAD, 1)
    efg.load_problem(eng, prob)
    eng.symbolic(prob.form.form_id, prob.quad)
    eng.numeric_with_load(prob.form.params(), Q)
    cp, rv, nz = eng.fetch_csc()
    F = eng.fetch_vec()
    assert np.array_equal(cp, cp0) and np.array_equal(rv, rv0)
    assert nz.tobytes() == nz0.tobytes()                      # K is the unfused kernel's K bit for bit
    assert F.tobytes() == oF.tobytes()                        # F is the CPU loop's vector bit for bit, in both FP modes
    # the plain numeric kernel still runs on the fused tiling, and a fused call can follow it
    eng.numeric(prob.form.params())
    assert eng.fetch_csc()[2].tobytes() == nz0.tobytes()
    eng.numeric_with_load(prob.form.params(), 2.5)
    assert eng.fetch_vec().tobytes() == oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, 2.5, prob.ndofs).tobytes()
    eng.close()


def test_fused_pass_needs_the_option_and_a_heat_form():
    prob = efg.heat_problem(efg.T3, 9)
    eng = efg.Engine(0)
    efg.load_problem(eng, prob)
    eng.symbolic(prob.form.form_id, prob.quad)
    with pytest.raises(_lib.EfgError):           # EFG_OPT_FUSE_LOAD was not set at symbolic time
        eng.numeric_with_load(prob.form.params(), 1.0)
    eng.close()
    pe = efg.elasticity_problem(7)
    e2 = efg.Engine(0)
    e2.set_option(_lib.OPT_FUSE_LOAD, 1)
    efg.load_problem(e2, pe)
    e2.symbolic(pe.form.form_id, pe.quad)
    with pytest.raises(_lib.EfgError):
        e2.numeric_with_load([1.0], 1.0)
    e2.close()


def test_heat_example_with_one_fused_pass(oracle):
    """examples/heat/poisson/t3.jl end to end: K and F from ONE pass through the mirrored API -> the golden vector of
    test/test_heat.jl:110 is reproduced by the same solve as in test_heat_examples_end_to_end_through_the_mirrored_api."""
    import scipy.sparse.linalg as spl
    prob = efg.heat_problem(efg.T3, 4)
    fesp = prob.spaces[0]
    xy = prob.meshes[0].xy
    tempf = lambda x, y: 1.0 + x ** 2 + 2.0 * y ** 2
    d = fesp.field.isdatum[:, 0]
    fesp.field.dofvals[d, 0] = tempf(xy[d, 0], xy[d, 1])
    elit, qpit = efg.FEIterator(fesp), efg.QPIterator(fesp, kind="default")
    am = efg.SysmatAssemblerGPU(0.0)
    av = efg.SysvecAssemblerGPU(0.0, like=am)
    efg.start(am, prob.ndofs, prob.ndofs); efg.start(av, prob.ndofs)
    efg.assemble_both(am, av, efg.HeatForm(1.0), efg.HeatLoadForm(-6.0), elit, qpit)
    K, F = efg.finish(am).to_scipy(), efg.finish(av)
    T = efg.gathersysvec(fesp)
    nu = efg.nunknowns(fesp)
    KT = efg.mul(am, T)
    T[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), F[:nu] - KT[:nu])
    assert np.abs(T[fesp.field.dofnums[:, 0] - 1] - tempf(xy[:, 0], xy[:, 1])).max() < 1e-13     # nodally exact (test/test_heat.jl:110)
