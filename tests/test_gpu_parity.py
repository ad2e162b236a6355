"""GPU parity: libelfelgpu.so (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): colptr/rowval bit-exact; nzval within 1e-12 relative / 1e-14 absolute.
With EFG_OPT_STRICT_FP the GPU result must additionally be numerically IDENTICAL to the oracle
(== on every value; only the sign of a zero may differ)."""
import numpy as np
import pytest

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-12, 1e-14     # tolerance stated by BASELINE.json north_star


def _cases():
    c = []
    for perturb in (False, True):
        c += [("heat_t3", lambda p=perturb: efg.heat_problem(efg.T3, 37, p)),
              ("heat_t3_3pt", lambda p=perturb: efg.heat_problem(efg.T3, 20, p, quad=3)),
              ("heat_t6", lambda p=perturb: efg.heat_problem(efg.T6, 33, p)),
              ("heat_q4", lambda p=perturb: efg.heat_problem(efg.Q4, 41, p)),
              ("heat_q4_o3", lambda p=perturb: efg.heat_problem(efg.Q4, 17, p, quad=3)),
              ("elast_t6", lambda p=perturb: efg.elasticity_problem(29, efg.T6, p)),
              ("elast_t3", lambda p=perturb: efg.elasticity_problem(31, efg.T3, p)),
              ("elast_q4", lambda p=perturb: efg.elasticity_problem(23, efg.Q4, p)),
              ("stokes_gen", lambda p=perturb: efg.stokes_problem(21, "gen", p)),
              ("stokes_reddy", lambda p=perturb: efg.stokes_problem(19, "reddy", p)),
              ("stokes_veclap_alt", lambda p=perturb: efg.stokes_problem(18, "veclap_alt", p)),
              ("stokes_veclap", lambda p=perturb: efg.stokes_problem(17, "veclap", p))]
    return c


CASES = _cases()
IDS = [f"{n}{'_jit' if i >= len(CASES) // 2 else ''}" for i, (n, _) in enumerate(CASES)]


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
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_parity_vs_oracle(oracle, case, path):
    prob = case[1]()
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    for strict in (1, 0):
        cp, rv, nz = _gpu(prob, path, strict)
        assert np.array_equal(cp, ocp), "colptr not bit-exact"
        assert np.array_equal(rv, orv), "rowval not bit-exact"
        if strict:
            assert np.array_equal(nz, onz), f"strict mode differs from the oracle: max |d| = {np.abs(nz - onz).max()}"
        else:
            assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"max |d| = {np.abs(nz - onz).max()}"


@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
def test_reference_api_heat_t3_n100(oracle, path):
    """BASELINE config 1 through the mirror of the reference's assembler API."""
    prob = efg.heat_problem(efg.T3, 100)
    fesp = prob.spaces[0]
    elit = efg.FEIterator(fesp)
    qpit = efg.QPIterator(fesp, kind="default")
    ass = efg.SysmatAssemblerGPU(0.0).set_option(_lib.OPT_PATH, path)
    efg.start(ass, efg.ndofs(fesp), efg.ndofs(fesp))
    efg.assemble(ass, efg.HeatForm(1.0), elit, qpit)
    K = efg.finish(ass)
    assert K.colptr.dtype == np.int64 and K.rowval.dtype == np.int64 and K.nzval.dtype == np.float64
    assert len(K.nzval) == 7 * 100 ** 2 + 6 * 100 + 1 == 70601
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    assert np.array_equal(K.colptr, ocp) and np.array_equal(K.rowval, orv)
    assert np.all(np.abs(K.nzval - onz) <= ATOL + RTOL * np.abs(onz))


@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
def test_column_range_blocks_concatenate(oracle, path):
    """Owner-computes column blocks (multi-GPU sharding): concatenation == the global matrix."""
    prob = efg.heat_problem(efg.T6, 24, True)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    n = prob.ndofs
    cuts = [0, n // 3, n // 3 + 1, (2 * n) // 3, n]
    cps, rvs, nzs = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_PATH, path)
        efg.load_problem(eng, prob, column_range=(a + 1, b))
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
        cp, rv, nz = eng.fetch_csc()
        eng.close()
        assert len(cp) == b - a + 1 and cp[0] == 1
        cps.append(cp); rvs.append(rv); nzs.append(nz)
    off, full = 0, [np.array([1], dtype=np.int64)]
    for cp in cps:
        full.append(cp[1:] + off)
        off += cp[-1] - 1
    assert np.array_equal(np.concatenate(full), ocp)
    assert np.array_equal(np.concatenate(rvs), orv)
    nz = np.concatenate(nzs)
    assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz))


def test_unnumbered_dofs_raise_argument_error():
    """A space that was never data-numbered holds dof number 0 -> sparse() ArgumentError (src/FEFields.jl:148)."""
    prob = efg.heat_problem(efg.T3, 8)
    efg.numberfreedofs(prob.spaces[0])          # zeroes the data dofs again, no numberdatadofs!
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_PATH, _lib.PATH_TWOPASS)
    efg.load_problem(eng, prob)
    with pytest.raises(efg.ArgumentError):
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    eng.close()


def test_call_order_and_bad_arguments():
    eng = efg.Engine(0)
    with pytest.raises(efg.EfgError):
        eng.symbolic(_lib.FORM_HEAT, 1)          # before start
    prob = efg.heat_problem(efg.T3, 4)
    efg.load_problem(eng, prob)
    with pytest.raises(efg.EfgError):
        eng.assemble(_lib.FORM_HEAT, 5, [1.0])   # no such triangle rule ("Unknown number of integration points", src/RefShapes.jl:227)
    with pytest.raises(efg.EfgError):
        eng.assemble(_lib.FORM_ELASTICITY, 1, [1.0] * 9)   # space has 1 component
    bad = prob.meshes[0].conn.copy(); bad[0, 0] = 10 ** 6
    with pytest.raises(efg.ArgumentError):
        eng.set_mesh(0, efg.T3, bad, prob.meshes[0].xy)
    eng.close()


def test_numeric_phase_is_deterministic_and_reusable():
    prob = efg.elasticity_problem(40, efg.T6, True)
    eng = efg.Engine(0)
    efg.load_problem(eng, prob)
    eng.symbolic(prob.form.form_id, prob.quad)
    eng.numeric(prob.form.params())
    _, _, a = eng.fetch_csc()
    eng.numeric(prob.form.params())
    _, _, b = eng.fetch_csc()
    assert a.tobytes() == b.tobytes()
    eng.numeric(2.0 * prob.form.params())        # linearity in D
    _, _, c = eng.fetch_csc()
    assert np.allclose(c, 2.0 * a, rtol=1e-14, atol=0)
    eng.close()


@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
@pytest.mark.parametrize("workload,n,world", [("heat_t6", 9, 3), ("heat_q4", 12, 2), ("elasticity_t6", 7, 2)])
def test_multirange_band_shards_merge_to_global(oracle, workload, n, world, path):
    """The multi-GPU decomposition on one GPU: every rank's band (a set of column ranges, halo elements
    replicated) is assembled by the engine; the interleaved blocks equal the oracle's global matrix."""
    import torch
    from elfel_jl_b200 import sharding as sh
    kind, conn, xy, dofnums, band, form, quad = sh.build_global(workload, n, world, "cpu")
    nd = dofnums.numel()
    ocp, orv, onz = oracle.assemble(form.form_id, quad, efg.Mesh(kind, conn.numpy(), xy.numpy()), None, [dofnums.numpy()],
                                    form.params(), nd, nd)
    blocks = []
    for r in range(world):
        # shard built on the CPU (same coordinates, bit for bit, as the oracle's mesh), handed over as DEVICE arrays
        s, (firsts, lasts), _ = sh.shard_problem(efg, workload, n, r, world, dev="cpu")
        s.meshes[0].conn, s.meshes[0].xy = s.meshes[0].conn.cuda(), s.meshes[0].xy.cuda()
        s.spaces[0].field.dofnums = s.spaces[0].field.dofnums.cuda()
        torch.cuda.synchronize()    # the ctx copies on its own stream: producer work must be complete
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_PATH, path)
        eng.set_option(_lib.OPT_STRICT_FP, 1)
        eng.set_mesh(0, s.meshes[0].kind, s.meshes[0].conn, s.meshes[0].xy)
        eng.set_space(0, 0, s.spaces[0].field.dofnums)
        eng.start(s.ndofs, s.ndofs)
        eng.set_column_ranges(firsts, lasts)
        eng.assemble(s.form.form_id, s.quad, s.form.params())
        cp, rv, nz = eng.fetch_csc()
        eng.close()
        blocks.append((firsts, lasts, cp, rv, nz))
    cp, rv, nz = sh.merge_blocks(nd, blocks)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
    assert np.array_equal(nz, onz)


# ---------------------------------------------------------------------------------------------------------
# larger meshes: closed-form pattern counts, tiled vs two-pass cross-check, properties
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,make,nnz_formula", [
    ("heat_t3", lambda: efg.heat_problem(efg.T3, 500), lambda N: 7 * N * N + 6 * N + 1),
    ("heat_t6", lambda: efg.heat_problem(efg.T6, 400), lambda N: 46 * N * N + 16 * N + 1),
    ("heat_q4", lambda: efg.heat_problem(efg.Q4, 700), lambda N: (3 * N + 1) ** 2),
    ("elasticity_t6", lambda: efg.elasticity_problem(250, efg.T6), lambda N: 4 * (46 * N * N + 16 * N + 1)),
    ("stokes_gen", lambda: efg.stokes_problem(200, "gen"), lambda N: 260 * N * N + 104 * N + 8),
])
def test_large_mesh_pattern_counts_and_path_agreement(name, make, nnz_formula):
    prob = make()
    N = int(prob.name.split("_N")[1])
    res = {}
    for path in (_lib.PATH_TILED, _lib.PATH_TWOPASS):
        cp, rv, nz = _gpu(prob, path, 1)
        res[path] = (cp, rv, nz)
        assert len(nz) == nnz_formula(N)
        assert cp[0] == 1 and cp[-1] == len(nz) + 1 and np.all(np.diff(cp) >= 0)
    a, b = res[_lib.PATH_TILED], res[_lib.PATH_TWOPASS]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]), "strict mode: the two independent GPU paths must agree bit for bit"
    cp, rv, nz = a
    import scipy.sparse as sp
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(prob.ndofs, prob.ndofs))
    # rows ascending inside every column (sparse()'s CSC invariant)
    d = np.diff(rv)
    col_start = np.zeros(len(rv), dtype=bool); col_start[(cp[1:-1] - 1)[cp[1:-1] - 1 < len(rv)]] = True
    assert np.all((d > 0) | col_start[1:])
    if name.startswith("heat"):
        # constants are in the null space of the (unconstrained) conductivity matrix; it is symmetric bit for bit
        assert np.abs(K @ np.ones(prob.ndofs)).max() < 1e-10
        assert (K - K.T).nnz == 0
    if name == "elasticity_t6":
        # rigid-body translations are in the null space
        ux = np.zeros(prob.ndofs); ux[prob.spaces[0].field.dofnums[:, 0] - 1] = 1.0
        assert np.abs(K @ ux).max() < 1e-10


def test_fast_mode_vs_strict_mode_large():
    """Default (FMA, shared reciprocal) vs strict arithmetic on a 1.3 M-element jittered T6 mesh: inside the parity bar."""
    prob = efg.heat_problem(efg.T6, 800, True)
    _, _, strict = _gpu(prob, _lib.PATH_TILED, 1)
    _, _, fast = _gpu(prob, _lib.PATH_TILED, 0)
    assert np.all(np.abs(fast - strict) <= ATOL + RTOL * np.abs(strict)), np.abs(fast - strict).max()


def test_auto_path_falls_back_for_high_valence_mesh(oracle):
    """A fan of 150 triangles around one node: that node's matrix column has 151 rows, beyond the tiled path's
    per-column limit -> EFG_OPT_PATH=2 reports EFG_ERR_LIMIT, auto mode silently takes the two-pass CUDA path."""
    nf = 150
    ang = 2 * np.pi * np.arange(nf) / nf
    xy = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    conn = np.stack([np.ones(nf, dtype=np.int64), 2 + np.arange(nf), 2 + (np.arange(nf) + 1) % nf], axis=1)
    mesh = efg.Mesh(efg.T3, conn.astype(np.int64), xy)
    fesp = efg.FESpace(mesh, efg.FEH1_T3())
    efg.numberdofs(fesp)
    n = efg.ndofs(fesp)
    ocp, orv, onz = oracle.assemble(oracle.FORM_HEAT, 1, mesh, None, [fesp.field.dofnums], [1.0], n, n)
    for path, expect_fail in ((_lib.PATH_TILED, True), (_lib.PATH_AUTO, False)):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_PATH, path)
        eng.set_option(_lib.OPT_STRICT_FP, 1)
        eng.set_mesh(0, efg.T3, mesh.conn, mesh.xy); eng.set_space(0, 0, fesp.field.dofnums); eng.start(n, n)
        if expect_fail:
            with pytest.raises(efg.EfgError) as ei:
                eng.assemble(_lib.FORM_HEAT, 1, [1.0])
            assert ei.value.code == _lib.ERR_LIMIT
        else:
            eng.assemble(_lib.FORM_HEAT, 1, [1.0])
            assert int(eng.stat(_lib.STAT_PATH)) == _lib.PATH_TWOPASS
            cp, rv, nz = eng.fetch_csc()
            assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
        eng.close()


def test_two_contexts_with_different_rules_interleave(oracle):
    """The quadrature tables / parameters sit in per-device __constant__ memory shared by every ctx of the process, and
    efg_numeric returns without synchronising: two engines on one device with different rules and parameters, their
    numeric calls interleaved with no host sync in between, must both produce their own matrix."""
    pa = efg.heat_problem(efg.T6, 120, True)                 # triangle rule, 3 points, kappa = 1
    pb = efg.heat_problem(efg.Q4, 150, True, kappa=2.5)      # Gauss order 2
    pc = efg.heat_problem(efg.T6, 120, True, kappa=0.5, quad=1)
    want = [oracle.assemble(*efg.oracle_args(p), p.ndofs, p.ndofs) for p in (pa, pb, pc)]
    engs = []
    for p in (pa, pb, pc):
        e = efg.Engine(0)
        efg.load_problem(e, p)
        e.symbolic(p.form.form_id, p.quad)
        engs.append(e)
    for _ in range(6):                                       # a -> b -> c -> a ... no synchronisation in between
        for e, p in zip(engs, (pa, pb, pc)):
            e.numeric(p.form.params())
    for e, p, (ocp, orv, onz) in zip(engs, (pa, pb, pc), want):
        cp, rv, nz = e.fetch_csc()
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
        assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"{p.name}: max |d| = {np.abs(nz - onz).max()}"
        e.close()


def test_pattern_first_then_overlapped_fetch(oracle):
    """efg_pattern -> (caller sizes its arrays) -> efg_fetch_pattern_async -> efg_numeric -> efg_fetch_csc(nzval): the sequence
    of the Julia shim's finish!; same matrix as the one-call sequence, also when repeated and with a column range."""
    prob = efg.elasticity_problem(45, efg.T6, True)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    eng = efg.Engine(0)
    for rep in range(2):
        efg.load_problem(eng, prob)
        nnz = eng.pattern(prob.form.form_id, prob.quad)
        assert nnz == len(orv)
        cp = np.empty(prob.ndofs + 1, dtype=np.int64); rv = np.empty(nnz, dtype=np.int64); nz = np.empty(nnz)
        eng.fetch_pattern_async(cp, rv)
        eng.numeric(prob.form.params())
        eng.fetch_csc(None, None, nz)
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
        assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz))
    # pattern only, fetched synchronously, values never computed
    efg.load_problem(eng, prob, column_range=(101, 4000))
    nnz = eng.pattern(prob.form.form_id, prob.quad)
    cp, rv, _ = eng.fetch_csc(np.empty(3901, dtype=np.int64), np.empty(nnz, dtype=np.int64), None)
    assert np.array_equal(cp, ocp[100:4001] - ocp[100] + 1) and np.array_equal(rv, orv[ocp[100] - 1: ocp[4000] - 1])
    with pytest.raises(efg.EfgError):
        eng.fetch_csc(None, None, np.empty(nnz))             # no values yet
    eng.close()


def test_remesh_without_renumbering_is_rejected():
    """efg_set_mesh with a different node count leaves the spaces stale: the next symbolic phase must say so
    (EFG_ERR_STATE) instead of reading past the dof map."""
    small, big = efg.heat_problem(efg.T3, 6), efg.heat_problem(efg.T3, 9)
    eng = efg.Engine(0)
    efg.load_problem(eng, small)
    eng.assemble(small.form.form_id, small.quad, small.form.params())
    eng.set_mesh(0, efg.T3, big.meshes[0].conn, big.meshes[0].xy)
    eng.start(big.ndofs, big.ndofs)
    with pytest.raises(efg.EfgError) as ei:
        eng.assemble(big.form.form_id, big.quad, big.form.params())
    assert ei.value.code == _lib.ERR_STATE
    with pytest.raises(efg.EfgError):
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, big.quad, [1.0], big.ndofs)
    eng.set_space(0, 0, big.spaces[0].field.dofnums)
    eng.assemble(big.form.form_id, big.quad, big.form.params())
    eng.close()


def test_deferred_coordinate_copy_gives_the_same_matrix(oracle):
    """EFG_OPT_DEFER_XY: efg_set_mesh borrows the host coordinates and the copy overlaps the pattern kernels; any entry point
    that reads coordinates first (here efg_qp_locations) brings them in on the main stream instead."""
    prob = efg.heat_problem(efg.T6, 37, True)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    m = prob.meshes[0]
    conn, xy, dof = (np.ascontiguousarray(m.conn, dtype=np.int64), np.ascontiguousarray(m.xy, dtype=np.float64),
                     np.ascontiguousarray(prob.spaces[0].field.dofnums, dtype=np.int64))      # kept alive below
    for seq in ("pattern-numeric", "assemble", "locations-first"):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_DEFER_XY, 1)
        eng.set_option(_lib.OPT_STRICT_FP, 1)
        eng.set_mesh(0, m.kind, conn, xy)
        eng.set_space(0, 0, dof)
        eng.start(prob.ndofs, prob.ndofs)
        if seq == "locations-first":
            loc = eng.qp_locations(0, prob.quad, m.nel)
            assert loc.tobytes() == oracle.qp_locations(prob.quad, m).tobytes()
        if seq == "pattern-numeric":
            nnz = eng.pattern(prob.form.form_id, prob.quad)
            cp, rv = np.empty(prob.ndofs + 1, dtype=np.int64), np.empty(nnz, dtype=np.int64)
            eng.fetch_pattern_async(cp, rv)
            eng.numeric(prob.form.params())
            nz = np.empty(nnz, dtype=np.float64)
            eng.fetch_csc(None, None, nz)
        else:
            eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
            cp, rv, nz = eng.fetch_csc()
        eng.close()
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz), seq


@pytest.mark.parametrize("host_widen", [0, 1], ids=["device-widened", "host-widened"])
def test_row_indices_reach_the_host_identically_on_both_routes(oracle, host_widen):
    """EFG_OPT_HOST_WIDEN: Int64 row indices widened on the device (many ranks per host) or sent as Int32 and widened in place by
    library threads (one or two GPUs per host) -- the caller's SparseMatrixCSC arrays must come out the same, for the plain
    fetch and for the overlapped pattern fetch, with page-locked and with pageable destinations."""
    import torch
    prob = efg.heat_problem(efg.T6, 61, True)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_HOST_WIDEN, host_widen)
    eng.set_option(_lib.OPT_STRICT_FP, 1)
    efg.load_problem(eng, prob)
    nnz = eng.pattern(prob.form.form_id, prob.quad)
    cp = torch.empty(prob.ndofs + 1, dtype=torch.int64).pin_memory()
    rv = torch.full((nnz,), -7, dtype=torch.int64).pin_memory()
    eng.fetch_pattern_async(cp, rv)
    eng.numeric(prob.form.params())
    nz = np.empty(nnz, dtype=np.float64)
    eng.fetch_csc(None, None, nz)
    assert np.array_equal(cp.numpy(), ocp) and np.array_equal(rv.numpy(), orv) and np.array_equal(nz, onz)
    cp2, rv2, nz2 = eng.fetch_csc()                       # pageable numpy arrays, everything in one call
    assert np.array_equal(cp2, ocp) and np.array_equal(rv2, orv) and np.array_equal(nz2, onz)
    eng.close()


@pytest.mark.parametrize("path", [_lib.PATH_TWOPASS, _lib.PATH_TILED], ids=["twopass", "tiled"])
@pytest.mark.parametrize("kind,N,quad", [(efg.T3, 21, 4), (efg.T3, 17, 6), (efg.T3, 13, 7), (efg.T3, 19, 9), (efg.T3, 11, 12), (efg.T3, 15, 13),
                                          (efg.T6, 14, 4), (efg.T6, 12, 6), (efg.T6, 10, 7), (efg.T6, 9, 12), (efg.T6, 11, 13),
                                          (efg.Q4, 18, 4), (efg.Q4, 13, 5)])
def test_heat_with_the_higher_quadrature_rules(oracle, kind, N, quad, path):
    """The rules beyond the examples' (triangles with 4 ... 13 points, Gauss orders 4 and 5 on the square: src/RefShapes.jl:85-110,
    120-230) run through one instantiation per element kind with the number of points read at run time; matrix and load vector
    against the oracle like the common rules."""
    prob = efg.heat_problem(kind, N, True, quad=quad)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    oF = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, 2.5, prob.ndofs)
    for strict in (1, 0):
        eng = efg.Engine(0)
        eng.set_option(_lib.OPT_PATH, path)
        eng.set_option(_lib.OPT_STRICT_FP, strict)
        efg.load_problem(eng, prob)
        eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
        cp, rv, nz = eng.fetch_csc()
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [2.5], prob.ndofs)
        F = eng.fetch_vec()
        loc = eng.qp_locations(0, prob.quad, prob.meshes[0].nel)
        eng.close()
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
        assert F.tobytes() == oF.tobytes()
        assert loc.tobytes() == oracle.qp_locations(prob.quad, prob.meshes[0]).tobytes()
        if strict:
            assert np.array_equal(nz, onz), f"max |d| = {np.abs(nz - onz).max()}"
        else:
            assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz)), f"max |d| = {np.abs(nz - onz).max()}"
