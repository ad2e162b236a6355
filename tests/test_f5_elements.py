"""SURVEY 8f row f5 -- FEH1_T3_BUBBLE / FEL2_T3 / FEL2_Q4 (src/FElements.jl:324-355, 394-448): the oracle's and the host
mirror's restatement against the known answers the reference's own tests hold (test/test_felements.jl:86-151,
test/test_fespaces.jl:117-170), and the p1b_p1 / q1_q0 colliding-flow examples solved end to end with the oracle matrix
(examples/stokes/colliding_flow/p1b_p1.jl, q1_q0.jl).  The reference asserts no (ep, ev) values for these two examples
(test/test_stokes.jl:156-312 only runs the bubble case), so entry values are pinned through the Reddy loop they share
with the T6/T3 goldens plus the element tables below; the convergence rates are the additional end-to-end check."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import elfel_jl_b200 as efg

trueux = lambda x, y: 20 * x * y ** 3
trueuy = lambda x, y: 5 * x ** 4 - 5 * y ** 4
truep = lambda x, y: 60 * x ** 2 * y - 20 * y ** 3


def test_bubble_and_l2_tables_match_the_reference_known_answers(oracle):
    # test/test_fespaces.jl:130,146: bfun(FEH1_T3_BUBBLE, [1/3, 1/3])
    want = [0.3333333333333334, 0.3333333333333333, 0.3333333333333333, 0.03703703703703704]
    assert np.allclose(oracle.bfun(oracle.FE_T3_BUBBLE, 1 / 3, 1 / 3), want, rtol=0, atol=1e-16)
    assert np.allclose(efg.bfun(efg.FEH1_T3_BUBBLE(), [1 / 3, 1 / 3]), want, rtol=0, atol=1e-16)
    g = oracle.bfungradpar(oracle.FE_T3_BUBBLE, 0.2, 0.3)
    assert np.array_equal(g[:3], [[-1., -1.], [1., 0.], [0., 1.]])          # the T3 rows, test/test_felements.jl:22-24
    assert np.allclose(g[3], [-0.2 * 0.3 + 0.5 * 0.3, -0.2 * 0.3 + 0.5 * 0.2], rtol=0, atol=1e-17)     # src/FElements.jl:354
    # test/test_felements.jl:105-107,149-151: FEL2_Q4 / FEL2_T3: bfun == [1.0], gradient [0 0]
    assert np.array_equal(oracle.bfun(oracle.FE_L2, 0.25, 0.25), [1.0])
    assert np.array_equal(oracle.bfungradpar(oracle.FE_L2, 0.25, 0.25), [[0.0, 0.0]])
    assert np.array_equal(efg.bfun(efg.FEL2_Q4(), [0.25, 0.25]), [1.0])


def test_host_mirror_of_spaces_with_cell_dofs():
    mesh = efg.T3block(1.0, 1.0, 2, 3)
    # test/test_fespaces.jl:119-128,150-170 (ndofperfeat, ndofsperel, edofmdim / edofbfnum / edofcompnt by the loop of :150-164)
    for ncopies in (1, 2):
        fesp = efg.FESpace(mesh, efg.FEH1_T3_BUBBLE(), ncopies)
        assert fesp.fe.ndofperfeat == (1, 0, 1) and fesp.fe.nbf == 4
        assert efg.ndofsperel(fesp) == 4 * ncopies
        emdim, bfnum, compnt, bfn = [], [], [], 1
        for m, nfeat, nd in ((0, 3, 1), (1, 3, 0), (2, 1, 1)):
            for _ in range(nfeat):
                for _ in range(nd):
                    for j in range(1, ncopies + 1):
                        emdim.append(m); bfnum.append(bfn); compnt.append(j)
                    bfn += 1
        assert list(efg.edofmdim(fesp)) == emdim and list(efg.edofbfnum(fesp)) == bfnum and list(efg.edofcompnt(fesp)) == compnt
        assert fesp.field.dofnums[0, 0] == 0          # :168 dofnum(fesp, 0, 1, 1) == 0 before numbering
    l2 = efg.FESpace(mesh, efg.FEL2_T3(), 1)
    assert l2.field is None and l2.cellfield.nterms == mesh.nel and efg.ndofsperel(l2) == 1       # test_felements.jl:138-141
    # numbering: free dofs of the vertex field, then of the cell field, then the data dofs (src/FESpaces.jl:141-173)
    ux = efg.FESpace(mesh, efg.FEH1_T3_BUBBLE(), 1)
    efg.setebc(ux, 0, 1, 1, 5.0); efg.setebc(ux, 0, 4, 1, 6.0)
    efg.numberdofs([ux, l2])
    nn, ne = mesh.nnodes, mesh.nel
    assert efg.nunknowns(ux) == nn - 2 + ne and efg.ndofs(ux) == nn + ne
    assert sorted(np.concatenate([ux.field.dofnums.ravel(), ux.cellfield.dofnums.ravel(), l2.cellfield.dofnums.ravel()])) == list(range(1, nn + 2 * ne + 1))
    assert ux.cellfield.dofnums[0, 0] == nn - 2 + 1 and l2.cellfield.dofnums[0, 0] == nn - 2 + ne + 1
    assert set(ux.field.dofnums[[0, 3], 0]) == {nn + 2 * ne - 1, nn + 2 * ne}
    U = efg.gathersysvec([ux, l2])
    assert U[ux.field.dofnums[0, 0] - 1] == 5.0 and U[ux.field.dofnums[3, 0] - 1] == 6.0


def _solve(oracle, pair, N):
    prob = efg.stokes_f5_problem(N, pair)
    cp, rv, nz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(prob.ndofs, prob.ndofs))
    xy = prob.meshes[0].xy
    for s_, f in zip(prob.spaces[:2], (trueux, trueuy)):
        d = s_.field.isdatum[:, 0]
        s_.field.dofvals[d, 0] = f(xy[d, 0], xy[d, 1])
    U = efg.gathersysvec(prob.spaces)
    nu = sum(efg.nunknowns(s_) for s_ in prob.spaces)
    KT = K @ U
    U[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), -KT[:nu])
    return prob, K, U


def test_triplet_counts_and_symmetry_of_the_f5_pairs(oracle):
    for pair, nt, nd in (("p1b_p1", 112, 11), ("q1_q0", 80, 9), ("p1_p0", 48, 7)):
        prob = efg.stokes_f5_problem(5, pair, perturb=True)
        row, col, val = oracle.assemble_coo(*efg.oracle_args(prob))
        assert len(row) == nt * prob.nel
        assert row.min() >= 1 and row.max() <= prob.ndofs
        cp, rv, nz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
        K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(prob.ndofs, prob.ndofs))
        assert abs(K - K.T).max() <= 1e-15 * abs(K).max()
        a = oracle.assemble_direct(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
        assert np.array_equal(a[0], cp) and np.array_equal(a[1], rv) and a[2].tobytes() == nz.tobytes()
        # the column of a cell dof holds the rows of its own element only: a bubble velocity couples to all 11 element dofs,
        # an L2 pressure to the 2 x NV velocity dofs (no pressure-pressure block)
        s_ = prob.spaces[0] if pair == "p1b_p1" else prob.spaces[2]
        c = int(s_.cellfield.dofnums[3, 0])
        assert cp[c] - cp[c - 1] == (11 if pair == "p1b_p1" else nd - 1)
    # veclap variant drops the two ux-uy blocks
    prob = efg.stokes_f5_problem(4, "p1b_p1", "veclap")
    assert len(oracle.assemble_coo(*efg.oracle_args(prob))[0]) == (112 - 32) * prob.nel


def test_p1b_p1_colliding_flow_converges_at_the_mini_element_rates(oracle):
    """examples/stokes/colliding_flow/p1b_p1.jl end to end with the oracle matrix: velocity error O(h^2), pressure better
    than O(h^1.5) between N = 8, 16, 32 (the example prints (ep, ev) for N = 4 ... 64; the reference asserts no values)."""
    errs = []
    for N in (8, 16, 32):
        prob, K, U = _solve(oracle, "p1b_p1", N)
        ux, uy, ph = prob.spaces
        m = prob.meshes[0]
        loc = oracle.qp_locations(prob.quad, m)
        ev = oracle.l2_error_fe(prob.quad, m, oracle.FE_T3_BUBBLE,
                                [(ux.field.dofnums, ux.cellfield.dofnums), (uy.field.dofnums, uy.cellfield.dofnums)], U,
                                np.stack([trueux(loc[..., 0], loc[..., 1]), trueuy(loc[..., 0], loc[..., 1])], -1))
        ep = oracle.l2_error(prob.quad, m, [(ph.field.dofnums, 0)], U, truep(loc[..., 0], loc[..., 1])[..., None])
        errs.append((ep, ev))
    for (ep0, ev0), (ep1, ev1) in zip(errs[:-1], errs[1:]):
        assert 3.6 < ev0 / ev1 < 4.4, (ev0, ev1)
        assert ep0 / ep1 > 2.8, (ep0, ep1)


def test_q1_q0_velocity_converges(oracle):
    """examples/stokes/colliding_flow/q1_q0.jl with the oracle matrix: the (unstable-pressure) Q1-Q0 pair still gives
    second-order velocities."""
    ev = []
    for N in (8, 16, 32):
        prob, K, U = _solve(oracle, "q1_q0", N)
        ux, uy, _ = prob.spaces
        m = prob.meshes[0]
        loc = oracle.qp_locations(prob.quad, m)
        ev.append(oracle.l2_error(prob.quad, m, [(ux.field.dofnums, 0), (uy.field.dofnums, 0)], U,
                                  np.stack([trueux(loc[..., 0], loc[..., 1]), trueuy(loc[..., 0], loc[..., 1])], -1)))
    assert 3.5 < ev[0] / ev[1] < 4.5 and 3.5 < ev[1] / ev[2] < 4.5, ev


# ---- FEH1_T4 (3-D): examples/heat/poisson/t4.jl ------------------------------------------------------------------------
def _solve_t4(oracle, N, quad, perturb=False):
    tempf = lambda x, y, z: 1.0 + x ** 2 + 2.0 * y ** 2          # t4.jl:19, Q = -6, kappa = 1
    prob = efg.heat_problem(efg.T4, N, perturb, quad=quad)
    cp, rv, nz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(prob.ndofs, prob.ndofs))
    F = oracle.assemble_vec_heat(prob.quad, prob.meshes[0], prob.spaces[0].field.dofnums, -6.0, prob.ndofs)
    fs, xyz = prob.spaces[0], prob.meshes[0].xy
    d = fs.field.isdatum[:, 0]
    fs.field.dofvals[d, 0] = tempf(xyz[d, 0], xyz[d, 1], xyz[d, 2])
    T = efg.gathersysvec(fs)
    nu = efg.nunknowns(fs)
    KT = K @ T
    T[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), F[:nu] - KT[:nu])
    return prob, K, float(np.abs(T[fs.field.dofnums[:, 0] - 1] - tempf(xyz[:, 0], xyz[:, 1], xyz[:, 2])).mean())


def test_t4_tables_match_the_reference_known_answers(oracle):
    # test/test_felements.jl:62-79: FEH1_T4 has vertex dofs only, bfun at the centroid = 1/4 each, constant gradient rows
    fe = efg.FEH1_T4()
    assert fe.ndofperfeat == (1, 0, 0) and fe.nbf == 4
    assert np.allclose(efg.bfun(fe, [0.25, 0.25, 0.25]), [0.25] * 4, rtol=0, atol=1e-16)
    import ctypes as C
    L = oracle.lib()
    pc, w = (C.c_double * 15)(), (C.c_double * 5)()
    for npts in (1, 4, 5):          # src/RefShapes.jl:232-259: every rule integrates a constant over the unit tetrahedron
        assert L.efo_quadrature_t4(npts, pc, w) == npts
        assert abs(sum(w[:npts]) - 1.0 / 6.0) < 1e-15
    # test/test_refshapes.jl:94-115 (mrs6): the tetrahedron rules' points and weights
    L.efo_quadrature_t4(1, pc, w)
    assert np.allclose(pc[:3], [0.25, 0.25, 0.25]) and np.isclose(w[0], 0.16666666666666666)
    L.efo_quadrature_t4(4, pc, w)
    assert np.allclose(np.array(pc[:12]).reshape(4, 3), [[0.1381966, 0.1381966, 0.1381966], [0.5854102, 0.1381966, 0.1381966],
                                                         [0.1381966, 0.5854102, 0.1381966], [0.1381966, 0.1381966, 0.5854102]])
    assert np.allclose(w[:4], [0.041666666666666664] * 4)
    L.efo_quadrature_t4(5, pc, w)
    a6 = 0.16666666666666666
    assert np.allclose(np.array(pc[:15]).reshape(5, 3), [[0.25, 0.25, 0.25], [0.5, a6, a6], [a6, 0.5, a6], [a6, a6, 0.5], [a6, a6, a6]])
    assert np.allclose(w[:5], [-0.13333333333333333, 0.075, 0.075, 0.075, 0.075])
    N = (C.c_double * 4)()
    L.efo_bfun_t4.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)]
    L.efo_bfun_t4(0.25, 0.25, 0.25, N)
    assert list(N) == [0.25] * 4


def test_t4_heat_example_meets_the_reference_correctness_bar(oracle):
    """examples/heat/poisson/t4.jl:78-87: mean nodal |T - tempf| <= 1e-9 (checkcorrectness) -- with the oracle's K and F on
    the Kuhn-split block, for the default 1-point rule and the 4- and 5-point rules."""
    for N, quad in ((4, 1), (7, 1), (5, 4), (5, 5)):
        prob, K, err = _solve_t4(oracle, N, quad)
        assert prob.nel == 6 * N ** 3 and prob.ndofs == (N + 1) ** 3
        assert abs(K - K.T).max() == 0.0                      # bitwise symmetric like the 2-D heat matrices
        assert err <= 1e-9, (N, quad, err)


def test_t4_patch_test_on_a_distorted_mesh(oracle):
    """A linear temperature field is reproduced to rounding on a jittered tetrahedral mesh (Q = 0): exercises the full 3x3
    Jacobian / cofactor solve of the restatement, which the axis-aligned block does not."""
    lin = lambda x, y, z: 1 + 2 * x - 3 * y + 0.5 * z
    prob = efg.heat_problem(efg.T4, 5, True, quad=4)
    cp, rv, nz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    K = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(prob.ndofs, prob.ndofs))
    fs, xyz = prob.spaces[0], prob.meshes[0].xy
    d = fs.field.isdatum[:, 0]
    fs.field.dofvals[d, 0] = lin(xyz[d, 0], xyz[d, 1], xyz[d, 2])
    T = efg.gathersysvec(fs)
    nu = efg.nunknowns(fs)
    T[:nu] = spl.spsolve(K[:nu, :nu].tocsc(), -(K @ T)[:nu])
    assert np.abs(T[fs.field.dofnums[:, 0] - 1] - lin(xyz[:, 0], xyz[:, 1], xyz[:, 2])).max() < 1e-13


def test_fel2_t4_tables_and_space():
    # test/test_felements.jl:117-135 (mfes4): one dof on the cell, none on vertices / edges / faces; bfun == [1.0]
    fe = efg.FEL2_T4()
    assert fe.ndofperfeat == (0, 0, 1) and fe.nbf == 1 and fe.nen == 4
    assert np.array_equal(efg.bfun(fe, [0.25, 0.25, 0.25]), [1.0])
    # a space on a tetrahedral block: one term per cell, numbered like every other field (src/FESpaces.jl:67-88)
    mesh = efg.T4block(1.0, 1.0, 1.0, 2, 2, 3)
    fesp = efg.FESpace(mesh, fe, 1)
    assert fesp.field is None and fesp.cellfield.nterms == mesh.nel == 6 * 2 * 2 * 3
    assert efg.ndofsperel(fesp) == 1
    efg.numberdofs([fesp])
    assert efg.ndofs(fesp) == mesh.nel and efg.nunknowns(fesp) == mesh.nel
    assert np.array_equal(np.sort(fesp.cellfield.dofnums[:, 0]), np.arange(1, mesh.nel + 1))
