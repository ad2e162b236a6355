"""Pins the CPU oracle (oracle/elfel_oracle.c) and the host mirror against every golden
vector / known answer the reference's own tests hold for the assembly path.
All reference values below are transcribed from /root/reference/test/*.jl (file:line cited);
nothing is read from /root/reference at run time."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import elfel_jl_b200 as efg


def _csc(colptr, rowval, nzval, n):
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(n, n))


def test_assembler_known_answer_7x7(oracle):
    # test/test_assemblers.jl:6-34 (scalar assemble!), :47-82 (LocalMatrixAssembler)
    m1 = np.array([[0.24406, 0.599773, 0.833404, 0.0420141],
                   [0.786024, 0.00206713, 0.995379, 0.780298],
                   [0.845816, 0.198459, 0.355149, 0.224996]])
    gi1, gj1 = [1, 7, 5], [5, 2, 1, 4]
    m2 = np.array([[0.146618, 0.53471, 0.614342, 0.737833],
                   [0.479719, 0.41354, 0.00760941, 0.836455],
                   [0.254868, 0.476189, 0.460794, 0.00919633],
                   [0.159064, 0.261821, 0.317078, 0.77646],
                   [0.643538, 0.429817, 0.59788, 0.958909]])
    gi2, gj2 = [2, 3, 1, 7, 5], [6, 7, 3, 4]
    row, col, val = [], [], []
    for m, gi, gj in ((m1, gi1, gj1), (m2, gi2, gj2)):
        for j in range(m.shape[1]):
            for i in range(m.shape[0]):
                row.append(gi[i]); col.append(gj[j]); val.append(m[i, j])
    colptr, rowval, nzval = oracle.sparse(np.array(row), np.array(col), np.array(val), 7, 7)
    A = _csc(colptr, rowval, nzval, 7).toarray()
    B = np.array([[0.833404, 0.599773, 0.460794, 0.0512104, 0.24406, 0.254868, 0.476189],
                  [0, 0, 0.614342, 0.737833, 0, 0.146618, 0.53471],
                  [0, 0, 0.00760941, 0.836455, 0, 0.479719, 0.41354],
                  [0, 0, 0, 0, 0, 0, 0],
                  [0.355149, 0.198459, 0.59788, 1.1839, 0.845816, 0.643538, 0.429817],
                  [0, 0, 0, 0, 0, 0, 0],
                  [0.995379, 0.00206713, 0.317078, 1.55676, 0.786024, 0.159064, 0.261821]])
    assert np.linalg.norm(A - B) / np.linalg.norm(B) < 1.0e-5
    # documented sparse() behaviour: 1-based Int64, rows ascending inside a column
    assert colptr[0] == 1 and colptr[-1] == len(rowval) + 1
    for j in range(7):
        r = rowval[colptr[j] - 1: colptr[j + 1] - 1]
        assert np.all(np.diff(r) > 0)


def test_sparse_keeps_explicit_zeros_and_folds_left_to_right(oracle):
    row = np.array([2, 1, 2, 2, 3]); col = np.array([1, 1, 1, 1, 3])
    val = np.array([1e16, 0.0, 1.0, -1e16, 0.0])
    colptr, rowval, nzval = oracle.sparse(row, col, val, 3, 3)
    assert colptr.tolist() == [1, 3, 3, 4]
    assert rowval.tolist() == [1, 2, 3]
    assert nzval.tolist() == [0.0, (1e16 + 1.0) + -1e16, 0.0]   # left-to-right: 0.0, not 1.0
    with pytest.raises(ValueError):
        oracle.sparse(np.array([0]), np.array([1]), np.array([1.0]), 3, 3)  # dof number 0


def test_refshapes_known_answers(oracle):
    # test/test_refshapes.jl:9-12 (Gauss order 3 via square rule), :26-29, :45-48, :64-67
    pc, w = oracle.quadrature(efg.T3, 3)
    assert np.allclose(pc, [[2 / 3, 1 / 6], [1 / 6, 2 / 3], [1 / 6, 1 / 6]], rtol=0, atol=1e-15)
    assert np.all(w == 0.16666666666666666)       # test/test_qpiterators.jl:58-61, exact
    pc, w = oracle.quadrature(efg.T3, 1)
    assert np.allclose(pc, [[1 / 3, 1 / 3]]) and w[0] == 0.5
    pc, w = oracle.quadrature(efg.Q4, 2)
    assert len(w) == 4 and np.all(w == 1.0)
    assert np.allclose(pc[2], [0.577350269189626, -0.577350269189626])  # point #3: i outer, j inner
    pc3, w3 = oracle.quadrature(efg.Q4, 3)
    assert len(w3) == 9
    assert np.allclose(sorted(set(np.round(pc3[:, 0], 12))), [-0.774596669241483, 0.0, 0.774596669241483])
    assert np.isclose(w3.sum(), 4.0)


def test_qpiterator_t3_tables(oracle):
    # test/test_qpiterators.jl:21-56; test/test_felements.jl:8-49 (T3 at centroid)
    pc, _ = oracle.quadrature(efg.T3, 1)
    assert np.allclose(oracle.bfun(efg.T3, *pc[0]), [0.3333333333333334, 0.3333333333333333, 0.3333333333333333], rtol=0, atol=1e-15)
    pc, _ = oracle.quadrature(efg.T3, 3)
    refN = [[0.1666666666666667, 0.6666666666666666, 0.16666666666666666],
            [0.16666666666666674, 0.16666666666666666, 0.6666666666666666],
            [0.6666666666666667, 0.16666666666666666, 0.16666666666666666]]
    for q in range(3):
        assert np.allclose(oracle.bfun(efg.T3, *pc[q]), refN[q], rtol=0, atol=1e-15)
        assert np.array_equal(oracle.bfungradpar(efg.T3, *pc[q]), [[-1, -1], [1, 0], [0, 1]])
    # partition of unity / zero-sum gradients for the T6 and Q4 tables (unpinned by the reference)
    for kind, rule in ((efg.T6, 3), (efg.Q4, 2)):
        pc, _ = oracle.quadrature(kind, rule)
        for q in range(len(pc)):
            assert np.isclose(oracle.bfun(kind, *pc[q]).sum(), 1.0)
            assert np.allclose(oracle.bfungradpar(kind, *pc[q]).sum(axis=0), 0.0, atol=1e-15)


def _qmesh():
    # test/qmesh.mesh == T3block(1, 1, 2, 3) orientation :a (test/qmesh-conn.dat, qmesh-xyz.dat)
    return efg.T3block(1.0, 1.0, 2, 3)


def test_qmesh_fixture_numbering():
    m = _qmesh()
    conn = [[1, 2, 5], [1, 5, 4], [4, 5, 8], [4, 8, 7], [7, 8, 11], [7, 11, 10],
            [2, 3, 6], [2, 6, 5], [5, 6, 9], [5, 9, 8], [8, 9, 12], [8, 12, 11]]
    assert m.conn.tolist() == conn                      # test/qmesh-conn.dat
    assert np.allclose(m.xy[4], [0.5, 0.3333333333333333])   # test/qmesh-xyz.dat
    mb = efg.T3block(2.0, 5.0, 2, 6, orientation="b")   # test/mt3gen3-conn.dat (first rows)
    assert mb.conn[:4].tolist() == [[1, 2, 4], [2, 5, 4], [4, 5, 7], [5, 8, 7]]


def test_dof_maps_with_ebc():
    # test/test_feiterators.jl:46-64 -- free first in node order, then data
    fesp = efg.FESpace(_qmesh(), efg.FEH1_T3())
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    assert efg.ndofs(fesp) == 12 and efg.nunknowns(fesp) == 12
    for i in (1, 4, 7, 10):
        efg.setebc(fesp, 0, i, 1, 0.0)
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    assert efg.ndofs(fesp) == 12 and efg.nunknowns(fesp) == 8
    refd = [9, 1, 2, 10, 3, 4, 11, 5, 6, 12, 7, 8]
    assert fesp.field.dofnums.ravel().tolist() == refd


def test_three_component_dof_lists():
    # test/test_feiterators.jl:136-172 -- component interleave, node-major element dof order
    fesp = efg.FESpace(_qmesh(), efg.FEH1_T3(), 3)
    for i in (1, 4, 7, 10):
        efg.setebc(fesp, 0, i, 1, 0.0)
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    assert efg.edofcompnt(fesp).tolist() == [1, 2, 3, 1, 2, 3, 1, 2, 3]
    ref = [[33, 1, 2, 3, 4, 5, 11, 12, 13], [33, 1, 2, 11, 12, 13, 34, 9, 10],
           [34, 9, 10, 11, 12, 13, 19, 20, 21], [34, 9, 10, 19, 20, 21, 35, 17, 18],
           [35, 17, 18, 19, 20, 21, 27, 28, 29], [35, 17, 18, 27, 28, 29, 36, 25, 26],
           [3, 4, 5, 6, 7, 8, 14, 15, 16], [3, 4, 5, 14, 15, 16, 11, 12, 13],
           [11, 12, 13, 14, 15, 16, 22, 23, 24], [11, 12, 13, 22, 23, 24, 19, 20, 21],
           [19, 20, 21, 22, 23, 24, 30, 31, 32], [19, 20, 21, 30, 31, 32, 27, 28, 29]]
    d = fesp.field.dofnums
    got = [d[fesp.mesh.conn[e] - 1].ravel().tolist() for e in range(12)]
    assert got == ref


def test_scatter_through_real_dof_maps(oracle):
    # test/test_feiterators.jl:85-120: S (sparse assembled) == D (dense accumulation)
    A = np.array([[0.6744582963441466, 0.2853043149927861, 0.27460710155821255],
                  [0.3781923479141225, 0.2838873430062512, 0.6316949656630075],
                  [0.19369805365903336, 0.8926164783344779, 0.07006905962860177]])
    fesp = efg.FESpace(_qmesh(), efg.FEH1_T3())
    for i in (1, 4, 7, 10):
        efg.setebc(fesp, 0, i, 1, 0.0)
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    row, col, val = [], [], []
    D = np.zeros((12, 12))
    for e in range(12):
        dofs = fesp.field.dofnums[fesp.mesh.conn[e] - 1, 0]
        for j in range(3):
            for i in range(3):
                row.append(dofs[i]); col.append(dofs[j]); val.append(A[i, j])
                D[dofs[i] - 1, dofs[j] - 1] += A[i, j]
    colptr, rowval, nzval = oracle.sparse(np.array(row), np.array(col), np.array(val), 12, 12)
    assert np.allclose(_csc(colptr, rowval, nzval, 12).toarray(), D, rtol=0, atol=1e-15)


def test_sysvec_assembler_semantics(oracle):
    # test/test_assemblers.jl:143-162 (mvass1): val[i] += v entry by entry, repeated dofs accumulate in call order.
    # Driven through the vector path with single-node "elements": a 1-point T3 rule would not give arbitrary
    # values, so the accumulation order is checked on the element loop itself: two T3 elements sharing nodes.
    mesh = efg.T3block(1.0, 1.0, 1, 1)             # 2 triangles, 4 nodes
    dof = np.array([[3], [1], [4], [2]], dtype=np.int64)
    F = oracle.assemble_vec_heat(1, mesh, dof, 2.5, 4)
    want = np.zeros(4)
    for e in range(mesh.nel):                       # Python restatement of assemble!(av, fe)
        x = mesh.xy[mesh.conn[e] - 1]
        J = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        for k in range(3):
            gi = dof[mesh.conn[e, k] - 1, 0]
            want[gi - 1] = want[gi - 1] + (0.0 + ((1 - 1 / 3 - 1 / 3 if k == 0 else 1 / 3) * 2.5) * (J * 0.5))
    assert np.array_equal(F, want)
    with pytest.raises(IndexError):                 # BoundsError: dof number beyond nrow
        oracle.assemble_vec_heat(1, mesh, dof, 2.5, 3)


def test_spmv_matches_scipy_and_julia_order(oracle):
    # K*T of examples/heat/poisson/t3.jl:78 -- column sweep, y[r] accumulated by ascending column
    rng = np.random.default_rng(7)
    prob = efg.heat_problem(efg.T6, 9, perturb=True)
    cp, rv, nz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    x = rng.standard_normal(prob.ndofs)
    y = oracle.spmv_csc(prob.ndofs, prob.ndofs, cp, rv, nz, x)
    K = _csc(cp, rv, nz, prob.ndofs)
    assert np.allclose(y, K @ x, rtol=1e-13, atol=1e-13)
    Kr = K.tocsr(); Kr.sort_indices()
    r = 17
    acc = 0.0
    for c, a in zip(Kr.indices[Kr.indptr[r]:Kr.indptr[r + 1]], Kr.data[Kr.indptr[r]:Kr.indptr[r + 1]]):
        acc = acc + a * x[c]
    assert y[r] == acc


def _solve(K, U, F, nu):
    # solve!: examples/heat/poisson/t3.jl:77-80
    KT = K @ U
    U[:nu] = spla.spsolve(K[:nu, :nu].tocsc(), F[:nu] - KT[:nu])
    return U


def test_heat_t3_n4_golden_solution(oracle):
    # test/test_heat.jl:22-110: T3block N=4, all boundary nodes prescribed, Q=-6, kappa=1
    N, kappa, Q = 4, 1.0, -6.0
    tempf = lambda x, y: 1.0 + x ** 2 + 2.0 * y ** 2
    mesh = efg.T3block(1.0, 1.0, N, N)
    fesp = efg.FESpace(mesh, efg.FEH1_T3())
    for i in efg.boundary_nodes(mesh):
        efg.setebc(fesp, 0, i, 1, tempf(*mesh.xy[i - 1]))
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    n = efg.ndofs(fesp)
    colptr, rowval, nzval = oracle.assemble(oracle.FORM_HEAT, 1, mesh, None, [fesp.field.dofnums],
                                            [kappa], n, n)
    assert len(rowval) == 7 * N * N + 6 * N + 1          # nnz closed form
    K = _csc(colptr, rowval, nzval, n)
    # F: SysvecAssembler + LocalVectorAssembler, fe[j] += N[j]*Q*JxW (test/test_heat.jl:40-57)
    F = oracle.assemble_vec_heat(1, mesh, fesp.field.dofnums, Q, n)
    # independent closed form: 1-pt rule, N = 1/3, J = 2*area
    F2 = np.zeros(n)
    d = fesp.field.dofnums[:, 0]
    for e in range(mesh.nel):
        nodes = mesh.conn[e] - 1
        x = mesh.xy[nodes]
        J = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        F2[d[nodes] - 1] += (1 / 3) * Q * (J * 0.5)
    assert np.allclose(F, F2, rtol=1e-14, atol=1e-16)
    T = _solve(K, efg.gathersysvec(fesp), F, efg.nunknowns(fesp))
    ref = [1.1875, 1.3749999999999998, 1.6874999999999998, 1.5624999999999998, 1.7499999999999998,
           2.0625, 2.1875, 2.375, 2.6875, 1.0, 1.0625, 1.25, 1.5625, 2.0, 1.125, 2.125, 1.5, 2.5,
           2.125, 3.125, 3.0, 3.0625, 3.25, 3.5625, 4.0]
    assert np.allclose(T, ref, rtol=0, atol=1e-13)


def test_heat_q4_n100_unknowns_and_accuracy(oracle):
    # test/test_heat.jl:141-221: Q4block N=100, Gauss order 2: 9801 unknowns, mean nodal error <= 1e-9
    N, kappa, Q = 100, 1.0, -6.0
    tempf = lambda x, y: 1.0 + x ** 2 + 2.0 * y ** 2
    mesh = efg.Q4block(1.0, 1.0, N, N)
    fesp = efg.FESpace(mesh, efg.FEH1_Q4())
    for i in efg.boundary_nodes(mesh):
        efg.setebc(fesp, 0, i, 1, tempf(*mesh.xy[i - 1]))
    efg.numberfreedofs(fesp); efg.numberdatadofs(fesp)
    assert efg.nunknowns(fesp) == 9801
    n = efg.ndofs(fesp)
    colptr, rowval, nzval = oracle.assemble(oracle.FORM_HEAT, 2, mesh, None, [fesp.field.dofnums], [kappa], n, n)
    assert len(rowval) == (3 * N + 1) ** 2
    K = _csc(colptr, rowval, nzval, n)
    # load vector through the restated SysvecAssembler path (test/test_heat.jl:154-171); on uniform squares
    # sum_q N_j * Q * JxW = Q * area / 4 per node and element
    F = oracle.assemble_vec_heat(2, mesh, fesp.field.dofnums, Q, n)
    F2 = np.zeros(n)
    d = fesp.field.dofnums[:, 0]
    area = (1.0 / N) ** 2
    np.add.at(F2, d[mesh.conn - 1].ravel() - 1, Q * area / 4)
    assert np.allclose(F, F2, rtol=1e-12, atol=1e-18)
    T = _solve(K, efg.gathersysvec(fesp), F, efg.nunknowns(fesp))
    efg.scattersysvec(fesp, T)
    err = np.abs(fesp.field.dofvals[:, 0] - tempf(mesh.xy[:, 0], mesh.xy[:, 1])).mean()
    assert err <= 1.0e-9


# ---------------------------------------------------------------------------------------------
# Stokes goldens: colliding flow, Taylor-Hood T6/T3 (test/test_stokes.jl)
# ---------------------------------------------------------------------------------------------
_trueux = lambda x, y: 20 * x * y ** 3
_trueuy = lambda x, y: 5 * x ** 4 - 5 * y ** 4
_truep = lambda x, y: 60 * x ** 2 * y - 20 * y ** 3


def _stokes_meshes(N, A=1.0):
    vmesh = efg.T6block(2 * A, 2 * A, N, N)
    efg.transform(vmesh, lambda x: x - A)
    pmesh = efg.T6toT3(vmesh)
    return vmesh, pmesh


def _ebc_velocity(spaces, vmesh, N, A=1.0):
    inflate = A / N / 100
    for box in ([-A, A, -A, -A], [-A, -A, -A, A], [A, A, -A, A], [-A, A, A, A]):
        for i in efg.vselect(vmesh.xy, box=box, inflate=inflate):
            x, y = vmesh.xy[i - 1]
            if len(spaces) == 1:
                efg.setebc(spaces[0], 0, i, 1, _trueux(x, y))
                efg.setebc(spaces[0], 0, i, 2, _trueuy(x, y))
            else:
                efg.setebc(spaces[0], 0, i, 1, _trueux(x, y))
                efg.setebc(spaces[1], 0, i, 1, _trueuy(x, y))


def _errors(oracle, vmesh, pmesh, ux, uy, p):
    """evaluate_pressure_error / evaluate_velocity_error (test/test_stokes.jl:438-496), numpy."""
    pc, w = oracle.quadrature(efg.T3, 3)
    Ep = 0.0
    Ev = 0.0
    Np = np.array([oracle.bfun(efg.T3, *q) for q in pc])
    Nu = np.array([oracle.bfun(efg.T6, *q) for q in pc])
    gp3 = np.array([oracle.bfungradpar(efg.T3, *q) for q in pc])
    gp6 = np.array([oracle.bfungradpar(efg.T6, *q) for q in pc])
    pc_ = pmesh.conn - 1
    vc_ = vmesh.conn - 1
    X3 = pmesh.xy[pc_]            # (nel, 3, 2)
    X6 = vmesh.xy[vc_]            # (nel, 6, 2)
    for q in range(3):
        J3 = np.einsum("eni,nk->eik", X3, gp3[q])
        det3 = J3[:, 0, 0] * J3[:, 1, 1] - J3[:, 1, 0] * J3[:, 0, 1]
        loc = np.einsum("eni,n->ei", X3, Np[q])
        pa = p[pc_] @ Np[q]
        Ep += np.sum(det3 * w[q] * (pa - _truep(loc[:, 0], loc[:, 1])) ** 2)
        J6 = np.einsum("eni,nk->eik", X6, gp6[q])
        det6 = J6[:, 0, 0] * J6[:, 1, 1] - J6[:, 1, 0] * J6[:, 0, 1]
        loc = np.einsum("eni,n->ei", X6, Nu[q])
        uxa = ux[vc_] @ Nu[q]
        uya = uy[vc_] @ Nu[q]
        Ev += np.sum(det6 * w[q] * ((uxa - _trueux(loc[:, 0], loc[:, 1])) ** 2
                                     + (uya - _trueuy(loc[:, 0], loc[:, 1])) ** 2))
    return np.sqrt(Ep), np.sqrt(Ev)


def _run_stokes(oracle, form, N, three_spaces, assemble=None, errors=None):
    vmesh, pmesh = _stokes_meshes(N)
    if three_spaces:
        U = [efg.FESpace(vmesh, efg.FEH1_T6(), 1), efg.FESpace(vmesh, efg.FEH1_T6(), 1)]
    else:
        U = [efg.FESpace(vmesh, efg.FEH1_T6(), 2)]
    _ebc_velocity(U, vmesh, N)
    Ph = efg.FESpace(pmesh, efg.FEH1_T3(), 1)
    efg.setebc(Ph, 0, efg.vselect(pmesh.xy, nearestto=[0.0, 0.0])[0], 1, 0.0)
    spaces = U + [Ph]
    efg.numberdofs(spaces)
    tndof = sum(efg.ndofs(s) for s in spaces)
    tnunk = sum(efg.nunknowns(s) for s in spaces)
    params = np.array([2.0, 0, 0, 0, 2.0, 0, 0, 0, 1.0]) if form == oracle.FORM_STOKES_GEN else np.array([1.0])
    if assemble is None:
        colptr, rowval, nzval = oracle.assemble(form, 3, vmesh, pmesh, [s.field.dofnums for s in spaces],
                                                params, tndof, tndof)
    else:
        colptr, rowval, nzval = assemble(form, spaces, params, tndof)
    K = _csc(colptr, rowval, nzval, tndof)
    Uv = _solve(K, efg.gathersysvec(spaces, tndof), np.zeros(tndof), tnunk)
    efg.scattersysvec(spaces, Uv)
    if three_spaces:
        ux, uy = U[0].field.dofvals[:, 0], U[1].field.dofvals[:, 0]
    else:
        ux, uy = U[0].field.dofvals[:, 0], U[0].field.dofvals[:, 1]
    ep_np, ev_np = _errors(oracle, vmesh, pmesh, ux, uy, Ph.field.dofvals[:, 0])
    # the restated evaluate_pressure_error / evaluate_velocity_error (oracle C, or the GPU through `errors`): eldofvals
    # come from the system vector through the dof numbers, the true solution is evaluated at location(el, qp)
    ep, ev = (errors or _oracle_errors)(oracle, vmesh, pmesh, spaces, Uv, three_spaces)
    assert np.allclose([ep, ev], [ep_np, ev_np], rtol=1e-11, atol=0)     # independent vectorised numpy evaluation
    return ep, ev, tndof, tnunk, len(rowval)


def _oracle_errors(oracle, vmesh, pmesh, spaces, Uv, three_spaces):
    lp = oracle.qp_locations(3, pmesh)
    ep = oracle.l2_error(3, pmesh, [(spaces[-1].field.dofnums, 0)], Uv, _truep(lp[..., 0], lp[..., 1])[..., None])
    lv = oracle.qp_locations(3, vmesh)
    truth = np.stack([_trueux(lv[..., 0], lv[..., 1]), _trueuy(lv[..., 0], lv[..., 1])], axis=-1)
    comps = ([(spaces[0].field.dofnums, 0), (spaces[1].field.dofnums, 0)] if three_spaces
             else [(spaces[0].field.dofnums, 0), (spaces[0].field.dofnums, 1)])
    ev = oracle.l2_error(3, vmesh, comps, Uv, truth)
    return ep, ev


def test_stokes_reddy_goldens(oracle):
    # test/test_stokes.jl:550-560
    ref = [(3.5171450671095306, 0.2968271617227661), (0.5999467323539439, 0.03781189670123018),
           (0.12350320261417459, 0.004741849976722882)]
    for N, r in zip((4, 8, 16), ref):
        ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_REDDY, N, True)
        assert np.allclose([ep, ev], r, rtol=1e-9, atol=0)   # reference: isapprox, rtol ~1.5e-8


def test_stokes_veclap_alt_golden(oracle):
    # test/test_stokes.jl:777-780
    ep, ev, _, _, nnz = _run_stokes(oracle, oracle.FORM_STOKES_VECLAP_ALT, 4, False)
    assert np.allclose([ep, ev], [2.596076907594511, 0.3001331486426876], rtol=1e-9, atol=0)
    assert nnz == 260 * 16 + 104 * 4 + 8     # explicit zeros of the skipped cross-component entries are stored


def test_stokes_gen_matches_reddy_golden_and_dof_counts(oracle):
    # gen formulation (examples/stokes/colliding_flow/ht_p2_p1_gen.jl) equals the Reddy golden at N=4
    ep, ev, *_ = _run_stokes(oracle, oracle.FORM_STOKES_GEN, 4, False)
    assert np.allclose([ep, ev], (3.5171450671095306, 0.2968271617227661), rtol=1e-9, atol=0)


def test_stokes_dof_counts_n100():
    # test/test_stokes.jl:130: (tndof, tnunk) == (91003, 89402) for the driven cavity T6/T3, N=100
    vmesh = efg.T6block_fast(1.0, 1.0, 100, 100)
    pmesh = efg.T6toT3(vmesh)
    Uh = efg.FESpace(vmesh, efg.FEH1_T6(), 2)
    for i in efg.boundary_nodes(vmesh):
        efg.setebc(Uh, 0, i, 1, 0.0); efg.setebc(Uh, 0, i, 2, 0.0)
    Ph = efg.FESpace(pmesh, efg.FEH1_T3(), 1)
    efg.setebc(Ph, 0, 1, 1, 0.0)
    efg.numberdofs([Uh, Ph])
    assert (efg.ndofs(Uh) + efg.ndofs(Ph), efg.nunknowns(Uh) + efg.nunknowns(Ph)) == (91003, 89402)


# ---- the direct-accumulate mode is the literal COO + sparse() mode, bit for bit -------------------------
@pytest.mark.parametrize("make", [
    lambda: efg.heat_problem(efg.T3, 9, True), lambda: efg.heat_problem(efg.T6, 7, True), lambda: efg.heat_problem(efg.Q4, 8, True),
    lambda: efg.heat_problem(efg.Q4, 5, True, quad=3), lambda: efg.elasticity_problem(6, efg.T6, True),
    lambda: efg.stokes_problem(5, "gen", True), lambda: efg.stokes_problem(5, "reddy", True),
    lambda: efg.stokes_problem(4, "veclap_alt", True), lambda: efg.stokes_problem(4, "veclap", True)])
def test_direct_accumulate_mode_equals_coo_mode(oracle, make):
    p = make()
    a = oracle.assemble(*efg.oracle_args(p), p.ndofs, p.ndofs)
    b = oracle.assemble_direct(*efg.oracle_args(p), p.ndofs, p.ndofs)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    pairs = [(p.meshes[ms].conn, s.field.dofnums) for s, ms in zip(p.spaces, p.space_mesh)]
    c = oracle.assemble_direct_parallel(efg.oracle_args(p), p.ndofs, p.ndofs, pairs, nblocks=5, threads=3)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, c))
    # a column block on its own, visiting only the elements that touch it
    c0, c1 = p.ndofs // 3, 2 * p.ndofs // 3
    el = oracle.elements_touching(pairs, c0, c1)
    assert len(el) <= p.nel
    cb = oracle.assemble_direct(*efg.oracle_args(p), p.ndofs, p.ndofs, c0=c0, c1=c1, elist=el)
    lo, hi = a[0][c0 - 1] - 1, a[0][c1] - 1
    assert np.array_equal(cb[0], a[0][c0 - 1:c1 + 1] - a[0][c0 - 1] + 1) and np.array_equal(cb[1], a[1][lo:hi])
    assert cb[2].tobytes() == a[2][lo:hi].tobytes()


def test_direct_accumulate_mode_rejects_unnumbered_dofs(oracle):
    p = efg.heat_problem(efg.T3, 5)
    efg.numberfreedofs(p.spaces[0])
    with pytest.raises(ValueError):
        oracle.assemble_direct(*efg.oracle_args(p), p.ndofs, p.ndofs)


def test_gauss_rules_integrate_exp_like_the_reference(oracle):
    """test/test_refshapes.jl:124-146 (mrs7): sum(exp.(param_coords) .* weights) of the interval rules, orders 1..5 (the orders
    the oracle restates with the reference's literals; higher orders are Golub-Welsch in the reference).  The square rule is
    the tensor product, so sum_ij w_ij exp(x_i) = 2 * the interval result."""
    ref = [2.0, 2.3426960879097307, 2.350336928680011, 2.3504020921563744, 2.350402386462827]
    for order, want in zip(range(1, 6), ref):
        pc, w = oracle.quadrature(efg.Q4, order)
        assert len(w) == order ** 2                                  # test/test_refshapes.jl:64-67 (mrs4)
        assert np.isclose(float((np.exp(pc[:, 0]) * w).sum()) / 2.0, want, rtol=1e-8, atol=0)


def test_higher_triangle_rules_as_in_the_reference(oracle):
    """src/RefShapes.jl:120-230, test/test_refshapes.jl:74-88 (mrs5: npts in [1, 3, 4, 6, 7, 9, 12, 13]): the literal tables --
    every rule sums to the triangle's area and integrates the monomials x^a y^b = a! b! / (a + b + 2)! up to its degree (to the
    precision of the reference's own 15-digit literals: the 7-point rule is only good to 3e-9, one of its coordinates is mistyped
    in the reference and restated as it is)."""
    from math import factorial
    degree = {1: 1, 3: 2, 4: 3, 6: 4, 7: 5, 9: 5, 12: 6, 13: 7}
    for n, deg in degree.items():
        pc, w = oracle.quadrature(efg.T3, n)
        assert len(w) == n and pc.shape == (n, 2)
        assert abs(w.sum() - 0.5) < 1e-14
        for d in range(deg + 1):
            for a in range(d + 1):
                exact = factorial(a) * factorial(d - a) / factorial(d + 2)
                assert abs(float((w * pc[:, 0] ** a * pc[:, 1] ** (d - a)).sum()) - exact) < 1e-8, (n, d, a)
    with pytest.raises(ValueError):
        oracle.quadrature(efg.T3, 5)            # "Unknown number of integration points"
