"""meshes, EBCs and numbering made on the device (efg_gen_mesh / efg_gen_space / efg_setebc_* /
efg_number_dofs) must be the arrays the host mirror of MeshSteward's generators and of numberdofs! produces, bit for bit --
and assembling from them must give the matrix assembled from the uploaded host arrays."""
import numpy as np
import pytest

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,gen", [(efg.T3, efg.T3block), (efg.Q4, efg.Q4block), (efg.T6, efg.T6block)], ids=["t3", "q4", "t6"])
@pytest.mark.parametrize("dims", [(1.0, 1.0, 7, 5), (2.0, 3.0, 1, 1), (1.0, 0.5, 13, 29), (3.0, 1.0, 40, 3)])
def test_generated_meshes_equal_the_host_mirror(kind, gen, dims):
    L, W, nL, nW = dims
    want = gen(L, W, nL, nW)
    eng = efg.Engine(0)
    eng.gen_mesh(0, kind, nL, nW, L, W)
    conn, xy = eng.fetch_mesh(0, kind)
    assert np.array_equal(conn, want.conn)
    assert xy.tobytes() == want.xy.tobytes()
    eng.close()


def _gen_heat(eng, kind, N):
    eng.gen_mesh(0, kind, N, N, 1.0, 1.0)
    eng.gen_space(0, 0, 1)
    tol = 1.0 / N / 100
    for box in ((0, 1, -tol, tol), (0, 1, 1 - tol, 1 + tol), (-tol, tol, 0, 1), (1 - tol, 1 + tol, 0, 1)):     # the four edges (vselect boxes)
        eng.setebc_box(0, 1, box[0] - tol, box[1] + tol, box[2], box[3])
    return eng.number_dofs([0])


@pytest.mark.parametrize("kind", [efg.T3, efg.Q4, efg.T6])
def test_heat_numbering_and_assembly_from_generated_inputs(oracle, kind):
    N = 37
    prob = efg.heat_problem(kind, N)
    eng = efg.Engine(0)
    nfree, nd = _gen_heat(eng, kind, N)
    assert nd == prob.ndofs and nfree == efg.nunknowns(prob.spaces[0])
    dof = eng.fetch_dofnums(0, prob.meshes[0].nnodes, 1)
    assert np.array_equal(dof, prob.spaces[0].field.dofnums)
    eng.start(nd, nd)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp, rv, nz = eng.fetch_csc()
    eng.close()
    e2 = efg.Engine(0)
    efg.load_problem(e2, prob)
    e2.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp2, rv2, nz2 = e2.fetch_csc()
    e2.close()
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2) and nz.tobytes() == nz2.tobytes()
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)


def test_stokes_setup_on_the_device_equals_the_host_mirror(oracle):
    """examples/stokes/colliding_flow/ht_p2_p1_gen.jl: T6block(2A, 2A, N, N) shifted by -A, T6toT3, velocity prescribed on the
    whole boundary, one pressure node pinned, numberdofs!([Uh, Ph])."""
    N, A = 14, 1.0
    prob = efg.stokes_problem(N, "gen")
    eng = efg.Engine(0)
    eng.gen_mesh(0, efg.T6, N, N, 2 * A, 2 * A, -A, -A)
    eng.gen_mesh_corners(1, 0)
    vconn, vxy = eng.fetch_mesh(0, efg.T6)
    pconn, pxy = eng.fetch_mesh(1, efg.T3)
    assert np.array_equal(vconn, prob.meshes[0].conn) and vxy.tobytes() == prob.meshes[0].xy.tobytes()
    assert np.array_equal(pconn, prob.meshes[1].conn) and pxy.tobytes() == prob.meshes[1].xy.tobytes()
    eng.gen_space(0, 0, 2)
    eng.gen_space(1, 1, 1)
    tol = 2 * A / N / 100
    for box in ((-A - tol, A + tol, -A - tol, -A + tol), (-A - tol, A + tol, A - tol, A + tol), (-A - tol, -A + tol, -A - tol, A + tol), (A - tol, A + tol, -A - tol, A + tol)):
        eng.setebc_box(0, 0, *box)
    pinned = int(np.argmin((prob.meshes[1].xy ** 2).sum(axis=1))) + 1
    eng.setebc_nodes(1, 1, [pinned])
    nfree, nd = eng.number_dofs([0, 1])
    assert nd == prob.ndofs
    assert np.array_equal(eng.fetch_dofnums(0, prob.meshes[0].nnodes, 2), prob.spaces[0].field.dofnums)
    assert np.array_equal(eng.fetch_dofnums(1, prob.meshes[1].nnodes, 1), prob.spaces[1].field.dofnums)
    eng.start(nd, nd)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    cp, rv, nz = eng.fetch_csc()
    eng.close()
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
    assert np.all(np.abs(nz - onz) <= 1e-14 + 1e-12 * np.abs(onz))


def test_gen_errors():
    eng = efg.Engine(0)
    with pytest.raises(efg.EfgError):
        eng.gen_space(0, 0, 1)                      # no mesh yet
    eng.gen_mesh(0, efg.T3, 4, 4)
    eng.gen_space(0, 0, 1)
    with pytest.raises(efg.ArgumentError):
        eng.setebc_nodes(0, 1, [1, 26])              # 25 nodes only
    with pytest.raises(efg.EfgError):
        eng.setebc_box(0, 2, 0, 1, 0, 1)            # no component 2
    prob = efg.heat_problem(efg.T3, 4)
    eng.set_space(1, 0, prob.spaces[0].field.dofnums)
    with pytest.raises(efg.EfgError):
        eng.number_dofs([1])                         # an uploaded space has no EBC flags on the device
    eng.close()
