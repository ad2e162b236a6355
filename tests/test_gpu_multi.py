"""Several GPUs behind one handle (efgm_*, include/elfel_gpu.h): global arrays in, global CSC out, the library shards
internally.  The device list may repeat a device, so the whole path -- bands, element selection, local
renumbering, owned column ranges, per-device symbolic + numeric phases on host threads, interleaving fetch -- is exercised
on a one-GPU box too; with two GPUs visible the same tests also run on devices (0, 1)."""
import numpy as np
import pytest
import torch

import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-12, 1e-14


def _device_lists():
    lists = [[0, 0], [0, 0, 0]]
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        lists.append([0, 1])
    return lists


def _single(prob, strict=0):
    eng = efg.Engine(0)
    eng.set_option(_lib.OPT_STRICT_FP, strict)
    efg.load_problem(eng, prob)
    eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
    out = eng.fetch_csc()
    eng.close()
    return out


def _multi(prob, devices, strict=0):
    m = efg.MultiEngine(devices)
    m.set_option(_lib.OPT_STRICT_FP, strict)
    for slot, mesh in enumerate(prob.meshes):
        m.set_mesh(slot, mesh.kind, np.ascontiguousarray(mesh.conn, dtype=np.int64), np.ascontiguousarray(mesh.xy, dtype=np.float64))
    for slot, (s, ms) in enumerate(zip(prob.spaces, prob.space_mesh)):
        m.set_space(slot, ms, np.ascontiguousarray(s.field.dofnums, dtype=np.int64))
    m.start(prob.ndofs, prob.ndofs)
    nnz = m.assemble(prob.form.form_id, prob.quad, prob.form.params())
    out = m.fetch_csc()
    owned = np.zeros(prob.ndofs, dtype=np.int64)
    for d in range(len(devices)):
        f, l = m.device_ranges(d)
        for a, b in zip(f, l):
            owned[a - 1: b] += 1
    return m, nnz, out, owned


PROBLEMS = [lambda: efg.heat_problem(efg.T6, 40, True), lambda: efg.heat_problem(efg.Q4, 61, True), lambda: efg.heat_problem(efg.T3, 50, True),
            lambda: efg.elasticity_problem(31, efg.T6, True), lambda: efg.stokes_problem(23, "gen", True), lambda: efg.stokes_problem(17, "reddy", True)]


@pytest.mark.parametrize("make", PROBLEMS, ids=["heat_t6", "heat_q4", "heat_t3", "elasticity_t6", "stokes_gen", "stokes_reddy"])
def test_multi_handle_equals_single_context(oracle, make):
    prob = make()
    scp, srv, snz = _single(prob)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    assert np.array_equal(scp, ocp) and np.array_equal(srv, orv)
    for devices in _device_lists():
        m, nnz, (cp, rv, nz), owned = _multi(prob, devices)
        assert np.all(owned == 1), "every column is owned by exactly one device"
        assert nnz == len(orv)
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv), f"pattern differs with devices {devices}"
        assert nz.tobytes() == snz.tobytes(), "the sharded result must be bit-identical to the single-GPU result"
        assert np.all(np.abs(nz - onz) <= ATOL + RTOL * np.abs(onz))
        # re-assembly on the cached shards
        m.numeric(prob.form.params())
        assert m.fetch_csc()[2].tobytes() == snz.tobytes()
        assert m.stat(_lib.STAT_NUMERIC_MS) > 0.0
        m.close()


def test_multi_handle_strict_mode_is_the_oracle(oracle):
    prob = efg.elasticity_problem(25, efg.T6, True)
    ocp, orv, onz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    m, nnz, (cp, rv, nz), _ = _multi(prob, [0, 0, 0, 0], strict=1)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
    m.close()


def test_multi_handle_errors():
    prob = efg.heat_problem(efg.T3, 8)
    m = efg.MultiEngine([0, 0])
    with pytest.raises(efg.EfgError):
        m.assemble(_lib.FORM_HEAT, 1, [1.0])                    # before start
    efg.numberfreedofs(prob.spaces[0])                          # data dofs unnumbered -> sparse()'s ArgumentError
    m.set_mesh(0, efg.T3, prob.meshes[0].conn, prob.meshes[0].xy)
    m.set_space(0, 0, prob.spaces[0].field.dofnums)
    m.start(prob.ndofs, prob.ndofs)
    with pytest.raises(efg.ArgumentError):
        m.assemble(_lib.FORM_HEAT, 1, [1.0])
    m.close()
