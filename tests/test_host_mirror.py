"""Host-side logic of the mirrored assembler API that needs no GPU: which (space slot, component) pairs and mesh slot
`evaluate_error` hands to the C ABI, and the life-cycle checks of the assembler objects (a recording stand-in replaces
the device context; no compute happens here)."""
import numpy as np
import pytest

import elfel_jl_b200 as efg
from elfel_jl_b200 import assemblers as asm


class _RecordingEngine:
    def __init__(self, its):
        self._keep = list(its)
        self.calls = []

    def qp_locations(self, mesh_slot, quad, nel):
        self.calls.append(("loc", mesh_slot, quad, nel))
        return np.zeros((nel, 3, 2))

    def l2_error(self, comps, quad, U, truth):
        self.calls.append(("err", list(comps), quad, truth.shape))
        return 0.0


class _Ass:
    def __init__(self, eng):
        self.engine = eng


def _stokes_spaces(three):
    vmesh = efg.T6block_fast(2.0, 2.0, 3, 3)
    pmesh = efg.T6toT3(vmesh)
    U = [efg.FESpace(vmesh, efg.FEH1_T6(), 1), efg.FESpace(vmesh, efg.FEH1_T6(), 1)] if three else [efg.FESpace(vmesh, efg.FEH1_T6(), 2)]
    return vmesh, pmesh, U + [efg.FESpace(pmesh, efg.FEH1_T3(), 1)]


@pytest.mark.parametrize("three", [True, False])
def test_evaluate_error_maps_iterators_to_slots_and_components(three):
    vmesh, pmesh, spaces = _stokes_spaces(three)
    its = [efg.FEIterator(s) for s in spaces]
    qp = efg.QPIterator(spaces[0], kind="default", npts=3)
    eng = _RecordingEngine(its)
    one = lambda x, y: np.ones_like(x)
    # pressure: the last space, on mesh slot 1 (the second distinct mesh of the assemble call)
    asm.evaluate_error(_Ass(eng), efg.FEIterator(spaces[-1]), efg.QPIterator(spaces[-1], kind="default", npts=3), np.zeros(5), one)
    assert eng.calls[0] == ("loc", 1, 3, pmesh.conn.shape[0])
    assert eng.calls[1][1] == [(len(spaces) - 1, 0)] and eng.calls[1][3] == (pmesh.conn.shape[0], 3, 1)
    # velocity: two scalar spaces -> components 0 of slots 0 and 1; one vector space -> components 0 and 1 of slot 0
    uel = (its[0], its[1]) if three else (its[0], its[0])
    asm.evaluate_error(_Ass(eng), uel, qp, np.zeros(5), (one, one))
    assert eng.calls[2] == ("loc", 0, 3, vmesh.conn.shape[0])
    assert eng.calls[3][1] == ([(0, 0), (1, 0)] if three else [(0, 0), (0, 1)])
    assert eng.calls[3][3] == (vmesh.conn.shape[0], 3, 2)
    # a space that was not part of the last assemble call is refused
    other = efg.FESpace(vmesh, efg.FEH1_T6(), 1)
    with pytest.raises(efg.EfgError):
        asm.evaluate_error(_Ass(eng), efg.FEIterator(other), qp, np.zeros(5), one)


def test_vector_forms_and_iterators_carry_what_the_abi_needs():
    f = efg.HeatLoadForm(-6.0)
    assert f.vform_id == 1 and np.array_equal(f.params(), [-6.0])
    mesh = efg.Q4block(1.0, 1.0, 2, 2)
    fesp = efg.FESpace(mesh, efg.FEH1_Q4())
    assert efg.QPIterator(fesp, kind="Gauss", order=2).rule == 2
    with pytest.raises(ValueError):
        efg.QPIterator(fesp, kind="Simpson")
    it = efg.FEIterator(fesp)
    assert len(it) == 4 and it._bir.shape == (4, 4) and it._geom.shape == (9, 2)


def test_in_place_widening_of_the_row_indices(tmp_path):
    """efg_hostcopy.cuh: the Int32 row indices arrive in the upper half of the caller's Int64 array and several host threads
    widen them in place -- every entry must come out as in32 + 1 for all sizes / chunkings (CUDA calls stubbed)."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_widen_test"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "elfel.jl_b200", "csrc"),
                           "-o", str(exe), os.path.join(root, "tests", "host_widen_test.cpp")])
    subprocess.check_call([str(exe)])
