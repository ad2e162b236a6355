"""Host-side sharding logic (owner-computes column ranges, halo replication, block merge) -- CPU tests,
including a world_size-2 gloo run.  The oracle stands in for the GPU assembler here (tests may use it)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import elfel_jl_b200 as efg
from elfel_jl_b200 import sharding as sh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torch_generators_match_numpy():
    for (nL, nW) in ((5, 7), (1, 1), (8, 3)):
        k, c, x = sh.t6block_torch(2.0, 3.0, nL, nW)
        m = efg.T6block_fast(2.0, 3.0, nL, nW)
        assert np.array_equal(c.numpy(), m.conn) and np.array_equal(x.numpy(), m.xy)
        k, c, x = sh.t3block_torch(2.0, 3.0, nL, nW)
        m = efg.T3block(2.0, 3.0, nL, nW)
        assert np.array_equal(c.numpy(), m.conn) and np.array_equal(x.numpy(), m.xy)
        k, c, x = sh.q4block_torch(2.0, 3.0, nL, nW)
        m = efg.Q4block(2.0, 3.0, nL, nW)
        assert np.array_equal(c.numpy(), m.conn) and np.array_equal(x.numpy(), m.xy)


def test_numbering_matches_host_mirror():
    p = efg.heat_problem(efg.T6, 6)
    kind, conn, xy, dofnums, band, form, quad = sh.build_global("heat_t6", 6, 1, "cpu")
    assert np.array_equal(conn.numpy(), p.meshes[0].conn)
    assert np.array_equal(dofnums.numpy(), p.spaces[0].field.dofnums)
    pe = efg.elasticity_problem(5, efg.T6)
    kind, conn, xy, dofnums, band, form, quad = sh.build_global("elasticity_t6", 5, 1, "cpu")
    assert np.array_equal(dofnums.numpy(), pe.spaces[0].field.dofnums)


def _global_oracle(oracle, workload, n, world):
    kind, conn, xy, dofnums, band, form, quad = sh.build_global(workload, n, world, "cpu")
    mesh = efg.Mesh(kind, conn.numpy(), xy.numpy())
    nd = dofnums.numel()
    return oracle.assemble(form.form_id, quad, mesh, None, [dofnums.numpy()], form.params(), nd, nd), nd


def _shard_block_via_oracle(oracle, workload, n, rank, world):
    s, (firsts, lasts), nel_global = sh.shard_problem(efg, workload, n, rank, world, dev="cpu")
    mesh = efg.Mesh(s.meshes[0].kind, s.meshes[0].conn.numpy(), s.meshes[0].xy.numpy())
    cp, rv, nz = oracle.assemble(s.form.form_id, s.quad, mesh, None, [s.spaces[0].field.dofnums.numpy()],
                                 s.form.params(), s.ndofs, s.ndofs)
    # keep the owned columns only
    cols = np.concatenate([np.arange(f, l + 1) for f, l in zip(firsts, lasts)])
    cnt = np.diff(cp)[cols - 1]
    lcp = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int64)
    idx = np.concatenate([np.arange(cp[c - 1] - 1, cp[c] - 1) for c in cols]) if len(cols) else np.zeros(0, np.int64)
    return (firsts, lasts, lcp, rv[idx], nz[idx]), s


@pytest.mark.parametrize("workload,n,world", [("heat_t6", 5, 3), ("heat_q4", 6, 2), ("heat_t3", 4, 4), ("elasticity_t6", 4, 2)])
def test_shards_merge_to_global_matrix(oracle, workload, n, world):
    (gcp, grv, gnz), nd = _global_oracle(oracle, workload, n, world)
    blocks, owned = [], np.zeros(nd, dtype=np.int64)
    for r in range(world):
        b, s = _shard_block_via_oracle(oracle, workload, n, r, world)
        for f, l in zip(b[0], b[1]):
            owned[f - 1: l] += 1
        assert s.nel < s.nel_global or world == 1          # a real sub-mesh
        blocks.append(b)
    assert np.all(owned == 1), "every column is owned by exactly one rank"
    cp, rv, nz = sh.merge_blocks(nd, blocks)
    assert np.array_equal(cp, gcp) and np.array_equal(rv, grv)
    assert nz.tobytes() == gnz.tobytes()       # halo elements replicated => identical sums, bit for bit


@pytest.mark.parametrize("workload,n,world", [("heat_t6", 5, 3), ("heat_q4", 6, 2), ("heat_t3", 4, 4)])
def test_sharded_load_vectors_merge_to_global_vector(oracle, workload, n, world):
    """System-vector assembly shards by the same owned ranges as the matrix columns: every rank sums the contributions of
    its sub-mesh (halo elements replicated) into its owned rows; the blocks scatter into the global vector bit for bit."""
    kind, conn, xy, dofnums, band, form, quad = sh.build_global(workload, n, world, "cpu")
    gmesh = efg.Mesh(kind, conn.numpy(), xy.numpy())
    nd = dofnums.numel()
    want = oracle.assemble_vec_heat(quad, gmesh, dofnums.numpy(), -6.0, nd)
    blocks = []
    for r in range(world):
        s, (firsts, lasts), _ = sh.shard_problem(efg, workload, n, r, world, dev="cpu")
        mesh = efg.Mesh(s.meshes[0].kind, s.meshes[0].conn.numpy(), s.meshes[0].xy.numpy())
        full = oracle.assemble_vec_heat(s.quad, mesh, s.spaces[0].field.dofnums.numpy(), -6.0, s.ndofs)
        rows = np.concatenate([np.arange(f, l + 1) for f, l in zip(firsts, lasts)])
        blocks.append((firsts, lasts, full[rows - 1]))
    assert sh.merge_vectors(nd, blocks).tobytes() == want.tobytes()


def test_world2_gloo(tmp_path):
    """Two processes over gloo: each assembles its shard (oracle stand-in), nnz and a checksum are reduced."""
    script = tmp_path / "w2.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import numpy as np, torch, torch.distributed as dist
import elfel_jl_b200 as efg
from oracle import oracle as orc
from test_sharding import _shard_block_via_oracle, _global_oracle
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
b, s = _shard_block_via_oracle(orc, "heat_t6", 6, r, w)
t = torch.tensor([float(len(b[3])), float(np.abs(b[4]).sum())], dtype=torch.float64)
dist.all_reduce(t)
(gcp, grv, gnz), nd = _global_oracle(orc, "heat_t6", 6, w)
assert int(t[0].item()) == len(grv), (t, len(grv))
assert abs(t[1].item() - np.abs(gnz).sum()) <= 1e-9 * np.abs(gnz).sum()
dist.barrier(); dist.destroy_process_group()
sys.stdout.write("done%d\\n" % r); sys.stdout.flush()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "done0" in r.stdout and "done1" in r.stdout


# ---- band-only generation (BASELINE config 5 strong scaling: no global arrays on any rank) ---------------------------
def _owned_block(cp, rv, nz, firsts, lasts):
    cols = np.concatenate([np.arange(f, l + 1) for f, l in zip(firsts, lasts)])
    cnt = np.diff(cp)[cols - 1]
    lcp = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int64)
    idx = np.concatenate([np.arange(cp[c - 1] - 1, cp[c] - 1) for c in cols]) if len(cols) else np.zeros(0, np.int64)
    return lcp, rv[idx], nz[idx]


@pytest.mark.parametrize("kind,N,world", [(efg.Q4, 7, 1), (efg.Q4, 8, 3), (efg.Q4, 5, 5), (efg.T3, 6, 4)])
def test_block_bands_reproduce_the_global_problem(oracle, kind, N, world):
    prob = efg.heat_problem(kind, N)
    quad = prob.quad
    gcp, grv, gnz = oracle.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
    blocks, owned, nnz_sum, chk_sum, sq_sum = [], np.zeros(prob.ndofs, dtype=np.int64), 0, 0, 0.0
    for r in range(world):
        b = sh.block_band(kind, N, r, world)
        assert b.ndofs == prob.ndofs and b.nel_global == prob.nel and b.nnz_global == len(grv)
        # the closed-form numbering is the host mirror's numbering, the band's nodes are a contiguous slice of the grid
        j0, j1 = b.rows
        ja = max(j0 - 1, 0)
        sl = slice(ja * (N + 1), ja * (N + 1) + b.xy.shape[0])
        assert np.array_equal(b.dofnums.numpy(), prob.spaces[0].field.dofnums[sl])
        assert np.array_equal(b.xy.numpy(), prob.meshes[0].xy[sl])
        for f, l in zip(b.firsts, b.lasts):
            owned[f - 1: l] += 1
        mesh = efg.Mesh(kind, b.conn.numpy(), b.xy.numpy())
        cp, rv, nz = oracle.assemble(1, quad, mesh, None, [b.dofnums.numpy()], [1.0], b.ndofs, b.ndofs)
        lcp, lrv, lnz = _owned_block(cp, rv, nz, b.firsts, b.lasts)
        blocks.append((b.firsts, b.lasts, lcp, lrv, lnz))
        if kind == efg.Q4:      # grid-derived expectations (what bench.py validates the sharded GPU result against)
            ennz, echk, esq = sh.q4_expected_checksums(N, j0, j1, chunk_rows=3)
            assert ennz == len(lrv)
            got = sh.pattern_checksum(torch.from_numpy(lcp), torch.from_numpy(lrv), b.firsts, b.lasts)
            assert (got - echk) % (1 << 64) == 0
            assert abs(esq - float((lnz ** 2).sum())) <= 1e-12 * esq
            nnz_sum += ennz
    assert np.all(owned == 1)
    cp, rv, nz = sh.merge_blocks(prob.ndofs, blocks)
    assert np.array_equal(cp, gcp) and np.array_equal(rv, grv)
    assert nz.tobytes() == gnz.tobytes()
    if kind == efg.Q4:
        assert nnz_sum == (3 * N + 1) ** 2
