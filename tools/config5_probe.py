"""Probe: BASELINE config 5 (heat FEH1_Q4, N x N cells) on ONE GPU, mesh generated on the device, device pointers through the ABI.
usage: python tools/config5_probe.py [N=16384] [rows=N]   (rows < N: only the first `rows` cell rows = a per-GPU share)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
from elfel_jl_b200.sharding import q4_band

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
t0 = time.perf_counter()
b = q4_band(N, rank, world, dev)
torch.cuda.synchronize()
print(f"band mesh: nel={b.conn.shape[0]} nnodes={b.xy.shape[0]} ranges={list(zip(b.firsts, b.lasts))} gen {time.perf_counter()-t0:.2f} s", flush=True)
eng = efg.Engine(0)
for it in range(2):
    t0 = time.perf_counter()
    eng.set_mesh(0, efg.Q4, b.conn, b.xy); eng.set_space(0, 0, b.dofnums); eng.start(b.ndofs, b.ndofs)
    eng.set_column_ranges(b.firsts, b.lasts)
    eng.synchronize()
    t1 = time.perf_counter()
    nnz = eng.symbolic(_lib.FORM_HEAT, 2)
    t2 = time.perf_counter()
    for _ in range(3):
        eng.numeric([1.0])
    eng.synchronize()
    print(f"iter {it}: load {1e3*(t1-t0):.1f} ms symbolic {1e3*(t2-t1):.1f} ms (events {eng.stat(_lib.STAT_SYMBOLIC_MS):.1f}) numeric {eng.stat(_lib.STAT_NUMERIC_MS):.3f} ms "
          f"nnz={nnz} expect_global={(3*N+1)**2} tiles={int(eng.stat(_lib.STAT_NTILES))} halo={eng.stat(_lib.STAT_TILE_ELEMS)/b.conn.shape[0]:.3f} "
          f"dev GB {eng.stat(_lib.STAT_DEVICE_BYTES)/1e9:.1f} torch GB {torch.cuda.memory_allocated()/1e9:.1f}", flush=True)
