"""Break the reference-facing call sequence (pinned host buffers) into its parts."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
prob = efg.heat_problem(efg.T6, n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
m = prob.meshes[0]
conn, xy, dof = pin(m.conn), pin(m.xy), pin(prob.spaces[0].field.dofnums)
eng = efg.Engine(0)
out = None
for it in range(5):
    t0 = time.perf_counter()
    eng.set_mesh(0, m.kind, conn, xy); eng.set_space(0, 0, dof); eng.start(prob.ndofs, prob.ndofs); eng.synchronize()
    t1 = time.perf_counter()
    nnz = eng.symbolic(prob.form.form_id, prob.quad); eng.synchronize()
    t2 = time.perf_counter()
    eng.numeric(prob.form.params()); eng.synchronize()
    t3 = time.perf_counter()
    if out is None:
        out = (torch.empty(prob.ndofs + 1, dtype=torch.int64).pin_memory(), torch.empty(nnz, dtype=torch.int64).pin_memory(),
               torch.empty(nnz, dtype=torch.float64).pin_memory())
        t3 = time.perf_counter()
    eng.fetch_csc(*out); eng.synchronize()
    t4 = time.perf_counter()
    print(f"iter {it}: load {1e3*(t1-t0):6.1f}  symbolic {1e3*(t2-t1):6.1f} (events {eng.stat(_lib.STAT_SYMBOLIC_MS):6.1f})  numeric {1e3*(t3-t2):5.1f}  fetch {1e3*(t4-t3):6.1f}  total {1e3*(t4-t0):7.1f} ms", flush=True)
