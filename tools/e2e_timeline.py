"""Wall-clock stamps of the overlapped reference-facing sequence (pinned host buffers), config 2."""
import sys, time, os
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
fid, quad, params = prob.form.form_id, prob.quad, prob.form.params()
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.set_mesh(0, m.kind, conn, xy); eng.set_space(0, 0, dof); eng.start(prob.ndofs, prob.ndofs)
    t1 = time.perf_counter()
    nnz = eng.pattern(fid, quad)
    t2 = time.perf_counter()
    if out is None:
        out = (torch.empty(prob.ndofs + 1, dtype=torch.int64).pin_memory(), torch.empty(nnz, dtype=torch.int64).pin_memory(),
               torch.empty(nnz, dtype=torch.float64).pin_memory())
        t2 = time.perf_counter()
    eng.fetch_pattern_async(out[0], out[1])
    t3 = time.perf_counter()
    eng.numeric(params)
    t4 = time.perf_counter()
    eng.synchronize()
    t5 = time.perf_counter()
    eng.fetch_csc(None, None, out[2])
    t6 = time.perf_counter()
    print(f"iter {it}: load {1e3*(t1-t0):6.1f} pattern {1e3*(t2-t1):6.1f} fetch_async {1e3*(t3-t2):5.1f} numeric(call) {1e3*(t4-t3):6.1f} sync {1e3*(t5-t4):5.1f} fetch {1e3*(t6-t5):6.1f} total {1e3*(t6-t0):7.1f}  (symbolic events {eng.stat(_lib.STAT_SYMBOLIC_MS):6.1f})", flush=True)
if os.environ.get("EFG_NOFETCH"):
    for it in range(3):
        t0 = time.perf_counter()
        eng.set_mesh(0, m.kind, conn, xy); eng.set_space(0, 0, dof); eng.start(prob.ndofs, prob.ndofs)
        nnz = eng.pattern(fid, quad); t2 = time.perf_counter()
        eng.numeric(params); eng.synchronize(); t4 = time.perf_counter()
        print(f"no fetch: numeric(call)+sync {1e3*(t4-t2):6.1f}")
