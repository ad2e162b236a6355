"""Repeat ingestion + symbolic phase in one process to separate first-call costs from steady state."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
prob = efg.heat_problem(efg.T6, n)
eng = efg.Engine(0)
for it in range(6):
    t0 = time.perf_counter()
    efg.load_problem(eng, prob)
    eng.synchronize()
    t1 = time.perf_counter()
    eng.symbolic(prob.form.form_id, prob.quad)
    t2 = time.perf_counter()
    eng.numeric(prob.form.params()); eng.synchronize()
    t3 = time.perf_counter()
    print(f"iter {it}: load {1e3*(t1-t0):7.1f} ms  symbolic wall {1e3*(t2-t1):7.1f} ms (events {eng.stat(_lib.STAT_SYMBOLIC_MS):7.1f})  numeric {1e3*(t3-t2):6.1f} ms  dev GB {eng.stat(_lib.STAT_DEVICE_BYTES)/1e9:.1f}", flush=True)
