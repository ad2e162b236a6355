#!/usr/bin/env python
"""heat FEH1_T4 on an N^3 block through the tiled path: numeric time and tile statistics (target of an ncu capture).
usage: python tools/t4_probe.py [N] [tile_elems]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
te = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prob = efg.heat_problem(efg.T4, N)
eng = efg.Engine(0)
if te:
    eng.set_option(_lib.OPT_TILE_ELEMS, te)
efg.load_problem(eng, prob)
eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", 0))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(5):
    eng.numeric(prob.form.params())
eng.synchronize()
with torch.cuda.stream(stream):
    ev[0].record(stream)
    for _ in range(20):
        eng.numeric(prob.form.params())
    ev[1].record(stream)
eng.synchronize()
nel = prob.meshes[0].nel
nt = int(eng.stat(_lib.STAT_NTILES))
print(f"T4 N={N}: {ev[0].elapsed_time(ev[1]) / 20:.4f} ms, path {int(eng.stat(_lib.STAT_PATH))}, tiles {nt}, elements/tile {nel / max(nt, 1):.0f}, "
      f"halo factor {eng.stat(_lib.STAT_TILE_ELEMS) / nel:.3f}")
