#!/usr/bin/env python
"""A/B of kernel builds: one process per library (EFG_LIB), all workloads timed in it with device-generated meshes.
usage: python tools/ab2.py lib_a.so lib_b.so:stokes_gen=64,elasticity_t6=112 ...   (name=te forces a tile size)
internal: python tools/ab2.py --one"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WL = [("heat_t6", 4000), ("elasticity_t6", 2000), ("stokes_gen", 1000), ("heat_q4", 5792), ("heat_t3", 4000)]
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch
    import elfel_jl_b200 as efg
    from elfel_jl_b200 import _lib
    import bench
    peak, _ = bench.load_peak()
    out = []
    for wl, n in WL:
        r = bench.bench_device_problem(torch, efg, _lib, 0, wl, n, 20, 5, peak)
        out.append(f"{wl} {r['ms_per_step']:.3f} ms ({r['roofline']['frac']:.3f}) te={r['tile_elems']}")
    print(" | ".join(out), flush=True)
else:
    for spec in sys.argv[1:]:
        lib, _, tes = spec.partition(":")
        env = dict(os.environ, EFG_LIB=os.path.abspath(lib))
        for kv in filter(None, tes.split(",")):
            env["EFG_BENCH_TE_" + kv.split("=")[0].upper()] = kv.split("=")[1]
        r = subprocess.run([sys.executable, __file__, "--one"], capture_output=True, text=True, env=env)
        print(f"{os.path.basename(spec):40s}", (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
