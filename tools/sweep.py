#!/usr/bin/env python
"""Run bench.py over a list of option sets and print one compact line each.
usage: python tools/sweep.py "--tile-elems 160" "--tile-elems 192" ...   (common args via SWEEP_COMMON env)"""
import json, os, subprocess, sys
common = os.environ.get("SWEEP_COMMON", "--steps 20 --no-cpu --no-e2e").split()
for opt in sys.argv[1:]:
    r = subprocess.run([sys.executable, "bench.py"] + common + opt.split(), capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{opt:40s} ms={d['ms_per_step']:.3f} frac={d['roofline']['frac']:.3f} halo={d['config']['halo_factor']:.3f} "
              f"tiles={d['config']['tiles']} sym_ms={d['phases']['symbolic_ms']:.1f} Gelem/s={d['value']/1e9:.3f}", flush=True)
    except Exception as e:
        print(opt, "FAILED", (r.stdout + r.stderr)[-400:], flush=True)
