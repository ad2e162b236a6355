#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu --no-config5 --no-e2e --no-callers --no-others --steps 5 > /dev/null 2>&1
python - <<'P'
import sys, json
sys.path.insert(0, '.')
import torch, bench
import elfel_jl_b200 as efg
from elfel_jl_b200 import _lib
peak, _ = bench.load_peak()
for r in bench.bench_widened_rows(torch, efg, _lib, 0, 20, 5, peak):
    print(json.dumps(r))
P
