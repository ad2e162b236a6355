#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu --no-others --no-config5 --e2e-steps 2 > gpurun_out/e8_bench.json 2> gpurun_out/e8_bench.err; tail -3 gpurun_out/e8_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e8_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['config']['tile_elems']); print(json.dumps(d['next_rows'], indent=1)[:2500])
"
