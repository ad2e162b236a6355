#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
python bench.py --no-cpu --no-config5 --no-e2e --no-callers > gpurun_out/h1_bench.json 2> gpurun_out/h1_bench.err; echo "rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/h1_bench.err
python -c "
import json; d=json.load(open('gpurun_out/h1_bench.json'))
for r in d['widened_rows']: print(json.dumps(r))
"
