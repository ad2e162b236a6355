#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_vector.py -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-others --no-config5 --no-e2e > gpurun_out/e9_bench.json 2> gpurun_out/e9_bench.err; tail -3 gpurun_out/e9_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e9_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['config']['tile_elems']); print(json.dumps(d['next_rows']['fused_matrix_and_load_vector'], indent=1))
"
for wl in heat_q4 heat_t3; do python bench.py --workload $wl --n 4000 --no-cpu --no-others --no-config5 --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', d['ms_per_step'], d['next_rows']['load_vector']['ms'], json.dumps(d['next_rows']['fused_matrix_and_load_vector']))
"; done
