#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/f2_bench.err
python -c "
import json; d=json.load(open('gpurun_out/f2_bench.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['e2e_first_call']['ms'], d['cpu_baseline']['value'])
for r in d['widened_rows']: print(r.get('row'), r.get('workload','')[:50], r.get('ms_per_step'), (r.get('roofline') or {}).get('frac'), r.get('generation_ms'), r.get('error'))
for o in d['other_configs']: print(o['workload'][:40], o['ms_per_step'], o['roofline']['frac'])
print(d['config5']['numeric_ms'], d['config5']['validation']['nnz_exact'])
print(json.dumps(d['next_rows']['fused_matrix_and_load_vector']))
"
t1=$(date +%s)
python bench.py --impl reference > gpurun_out/f2_ref.json 2> gpurun_out/f2_ref.err; echo "ref rc=$? $(( $(date +%s) - t1 )) s"; cat gpurun_out/f2_ref.json | head -c 600
