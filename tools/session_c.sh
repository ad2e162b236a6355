#!/bin/bash
# GPU session: all GPU tests + the e2e record with host-side widening of the row indices (thread-count sweep)
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --no-others --no-config5 --no-callers > gpurun_out/e3_bench.json 2> gpurun_out/e3_bench.err; tail -3 gpurun_out/e3_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e3_bench.json')); e=d['e2e']; print(e['ms_per_step'], e['all_calls_ms'], e['plain_sequence_ms'], e['e2e_first_call']['ms'], e['e2e_reassembly']['ms'], d['ms_per_step'])
"
for t in 1 4 16; do EFG_HOST_THREADS=$t python bench.py --no-cpu --no-others --no-config5 --no-callers --steps 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('threads', $t, e['ms_per_step'], e['plain_sequence_ms'])
"; done
