#!/bin/bash
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|MHz"
run() { python bench.py --no-cpu --no-others --no-config5 --no-callers --steps 5 --e2e-steps 3 --e2e-warmup 2 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$LABEL', round(e['ms_per_step'],1), e['all_calls_ms'], 'reasm', round(e['e2e_reassembly']['ms'],1), 'first', round(e['e2e_first_call']['ms'],1))"; }
LABEL=default run
LABEL=no-defer run --no-defer-xy
LABEL=threads4 EFG_HOST_THREADS=4 run
LABEL=threads12 EFG_HOST_THREADS=12 run
LABEL=default-again run
