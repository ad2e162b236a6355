#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for hw in 0 1; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$hw bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-config5 --no-others --no-callers --host-widen $hw 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('N=$N host_widen=$hw', round(e['ms_per_step'],1), e['all_calls_ms'], 'reasm', round(e['e2e_reassembly']['ms'],1), 'first', round(e['e2e_first_call']['ms'],1), 'plain', round(e['plain_sequence_ms'],1))
"
done
