#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/g_bench_n$N.err
python -c "
import json
for l in open('gpurun_out/g_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value']); c=d['config5']; print(c.get('numeric_ms'), c.get('value'), c.get('rank_numeric_ms'), c.get('validation',{}).get('nnz_exact'), c.get('validation',{}).get('pattern_checksum_ok'), c.get('error'))
"
