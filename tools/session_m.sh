#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/ab2.py tools/ab/lib_prev.so elfel.jl_b200/libelfelgpu.so 2>&1 | tee gpurun_out/e16_ab.log
timeout 900 python bench.py --verify > gpurun_out/e16_verify.json 2> gpurun_out/e16_verify.err; tail -2 gpurun_out/e16_verify.err
python -c "
import json; d=json.load(open('gpurun_out/e16_verify.json'))
for r in d['verify']: print(r['workload'][:30], r['pattern_bit_exact'], r['default_fp'], r['strict_fp'])
"
