#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-others --no-config5 --no-callers --steps 10 > gpurun_out/e12_bench.json 2> gpurun_out/e12_bench.err; tail -3 gpurun_out/e12_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e12_bench.json')); e=d['e2e']; print(e['ms_per_step'], e['all_calls_ms'], e['plain_sequence_ms'], e['e2e_first_call'], e['e2e_reassembly']['ms'], d['ms_per_step'])"
python tools/ab2.py elfel.jl_b200/libelfelgpu.so 2>&1 | tee gpurun_out/e13_ab.log
