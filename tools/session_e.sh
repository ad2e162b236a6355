#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/e2e_timeline.py 2>&1 | tail -4
python tools/ab2.py elfel.jl_b200/libelfelgpu.so tools/ab/lib_m5b128.so tools/ab/lib_m4b192.so 2>&1 | tee gpurun_out/e4_ab.log
python bench.py --no-cpu --no-others --no-config5 --no-callers > gpurun_out/e4_bench.json 2> gpurun_out/e4_bench.err; tail -3 gpurun_out/e4_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e4_bench.json')); e=d['e2e']; print(e['ms_per_step'], e['all_calls_ms'], e['plain_sequence_ms'], e['e2e_first_call']['ms'], e['e2e_reassembly']['ms'], d['ms_per_step'], e['symbolic_ms_per_call'])
"
