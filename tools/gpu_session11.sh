#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
N=elfel.jl_b200/libelfelgpu.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s11_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s11_tests.log
{
for wl in elasticity_t6 stokes_gen; do
timeout 300 tools/ab.sh "--workload $wl --no-callers" $A/lib_base.so $N
done
} 2>&1 | tee gpurun_out/s11_ab.log
python bench.py --no-cpu > gpurun_out/s11_bench_heat_t6.json 2> gpurun_out/s11_bench_err.log; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s11_bench_heat_t6.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['all_calls_ms'], d['e2e']['symbolic_ms_per_call'])"
