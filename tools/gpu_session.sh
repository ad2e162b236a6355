#!/bin/bash
# One GPU session: parity tests, A/B of kernel builds, the default bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/s_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - t0 )) s"; tail -14 gpurun_out/s_tests.log
for wl in heat_t6 elasticity_t6; do
    tools/ab.sh "--workload $wl --no-callers" tools/ab/lib_base.so elfel.jl_b200/libelfelgpu.so
done 2>&1 | tee gpurun_out/s_ab.log
python bench.py > gpurun_out/s_bench_heat_t6.json 2> gpurun_out/s_bench_err.log
echo "bench rc=$?"; tail -c 3000 gpurun_out/s_bench_heat_t6.json; tail -5 gpurun_out/s_bench_err.log
echo "total $(( $(date +%s) - t0 )) s"
