#!/bin/bash
# GPU session: all GPU tests, default bench line, launch list, ncu full capture of the three main kernels
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/e1_tests.log 2>&1
echo "tests rc=$? $(( $(date +%s) - t0 )) s"; tail -14 gpurun_out/e1_tests.log
python bench.py > gpurun_out/e1_bench.json 2> gpurun_out/e1_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/e1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_heat_t6_N4000.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-warmup 1 --no-others > gpurun_out/e1_launch.log 2>&1; echo "launch list rc=$?"
for wl in heat_t6 elasticity_t6 stokes_gen; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r2_k_tl_numeric_${wl} -f python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers --no-others > gpurun_out/e1_ncu_$wl.log 2>&1; echo "ncu $wl rc=$?"
done
echo "total $(( $(date +%s) - t0 )) s"
