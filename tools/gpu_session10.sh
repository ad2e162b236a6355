#!/bin/bash
# final records: default bench line, launch list of the same command, ncu --set full of the numeric kernel
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_heat_t6_n1.json 2> gpurun_out/s10_bench_err.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_launches_bench_heat_t6_N4000.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-warmup 1 > gpurun_out/s10_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r1b_k_tl_numeric_heat_t6_N4000 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers > gpurun_out/s10_ncu.log 2>&1; echo "ncu rc=$?"
python bench.py --workload elasticity_t6 --no-e2e --no-cpu > gpurun_out/r1_bench_elasticity_t6_n1.json 2>> gpurun_out/s10_bench_err.log
python bench.py --workload heat_q4 --no-e2e --no-cpu > gpurun_out/r1_bench_heat_q4_n1.json 2>> gpurun_out/s10_bench_err.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/s10_bench_err.log
ls -la gpurun_out | tail -12
