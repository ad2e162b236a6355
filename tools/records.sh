#!/bin/bash
# round records: per-workload bench lines (N=1), launch list and ncu --set full captures -> gpurun_out/ (copy what is to be kept into profiles/)
mkdir -p gpurun_out
python bench.py > gpurun_out/r1_bench_heat_t6_n1.json 2> gpurun_out/s12_err.log; echo "bench rc=$?"
for wl in elasticity_t6 stokes_gen heat_q4; do python bench.py --workload $wl --no-e2e --no-cpu > gpurun_out/r1_bench_${wl}_n1.json 2>> gpurun_out/s12_err.log; done
python bench.py --workload heat_t3 --n 4000 --no-e2e --no-cpu > gpurun_out/r1_bench_heat_t3_N4000_n1.json 2>> gpurun_out/s12_err.log
python bench.py --workload heat_t3 --no-e2e --no-cpu > gpurun_out/r1_bench_heat_t3_N100_n1.json 2>> gpurun_out/s12_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_launches_bench_heat_t6_N4000.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-warmup 1 > gpurun_out/s12_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r1b_k_tl_numeric_heat_t6_N4000 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers > gpurun_out/s12_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r1b_k_tl_numeric_elasticity_t6_N2000 -f python bench.py --workload elasticity_t6 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/s12_ncu2.log 2>&1; echo "ncu2 rc=$?"
