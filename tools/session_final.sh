#!/bin/bash
# final records of the round: all GPU tests, the default bench line, the reference arm is skipped here (unchanged), launch list, ncu summaries
mkdir -p gpurun_out
t0=$(date +%s)
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/z_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench_heat_t6_N4000.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-warmup 1 --no-others --no-config5 > gpurun_out/z_launch.log 2>&1; echo "launch list rc=$?"
for wl in heat_t6 elasticity_t6 stokes_gen heat_q4; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o /tmp/z_${wl} -f python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers --no-others --no-config5 > gpurun_out/z_ncu_$wl.log 2>&1; echo "ncu $wl rc=$?"
python profiles/ncu_summary.py /tmp/z_${wl}.ncu-rep > gpurun_out/r2_ncu_k_tl_numeric_${wl}.txt 2>&1
done
echo "total $(( $(date +%s) - t0 )) s"
