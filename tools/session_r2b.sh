#!/bin/bash
# records after the Stokes tail-pass change: all GPU tests, the default bench line, the Stokes gen record and its ncu summary
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/z_bench.err
python bench.py --workload stokes_gen --no-cpu --no-e2e --no-callers --no-others --no-config5 --no-widened > gpurun_out/z_bench_stokes.json 2>> gpurun_out/z_bench.err; echo "stokes rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
    -o /tmp/z_stokes_gen -f python bench.py --workload stokes_gen --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers --no-others --no-config5 --no-widened > gpurun_out/z_ncu_stokes_gen.log 2>&1; echo "ncu rc=$?"
python profiles/ncu_summary.py /tmp/z_stokes_gen.ncu-rep > gpurun_out/r2_ncu_k_tl_numeric_stokes_gen.txt 2>&1
echo "total $(( $(date +%s) - t0 )) s"
