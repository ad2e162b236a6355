#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
N=elfel.jl_b200/libelfelgpu.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s9_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s9_tests.log
{
for wl in heat_t6 elasticity_t6 stokes_gen heat_q4 "heat_t3 --n 4000"; do
timeout 300 tools/ab.sh "--workload $wl --no-callers" $A/lib_base.so $N
done
timeout 200 tools/ab.sh "--workload heat_t6 --no-callers --tile-elems 228" $N
} 2>&1 | tee gpurun_out/s9_ab.log
