#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s5_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s5_tests.log
{
tools/ab.sh "--workload heat_t6 --no-callers" $A/lib_base.so elfel.jl_b200/libelfelgpu.so
tools/ab.sh "--workload heat_q4 --no-callers" $A/lib_base.so elfel.jl_b200/libelfelgpu.so
tools/ab.sh "--workload heat_t3 --n 4000 --no-callers" $A/lib_base.so elfel.jl_b200/libelfelgpu.so
} 2>&1 | tee gpurun_out/s5_ab.log
