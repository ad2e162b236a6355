#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
N=elfel.jl_b200/libelfelgpu.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s7_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s7_tests.log
{
timeout 300 tools/ab.sh "--workload heat_t6 --no-callers" $A/lib_base.so $N
for te in 208 224 232; do timeout 200 tools/ab.sh "--workload heat_t6 --no-callers --tile-elems $te" $N; done
timeout 200 tools/ab.sh "--workload heat_q4 --no-callers" $N
timeout 200 tools/ab.sh "--workload heat_t3 --n 4000 --no-callers" $N
} 2>&1 | tee gpurun_out/s7_ab.log
