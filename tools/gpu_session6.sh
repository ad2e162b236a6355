#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
N=elfel.jl_b200/libelfelgpu.so
{
tools/ab.sh "--workload heat_t6 --no-callers" $A/lib_base.so $N
for te in 192 208 216 224; do tools/ab.sh "--workload heat_t6 --no-callers --tile-elems $te" $N; done
tools/ab.sh "--workload heat_q4 --no-callers" $N
tools/ab.sh "--workload heat_q4 --no-callers --tile-elems 448" $N
tools/ab.sh "--workload heat_q4 --no-callers --tile-elems 384" $N
tools/ab.sh "--workload elasticity_t6 --no-callers" $N
} 2>&1 | tee gpurun_out/s6_ab.log
