#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
{
timeout 200 tools/ab.sh "--workload elasticity_t6 --no-callers" $A/lib_final.so $A/lib_s288.so $A/lib_s320.so
timeout 200 tools/ab.sh "--workload elasticity_t6 --no-callers --tile-elems 56" $A/lib_s288.so $A/lib_s320.so
timeout 200 tools/ab.sh "--workload stokes_gen --no-callers" $A/lib_s288.so
} 2>&1 | tee gpurun_out/s14_ab.log
