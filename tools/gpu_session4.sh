#!/bin/bash
mkdir -p gpurun_out
A=tools/ab
{
tools/ab.sh "--workload heat_t6 --no-callers" $A/lib_base.so $A/lib_d0.so $A/lib_d1.so
tools/ab.sh "--workload heat_q4 --no-callers" $A/lib_base.so $A/lib_d1.so
} 2>&1 | tee gpurun_out/s4_ab.log
