#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
A=tools/ab
{
tools/ab.sh "--workload heat_t6 --no-callers" $A/lib_base.so $A/lib_c.so $A/lib_u8.so
tools/ab.sh "--workload heat_t6 --no-callers --tile-elems 144" $A/lib_b224.so
tools/ab.sh "--workload heat_t6 --no-callers --tile-elems 160" $A/lib_b224.so $A/lib_b256m3.so
tools/ab.sh "--workload heat_t6 --no-callers --tile-elems 168" $A/lib_b224.so
tools/ab.sh "--workload elasticity_t6 --no-callers" $A/lib_base.so $A/lib_u8.so
tools/ab.sh "--workload heat_q4 --no-callers" $A/lib_base.so $A/lib_c.so $A/lib_u8.so
} 2>&1 | tee gpurun_out/s3_ab.log
echo "total $(( $(date +%s) - t0 )) s"
