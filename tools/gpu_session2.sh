#!/bin/bash
# A/B of kernel builds + one ncu --set full capture of the numeric kernel per build (heat T6 N=4000).
mkdir -p gpurun_out
t0=$(date +%s)
for wl in heat_t6 elasticity_t6; do
    tools/ab.sh "--workload $wl --no-callers" $LIBS
done 2>&1 | tee gpurun_out/s2_ab.log
for lib in $NCU_LIBS; do
    tag=$(basename $lib .so)
    EFG_LIB=$lib timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tl_numeric --launch-skip 4 --launch-count 1 \
        -o gpurun_out/s2_${tag}_heat_t6 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-callers ${NCU_ARGS:-} > gpurun_out/s2_ncu_${tag}.log 2>&1
    echo "ncu $tag rc=$?"; ls -la gpurun_out/s2_${tag}_heat_t6.ncu-rep
done
echo "total $(( $(date +%s) - t0 )) s"
