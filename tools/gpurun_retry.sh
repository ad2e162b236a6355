#!/bin/bash
# tools/gpurun_retry.sh <timeout> <command...>: retry while the pod answers "transient" (nothing is charged for those)
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
    out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
    if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
    echo "$out"; exit 0
done
echo "$out"
