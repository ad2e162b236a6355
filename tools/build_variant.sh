#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>  ->  tools/ab/lib_<name>.so  (+ /tmp/build_<name>.log with the ptxas report)
name=$1; shift
mkdir -p "$(dirname "$0")/ab"
EFG_NVCC_EXTRA="$*" EFG_LIB="$(cd "$(dirname "$0")" && pwd)/ab/lib_$name.so" python -c "
from elfel_jl_b200 import _lib
_lib.build(force=True, verbose=True)" > /tmp/build_$name.log 2>&1
echo "built $name rc=$?"
