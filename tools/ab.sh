#!/bin/bash
# A/B kernel variants in one GPU session: tools/ab.sh "<bench args>" lib1.so lib2.so ...
args="$1"; shift
for lib in "$@"; do
    EFG_LIB=$lib python bench.py --steps 20 --no-cpu --no-e2e $args 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$args', 'ms=%.3f frac=%.3f halo=%.3f' % (d['ms_per_step'], d['roofline']['frac'], d['config']['halo_factor']))"
done
