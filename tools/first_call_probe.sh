#!/bin/bash
# e2e_first_call of three consecutive processes on one fresh box (is the first one slower because the library file is cold?)
for i in 1 2 3; do
  python bench.py --no-cpu --no-callers --no-others --no-config5 --no-widened --e2e-steps 1 --e2e-warmup 1 --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=r['e2e']
print('process $i: first call %.1f ms (symbolic %.1f, efg_create %.1f), warm %.1f ms' % (e['e2e_first_call']['ms'], e['e2e_first_call']['symbolic_ms'], e['e2e_first_call'].get('efg_create_ms', -1), e['ms_per_step']))"
done
