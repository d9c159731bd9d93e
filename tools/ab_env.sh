#!/bin/bash
# A/B of environment switches inside ONE gpurun call (same box, same clocks): tools/ab_env.sh WORKLOAD "VAR=a" "VAR=b" ...
wl=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do for kv in "$@"; do
env $kv python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl', '$kv', 'rep $rep', 'ms/step', round(d['ms_per_step'], 4), 'cells/s', int(d['value']))"
done; done
