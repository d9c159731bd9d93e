#!/bin/bash
# A/B of environment switches for the N-GPU step inside ONE gpurun call: tools/ab_env_dp.sh N WORKLOAD "VAR=a" "VAR=b" ...
n=$1; wl=$2; shift; shift
for rep in 1 2; do for kv in "$@"; do
env $kv timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('$wl N=$n', '$kv', 'rep $rep', 'ms/step', round(d['ms_per_step'], 4), 'cells/s', int(d['value']))"
done; done
