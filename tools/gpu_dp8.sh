#!/bin/bash
# N-GPU check of the data-parallel step (in-graph NVLink all-reduce): bench lines at C5 and C2; usage: tools/gpu_dp8.sh N
n=${1:-8}
mkdir -p gpurun_out
for wl in C5 C2; do
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline > gpurun_out/dp${n}_bench_${wl}.json 2> gpurun_out/dp${n}_bench_${wl}.err
python - gpurun_out/dp${n}_bench_${wl}.json $wl $n <<'PY'
import sys, json
ok = False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); ok = True
        print(sys.argv[2], 'N', sys.argv[3], 'value', int(d['value']), 'ms', round(d['ms_per_step'], 4), d['config'].get('grad_sync'), 'loss', d.get('final_loss'))
if not ok:
    print(sys.argv[2], 'NO LINE')
PY
tail -3 gpurun_out/dp${n}_bench_${wl}.err | cut -c1-300
done
