#!/bin/bash
# 2-GPU check of the data-parallel step: all-reduce kernel + captured-step tests, then bench lines (C2, C5) per synchroniser
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/dp2_topo.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_nccl.py -x -q -s > gpurun_out/dp2_pytest.log 2>&1; tail -5 gpurun_out/dp2_pytest.log
for wl in C2 C5; do for sync in nvlink nccl; do
SPV_DP_SYNC=$sync timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline > gpurun_out/dp2_bench_${wl}_$sync.json 2> gpurun_out/dp2_bench_${wl}_$sync.err
python - gpurun_out/dp2_bench_${wl}_$sync.json $wl $sync <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print(sys.argv[2], sys.argv[3], 'value', int(d['value']), 'ms', round(d['ms_per_step'], 4), d['config'].get('grad_sync'), 'loss', d.get('final_loss'))
PY
done; done
python bench.py --workload C2 --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 1gpu', int(d['value']), round(d['ms_per_step'],4))"
python bench.py --workload C5 --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 1gpu', int(d['value']), round(d['ms_per_step'],4))"
