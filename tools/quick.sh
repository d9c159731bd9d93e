#!/bin/bash
# quick GPU check used during kernel work: the GPU test suite, then device-resident bench lines for C5 and C2
tag=${1:-q}
if [ "$2" != "nopytest" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log; fi
for wl in C5 C2; do
python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-configs --no-cpu-baseline > gpurun_out/${tag}_bench_$wl.log 2> gpurun_out/${tag}_bench_$wl.err
python - gpurun_out/${tag}_bench_$wl.log $wl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(sys.argv[2], 'value', int(d['value']), 'ms', round(d['ms_per_step'], 4), 'nb fwd-only', round(r['avg_launch_ms'], 4), 'other', [(o['kernel'][:14], o['bound'], round(o['avg_launch_ms'], 4)) for o in r['other']], 'loss', d.get('final_loss'))
PY
done
