#!/bin/bash
# evidence capture on one B200 (no tests): ncu launch list of C5 training steps + `ncu --set full` of the likelihood sweeps
tag=${1:-prof}
mkdir -p gpurun_out
SPV_PROFILE_RANGE=1 python bench.py --workload C5 --steps 2 --warmup 3 --no-e2e --no-configs --no-cpu-baseline --no-graph > gpurun_out/${tag}_plain_C5.log 2>&1 &&
SPV_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv --log-file gpurun_out/${tag}_launches_C5.csv python bench.py --workload C5 --steps 2 --warmup 3 --no-e2e --no-configs --no-cpu-baseline --no-graph > gpurun_out/${tag}_ncu_C5.log 2>&1
python tools/nb_profile_run.py C5 > gpurun_out/${tag}_prof_plain_C5.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"nb_tc_(fwd|bwd|train|stats)" -c 6 -o gpurun_out/${tag}_nb_C5 python tools/nb_profile_run.py C5 > gpurun_out/${tag}_ncufull_C5.log 2>&1
ls -la gpurun_out | grep ${tag}
