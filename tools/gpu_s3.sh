#!/bin/bash
# round-2 state check on one B200: GPU tests, the default bench line, launch lists and ncu captures of the likelihood kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 600 gpurun_out/s3_bench.err
for wl in C2 C5; do
python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-configs --no-cpu-baseline --no-graph > gpurun_out/s3_plain_$wl.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-configs --no-cpu-baseline --no-graph > gpurun_out/s3_ncu_$wl.log 2>&1
python tools/nb_profile_run.py $wl > gpurun_out/s3_prof_plain_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nb_tc -s 8 -c 4 -o gpurun_out/s3_nb_$wl python tools/nb_profile_run.py $wl > gpurun_out/s3_ncufull_$wl.log 2>&1
done
ls -la gpurun_out
