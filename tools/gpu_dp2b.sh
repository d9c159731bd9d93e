#!/bin/bash
# 2-GPU check: data-parallel tests + bench lines with the in-graph NVLink all-reduce (C5, C2) and the single-GPU lines beside them
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_nccl.py -x -q -s > gpurun_out/dp2b_pytest.log 2>&1; tail -3 gpurun_out/dp2b_pytest.log; grep "max rel" gpurun_out/dp2b_pytest.log
bash tools/gpu_dp8.sh 2
bash tools/ab_env.sh C5 X=1 | head -1
bash tools/ab_env.sh C2 X=1 | head -1
