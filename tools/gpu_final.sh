#!/bin/bash
# what the driver runs at round end, on one B200: GPU tests, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; tail -2 gpurun_out/fin_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; tail -6 gpurun_out/fin_smoke.log
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; tail -3 gpurun_out/fin_ref.err
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; tail -4 gpurun_out/fin_bench.err
