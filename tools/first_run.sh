#!/bin/bash
# First GPU call of a round: the first-run tests (written without a GPU), then the verified suite, a bench line
# and a launch list.   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_run.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_first_run_gpu.py -m gpu -q -rxX > gpurun_out/first_run_tests.log 2>&1
tail -40 gpurun_out/first_run_tests.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_zz_first_run_gpu.py > gpurun_out/gpu_suite.log 2>&1
tail -5 gpurun_out/gpu_suite.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.json
