#!/bin/bash
# Round 2, GPU call J: the whole GPU suite on the build with shock flattening, vector potential, stage hooks, the dt
# reduction folded into the epilogue; then the bench line.
mkdir -p gpurun_out
L=gpurun_out/r2j.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 100 python -m pytest tests/test_tmem_gpu.py tests/test_oracle_gpu.py -q -x --timeout 60
if grep -q "rc=0" <(tail -1 $L); then
  run 900 python -m pytest tests -m gpu -q --timeout 150 -rfEs
  timeout 200 bash tools/ab_run.sh default 512 >> $L 2>&1
  timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
  echo "bench rc=$?" >> $L; tail -c 600 gpurun_out/r2j_bench.err >> $L
fi
grep -v "^$" $L | cut -c1-600 | tail -120
