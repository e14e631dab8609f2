#!/bin/bash
# Round 2, GPU call F: x1 sweep with TMA boxes, x2/x3 marches with the TMEM carry-over.  Quick parity first, then
# the A/B legs, the bench line, and the whole GPU suite last (per-test time-out) -- every step under its own time-out.
mkdir -p gpurun_out
L=gpurun_out/r2f.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 100 python -m pytest tests/test_tmem_gpu.py -q
run 200 python -m pytest tests/test_oracle_gpu.py tests/test_parity_gpu.py -q -x --timeout 60 -k "not full"
if grep -q "rc=0" <(tail -1 $L); then
  {
    timeout 200 bash tools/ab_run.sh row-tma 512
    timeout 200 bash tools/ab_run.sh row-no-tma 512 IDFX_NO_TMA=1
    timeout 200 bash tools/ab_run.sh row4-tma 512 IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_row4.so
    timeout 200 bash tools/ab_run.sh row4-no-tma 512 IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_row4.so IDFX_NO_TMA=1
  } > gpurun_out/r2f_ab.log 2>&1
  cat gpurun_out/r2f_ab.log >> $L
  timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
  echo "bench rc=$?" >> $L; tail -c 600 gpurun_out/r2f_bench.err >> $L
  run 800 python -m pytest tests -m gpu -q --timeout 150 -rfEs
fi
grep -v "^$" $L | cut -c1-400 | tail -90
