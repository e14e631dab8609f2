#!/bin/bash
# Round 2, GPU call D: TMA with even box starts.  Self-tests, quick parity with / without TMA, the whole GPU suite
# (per-test time-out), then the A/B legs -- every step under its own time-out.
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 120 python -m pytest tests/test_tmem_gpu.py -q
run 200 python -m pytest tests/test_oracle_gpu.py -q -x --timeout 60
if grep -q "rc=0" <(tail -1 $L); then
  run 900 python -m pytest tests -m gpu -q --timeout 150 -rfEs
  {
    timeout 250 bash tools/ab_run.sh tma+tmem 512
    timeout 250 bash tools/ab_run.sh no-tma 512 IDFX_NO_TMA=1
    timeout 250 bash tools/ab_run.sh row4 512 IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_row4.so
    timeout 250 bash tools/ab_run.sh row4-no-tma 512 IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_row4.so IDFX_NO_TMA=1
  } > gpurun_out/r2d_ab.log 2>&1
  cat gpurun_out/r2d_ab.log >> $L
  timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
  echo "bench rc=$?" >> $L; tail -c 600 gpurun_out/r2d_bench.err >> $L
fi
grep -v "^$" $L | tail -80
