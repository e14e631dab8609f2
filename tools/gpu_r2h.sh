#!/bin/bash
# Round 2, GPU call H: overlap of the ghost-zone work with the interior epilogue (A/B), then the ncu evidence:
# launch list of the bench command and --set full of the six kernels of stage 0 and stage 1.
mkdir -p gpurun_out
L=gpurun_out/r2h.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 300 python -m pytest tests/test_oracle_gpu.py tests/test_parity_gpu.py tests/test_p2p_gpu.py -q -x --timeout 100 -k "not full"
if grep -q "rc=0" <(tail -1 $L); then
  {
    timeout 200 bash tools/ab_run.sh overlap 512
    timeout 200 bash tools/ab_run.sh no-overlap 512 IDFX_NO_OVERLAP=1
    timeout 200 bash tools/ab_run.sh overlap-256 256
    timeout 200 bash tools/ab_run.sh no-overlap-256 256 IDFX_NO_OVERLAP=1
  } > gpurun_out/r2h_ab.log 2>&1
  cat gpurun_out/r2h_ab.log >> $L
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2h_ncu_launches.log 2>&1
  echo "ncu launches rc=$?" >> $L
  IDFX_NO_OVERLAP=1 IDFX_NO_GRAPH=1 timeout 700 ncu --set full --clock-control none --import-source on \
    -k regex:'pencil|contactAverage|evolveMagField|stageEpilogue' -s 36 -c 12 -o gpurun_out/r2h_kernels \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2h_ncu_full.log 2>&1
  echo "ncu full rc=$?" >> $L
fi
grep -v "^$" $L | cut -c1-400 | tail -60
