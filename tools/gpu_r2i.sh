#!/bin/bash
# Round 2, GPU call I: dt reduction folded into the stage-0 epilogue (parity + timing), then the ncu evidence of the
# round: launch list of the bench command, --set full of the six kernels of stage 0 and of stage 1 (summarised on the
# box: the report itself is too large to travel).
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 300 python -m pytest tests/test_oracle_gpu.py tests/test_parity_gpu.py tests/test_p2p_gpu.py -q -x --timeout 100 -k "not full"
if grep -q "rc=0" <(tail -1 $L); then
  timeout 200 bash tools/ab_run.sh default 512 >> $L 2>&1
  CMD1="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 1"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches.csv $CMD1 > /tmp/ncu1.log 2>&1
  echo "ncu launches rc=$?" >> $L
  python tools/ncu_summary.py launches /tmp/launches.csv "Round 2, final kernels: ncu launch list at 512^3 (bench.py --steps 2 --warmup 3)" \
    "ncu --metrics gpu__time_duration.sum --clock-control none --csv $CMD1" > gpurun_out/r2i_launches_512cube.md 2>> $L
  CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 1"
  IDFX_NO_GRAPH=1 timeout 800 ncu --set full --clock-control none --import-source on \
    -k regex:'pencil|contactAverage|evolveMagField|stageEpilogue' -s 36 -c 12 -o /tmp/kernels $CMD2 > /tmp/ncu2.log 2>&1
  echo "ncu full rc=$?" >> $L
  tail -3 /tmp/ncu2.log >> $L
  python tools/ncu_summary.py full /tmp/kernels.ncu-rep "Round 2, final kernels: ncu --set full at 512^3, first timed cycle: six kernels of RK stage 0, then six of stage 1" \
    "IDFX_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:'pencil|contactAverage|evolveMagField|stageEpilogue' -s 36 -c 12 $CMD2" \
    > gpurun_out/r2i_kernels_ncu_full_512cube.md 2>> $L
  ncu -i /tmp/kernels.ncu-rep --page raw --csv > gpurun_out/r2i_kernels_raw.csv 2>> $L
  ncu -i /tmp/kernels.ncu-rep --page source --csv --kernel-name regex:pencilRow > gpurun_out/r2i_x1_source.csv 2>> $L
  ncu -i /tmp/kernels.ncu-rep --page source --csv --kernel-name regex:pencilMarch > gpurun_out/r2i_x23_source.csv 2>> $L
  ls -la gpurun_out/r2i_* >> $L
fi
grep -v "^$" $L | cut -c1-400 | tail -50
