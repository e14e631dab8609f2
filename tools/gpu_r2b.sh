#!/bin/bash
# Round 2, GPU call B: full suite on the TMA + TMEM build, A/B of the TMA path and of 4 CTAs/SM in the x1 sweep,
# the new bench line, and one ncu --set full capture of the three sweeps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
{
  bash tools/ab_run.sh tma+tmem 512
  bash tools/ab_run.sh no-tma 512 IDFX_NO_TMA=1
  bash tools/ab_run.sh row4 512 IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_row4.so
} > gpurun_out/r2b_ab.log 2>&1
cat gpurun_out/r2b_ab.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 1500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pencil' -s 18 -c 3 -o gpurun_out/r2b_sweeps \
  python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
