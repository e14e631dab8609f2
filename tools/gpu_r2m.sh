#!/bin/bash
# Round 2, GPU call M: A/B of the instruction-count trims found with the ncu source page of the x1 sweep
# (FLAT as a template parameter of the row kernel, stepped cell / face indices, unguarded limiter operands).
mkdir -p gpurun_out
L=gpurun_out/r2m.log
: > $L
P=$PWD/idefix_b200
timeout 200 bash tools/ab_run.sh committed 512 >> $L 2>&1
timeout 200 bash tools/ab_run.sh t1-all 512 IDEFIX_B200_LIB=$P/libidefix_b200_t1.so >> $L 2>&1
timeout 200 bash tools/ab_run.sh t4-guards-kept 512 IDEFIX_B200_LIB=$P/libidefix_b200_t4.so >> $L 2>&1
timeout 200 bash tools/ab_run.sh committed-again 512 >> $L 2>&1
grep -v "^$" $L | cut -c1-400 | tail -40
