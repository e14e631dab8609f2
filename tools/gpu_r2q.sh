#!/bin/bash
# Round 2, GPU call Q: axis diagnostics over grid sizes / RK orders (the axisdiag_* cases existed only for this call:
# the answer was the mirrored theta spacing of the ghost cells, gridHost.cpp:246-262) and summation-order switches.
mkdir -p gpurun_out
L=gpurun_out/r2q.log
timeout 200 python tools/axis_diag.py axisdiag_v1 axisdiag_v2 axisdiag_v3 axisdiag_v4 axisdiag_v5 axisdiag_v6 axisdiag_v7 > $L 2>&1
for m in 1 2; do echo "##### IDFX_AXIS_SUM=$m" >> $L; IDFX_AXIS_SUM=$m timeout 100 python tools/axis_diag.py axisdiag_v6 axisdiag_v7 >> $L 2>&1; done
grep -v "equal$" $L | cut -c1-300 | tail -70
