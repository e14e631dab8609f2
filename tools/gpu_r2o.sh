#!/bin/bash
# Round 2, GPU call O: the polar-axis cases through the C++ host against the reference build's dumps.
mkdir -p gpurun_out
L=gpurun_out/r2o.log
: > $L
timeout 400 python -m pytest tests/test_host_cpp.py -m gpu -q --timeout 150 -k "axis" -rfE 2>&1 | cut -c1-900 > $L
grep -v "^$" $L | grep -v "^E  *\(\[\|\t\|Input\|Grid\|-----\|Fluid\|Time\|Main\)" | tail -70
