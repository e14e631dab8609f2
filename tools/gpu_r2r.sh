#!/bin/bash
# Round 2, GPU call R (final validation): the whole GPU suite incl. the polar-axis cases, then the bench line.
mkdir -p gpurun_out
L=gpurun_out/r2r.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 290 python -m pytest tests -m gpu -q --timeout 120 -x -rfEs 2>&1 | cut -c1-700 | tail -45 >> $L
echo "=== bench" >> $L
timeout 95 python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?" >> $L; tail -c 400 gpurun_out/r2r_bench.err >> $L
grep -v "^$" $L | tail -60
