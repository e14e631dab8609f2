#!/bin/bash
# Round 2, GPU call A: correctness of the full build (TMEM carry-over in the x2/x3 marches, HD Roe stale flux, dump
# numbering), then A/B of the TMEM variants at 512^3.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 bash tools/ab.sh 512 full notmem tmem5 > gpurun_out/r2a_ab.log 2>&1
cat gpurun_out/r2a_ab.log
