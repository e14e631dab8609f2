#!/bin/bash
# Round 2, GPU call K: A/B of (1) shock-flattening code in the sweeps, (2) the dt reduction of the stage-0 epilogue
# per warp without a barrier vs per CTA; plus the two C++-host vector-potential cases that failed in call J.
mkdir -p gpurun_out
L=gpurun_out/r2k.log
: > $L
P=$PWD/idefix_b200
timeout 200 bash tools/ab_run.sh flat+ctadt 512 IDEFIX_B200_LIB=$P/libidefix_b200_flat.so >> $L 2>&1
timeout 200 bash tools/ab_run.sh noflat+ctadt 512 IDEFIX_B200_LIB=$P/libidefix_b200_noflat.so >> $L 2>&1
timeout 200 bash tools/ab_run.sh flat+warpdt 512 IDEFIX_B200_LIB=$P/libidefix_b200_warpdt.so >> $L 2>&1
timeout 200 bash tools/ab_run.sh full-lib 512 >> $L 2>&1
echo "=== vp host" >> $L
timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -q --timeout 150 -k "vector_potential" >> $L 2>&1
echo "rc=$?" >> $L
grep -v "^$" $L | cut -c1-400 | tail -60
