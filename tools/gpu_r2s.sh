#!/bin/bash
# Round 2, GPU call S (the last 78 seconds of the budget): the 2.5-D axis configuration through the C++ host; the
# dumps travel back and are compared with the reference build's on the CPU.
mkdir -p gpurun_out/a25
R=$PWD
timeout 45 make -s -C idefix_b200/host PROBLEM=$R/oracle/ref_build/problems/Axis25D MHD=YES ORDER=2 OUT=/tmp/a25_idefix > gpurun_out/a25/make.log 2>&1
echo "make rc=$?"
for c in axis25d_hlld_contact_plm axis25d_roe_uct0_plm_rk3; do
  mkdir -p /tmp/a25_$c && cp tools/a25/$c.ini /tmp/a25_$c/idefix.ini
  (cd /tmp/a25_$c && timeout 15 /tmp/a25_idefix -i idefix.ini -maxcycles 10 > run.log 2>&1; echo "$c rc=$?"; tail -3 run.log | cut -c1-160)
  for n in 0000 0001 0010; do cp /tmp/a25_$c/dump.$n.dmp gpurun_out/a25/$c.c$n.dmp 2>/dev/null; done
done
ls gpurun_out/a25 | wc -l
