#!/bin/bash
# Round 2, GPU call L: whole GPU suite on the build with FLAT as a template parameter of the marches and the per-warp
# dt minimum; bench line; one source-level ncu capture of the x1 sweep (stage 1 launch).
mkdir -p gpurun_out
L=gpurun_out/r2l.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 100 python -m pytest tests/test_tmem_gpu.py tests/test_oracle_gpu.py -q -x --timeout 60
if grep -q "rc=0" <(tail -1 $L); then
  run 900 python -m pytest tests -m gpu -q --timeout 150 -rfEs
  timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
  echo "bench rc=$?" >> $L; tail -c 600 gpurun_out/r2l_bench.err >> $L
  IDFX_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pencilRow -s 7 -c 1 \
      -o /tmp/r2l_x1 python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2l_ncu.log 2>&1
  echo "ncu rc=$?" >> $L
  ls -la /tmp/r2l_x1.ncu-rep >> $L 2>&1
  sz=$(stat -c %s /tmp/r2l_x1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 45000000 ]; then cp /tmp/r2l_x1.ncu-rep gpurun_out/; fi
  ncu -i /tmp/r2l_x1.ncu-rep --page source --csv > gpurun_out/r2l_x1_source.csv 2>> $L
  gzip -f gpurun_out/r2l_x1_source.csv
  ls -la gpurun_out >> $L
fi
grep -v "^$" $L | cut -c1-600 | tail -80
