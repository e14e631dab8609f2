#!/bin/bash
# Round 2, GPU call G (2 GPUs): decomposition parity over NCCL and over the peer-to-peer kernels, then the bench at N = 2
# with both transports.
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
run() { local t=$1; shift; echo "=== $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi topo -m >> $L 2>&1
run 500 python -m pytest tests/test_multigpu.py -q -k "two" --timeout 200 -rfE
for halo in p2p nccl; do
  echo "=== bench --gpus 2 --halo $halo" >> $L
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 6 --warmup 3 --halo $halo > gpurun_out/r2g_bench_n2_$halo.json 2> gpurun_out/r2g_bench_n2_$halo.err
  echo "rc=$?" >> $L
  python - >> $L <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_n2_$halo.json").read().strip().splitlines()[-1])
    print("value %.4g ms/step %.2f per_gpu %.4g parity %s stated %s" % (d["value"], d["ms_per_step"], d["per_gpu"], d["parity_ngpu_max_abs"], json.dumps(d.get("c4_as_stated"))[:300]))
except Exception as e:
    print("no bench line:", e)
PY
  tail -c 800 gpurun_out/r2g_bench_n2_$halo.err >> $L
done
grep -v "^$" $L | cut -c1-400 | tail -70
