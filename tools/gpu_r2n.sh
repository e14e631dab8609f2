#!/bin/bash
# Round 2, GPU call N (8 GPUs, charged 8x: every leg has a hard limit and optional legs are skipped when time is short):
# decomposition parity on 2x2x2 blocks over the peer-to-peer kernels and over NCCL, then the bench at N = 8 and N = 4.
mkdir -p gpurun_out
L=gpurun_out/r2n.log
: > $L
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
run() { local t=$1; shift; echo "=== [$(el)s] $*" >> $L; timeout $t "$@" >> $L 2>&1; echo "rc=$? [$(el)s]" >> $L; }
bench() {  # bench <N> <halo> <limit>
  echo "=== [$(el)s] bench --gpus $1 --halo $2" >> $L
  timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2954$1 \
    bench.py --gpus $1 --steps 5 --warmup 3 --halo $2 > gpurun_out/r2n_bench_n$1_$2.json 2> gpurun_out/r2n_bench_n$1_$2.err
  echo "rc=$? [$(el)s]" >> $L
  python - >> $L <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench_n$1_$2.json").read().strip().splitlines()[-1])
    print("value %.4g ms/step %.2f per_gpu %.4g parity %s stated %s" % (d["value"], d["ms_per_step"], d["per_gpu"], d["parity_ngpu_max_abs"], json.dumps(d.get("c4_as_stated"))[:300]))
except Exception as e:
    print("no bench line:", e)
PY
  tail -c 600 gpurun_out/r2n_bench_n$1_$2.err >> $L
}
nvidia-smi -L | wc -l >> $L
run 130 python -m pytest tests/test_multigpu.py -q -k "eight_gpus_p2p" --timeout 120 -rfE
bench 8 p2p 110
[ $(el) -lt 215 ] && run 70 python -m pytest tests/test_multigpu.py -q -k "eight_gpus and not p2p" --timeout 60 -rfE
[ $(el) -lt 225 ] && bench 4 p2p 70
[ $(el) -lt 200 ] && bench 8 nccl 90
grep -v "^$" $L | cut -c1-400 | tail -60
