#!/bin/bash
# One A/B leg on the GPU box: tools/ab_run.sh <label> <block> [ENV=VALUE ...]
# (IDEFIX_B200_LIB=... selects an experiment build of the library, IDFX_NO_TMA=1 / IDFX_NO_GRAPH=1 are run-time switches)
label=$1; block=$2; shift 2
echo "=== $label $*"
env "$@" python -m pytest tests/test_oracle_gpu.py -q -x -k 3d3c 2>&1 | tail -2
env "$@" python bench.py --block $block --steps 4 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_total']
        print('value %.4g  ms/step %.2f' % (d['value'], d['ms_per_step']), {c: round(v['ms_total']/max(v['launch_brackets'],1),3) for c,v in k.items()})
    else: print(l.rstrip())
"
