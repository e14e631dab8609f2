#!/bin/bash
# A/B of experiment builds on the GPU box: tools/ab.sh <block> <tag> [<tag> ...]
# (tags name idefix_b200/libidefix_b200_<tag>.so built with `make DEV=1 DEVTAG=<tag>`; "full" = the product library)
block=$1; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = full ]; then unset IDEFIX_B200_LIB; else export IDEFIX_B200_LIB=$PWD/idefix_b200/libidefix_b200_$tag.so; fi
  echo "=== $tag"
  python -m pytest tests/test_oracle_gpu.py -q -x -k 3d3c 2>&1 | tail -2
  python bench.py --block $block --steps 4 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_total']
        print('value %.4g  ms/step %.2f' % (d['value'], d['ms_per_step']), {c: round(v['ms_total']/v['launch_brackets'],3) for c,v in k.items()})
    else: print(l.rstrip())
"
done
