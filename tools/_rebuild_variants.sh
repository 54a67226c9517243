#!/bin/bash
# usage: tools/_rebuild_variants.sh outdir "NVCCFLAGS|ENV..." ...   (rebuilds the library on the box for each variant)
out=$1; shift
mkdir -p $out
for v in "$@"; do
  flags="${v%%|*}"; envs="${v#*|}"
  echo "== build [$flags] run [$envs]" >> $out/variants.log
  CMX_NVCC_FLAGS="$flags" python -c "from casmcode_clexmonte_b200 import build; build.build(force=True)" >> $out/build.log 2>&1
  env $envs timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g  kernel_ms %.4f  frac %.3f accept %.6f %s'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['accept_rate'],d['schedule']))
    elif l: print(l[:300])
" >> $out/variants.log
done
cat $out/variants.log
