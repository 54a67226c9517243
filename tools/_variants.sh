#!/bin/bash
# usage: tools/_variants.sh outdir "ENV1=a ENV2=b" "ENV1=c" ...
out=$1; shift
mkdir -p $out
for v in "$@"; do
  echo "== $v" >> $out/variants.log
  env $v timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $BENCH_ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g  kernel_ms %.4f  frac %.3f accept %.6f %s'%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['accept_rate'],d['schedule']))
    elif l: print(l[:300])
" >> $out/variants.log
done
cat $out/variants.log
