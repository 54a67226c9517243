#!/bin/bash
# usage: tools/_traffic.sh outdir "ENV..." ...  : DRAM traffic + duration of one stream-kernel launch per variant
out=$1; shift
mkdir -p $out
for v in "$@"; do
  echo "== $v" >> $out/traffic.log
  env $v ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_sweep_ -s 3 -c 1 --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' >> $out/traffic.log
done
cat $out/traffic.log
