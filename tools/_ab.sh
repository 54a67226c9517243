#!/bin/bash
# usage: tools/_ab.sh outdir libA libB ... : bench each library variant on the same box, twice
out=$1; shift
mkdir -p $out
P=casmcode_clexmonte_b200
for rep in 1 2; do
for lib in "$@"; do
  cp $P/$lib $P/libcmx_b200.so
  echo "== $lib (rep $rep)" >> $out/variants.log
  tools/_variants.sh $out "X=0" > /dev/null
done
done
cat $out/variants.log
