#!/bin/bash
# usage: tools/_ab2.sh outdir "lib|ENV=.. ENV=.." ... : bench each (library, environment) pair on the same box
out=$1; shift
mkdir -p $out
P=casmcode_clexmonte_b200
for spec in "$@"; do
  lib=${spec%%|*}; envs=${spec#*|}
  cp $P/$lib $P/libcmx_b200.so
  echo "== $lib $envs" >> $out/variants.log
  tools/_variants.sh $out "X=0 $envs" > /dev/null
done
cat $out/variants.log
