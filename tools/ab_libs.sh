#!/bin/bash
# A/B of prebuilt library variants on one GPU box: tools/ab_libs.sh "<lib file> <ds_timeline options>" ...
# (each argument: path of a libquipb200.so variant followed by option=value pairs for tools/ds_timeline.py)
LIB=quip_for_all_b200/lib/libquipb200.so
cp $LIB /tmp/lib_orig.so
for spec in "$@"; do
  set -- $spec
  f=$1; shift
  cp $f $LIB
  echo "=== $f $*"
  python tools/ds_timeline.py 4 128 0 "$@" 2>&1 | grep -E "^\s+\[( 2| 3| 4| 6|11|12|13|14|15|16|17|18|19|22|23|24)\]|eager" | awk '{printf "%s ", $0} END {print ""}' | sed 's/(t = *[0-9.]* us)//g; s/  */ /g'
done
cp /tmp/lib_orig.so $LIB
