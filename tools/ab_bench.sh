#!/bin/bash
# A/B of prebuilt library variants with the real benchmark loop: tools/ab_bench.sh "<lib file> <ds_flags>" ...
LIB=quip_for_all_b200/lib/libquipb200.so
cp $LIB /tmp/lib_orig.so
for spec in "$@"; do
  set -- $spec
  cp $1 $LIB
  QUIPB200_OPTIONS=ds_flags=${2:-0} python bench.py --steps 96 --warmup 8 --no-cpu-baseline --no-ref-cuda --no-kernel-bench --no-hf-dropin --no-70b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 flags=${2:-0}: %.1f tok/s  %.3f ms/step  kernel %.1f us' % (d['value'], d['ms_per_step'], d.get('roofline',{}).get('kernel_us', 0)))"
done
cp /tmp/lib_orig.so $LIB
