#!/bin/bash
# A/B several builds of libacg.so on the same box: tools/ab_multi.sh "<bench args>" so1 so2 ...
set -u
ARGS=$1; shift
cp arithmetic-circuits_b200/libacg.so /tmp/libacg_keep.so
for rep in 1 2; do for so in "$@"; do
  cp "$so" arithmetic-circuits_b200/libacg.so
  printf "%s rep%d: " "$(basename $so)" $rep
  timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 $ARGS 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print('roofline', round(j['roofline']['frac'],4), 'kernel_ms', round(j['roofline']['kernel_ms_mean'],4), 'sm_mhz', j['clocks']['sm_mhz'])
"
done; done
cp /tmp/libacg_keep.so arithmetic-circuits_b200/libacg.so
