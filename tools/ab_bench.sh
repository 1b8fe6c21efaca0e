#!/bin/bash
# A/B two builds of libacg.so on the same box: tools/ab_bench.sh <so_a> <so_b> [bench args...]
set -u
A=$1; B=$2; shift 2
cp arithmetic-circuits_b200/libacg.so /tmp/libacg_keep.so
for rep in 1 2; do for so in "$A" "$B"; do
  cp "$so" arithmetic-circuits_b200/libacg.so
  for lr in 20 22; do
    printf "%s rep%d 2^%d: " "$(basename $so)" $rep $lr
    timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 --log-rows $lr "$@" 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print('roofline', round(j['roofline']['frac'],4), 'kernel_ms', round(j['roofline']['kernel_ms_mean'],4), 'sm_mhz', j['clocks']['sm_mhz'])
"
  done
done; done
cp /tmp/libacg_keep.so arithmetic-circuits_b200/libacg.so
