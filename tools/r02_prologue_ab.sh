#!/bin/bash
# A/B of the start-up changes of the tiled kernel on ONE box (box-to-box variance is ~1 us): rebuilds the library with
# each switch and times the isolated check.
set -u
B="--steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot --tune-rounds 0 --no-overlap"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('   isolated %.5f ms frac %.4f' % (r['kernel_ms_mean'], r['frac']))
    elif 'rror' in l or 'ssert' in l: print(l.strip())
"; }
run() { for i in 1 2 3; do timeout 300 python bench.py $B "$@" 2>&1 | show; done; }
echo "=== default build (run_far + first tile evict-last), direct"; run
echo "=== default build, ticket"; ACG_K2_TICKET=1 run
for cfg in "-DACG_FIRST_TILE_KEEP=0" "-DACG_RUN_FAR=0" "-DACG_RUN_FAR=0 -DACG_FIRST_TILE_KEEP=0"; do
  echo "=== build $cfg"; ACG_NVCC_EXTRA="$cfg" python arithmetic-circuits_b200/build.py --force > /dev/null 2>&1 || echo build failed
  run
done
echo "=== 2^22, last build (old prologue)"; timeout 300 python bench.py $B --log-rows 22 2>&1 | show
python arithmetic-circuits_b200/build.py --force > /dev/null 2>&1
echo "=== 2^22, default build"; timeout 300 python bench.py $B --log-rows 22 2>&1 | show
