#!/bin/bash
# round 2, GPU call 3: per-SM CTAs with dynamic group scheduling -- parity, A/B of the geometries, timeline
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (K2 tests first)"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40
B="--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('ms/step %.5f  sampled %.5f  min %.5f  frac %.4f  iso %.4f' % (j['ms_per_step'], r['kernel_ms_sampled_mean'], r['kernel_ms_min'], r['frac'], r['frac_isolated_launch']))
    elif 'timeline' in l or 'phase cycles' in l or 'rror' in l: print(l.strip())
"; }
for v in 0 2 4 6 7 3 1; do
  echo "=== variant $v no-overlap"; timeout 600 python bench.py $B --variant $v --no-overlap 2>&1 | show
done
for v in 0 2; do
  echo "=== variant $v overlap"; timeout 600 python bench.py $B --variant $v 2>&1 | show
done
echo "=== timeline (variant 0)"; ACG_TILED_TIMING=1 ACG_TILED_TIMING_DUMP=gpurun_out/cta_marks_v12.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-overlap 2>&1 | show
for lr in 18 22; do
  echo "=== rows 2^$lr variant 0 no-overlap"; timeout 600 python bench.py $B --log-rows $lr --no-overlap 2>&1 | show
done
echo "=== dense v0"; timeout 600 python bench.py $B --dense --no-overlap 2>&1 | show
echo "=== bls v0"; timeout 600 python bench.py $B --field bls12_381 --no-overlap 2>&1 | show
echo "=== ncu full tiled v0"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/r02_k2_v12 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-overlap > gpurun_out/ncu_k2_v12.log 2>&1; tail -1 gpurun_out/ncu_k2_v12.log
