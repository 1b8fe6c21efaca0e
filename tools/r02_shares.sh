#!/bin/bash
# round 2: A/B of the per-wave tile shares of the static split (k_r1cs_tiled), then a timeline of the instrumented build
set -u
mkdir -p gpurun_out
B="--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot --no-overlap"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('isolated %.5f ms (pairs mean %.5f min %.5f) frac %.4f' % (r['kernel_ms_mean'], r['kernel_ms_event_pair_mean'], r['kernel_ms_event_pair_min'], r['frac']))
    elif 'timeline' in l or 'phase cycles' in l or 'rror' in l: print(l.strip())
"; }
echo "=== pytest (full-size K2 tests)"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "full_size or config5 or row_shards or overlapping or gate_mix" 2>&1 | tail -4
for sh in default 0 "1.3,1.25,1.1,0.93,0.8" "1.3,1.26,1.15,0.98,0.8" "1.35,1.3,1.15,0.95,0.78" "1.28,1.22,1.12,1.0,0.85"; do
  echo "=== shares $sh"
  if [ "$sh" = default ]; then timeout 600 python bench.py $B 2>&1 | show; else ACG_K2_WAVE_SHARES=$sh timeout 600 python bench.py $B 2>&1 | show; fi
done
echo "=== 2^22 default vs equal"; timeout 600 python bench.py $B --log-rows 22 --scaling weak 2>&1 | show; ACG_K2_WAVE_SHARES=0 timeout 600 python bench.py $B --log-rows 22 --scaling weak 2>&1 | show
echo "=== bls default vs equal"; timeout 600 python bench.py $B --field bls12_381 2>&1 | show; ACG_K2_WAVE_SHARES=0 timeout 600 python bench.py $B --field bls12_381 2>&1 | show
echo "=== instrumented build"; ACG_NVCC_EXTRA=-DACG_TILED_TIMING_BUILD timeout 600 python arithmetic-circuits_b200/build.py --force > /dev/null 2>&1; echo rc=$?
for sh in default 0; do
  echo "=== timeline shares $sh"
  if [ "$sh" = default ]; then ACG_TILED_TIMING=1 ACG_TILED_TIMING_DUMP=gpurun_out/cta_marks_shares_default.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-overlap --no-qap --no-one-shot 2>&1 | show | tail -6
  else ACG_K2_WAVE_SHARES=0 ACG_TILED_TIMING=1 ACG_TILED_TIMING_DUMP=gpurun_out/cta_marks_shares_equal.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-overlap --no-qap --no-one-shot 2>&1 | show | tail -6; fi
done
