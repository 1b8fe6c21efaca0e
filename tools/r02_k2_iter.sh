#!/bin/bash
# round 2: quick K2 iteration -- K2 parity subset, A/B of geometries / tapering, timeline
set -u
mkdir -p gpurun_out
[ "${SKIP_TESTS:-0}" = 1 ] || { echo "=== pytest gpu (K2 subset)"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "synth_parity or edge_shapes or mixed or overlapping or full_size or constant_wire or row_shards or reference_unit or eq_and_split or linear_constraints or gate_mix or qap_witness_vs or device_witness" 2>&1 | tail -5; }
B="--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('isolated %.5f ms (event pairs: mean %.5f min %.5f)  frac %.4f | region-1 step %.5f ms frac_ovl %.4f' % (r['kernel_ms_mean'], r['kernel_ms_event_pair_mean'], r['kernel_ms_event_pair_min'], r['frac'], j['ms_per_step'], r['frac_overlapped']))
    elif 'timeline' in l or 'phase cycles' in l or 'rror' in l: print(l.strip())
"; }
for v in 0 2 6 4; do
  echo "=== variant $v no-overlap"; timeout 600 python bench.py $B --variant $v --no-overlap 2>&1 | show
done
for t in; do
  echo "=== variant 0 no-overlap ACG_TAPER=$t"; ACG_TAPER=$t timeout 600 python bench.py $B --variant 0 --no-overlap 2>&1 | show
done
echo "=== variant 0 overlap"; timeout 600 python bench.py $B --variant 0 2>&1 | show
[ "${SKIP_TIMELINE:-1}" = 1 ] || echo "=== timeline (variant 0)"; [ "${SKIP_TIMELINE:-1}" = 1 ] || ACG_TILED_TIMING=1 ACG_TILED_TIMING_DUMP=gpurun_out/cta_marks_iter.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-overlap --no-qap --no-one-shot 2>&1 | show | tail -7
for lr in 18 22; do
  echo "=== rows 2^$lr variant 0 no-overlap"; timeout 600 python bench.py $B --log-rows $lr --no-overlap 2>&1 | show
done
echo "=== dense v0"; timeout 600 python bench.py $B --dense --no-overlap 2>&1 | show
echo "=== bls v0"; timeout 600 python bench.py $B --field bls12_381 --no-overlap 2>&1 | show
