#!/bin/bash
# same-box A/B: a full copy of an earlier commit under _ab/head (built there) against the working tree
set -u
B="--steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot --no-overlap"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('   isolated %.5f ms frac %.4f' % (r['kernel_ms_mean'], r['frac']))
    elif 'rror' in l or 'ssert' in l: print(l.strip())
"; }
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "synth_parity or edge_shapes or handover or gate_mix or full_size" 2>&1 | tail -3
for rep in 1 2; do
echo "=== HEAD copy"; (cd _ab/head && timeout 300 python bench.py $B 2>&1 | show)
echo "=== working tree"; timeout 300 python bench.py $B 2>&1 | show
done
for a in "--workload mix" "--dense" "--field bls12_381" "--log-rows 22"; do
echo "=== HEAD copy $a"; (cd _ab/head && timeout 300 python bench.py $B $a 2>&1 | show)
echo "=== working tree $a"; timeout 300 python bench.py $B $a 2>&1 | show
done
