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
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "synth_parity or edge_shapes or handover or overlap or pipelined or full_size" 2>&1 | tail -3
for rep in 1 2; do
echo "=== HEAD copy (ticket)"; (cd _ab/head && timeout 300 python bench.py $B 2>&1 | show)
echo "=== working tree, ticket"; ACG_K2_TICKET=1 timeout 300 python bench.py $B 2>&1 | show
echo "=== working tree, direct"; timeout 300 python bench.py $B "$@" 2>&1 | show
done
echo "=== working tree direct 2^22"; timeout 300 python bench.py $B --log-rows 22 2>&1 | show
echo "=== working tree direct bls"; timeout 300 python bench.py $B --field bls12_381 2>&1 | show
