#!/bin/bash
# same-box A/B: a full copy of an earlier commit under _ab/head (built there) against the working tree
#   usage: tools/r02_head_ab.sh "<pytest -k expression>" "<bench args>" ["<bench args>" ...]
#   the copy (here, before the gpurun call; _ab/ is git-ignored but travels to the GPU box):
#     rm -rf _ab/head && mkdir -p _ab/head && git archive <commit> arithmetic-circuits_b200 arithmetic_circuits_b200 \
#       bench.py oracle include tests/helpers.py | tar -x -C _ab/head/ && cp MEASURED_PEAKS.json _ab/head/ && \
#       python _ab/head/arithmetic-circuits_b200/build.py
set -u
B="--steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot --no-overlap"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('   isolated %.5f ms frac %.4f (%s launches per check)' % (r['kernel_ms_mean'], r['frac'], r.get('launches_per_check')))
    elif 'rror' in l or 'ssert' in l: print(l.strip())
"; }
K=${1:-synth_parity}; shift || true
echo "=== pytest -k '$K'"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "$K" 2>&1 | tail -3
[ $# -eq 0 ] && set -- ""
for a in "$@"; do
  for rep in 1 2; do
    echo "=== copy of the earlier commit: $a"; (cd _ab/head && timeout 300 python bench.py $B $a 2>&1 | show)
    echo "=== working tree: $a"; timeout 300 python bench.py $B $a 2>&1 | show
  done
done
