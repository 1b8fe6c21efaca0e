#!/bin/bash
# round 2: two GPUs -- peer all-reduce parity test, strong-scaling bench line (S(2^24) split in 2), weak line
set -u
mkdir -p gpurun_out
echo "=== pytest peer exchange"; timeout 900 python -m pytest tests -m gpu -q --tb=short -k "peer_exchange" 2>&1 | tail -5
echo "=== bench --gpus 2 (strong, 2^24)"; timeout 1500 python bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/scale_n2.err | tail -1 > gpurun_out/scale_n2.json; tail -5 gpurun_out/scale_n2.err; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/scale_n2.json'))
    r=j['roofline']
    print(j['config']['workload'])
    print('value %.4g ms/step %.5f | iso %.5f frac %.4f ovl %.4f per-rank %s | e2e %.4g (%.4f ms)' % (j['value'], j['ms_per_step'], r['kernel_ms_mean'], r['frac'], r['frac_overlapped'], r['per_rank_kernel_ms'], j['e2e']['value'], j['e2e']['ms_per_step']))
    print(j['config'].get('peer_allreduce_check'), j['config']['setup_s'])
except Exception as e: print('ERR', e); print(open('gpurun_out/scale_n2.json').read()[:2000])
PY
echo "=== bench --gpus 2 weak 2^20/GPU"; timeout 900 python bench.py --gpus 2 --scaling weak --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_n2_weak.json; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/scale_n2_weak.json'))
    r=j['roofline']
    print('value %.4g ms/step %.5f | iso %.5f frac %.4f | e2e %.4g (%.4f ms)' % (j['value'], j['ms_per_step'], r['kernel_ms_mean'], r['frac'], j['e2e']['value'], j['e2e']['ms_per_step']))
except Exception as e: print('ERR', e); print(open('gpurun_out/scale_n2_weak.json').read()[:2000])
PY
