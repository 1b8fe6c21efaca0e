#!/bin/bash
# round 2: N GPUs -- strong-scaling bench line (S(2^24) split in N), as the driver launches it
set -u
N=${1:-8}
mkdir -p gpurun_out
echo "=== bench --gpus $N (strong, 2^24)"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/scale_n$N.err | tail -1 > gpurun_out/scale_n$N.json; tail -3 gpurun_out/scale_n$N.err | cut -c1-300; python - "$N" <<'PY'
import json,sys
N=sys.argv[1]
try:
    j=json.load(open('gpurun_out/scale_n%s.json'%N))
    r=j['roofline']
    print(j['config']['workload'])
    print('value %.4g ms/step %.5f | iso %.5f frac %.4f ovl %.4f | e2e %.4g (%.4f ms)' % (j['value'], j['ms_per_step'], r['kernel_ms_mean'], r['frac'], r['frac_overlapped'], j['e2e']['value'], j['e2e']['ms_per_step']))
    print('per-rank kernel ms', [round(x,4) for x in r['per_rank_kernel_ms']])
    print(j['config'].get('peer_allreduce_check'), j['config']['setup_s'])
except Exception as e: print('ERR', e); print(open('gpurun_out/scale_n%s.json'%N).read()[:2000])
PY
