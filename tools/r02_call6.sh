#!/bin/bash
# round 2, GPU call 6: full parity suite (new config tests), reworked bench (default line, Split mix)
set -u
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short --durations=8 2>&1 | tail -40
echo "=== bench ours (default)"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_ours.err | tail -1 > gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_ours.json'))
r=j['roofline']
print('value %.4g ms/step %.5f | iso %.5f ms frac %.4f frac_ovl %.4f | e2e %.4g (%.4f ms) one_shot %s' % (j['value'], j['ms_per_step'], r['kernel_ms_mean'], r['frac'], r['frac_overlapped'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['one_shot'] and '%.3g' % j['e2e']['one_shot']['value']))
print('setup', j['config']['setup_s'], 'cpu', j.get('cpu_baseline',{}).get('value'), 'qap', j.get('qap'))
PY
echo "=== bench mix"; timeout 900 python bench.py --workload mix --steps 50 --warmup 5 --no-qap --no-one-shot 2>&1 | tail -1 > gpurun_out/bench_mix.json; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_mix.json'))
r=j['roofline']
print(j['config']['workload']); print('value %.4g ms/step %.5f | iso %.5f ms frac %.4f | launches %d per %d steps | e2e %.4g | cpu %s' % (j['value'], j['ms_per_step'], r['kernel_ms_mean'], r['frac'], j['gpu_launches'], j['steps'], j['e2e']['value'], j.get('cpu_baseline',{}).get('value')))
PY
