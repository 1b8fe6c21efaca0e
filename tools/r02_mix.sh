#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (K2 subset)"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "synth_parity or edge_shapes or mixed or gate_mix or eq_and_split or full_size or device_witness or row_shards or reference_unit" 2>&1 | tail -4
echo "=== bench mix"; timeout 900 python bench.py --workload mix --steps 50 --no-qap --no-one-shot --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours_mix.json | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('mix: step %.5f ms, isolated %.5f ms, launches/check %.1f, rows %d' % (j['ms_per_step'], r['kernel_ms_mean'], r['launches_per_check'], j['config']['rows_per_gpu']))"
echo "=== ncu launch list mix"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_mix.csv python bench.py --workload mix --steps 3 --warmup 3 --no-qap --no-one-shot --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1; python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_mix.csv')) if len(r)>10]
hdr=rows[0]; ki,vi=hdr.index('Kernel Name'),hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',','')))
for k,v in agg.items(): print('%-62s n=%3d mean %.1f us'%(k,len(v),sum(v)/len(v)/1000))
PY
