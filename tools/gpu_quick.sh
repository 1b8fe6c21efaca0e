#!/bin/bash
# quick GPU iteration: smoke, parity tests, bench variants, one ncu capture of the tiled kernel
set -u
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
for cfg in "--variant 0" "--variant 1" "--variant 2" "--variant 3" "--variant 0 --dense" "--variant 0 --log-rows 22" "--variant 0 --field bls12_381"; do
  echo "=== bench $cfg"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 $cfg 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print({k:j[k] for k in ('value','ms_per_step')}, 'roofline', round(j['roofline']['frac'],4), 'kernel_ms', round(j['roofline']['kernel_ms_mean'],4), 'stream MB', round(j['roofline']['device_stream_bytes_per_launch']/1e6,1))
"; done
echo "=== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/prof_tiled_quick -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant 0 > gpurun_out/ncu_quick.log 2>&1; tail -1 gpurun_out/ncu_quick.log
