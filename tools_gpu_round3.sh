#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== sanitizer (memcheck + racecheck, small)"; for tool in memcheck racecheck; do timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "
import arithmetic_circuits_b200 as acg
g,w=acg.synth_r1cs(0,3000,5)
ctx=acg.Context(0,0)
m,dw=ctx.upload_r1cs(g),ctx.upload_witness(w)
for k,s in ((1,0),(2,0),(2,1)):
    ctx.set_check_kernel(k); ctx.set_tiled_variant(s); print(k,s, ctx.r1cs_check(m,dw))
print(ctx.r1cs_check_host(g,w))
" 2>&1 | tail -8; done
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30
for cfg in "--variant 0" "--variant 1" "--variant 0 --dense" "--variant 1 --dense" "--variant 0 --log-rows 22" "--variant 1 --log-rows 22" "--variant 0 --field bls12_381"; do
  echo "=== bench $cfg"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $cfg 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print({k:j[k] for k in ('value','ms_per_step')}, 'roofline', round(j['roofline']['frac'],4), 'kernel_ms', round(j['roofline']['kernel_ms_mean'],4), 'e2e', round(j['e2e']['ms_per_step'],2), 'ms', round(j['e2e']['value']/1e6,1),'Mc/s', 'w-only', round(j['e2e']['witness_only_ms_per_step'],3))
"; done
echo "=== ncu full tiled s1"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/prof_tiled_v4v0 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant 0 > gpurun_out/ncu_full_s1.log 2>&1; tail -1 gpurun_out/ncu_full_s1.log
echo "=== ncu full tiled s2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/prof_tiled_v4v1 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant 1 > gpurun_out/ncu_full_s2.log 2>&1; tail -1 gpurun_out/ncu_full_s2.log
