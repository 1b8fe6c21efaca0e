#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== sanitizer"; for tool in memcheck racecheck; do timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "
import arithmetic_circuits_b200 as acg
g,w=acg.synth_r1cs(0,3000,5)
ctx=acg.Context(0,0)
dw=ctx.upload_witness(w)
for v in (0,1):
    ctx.set_tiled_variant(v); m=ctx.upload_r1cs(g)
    for k,s in ((1,1),(2,1),(2,2)):
        ctx.set_check_kernel(k); ctx.set_tiled_stages(s); print(v,k,s, ctx.r1cs_check(m,dw))
print(ctx.r1cs_check_host(g,w))
" 2>&1 | tail -9; done
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30
for cfg in "--variant 0 --stages 1" "--variant 0 --stages 2" "--variant 1 --stages 1" "--variant 1 --stages 2" "--variant 0 --dense" "--variant 1 --stages 2 --dense" "--variant 0 --log-rows 22" "--variant 1 --stages 2 --log-rows 22" "--variant 0 --field bls12_381"; do
  echo "=== bench $cfg"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 $cfg 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print({k:j[k] for k in ('value','ms_per_step')}, 'roofline', round(j['roofline']['frac'],4), 'kernel_ms', round(j['roofline']['kernel_ms_mean'],4), 'stream MB', round(j['roofline']['device_stream_bytes_per_launch']/1e6,1))
"; done
echo "=== ncu A"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/prof_tiled_v5a -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant 0 --stages 1 > gpurun_out/ncu_full_a.log 2>&1; tail -1 gpurun_out/ncu_full_a.log
echo "=== ncu B"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 1 -o gpurun_out/prof_tiled_v5b -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant 1 --stages 2 > gpurun_out/ncu_full_b.log 2>&1; tail -1 gpurun_out/ncu_full_b.log
