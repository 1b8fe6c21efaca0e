#!/bin/bash
# first GPU round: smoke -> sanitizer -> parity tests -> bench variants -> ncu
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== sanitizer (memcheck, small)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "
import arithmetic_circuits_b200 as acg
g,w=acg.synth_r1cs(0,3000,5)
ctx=acg.Context(0,0)
m,dw=ctx.upload_r1cs(g),ctx.upload_witness(w)
for k in (1,2):
    ctx.set_check_kernel(k); print(k, ctx.r1cs_check(m,dw))
print(ctx.qap_witness(m,dw,(1,2,3),want=())[1])
" 2>&1 | tail -12
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -40
echo "=== bench tiled"; timeout 600 python bench.py --steps 100 --warmup 10 2>&1 | tail -3 | tee gpurun_out/bench_tiled.json
echo "=== bench rowwise"; timeout 600 python bench.py --steps 100 --warmup 10 --kernel rowwise --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_rowwise.json
echo "=== bench tiled dense"; timeout 600 python bench.py --steps 100 --warmup 10 --dense --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_tiled_dense.json
echo "=== bench tiled 2^22"; timeout 600 python bench.py --steps 50 --warmup 5 --log-rows 22 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_tiled_22.json
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
echo "=== ncu full tiled"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 3 -c 2 -o gpurun_out/prof_tiled -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== ncu full rowwise"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_rowwise -s 3 -c 1 -o gpurun_out/prof_rowwise -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --kernel rowwise > gpurun_out/ncu_full_rw.log 2>&1; tail -2 gpurun_out/ncu_full_rw.log
ls -la gpurun_out
