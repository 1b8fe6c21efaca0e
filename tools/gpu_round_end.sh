#!/bin/bash
# round-end measurement pass (one GPU): parity tests, contract bench (both arms), other workloads, ncu evidence,
# instrumented timeline.  Outputs land in gpurun_out/ and are copied to profiles/ by tools/collect_profiles.sh <round>.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1; nproc >> gpurun_out/gpu_info.txt; lscpu | grep "Model name" >> gpurun_out/gpu_info.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
echo "=== bench ours (driver command)"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_ours_driver_cmd.json | cut -c1-600
echo "=== bench ours (defaults)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_ours.json | cut -c1-300
S="--no-cpu-baseline --no-qap --no-one-shot"
echo "=== bench ours bls"; timeout 900 python bench.py --field bls12_381 $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_bls.json | cut -c1-200
echo "=== bench ours equal runs"; ACG_K2_WAVE_SHARES=0 timeout 900 python bench.py $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_equal_runs.json | cut -c1-200
echo "=== bench ours, ticket hand-over (A/B of the direct hand-over)"; ACG_K2_TICKET=1 timeout 900 python bench.py $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_ticket.json | cut -c1-200
for v in 2 4 6; do echo "=== bench ours variant $v"; timeout 600 python bench.py --variant $v $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_variant$v.json | cut -c1-120; done
echo "=== bench ours 2^21"; timeout 900 python bench.py --scaling weak --log-rows 21 $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_21.json | cut -c1-200
echo "=== bench ours 2^22"; timeout 900 python bench.py --scaling weak --log-rows 22 --steps 50 $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_22.json | cut -c1-200
echo "=== bench ours dense"; timeout 900 python bench.py --dense $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_dense.json | cut -c1-200
echo "=== bench ours rowwise"; timeout 900 python bench.py --kernel rowwise $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_rowwise.json | cut -c1-200
echo "=== bench ours mix"; timeout 900 python bench.py --workload mix --steps 50 --no-qap --no-one-shot 2>&1 | tail -1 | tee gpurun_out/bench_ours_mix.json | cut -c1-200
echo "=== bench ours mix, separate long-row launch"; ACG_K2_SEPARATE_LONGROWS=1 timeout 900 python bench.py --workload mix --steps 50 $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_mix_separate.json | cut -c1-200
echo "=== bench qap 2^20"; timeout 900 python bench_qap.py --log-n 20 2>&1 | tail -6 | tee gpurun_out/bench_qap_20.json | cut -c1-300
echo "=== K1 ceiling"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I arithmetic-circuits_b200/csrc -I tools/microbench tools/microbench/fr_mul_throughput.cu -o /tmp/fr_mul_throughput 2>/dev/null; timeout 300 /tmp/fr_mul_throughput > gpurun_out/fr_mul_throughput.txt 2>&1; tail -2 gpurun_out/fr_mul_throughput.txt | cut -c1-160
N="--steps 5 --warmup 3 --no-cpu-baseline --no-qap --no-one-shot --e2e-steps 2"
echo "=== ncu launches (bench)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_bench.csv python bench.py $N > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
echo "=== ncu full tiled"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_r1cs_tiled -s 6 -c 1 -o gpurun_out/prof_k2_final -f python bench.py $N > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
echo "=== ncu full ntt pass"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 8 -c 3 -o gpurun_out/prof_k3_ntt -f python bench_qap.py --log-n 22 --reps 1 --no-cpu > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
echo "=== instrumented build: phase cycles + CTA timeline"; ACG_NVCC_EXTRA=-DACG_TILED_TIMING_BUILD timeout 600 python arithmetic-circuits_b200/build.py --force > /dev/null 2>&1
ACG_TILED_TIMING=1 ACG_TILED_TIMING_DUMP=gpurun_out/cta_marks_final.txt timeout 300 python bench.py $N --no-overlap 2>&1 | grep "phase cycles\|cta timeline" | tail -6 | tee gpurun_out/phase_cycles.txt
ls -la gpurun_out | tail -30
