#!/bin/bash
# last pass of the round: the whole GPU suite, sanitizers on the mixed workload, the bench lines that changed
set -u
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4
PARTS="mix" bash tools/r02_sanitize.sh 2>&1 | grep -E "^===|rc=|SUMMARY|ok$" | grep -v "K1 ceiling"
S="--no-cpu-baseline --no-qap --no-one-shot"
echo "=== bench ours mix"; timeout 900 python bench.py --workload mix --steps 50 --no-qap --no-one-shot 2>&1 | tail -1 | tee gpurun_out/bench_ours_mix.json | cut -c1-200
echo "=== bench ours mix, separate long-row launch"; ACG_K2_SEPARATE_LONGROWS=1 timeout 900 python bench.py --workload mix --steps 50 $S 2>&1 | tail -1 | tee gpurun_out/bench_ours_mix_separate.json | cut -c1-200
echo "=== bench ours (driver command)"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_ours_driver_cmd.json | cut -c1-300
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-200
