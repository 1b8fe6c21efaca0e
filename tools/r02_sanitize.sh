#!/bin/bash
# compute-sanitizer evidence (SURVEY section 5 hook): memcheck, racecheck, synccheck on the small configurations
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in ${PARTS:-k2 mix qap}; do
    [ "$tool" != "memcheck" ] && [ "$part" = "qap" ] && continue   # racecheck / synccheck: the shared-memory kernels (K2, long rows) + mix
    echo "=== $tool $part"
    timeout 1500 $CS --tool $tool --error-exitcode 9 python tools/sanitize_cases.py $part > gpurun_out/sanitizer_${tool}_${part}.log 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$|Error|error" gpurun_out/sanitizer_${tool}_${part}.log | head -8
  done
done
echo "=== K1 ceiling + 4xu64"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I arithmetic-circuits_b200/csrc -I tools/microbench tools/microbench/fr_mul_throughput.cu -o /tmp/fr_mul_throughput 2>/dev/null; timeout 300 /tmp/fr_mul_throughput > gpurun_out/fr_mul_throughput.txt 2>&1; grep -E "chains=1 warps/SM=(16|32)|mismatch" gpurun_out/fr_mul_throughput.txt | cut -c1-170
