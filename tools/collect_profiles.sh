#!/bin/bash
# copy the outputs of tools/gpu_round_end.sh from gpurun_out/ (scratch) into profiles/ (tracked), named per round
set -u
R=${1:-r02}
G=gpurun_out; P=profiles
for f in bench_ours bench_ours_driver_cmd bench_reference bench_ours_bls bench_ours_equal_runs bench_ours_ticket bench_ours_variant2 bench_ours_variant4 bench_ours_variant6 bench_ours_21 bench_ours_22 bench_ours_dense bench_ours_rowwise bench_ours_mix bench_ours_mix_separate bench_qap_20; do
  [ -s $G/$f.json ] && cp $G/$f.json $P/${R}_$f.json
done
[ -s $G/fr_mul_throughput.txt ] && cp $G/fr_mul_throughput.txt $P/${R}_fr_mul_throughput.txt
[ -s $G/phase_cycles.txt ] && cp $G/phase_cycles.txt $P/${R}_k2_phase_cycles.txt
[ -s $G/cta_marks_final.txt ] && cp $G/cta_marks_final.txt $P/${R}_cta_timeline_final.txt
[ -s $G/gpu_info.txt ] && cp $G/gpu_info.txt $P/${R}_gpu_info.txt
[ -s $G/launches_bench.csv ] && cp $G/launches_bench.csv $P/${R}_ncu_launches_bench.csv
python tools/ncu_summarize.py $P/${R}_ncu_summary.json k2_r1cs_tiled=$G/prof_k2_final.ncu-rep k3_ntt_pass=$G/prof_k3_ntt.ncu-rep
ls $P | grep "^${R}_"
