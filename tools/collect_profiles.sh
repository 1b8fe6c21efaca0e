#!/bin/bash
# copy the outputs of tools/gpu_round_end.sh from gpurun_out/ (scratch) into profiles/ (tracked), named per round
set -u
R=${1:-r01}
G=gpurun_out; P=profiles
cp $G/bench_ours.json $P/${R}_bench_ours.json
cp $G/bench_reference.json $P/${R}_bench_reference.json
cp $G/bench_ours_bls.json $P/${R}_bench_ours_bls12_381.json
cp $G/bench_ours_no_overlap.json $P/${R}_bench_ours_no_overlap.json
cp $G/bench_ours_21.json $P/${R}_bench_ours_2p21.json
cp $G/bench_ours_22.json $P/${R}_bench_ours_2p22.json
cp $G/bench_ours_dense.json $P/${R}_bench_ours_dense.json
cp $G/bench_ours_rowwise.json $P/${R}_bench_ours_rowwise.json
cp $G/bench_qap_20.json $P/${R}_bench_qap_2p20.json
cp $G/bench_qap_22.json $P/${R}_bench_qap_2p22.json
cp $G/fr_mul_throughput.txt $P/${R}_fr_mul_throughput.txt
cp $G/phase_cycles.txt $P/${R}_k2_phase_cycles.txt
cp $G/gpu_info.txt $P/${R}_gpu_info.txt
cp $G/launches_bench.csv $P/${R}_ncu_launches_bench.csv
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$G/launches_qap20.csv")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    agg[name][r[mi]] += float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum": cnt[name] += 1
with open("$P/${R}_ncu_launches_qap_2p20_summary.txt", "w") as f:
    f.write("kernel, launches, total time (ns, cold-cache serialised under ncu), dram read (B), dram write (B)\n")
    for k in sorted(agg, key=lambda k: -agg[k]["gpu__time_duration.sum"]):
        f.write("%s, %d, %.0f, %.0f, %.0f\n" % (k, cnt[k], agg[k]["gpu__time_duration.sum"], agg[k]["dram__bytes_read.sum"], agg[k]["dram__bytes_write.sum"]))
PY
python tools/ncu_summarize.py $P/${R}_ncu_summary.json k2_r1cs_tiled=$G/prof_k2_final.ncu-rep k3_ntt_pass=$G/prof_k3_ntt.ncu-rep
ls $P
