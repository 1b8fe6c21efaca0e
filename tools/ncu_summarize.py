#!/usr/bin/env python3
"""Summarise `ncu --set full` captures into profiles/<round>_ncu_summary.json.

    python tools/ncu_summarize.py OUT.json key1=report1.ncu-rep [key2=report2.ncu-rep ...]

Each report is read here with `ncu -i REPORT --page raw --csv`; one record per captured launch with the
metrics the DESIGN/bench refer to (duration, DRAM bytes, pipe and issue utilisation, occupancy, stall mix)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    recs = []
    for d in rows[2:]:
        if len(d) < len(hdr):
            continue
        m = dict(zip(hdr, d))
        u = dict(zip(hdr, units))
        rec = {k: m[k] for k in ("Kernel Name", "Grid Size", "Block Size") if k in m}
        rec.update({k: m[k] for k in KEEP if k in m})
        rec["stall_cycles_per_issue"] = {h[len(STALL):-len("_per_issue_active.ratio")]: round(float(m[h]), 3)
                                         for h in hdr if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")
                                         and float(m[h] or 0) >= 0.05}
        rec["units"] = {k: u[k] for k in KEEP if k in u}
        recs.append(rec)
    return recs


def main():
    out_path = sys.argv[1]
    res = {}
    for arg in sys.argv[2:]:
        key, path = arg.split("=", 1)
        res[key] = summarize(path)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out_path, {k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
