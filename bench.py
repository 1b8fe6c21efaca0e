#!/usr/bin/env python3
"""Benchmark of the hot path: R1CS witness check (A.w o B.w - C.w == 0 over BN254 Fr) on synthetic
circuits S(n, seed, field) of SURVEY.md 8(d).  Metric: R1CS constraints / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-rows 20] [--field bn254]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one check of every constraint of the (sharded) system.  Weak scaling: each rank owns
2^log_rows consecutive rows of a global system of N * 2^log_rows rows (witness replicated), checks
them with the tiled CUDA kernel -- ONE kernel launch per step: the kernel's last CTA finalises the result pair
and, for N > 1, all-reduces it over peer memory (NVLink P2P stores; `--collective nccl` uses one NCCL all-reduce).

  value     device-resident throughput: K steps timed with CUDA events on the launching stream, barrier +
            synchronize on both sides, max over ranks.
  e2e       same metric end to end per witness: the system stays resident on the device (the reference keeps its
            QAP value in memory between calls, too) and every step copies a new witness from pinned host memory
            (acg_witness_update: H2D, range check, Montgomery conversion), checks it and reads the result pair back.
            e2e.one_shot: everything from host buffers every step (acg_r1cs_check_host: pinned host CSR + witness
            -> H2D -> kernels -> D2H), PCIe-bound.
  roofline  tiled check kernel: SURVEY 8(d) algorithmic bytes of the shard / average per-launch duration (CUDA
            event pairs around every 8th launch inside the timed region; bounded by the step time, a step being
            exactly one launch) vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference
            the oracle's C restatement of the reference algorithm (oracle/r1cs_oracle.c, "port": the Haskell
            reference cannot be built in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEEDS = {20: 20260002, 22: 20260003, 24: 20260004}
FIELD_IDS = {"bn254": 0, "bls12_381": 1}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-rows", type=int, default=20, help="log2 of constraints per GPU")
    ap.add_argument("--field", default="bn254", choices=list(FIELD_IDS))
    ap.add_argument("--dense", action="store_true", help="all coefficients uniform (stress variant)")
    ap.add_argument("--kernel", default="tiled", choices=["tiled", "rowwise"])
    ap.add_argument("--no-overlap", action="store_true",
                    help="do not overlap consecutive checks (programmatic dependent launch of the next check kernel)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: all-reduce of the result pair over peer memory (default) or by NCCL")
    ap.add_argument("--variant", type=int, default=0, choices=list(range(8)), help="tiled kernel geometry (kernels.h kTileGeom): 0-3 rows per tile 128/256/64/32, 4/5 = 128/64 with the natural term layout, 6 = products in place, 7 = 6 + one far buffer (6 CTAs per SM)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args, world):
    total = world << args.log_rows
    return "S(n=%d, seed=%d, %s%s): %d Mul-gate R1CS rows, ~5.5 nnz/row, %d rows/GPU, witness replicated" % (
        total, SEEDS.get(args.log_rows, 20260000 + args.log_rows), args.field, ", dense" if args.dense else "", total,
        1 << args.log_rows)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(j.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------
def cpu_check_throughput(g, w, field_id, n_threads, min_seconds, max_reps):
    from oracle import c_oracle as CO
    mats = [(m[0], m[1], m[2]) for m in g.mats]
    CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *mats, w, False, n_threads)  # warm: page in, init tables
    times = []
    t_all = time.perf_counter()
    while len(times) < max_reps and (len(times) < 1 or time.perf_counter() - t_all < min_seconds):
        t0 = time.perf_counter()
        res = CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *mats, w, False, n_threads)
        times.append(time.perf_counter() - t0)
        assert res["n_violations"] == 0
    return g.n_rows / min(times), times


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Haskell original has no toolchain
    here) on rank 0's host cores, all threads, same workload and metric."""
    if rank != 0:
        return
    import arithmetic_circuits_b200 as acg  # host-side generator only (no GPU use in this arm)
    from oracle import c_oracle as CO
    CO.build()
    field_id = FIELD_IDS[args.field]
    n = 1 << args.log_rows
    total = world * n
    seed = SEEDS.get(args.log_rows, 20260000 + args.log_rows)
    # the bounded sample: rank 0's shard of the global system (all of it at N = 1)
    g, w = acg.synth_r1cs(field_id, total, seed, args.dense, rows=(0, n))
    threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1: ask for every host thread explicitly
    mats = [(m[0], m[1], m[2]) for m in g.mats]
    for _ in range(max(1, min(args.warmup, 3))):
        CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *mats, w, False, threads)
    steps = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(steps):
        res = CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *mats, w, False, threads)
    dt = time.perf_counter() - t0
    assert res["n_violations"] == 0
    value = g.n_rows * steps / dt
    line = {
        "impl": "reference", "metric": "R1CS constraints/sec (BN254 Fr)" if field_id == 0 else "R1CS constraints/sec (BLS12-381 Fr)",
        "value": value, "unit": "constraints/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (4x64-bit Montgomery, unsigned __int128)", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "note": "CPU arm checks one shard (2^%d rows) per step on rank 0" % args.log_rows},
        "cpu_baseline": {"value": value, "unit": "constraints/s", "cores": threads, "kind": "port",
                         "sample": "%d full checks of a 2^%d-row shard, OpenMP over rows" % (steps, args.log_rows)},
        "e2e": {"value": value, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import arithmetic_circuits_b200 as acg
    from arithmetic_circuits_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    field_id = FIELD_IDS[args.field]
    n = 1 << args.log_rows
    total = world * n
    seed = SEEDS.get(args.log_rows, 20260000 + args.log_rows)
    rb, re = sharding.row_shard(total, world, rank)
    t_gen = time.perf_counter()
    g, w = acg.synth_r1cs(field_id, total, seed, args.dense, rows=(rb, re))
    t_gen = time.perf_counter() - t_gen

    ctx = acg.Context(field_id, local_rank)
    ctx.set_check_kernel(acg.CHECK_TILED if args.kernel == "tiled" else acg.CHECK_ROWWISE)
    ctx.set_tiled_variant(args.variant)
    ctx.set_overlap_checks(not args.no_overlap)
    m = ctx.upload_r1cs(g)
    dw = ctx.upload_witness(w)
    algo_bytes = m.algorithmic_bytes
    if world > 1:  # shards differ (later rows reference a larger part of the witness): the mean over the ranks
        tb = torch.tensor([float(algo_bytes)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        algo_bytes = int(tb.item() / world)
    result = torch.zeros(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    # the one exchange of the sharded check: the result pair, all-reduced over peer memory by the last CTA of the
    # check kernel itself (NVLink P2P stores, include/acg.h) -- or, with --collective nccl, by one NCCL all-reduce
    peer = None
    collective = "none"
    if world > 1:
        collective = args.collective
        if collective == "p2p":
            try:
                peer = sharding.connect_peers(ctx)
            except Exception as e:  # no peer access between these GPUs: say so and use NCCL
                sys.stderr.write("bench.py: peer exchange unavailable (%s); using NCCL\n" % (e,))
                collective = "nccl"
            ok = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                peer, collective = None, "nccl"

    def step():
        if peer is not None:
            ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
            return
        ctx.r1cs_check_async(m, dw, result.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(result[0:1], op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    assert int(result[0].item()) == 0, "honest witness must verify"

    # ---- device-resident timed region.  One step is ONE kernel launch (the check kernel finalises its own result
    # and, for N > 1, all-reduces it over peer memory in its last CTA), so consecutive steps run back to back.
    # Per-launch durations are sampled with CUDA event pairs on the launching stream around every 8th step: an
    # event pair around EVERY launch costs ~6 us of front-end time per step (measured), i.e. it would perturb the
    # very number it measures.
    launches0 = ctx.kernel_launch_count()
    sample_every = 8
    pairs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record(stream)
    for i in range(args.steps):
        if i % sample_every == sample_every // 2:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            step()
            eb.record(stream)
            pairs.append((ea, eb))
        else:
            step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    kernel_ms = [a.elapsed_time(b) for a, b in pairs]
    launches = ctx.kernel_launch_count() - launches0
    assert int(result[0].item()) == 0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer call (pinned host inputs, H2D + D2H every step)
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
    pinned = []
    mats = []
    for rp, col, val in g.mats:
        trip = []
        for a in (rp, col, val):
            tt = torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype == np.uint32 else np.int64)).pin_memory()
            pinned.append(tt)
            trip.append(tt.numpy().view(a.dtype).reshape(a.shape))
        mats.append(tuple(trip))
    wt = torch.from_numpy(w.view(np.int64)).pin_memory()
    pinned.append(wt)
    gp = acg.GenQAP(field_id, g.n_rows, g.n_cols, g.layout, mats)
    wp = wt.numpy().view(np.uint64)
    h2d_bytes = sum(int(a.nbytes) for tr in mats for a in tr) + int(wp.nbytes)
    ctx.r1cs_check_host(gp, wp)  # warm-up (allocator, page tables)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    l_e2e0 = ctx.kernel_launch_count()
    for _ in range(e2e_steps):
        nv, _fb = ctx.r1cs_check_host(gp, wp)
        assert nv == 0
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    l_e2e = ctx.kernel_launch_count() - l_e2e0
    sampler.stop()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = total * e2e_steps / e2e_s
    # the end-to-end step of a prover: the circuit (the R1CS) is resident on the device -- as the reference keeps its QAP
    # value in memory between calls of verifyAssignment -- and every step brings a new witness from pinned host memory
    # (H2D + canonical-range check + Montgomery conversion), checks it and reads the result pair back (D2H).
    # N > 1: every rank uploads only its slice of the new witness over its own PCIe link and the slices are exchanged
    # over NVLink (sharding.upload_witness_sliced); falls back to a full upload per rank if that is unavailable
    sliced = world > 1
    dw_bytes = None
    if sliced:
        try:
            dw_bytes = dw.as_torch_bytes()
            sharding.upload_witness_sliced(dw, wp.reshape(-1, 4), None, dw_bytes)
            torch.cuda.synchronize()
        except Exception as e:
            sys.stderr.write("bench.py: sliced witness upload unavailable (%s); every rank uploads the whole witness\n" % (e,))
            sliced = False
        ok = torch.tensor([1 if sliced else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sliced = bool(int(ok.item()))

    def e2e_witness_step():
        if sliced:
            sharding.upload_witness_sliced(dw, wp.reshape(-1, 4), None, dw_bytes)
        else:
            dw.update(wp)
        if peer is not None:
            ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
            return int(result[0].item())
        if world > 1:
            ctx.r1cs_check_async(m, dw, result.data_ptr(), stream.cuda_stream)
            dist.all_reduce(result[0:1], op=dist.ReduceOp.SUM)
            return int(result[0].item())
        return ctx.r1cs_check(m, dw)[0]

    e2e_w_steps = max(e2e_steps, 20) if not args.e2e_steps else e2e_steps
    assert e2e_witness_step() == 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l_w0 = ctx.kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_w_steps):
        assert e2e_witness_step() == 0
    torch.cuda.synchronize()
    e2e_w_s = time.perf_counter() - t0
    l_w = ctx.kernel_launch_count() - l_w0
    t = torch.tensor([e2e_w_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_w_s = float(t.item())
    e2e_w_value = total * e2e_w_steps / e2e_w_s

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this same command
        # (profiles/r01_ncu_summary.json; never measured under the profiler here), only for the default workload
        traffic = None
        if world == 1 and args.log_rows == 20 and args.field == "bn254" and not args.dense and args.kernel == "tiled":
            try:
                with open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")) as f:
                    k2 = json.load(f)["k2_r1cs_tiled"][0]
                traffic = (float(k2["dram__bytes_read.sum"]) + float(k2["dram__bytes_write.sum"])) * 1e6
            except Exception:
                traffic = None
        # average launch duration of the dominant kernel: the sampled event pairs (each pair adds ~3 us of event
        # latency to what it brackets), bounded by the step time when a step is exactly one launch of that kernel
        k_ms = statistics.mean(kernel_ms) if kernel_ms else float("nan")
        one_launch_per_step = launches == args.steps
        if one_launch_per_step and (not kernel_ms or total_ms / args.steps < k_ms):
            k_ms = total_ms / args.steps
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "R1CS constraints/sec (BN254 Fr)" if field_id == 0 else "R1CS constraints/sec (BLS12-381 Fr)",
            "value": value, "unit": "constraints/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (8x32-bit-limb Montgomery Fr, integer)", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "kernel": args.kernel, "tiled_variant": args.variant,
                       "overlap": "consecutive checks overlap (next check kernel launched as a programmatic dependent)"
                                  if not args.no_overlap else "off",
                       "l2": "inputs streamed per step (%.0f MB CSR + %.0f MB witness per GPU) exceed the 126 MB L2; no explicit flush"
                             % ((sum(36 * k for k in g.nnz) + 12 * (g.n_rows + 1)) / 1e6, 32 * g.n_cols / 1e6),
                       "parallelism": "rows sharded over %d rank(s), 1 all-reduce of the result pair per step (%s)" % (
                           world, {"p2p": "fused: the check kernel's last CTA stores the pair into every peer's memory over NVLink (CUDA IPC) and reduces", "nccl": "NCCL",
                                   "none": "single GPU: none"}[collective]),
                       "setup_s": {"generate": round(t_gen, 2)}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "k_r1cs_tiled" if args.kernel == "tiled" else "k_r1cs_rowwise",
                         "algorithmic_bytes_per_launch": algo_bytes, "device_stream_bytes_per_launch": m.stream_bytes + 32 * g.n_cols,
                         "frac_isolated_launch": (algo_bytes / (statistics.mean(kernel_ms) * 1e-3) / 1e9 / peak) if kernel_ms else None,
                         "kernel_ms_mean": k_ms,
                         "kernel_ms_sampled_mean": statistics.mean(kernel_ms) if kernel_ms else None,
                         "kernel_ms_min": min(kernel_ms) if kernel_ms else None,
                         "timing": "kernel_ms_mean = min(sampled mean, step time): a step is exactly one launch of this kernel, "
                                   "and consecutive launches overlap (the next check moves onto the SMs as this one's CTAs "
                                   "run out of tiles), so the steady-state duration per launch is the step time; "
                                   "kernel_ms_sampled_mean / frac_isolated_launch = event pairs around every 8th launch "
                                   "(%d samples), which serialise that launch and add ~3 us of event latency" % len(kernel_ms)
                                   if one_launch_per_step else "event pairs around every 8th step (%d samples)" % len(kernel_ms)},
            "e2e": {"value": e2e_w_value, "unit": "constraints/s",
                    "h2d_bytes_per_step": int(wp.nbytes) // world if sliced else int(wp.nbytes), "d2h_bytes_per_step": 16,
                    "witness_upload": ("each rank uploads 1/%d of the witness over its own PCIe link, slices exchanged over "
                                       "NVLink (one NCCL broadcast per rank)" % world) if sliced else "whole witness per rank",
                    "steps": e2e_w_steps, "ms_per_step": 1e3 * e2e_w_s / e2e_w_steps,
                    "call": "acg_witness_update (new witness from pinned host memory) + acg_r1cs_check against the system "
                            "resident on the device -- the reference, too, keeps its QAP value in memory between calls",
                    "one_shot": {"value": e2e_value, "unit": "constraints/s", "h2d_bytes_per_step": h2d_bytes,
                                 "d2h_bytes_per_step": 16, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                                 "call": "acg_r1cs_check_host: CSR matrices AND witness from pinned host memory every step "
                                         "(PCIe-bound)"}},
            "gpu_launches": launches, "gpu_launches_e2e": l_w, "gpu_launches_e2e_one_shot": l_e2e,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle as CO
            CO.build()
            threads = host_threads()
            v_all, times = cpu_check_throughput(g, w, field_id, threads, 6.0, 20)
            v_one, _ = cpu_check_throughput(g, w, field_id, 1, 4.0, 5)
            line["cpu_baseline"] = {"value": v_all, "unit": "constraints/s", "cores": threads, "kind": "port",
                                    "sample": "%d full checks of the same 2^%d-row system, best of" % (len(times), args.log_rows),
                                    "single_thread_value": v_one}
        print(json.dumps(line), flush=True)
    dw.free()
    m.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so there is one process per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
