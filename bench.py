#!/usr/bin/env python3
"""Benchmark of the hot path: R1CS witness check (A.w o B.w - C.w == 0 over BN254 Fr) on synthetic
circuits S(n, seed, field) of SURVEY.md 8(d).  Metric: R1CS constraints / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log-rows L] [--field bn254]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload.  N = 1: BASELINE configs[1], S(2^20, 20260002, BN254) on one GPU.  N > 1: BASELINE configs[3], STRONG
scaling: S(2^24, 20260004, BN254), rows split in N contiguous blocks, witness replicated (`--scaling weak --log-rows L`
gives 2^L rows per GPU instead).  `--workload mix`: a circuit with the reference generator's gate mix (Mul : Equal :
Split = 50 : 10 : 1, 256-bit Split rows) instead of the Split-free family.

One "step" = one check of every constraint of the (sharded) system = ONE launch of the tiled CUDA kernel per rank (the
kernel's last CTA finalises the result pair and, for N > 1, all-reduces it over peer memory -- NVLink P2P stores;
`--collective nccl` uses one NCCL all-reduce instead).

  value     device-resident throughput: K steps between CUDA events on the launching stream, barrier + synchronize on
            both sides, max over ranks.  The steps alternate between two resident witnesses and consecutive checks
            overlap (acg_ctx_set_overlap_checks, opt-in: the next check kernel is a programmatic dependent launch).
  roofline  k_r1cs_tiled: SURVEY 8(d) algorithmic bytes of the shard / the ISOLATED per-launch duration -- a second
            timed region of K plain launches, each bracketed by its own CUDA event pair on the launching stream
            (acg_profile_begin/end) -- vs MEASURED_PEAKS.json.  frac_overlapped is the same with the step time of the
            first region.
  e2e       same metric end to end per witness through the C ABI with HOST buffers: the system stays resident on the
            device (the reference keeps its QAP value in memory between calls, too); every step copies a new witness
            from pinned host memory (acg_witness_update_async on the copy stream: H2D, range check, Montgomery
            conversion; two device vectors, so the copy of witness i + 1 overlaps the check of witness i), checks it
            and reads the result pair back.  N > 1: every rank copies 1/N of the witness over its own PCIe link and one
            NCCL all-gather over NVLink completes it.  e2e.one_shot: everything from host buffers every step.
  cpu_baseline / --impl reference
            the oracle's C restatement of the reference algorithm (oracle/r1cs_oracle.c, "port": the Haskell
            reference cannot be built in this image) on the host cores.
  qap       (N = 1) BASELINE configs[2] as a secondary object: acg_qap_witness on S(2^22, 20260003): wall and kernel
            time of the call, h checked against the C oracle.
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
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="auto: N = 1 -> 2^20 rows; N > 1 -> strong scaling of 2^24 rows (BASELINE configs[3])")
    ap.add_argument("--log-rows", type=int, default=0,
                    help="log2 of constraints: per GPU (weak) or in total (strong); 0 = 20 / 24 per --scaling")
    ap.add_argument("--field", default="bn254", choices=list(FIELD_IDS))
    ap.add_argument("--workload", default="s", choices=["s", "mix"],
                    help="s: family S(n) of SURVEY 8(d); mix: the reference generator's gate mix with 256-bit Split gates")
    ap.add_argument("--dense", action="store_true", help="all coefficients uniform (stress variant)")
    ap.add_argument("--kernel", default="tiled", choices=["tiled", "rowwise"])
    ap.add_argument("--no-overlap", action="store_true",
                    help="first timed region: do not overlap consecutive checks (plain launches)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: all-reduce of the result pair over peer memory (default) or by NCCL")
    ap.add_argument("--variant", type=int, default=0, choices=list(range(9)), help="tiled kernel geometry (kernels.h kTileGeom): 0-3 rows per tile 128/256/64/32, 4/5 = 128/64 with the natural term layout, 6 = products in place, 7 = 6 + one far buffer (6 CTAs per SM), 8 = room for 4 products per row (chosen automatically for dense systems)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = max(steps, 50) capped at 200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-qap", action="store_true", help="skip the secondary configs[2] measurement (N = 1)")
    ap.add_argument("--no-one-shot", action="store_true", help="skip e2e.one_shot")
    ap.add_argument("--no-balance", action="store_true",
                    help="N > 1: keep equal row blocks (default: one timing pass, then cost-balanced contiguous blocks)")
    return ap.parse_args()


def plan(args, world):
    """(rows in total, rows per rank or None, seed, scaling label)"""
    scaling = args.scaling
    if scaling == "auto":
        scaling = "strong" if world > 1 else "weak"
    if scaling == "strong":
        log_total = args.log_rows or (24 if world > 1 else 20)
        total = 1 << log_total
        return total, SEEDS.get(log_total, 20260000 + log_total), scaling
    log_per = args.log_rows or 20
    total = world << log_per
    return total, SEEDS.get(log_per, 20260000 + log_per), scaling


def workload_name(args, world):
    total, seed, scaling = plan(args, world)
    if args.workload == "mix":
        return "M(n>=%d, seed=%d, %s): gate mix of the reference generator (Mul:Equal:Split = 50:10:1, 256-bit Split rows), %d GPU(s)" % (
            total, seed, args.field, world)
    return "S(n=%d, seed=%d, %s%s): %d Mul-gate R1CS rows, ~5.5 nnz/row, split in %d contiguous row block(s) (%s scaling), witness replicated" % (
        total, seed, args.field, ", dense" if args.dense else "", total, world, scaling)


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


def make_workload(acg, args, world, rank, rows=None):
    """-> (GenQAP of this rank's rows, full honest witness, (row_begin, row_end), total rows, generate seconds).
    rows: this rank's block when it is not the equal split."""
    from arithmetic_circuits_b200 import sharding
    field_id = FIELD_IDS[args.field]
    total, seed, _ = plan(args, world)
    t0 = time.perf_counter()
    if args.workload == "mix":
        if world > 1:
            raise SystemExit("bench.py: --workload mix is a single-GPU workload")
        g, w = acg.synth_mixed_r1cs(field_id, total, seed)
        return g, w, (0, g.n_rows), g.n_rows, time.perf_counter() - t0
    rb, re = rows if rows is not None else sharding.row_shard(total, world, rank)
    g, w = acg.synth_r1cs(field_id, total, seed, args.dense, rows=(rb, re))
    return g, w, (rb, re), total, time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Haskell original has no toolchain
    here) on rank 0's host cores, all threads, same workload and metric."""
    if rank != 0:
        return
    import arithmetic_circuits_b200 as acg  # host-side generator only (no GPU use in this arm)
    from oracle import c_oracle as CO
    CO.build()
    field_id = FIELD_IDS[args.field]
    # the bounded sample: rank 0's row block of the system (all of it at N = 1)
    g, w, (rb, re), total, _ = make_workload(acg, args, world, 0)
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
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": plan(args, world)[2], "vs_baseline": None,
        "dtype": "u256 (4x64-bit Montgomery, unsigned __int128)", "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "note": "CPU arm checks one row block (%d rows: 1/%d of the system) per step on rank 0" % (g.n_rows, world)},
        "cpu_baseline": {"value": value, "unit": "constraints/s", "cores": threads, "kind": "port",
                         "sample": "%d full checks of a %d-row block, OpenMP over rows" % (steps, g.n_rows)},
        "e2e": {"value": value, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def qap_secondary(acg, ctx, field_id):
    """BASELINE configs[2]: S(2^22, 20260003) -- acg_qap_witness (check + A.w, B.w, C.w + 7 NTTs of 2^22 points + coset
    quotient): wall and kernel time of the call with pinned host outputs, h compared with the C oracle."""
    import numpy as np
    import torch
    from oracle import c_oracle as CO
    n = 1 << 22
    g, w = acg.synth_r1cs(field_id, n, SEEDS[22])
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    h_pin = torch.empty((n + 1, 4), dtype=torch.int64).pin_memory()
    L = acg._lib.lib()
    import ctypes as C
    div = C.c_int()
    hp = C.c_void_p(h_pin.data_ptr())

    def call():
        acg.qap._check(L.acg_qap_witness(ctx._h, m._h, dw._h, None, None, None, None, hp, C.byref(div)), ctx)
    call()
    walls, kernels = [], []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        walls.append(1e3 * (time.perf_counter() - t0))
        kernels.append(ctx.last_timing()["kernel_ms"])
    ref = CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *[(x[0], x[1], x[2]) for x in g.mats], w, True, host_threads())
    t0 = time.perf_counter()
    _a, _b, _c, h_ref, ok_ref = CO.qap_witness(field_id, ref["Aw"], ref["Bw"], ref["Cw"], (0, 0, 0), host_threads())
    cpu_s = time.perf_counter() - t0
    exact = bool(div.value) and ok_ref and bool((h_pin.numpy().view(np.uint64) == h_ref).all())
    # isolated 2^22-point NTT (device resident, natural order in and out)
    stream = torch.cuda.current_stream()
    dv = ctx.upload_witness(w[:n])
    ntt_ms = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        acg.qap._check(L.acg_ntt_device(ctx._h, dv._h, 22, i & 1, C.c_void_p(stream.cuda_stream)), ctx)
        e1.record(stream)
        torch.cuda.synchronize()
        ntt_ms.append(e0.elapsed_time(e1))
    dv.free()
    out = {"workload": "S(n=%d, seed=%d): acg_qap_witness = R1CS check + A.w, B.w, C.w + 7 NTTs of 2^22 points + coset quotient -> h" % (n, SEEDS[22]),
           "wall_ms": min(walls), "kernel_ms": min(kernels), "wall_over_kernel": min(walls) / min(kernels),
           "d2h_bytes": int(h_pin.numel() * 8), "h_bit_exact_vs_oracle": exact,
           "ntt_2p22_ms": min(ntt_ms[1:]), "cpu_oracle_s": cpu_s, "cpu_threads": host_threads(),
           "constraints_per_s": n / (min(walls) * 1e-3)}
    dw.free()
    m.free()
    return out


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
    numa = None
    if world > 1:
        # one process per GPU: run (and allocate the pinned witness buffers) on the CPUs next to this rank's GPU, so that
        # the N host-to-device copies of a step do not all pull from one NUMA node
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "cpu affinity set to the GPU's NUMA node (%d cpus)" % len(os.sched_getaffinity(0))
        except Exception as e:
            numa = "not set (%s)" % (e,)
        dist.init_process_group("nccl", device_id=dev)
    field_id = FIELD_IDS[args.field]
    scaling = plan(args, world)[2]
    g, w, (rb, re), total, t_gen = make_workload(acg, args, world, rank)

    ctx = acg.Context(field_id, local_rank)
    ctx.set_check_kernel(acg.CHECK_TILED if args.kernel == "tiled" else acg.CHECK_ROWWISE)
    ctx.set_tiled_variant(args.variant)
    t_up = time.perf_counter()
    m = ctx.upload_r1cs(g)
    m.set_row_offset(rb)   # this rank's rows are [rb, re) of the global system
    t_up = time.perf_counter() - t_up
    # N > 1: the row blocks are contiguous but need not be equal -- later rows of a circuit built gate by gate gather
    # from a wider part of the (replicated) witness and cost more.  One timing pass over the equal blocks, then the
    # boundaries are moved so that every rank gets the same COST (sharding.rebalance_cuts) and the blocks are rebuilt.
    balance = None
    if world > 1 and not args.no_balance and args.workload == "s":
        dw0 = ctx.upload_witness(w)
        r0 = torch.zeros(2, dtype=torch.int64, device=dev)
        st0 = torch.cuda.current_stream()
        for _ in range(3):
            ctx.r1cs_check_async(m, dw0, r0.data_ptr(), st0.cuda_stream)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ea.record(st0)
        for _ in range(10):
            ctx.r1cs_check_async(m, dw0, r0.data_ptr(), st0.cuda_stream)
        eb.record(st0)
        torch.cuda.synchronize()
        mine = torch.tensor([ea.elapsed_time(eb) / 10, float(rb), float(re)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        times = [float(x[0].item()) for x in allr]
        cuts = [int(x[1].item()) for x in allr] + [int(allr[-1][2].item())]
        new_cuts = sharding.rebalance_cuts(cuts, times)
        balance = {"equal_blocks_kernel_ms": [round(t, 5) for t in times], "rows_per_rank": [new_cuts[i + 1] - new_cuts[i] for i in range(world)]}
        dw0.free()
        m.free()
        del g, w
        g, w, (rb, re), total, t_gen2 = make_workload(acg, args, world, rank, rows=(new_cuts[rank], new_cuts[rank + 1]))
        t_gen += t_gen2
        t1 = time.perf_counter()
        m = ctx.upload_r1cs(g)
        m.set_row_offset(rb)
        t_up += time.perf_counter() - t1
    dws = [ctx.upload_witness(w), ctx.upload_witness(w)]   # two resident witnesses, checked alternately
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

    def step(dw):
        if peer is not None:
            ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
            return
        ctx.r1cs_check_async(m, dw, result.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(result[0:1], op=dist.ReduceOp.SUM)

    # ---- N > 1: correctness of the fused all-reduce before anything is timed: a tampered witness must give, on every
    # rank, the global (count, first bad row) the CPU oracle finds over the shards
    peer_check = None
    if world > 1:
        from oracle import c_oracle as CO
        CO.build()
        wb = w.copy()
        wb[1025 + total // 3, 0] += np.uint64(1)      # SURVEY 8(d) negative variant
        wb[1025 + 7, 1] ^= np.uint64(1 << 5)          # .. and one wire every shard references
        ref = CO.r1cs_eval_check(field_id, g.n_rows, g.n_cols, *[(x[0], x[1], x[2]) for x in g.mats], wb, False, 8)
        loc = torch.tensor([ref["n_violations"], (rb + ref["first_bad_row"]) if ref["first_bad_row"] >= 0 else (1 << 62)],
                           dtype=torch.int64, device=dev)
        cnt, first = loc[0:1].clone(), loc[1:2].clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(first, op=dist.ReduceOp.MIN)
        dws[1].update(wb)
        step(dws[1])
        torch.cuda.synchronize()
        want = (int(cnt.item()), int(first.item()))
        got = (int(result[0].item()), int(result[1].item()) if peer is not None else want[1])  # (NCCL path: count only)
        assert got == want and want[0] > 0, "rank %d: all-reduced result %r != oracle %r" % (rank, got, want)
        peer_check = {"tampered_witness": {"violations": got[0], "first_bad_row": got[1]},
                      "cpu_oracle_over_all_shards": {"violations": want[0], "first_bad_row": want[1]},
                      "asserted_equal_on_every_rank": True, "collective": collective}
        dws[1].update(w)

    sampler = ClockSampler(local_rank)
    ctx.set_overlap_checks(not args.no_overlap)
    for i in range(max(args.warmup, 3)):
        step(dws[i & 1])
    torch.cuda.synchronize()
    assert int(result[0].item()) == 0, "honest witness must verify"

    # ---- timed region 1 (value): K steps back to back, alternating the two resident witnesses; one launch per step
    launches0 = ctx.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record(stream)
    for i in range(args.steps):
        step(dws[i & 1])
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launch_count() - launches0
    assert int(result[0].item()) == 0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = total * args.steps / (total_ms * 1e-3)

    # ---- timed region 2 (roofline): the ISOLATED launch duration of the check kernel -- K plain launches (no overlap of
    # consecutive checks) of the local shard, back to back between two CUDA events on the launching stream: the average
    # launch duration over the timed region, launch gaps included.  For N > 1 this is the local shard without the peer
    # exchange, i.e. the per-rank kernel time without the epilogue spin.  A few launches are also bracketed one by one
    # by their own event pair (acg_profile_begin/end; the pair adds ~3 us of event latency to what it brackets).
    ctx.set_overlap_checks(False)
    k_iso = max(args.steps, 10)
    for i in range(3):
        ctx.r1cs_check_async(m, dws[i & 1], result.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    l_iso0 = ctx.kernel_launch_count()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for i in range(k_iso):
        ctx.r1cs_check_async(m, dws[i & 1], result.data_ptr(), stream.cuda_stream)
    f1.record(stream)
    torch.cuda.synchronize()
    iso_ms = f0.elapsed_time(f1) / k_iso
    launches_per_check = (ctx.kernel_launch_count() - l_iso0) / k_iso
    k_pair = min(k_iso, 16)
    ctx.profile_begin(k_pair)
    for i in range(k_pair):
        ctx.r1cs_check_async(m, dws[i & 1], result.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    kernel_ms = ctx.profile_end(k_pair)
    per_rank_ms = [iso_ms]
    if world > 1:
        tt = torch.tensor([iso_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allr, tt)
        per_rank_ms = [float(x.item()) for x in allr]
        iso_ms = max(per_rank_ms)   # the slowest shard bounds the step

    # ---- end to end, one shot: everything from pinned host buffers every step (PCIe-bound)
    e2e_steps = min(max(args.e2e_steps or max(args.steps, 50), 1), 200)
    pinned = []
    one_shot = None
    wt = torch.from_numpy(w.view(np.int64)).pin_memory()
    pinned.append(wt)
    wp = wt.numpy().view(np.uint64)
    if not args.no_one_shot and world == 1:
        mats = []
        for rp, col, val in g.mats:
            trip = []
            for a in (rp, col, val):
                tt = torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype == np.uint32 else np.int64)).pin_memory()
                pinned.append(tt)
                trip.append(tt.numpy().view(a.dtype).reshape(a.shape))
            mats.append(tuple(trip))
        gp = acg.GenQAP(field_id, g.n_rows, g.n_cols, g.layout, mats)
        h2d_bytes = sum(int(a.nbytes) for tr in mats for a in tr) + int(wp.nbytes)
        ctx.r1cs_check_host(gp, wp)  # warm-up (allocator, page tables)
        torch.cuda.synchronize()
        os_steps = min(e2e_steps, 5)
        t0 = time.perf_counter()
        l0 = ctx.kernel_launch_count()
        for _ in range(os_steps):
            nv, _fb = ctx.r1cs_check_host(gp, wp)
            assert nv == 0
        torch.cuda.synchronize()
        os_s = time.perf_counter() - t0
        one_shot = {"value": total * os_steps / os_s, "unit": "constraints/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 16, "steps": os_steps, "ms_per_step": 1e3 * os_s / os_steps,
                    "gpu_launches": ctx.kernel_launch_count() - l0,
                    "call": "acg_r1cs_check_host: CSR matrices AND witness from pinned host memory every step (PCIe-bound)"}

    # ---- end to end per witness (the headline): the circuit is resident on the device -- as the reference keeps its QAP
    # value in memory between calls of verifyAssignment -- and every step brings a NEW witness from pinned host memory
    # through the C ABI (H2D + range check + Montgomery conversion), checks it and reads the result pair back (D2H).
    # Two device vectors: the copy of witness i + 1 (copy stream) overlaps the check of witness i.
    # N > 1: every rank copies only its 1/N slice over its own PCIe link; one NCCL all-gather over NVLink completes it.
    wt2 = torch.from_numpy(w.view(np.int64).copy()).pin_memory()   # a second host witness (same values, other buffer)
    pinned.append(wt2)
    hosts = [wp.reshape(-1, 4), wt2.numpy().view(np.uint64).reshape(-1, 4)]
    sliced = world > 1
    views = [d.as_torch_bytes() for d in dws] if sliced else None

    def e2e_upload(i):
        if sliced:
            sharding.upload_witness_allgather(dws[i & 1], hosts[i & 1], stream, None, views[i & 1])
        else:
            dws[i & 1].update_async(hosts[i & 1])

    host_res = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]
    done_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_enqueue_check(i):   # check of witness i + read-back of the result pair, enqueue only
        if peer is not None:
            ctx.r1cs_check_async_allreduce(m, dws[i & 1], peer, result.data_ptr(), stream.cuda_stream)
        else:
            ctx.r1cs_check_async(m, dws[i & 1], result.data_ptr(), stream.cuda_stream)
            if world > 1:
                dist.all_reduce(result[0:1], op=dist.ReduceOp.SUM)
        host_res[i & 1].copy_(result, non_blocking=True)
        done_ev[i & 1].record(stream)

    def e2e_run(k):
        e2e_upload(0)
        for i in range(k):
            e2e_enqueue_check(i)
            if i + 1 < k:
                e2e_upload(i + 1)       # enqueue only: the copy overlaps the check just enqueued
            done_ev[i & 1].synchronize()   # the step's result is on the host
            assert int(host_res[i & 1][0]) == 0
    e2e_run(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The region is run e2e_trials times and the MEDIAN trial is reported (all trials are listed): the host-to-device
    # copies share the box's PCIe switches and host memory with whatever else runs on it, and single trials were seen
    # 40 % apart on the same box within a minute.
    e2e_trials = 5
    trial_s = []
    for _ in range(e2e_trials):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l_w0 = ctx.kernel_launch_count()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        trial_s.append(time.perf_counter() - t0)
        l_w = ctx.kernel_launch_count() - l_w0
    for d in dws:
        d.status()   # (raises if an asynchronous update had met a non-canonical element)
    # the floor of that step: the host-to-device copies alone (every rank its slice, all ranks at once), no all-gather,
    # no check -- what the PCIe links / host memory of the box deliver when N GPUs pull at the same time
    h2d_only_ms = None
    if sliced:
        s_len, rem0 = sharding.gather_plan(len(dws[0]), world)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(10):
            dws[i & 1].update_async(hosts[i & 1][rank * s_len:(rank + 1) * s_len], rank * s_len)
        for d in dws:
            d.status()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d_only_ms = 1e3 * float(tt.item()) / 10
    sampler.stop()
    t = torch.tensor(trial_s, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # every trial: the slowest rank
    trial_s = sorted(float(x) for x in t.tolist())
    e2e_w_s = trial_s[len(trial_s) // 2]
    e2e_w_value = total * e2e_steps / e2e_w_s

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this same command
        # (never measured under the profiler here), only for the default workload
        traffic, traffic_src = None, None
        if world == 1 and total == (1 << 20) and args.field == "bn254" and not args.dense and args.kernel == "tiled" \
                and args.workload == "s":
            for name in ("r02_ncu_summary.json", "r01_ncu_summary.json"):
                try:
                    with open(os.path.join(ROOT, "profiles", name)) as f:
                        k2 = json.load(f)["k2_r1cs_tiled"][0]
                    traffic = (float(k2["dram__bytes_read.sum"]) + float(k2["dram__bytes_write.sum"])) * 1e6
                    traffic_src = "profiles/%s (committed ncu --set full capture of this command; not measured in this run)" % name
                    break
                except Exception:
                    continue
        step_ms = total_ms / args.steps
        achieved = algo_bytes / (iso_ms * 1e-3) / 1e9
        achieved_ovl = algo_bytes / (step_ms * 1e-3) / 1e9
        line = {
            "metric": "R1CS constraints/sec (BN254 Fr)" if field_id == 0 else "R1CS constraints/sec (BLS12-381 Fr)",
            "value": value, "unit": "constraints/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "u256 (8x32-bit-limb Montgomery Fr, integer)", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "kernel": args.kernel, "tiled_variant": args.variant,
                       "rows_per_gpu": g.n_rows,
                       "timed_region": "K checks back to back, alternating two resident witnesses; "
                                       + ("consecutive checks overlap (opt-in acg_ctx_set_overlap_checks: programmatic dependent launch)"
                                          if not args.no_overlap else "plain launches (no overlap)"),
                       "l2": "inputs streamed per step (%.0f MB CSR + %.0f MB witness per GPU) exceed the 126 MB L2; no explicit flush"
                             % ((sum(36 * k for k in g.nnz) + 12 * (g.n_rows + 1)) / 1e6, 32 * g.n_cols / 1e6),
                       "parallelism": "rows split over %d rank(s), 1 all-reduce of the result pair per step (%s)" % (
                           world, {"p2p": "fused: the check kernel's last CTA stores the pair into every peer's memory over NVLink (CUDA IPC) and reduces", "nccl": "NCCL",
                                   "none": "single GPU: none"}[collective]),
                       "setup_s": {"generate": round(t_gen, 2), "upload": round(t_up, 2)}, "host_numa": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "k_r1cs_tiled" if args.kernel == "tiled" else "k_r1cs_rowwise",
                         "algorithmic_bytes_per_launch": algo_bytes, "device_stream_bytes_per_launch": m.stream_bytes + 32 * g.n_cols,
                         "kernel_ms_mean": iso_ms, "launches_timed": k_iso, "launches_per_check": launches_per_check,
                         "kernel_ms_event_pair_mean": statistics.mean(kernel_ms), "kernel_ms_event_pair_min": min(kernel_ms),
                         "frac_overlapped": achieved_ovl / peak, "step_ms_overlapped": step_ms,
                         "per_rank_kernel_ms": per_rank_ms,
                         "timing": "frac = algorithmic bytes / mean ISOLATED launch duration: %d plain launches of the check kernel back "
                                   "to back (no overlap of consecutive checks) between two CUDA events on the launching stream "
                                   "(second timed region; for N > 1 the local shard without the peer exchange, slowest rank); "
                                   "kernel_ms_event_pair_*: %d launches bracketed one by one by their own event pair; frac_overlapped "
                                   "= the same bytes / the step time of the first timed region, where consecutive checks overlap"
                                   % (k_iso, len(kernel_ms))},
            "e2e": {"value": e2e_w_value, "unit": "constraints/s",
                    "h2d_bytes_per_step": (int(wp.nbytes) // world) if sliced else int(wp.nbytes), "d2h_bytes_per_step": 16,
                    "witness_upload": ("each rank copies 1/%d of the witness over its own PCIe link (copy stream), one NCCL "
                                       "all-gather over NVLink completes it; two device vectors, copy i+1 overlaps all-gather + check i" % world)
                                      if sliced else "whole witness, copy stream; two device vectors, copy i+1 overlaps check i",
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_w_s / e2e_steps,
                    "trials_ms_per_step": [round(1e3 * x / e2e_steps, 5) for x in trial_s], "reported": "median trial",
                    "h2d_only_ms_per_step": h2d_only_ms,
                    "call": "acg_witness_update_async (new witness from pinned host memory) + acg_r1cs_check against the system "
                            "resident on the device -- the reference, too, keeps its QAP value in memory between calls",
                    "one_shot": one_shot},
            "gpu_launches": launches, "gpu_launches_isolated_region": int(launches_per_check * k_iso), "gpu_launches_e2e": l_w,
            "clocks": sampler.summary(),
        }
        if peer_check is not None:
            line["config"]["peer_allreduce_check"] = peer_check
        if balance is not None:
            line["config"]["row_blocks"] = dict(balance, note="contiguous blocks balanced by measured cost (one timing pass over equal blocks)")
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle as CO
            CO.build()
            threads = host_threads()
            v_all, times = cpu_check_throughput(g, w, field_id, threads, 6.0, 20)
            v_one, _ = cpu_check_throughput(g, w, field_id, 1, 4.0, 5)
            line["cpu_baseline"] = {"value": v_all, "unit": "constraints/s", "cores": threads, "kind": "port",
                                    "sample": "%d full checks of the same %d-row system, best of" % (len(times), g.n_rows),
                                    "single_thread_value": v_one}
        if world == 1 and not args.no_qap and args.workload == "s" and not args.dense:
            for d in dws:
                d.free()
            m.free()
            dws, m = [], None
            try:
                line["qap"] = qap_secondary(acg, ctx, field_id)
            except Exception as e:  # secondary: never lose the main line
                line["qap"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    for d in dws:
        d.free()
    if m is not None:
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
