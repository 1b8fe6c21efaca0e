#!/usr/bin/env python3
"""Multi-GPU parity check of the peer-memory all-reduce (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        tests/peer_exchange_ranks.py

Every rank checks its row shard of S(n) with acg_r1cs_check_async_allreduce and must see the GLOBAL violation count and
first bad row that the C oracle computes for the whole system -- for the honest witness and for tampered ones whose
violations fall into different shards -- and the same numbers as the NCCL path (sharding.reduce_check_result)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repository root (this file lives in tests/)
sys.path.insert(0, ROOT)
import arithmetic_circuits_b200 as acg  # noqa: E402
from arithmetic_circuits_b200 import sharding  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 40000 + 13
    g_all, w = acg.synth_r1cs(0, n, 4242)
    rb, re = sharding.row_shard(n, world, rank)
    ctx = acg.Context(0, local)
    m = ctx.upload_r1cs(g_all, rb, re)
    dw = ctx.upload_witness(w)
    peer = sharding.connect_peers(ctx)
    stream = torch.cuda.current_stream()
    result = torch.zeros(2, dtype=torch.int64, device=dev)
    CO.build()
    mats = [(x[0], x[1], x[2]) for x in g_all.mats]
    rng = np.random.default_rng(7)
    cases = [w]
    for k in range(4):
        wb = w.copy()
        for t in rng.integers(1025, 1025 + n, size=1 + 3 * k):
            wb[int(t), 0] ^= np.uint64(1 << 9)
        cases.append(wb)
    for it, wc in enumerate(cases * 3):   # repeated: exercises the sequence-parity buffers
        ref = CO.r1cs_eval_check(0, n, g_all.n_cols, *mats, wc, False, 4)
        want = (ref["n_violations"], ref["first_bad_row"])
        dw.update(wc)
        ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        got = (int(result[0].item()), int(result[1].item()))
        assert got == want, (rank, it, got, want)
        ctx.r1cs_check_async(m, dw, result.data_ptr(), stream.cuda_stream)
        assert sharding.reduce_check_result(result) == want
    # back-to-back launches without host synchronisation in between
    dw.update(w)
    for _ in range(200):
        ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    assert (int(result[0].item()), int(result[1].item())) == (0, -1)
    # sliced witness upload: every rank uploads 1/world of the witness, the slices travel device to device
    dw_bytes = dw.as_torch_bytes()
    for wc in (cases[1], w, cases[3]):
        ref = CO.r1cs_eval_check(0, n, g_all.n_cols, *mats, wc, False, 4)
        sharding.upload_witness_sliced(dw, wc, None, dw_bytes)
        ctx.r1cs_check_async_allreduce(m, dw, peer, result.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        assert (int(result[0].item()), int(result[1].item())) == (ref["n_violations"], ref["first_bad_row"]), rank
        assert (dw.download() == wc).all()
    # a system with fewer rows than ranks: the ranks with an empty shard launch no check kernel but still take part
    n_tiny = max(1, world - 1)
    g_t, w_t = acg.synth_r1cs(0, n_tiny, 99)
    rb, re = sharding.row_shard(n_tiny, world, rank)
    m_t, dw_t = ctx.upload_r1cs(g_t, rb, re), ctx.upload_witness(w_t)
    mats_t = [(x[0], x[1], x[2]) for x in g_t.mats]
    wt_bad = w_t.copy()
    wt_bad[1025, 0] ^= np.uint64(2)   # the first gate's output
    for wc in (w_t, wt_bad, w_t):
        ref = CO.r1cs_eval_check(0, n_tiny, g_t.n_cols, *mats_t, wc, False, 1)
        dw_t.update(wc)
        ctx.r1cs_check_async_allreduce(m_t, dw_t, peer, result.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        assert (int(result[0].item()), int(result[1].item())) == (ref["n_violations"], ref["first_bad_row"]), rank
    dist.barrier()
    if rank == 0:
        print("peer exchange ok: world=%d, %d cases, global count / first bad row == oracle == NCCL path" % (world, 3 * len(cases)))
    peer.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
