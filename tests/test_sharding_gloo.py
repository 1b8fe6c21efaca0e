"""world_size-2 gloo test of the multi-GPU plumbing: contiguous row shards, one all-reduce of
the per-shard violated-row count, min-reduction of the first bad row.  On CPU the per-shard count comes
from the C oracle (the product has no CPU compute path); what is under test is the sharding and the
reduction, which the GPU path uses unchanged."""
import os
import socket

import numpy as np
import pytest

from oracle import c_oracle as CO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, tamper, q):
    import torch
    import torch.distributed as dist
    import arithmetic_circuits_b200 as acg
    from arithmetic_circuits_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, w = acg.synth_r1cs(0, n, seed)
        for t in tamper:
            w[t, 0] ^= np.uint64(1)
        rb, re = sharding.row_shard_balanced([m[0] for m in g.mats], world, rank) if seed % 2 else \
            sharding.row_shard(g.n_rows, world, rank)
        # shard-local CSR slices, as acg_r1cs_upload(row_begin, row_end) takes them
        mats = []
        for rp, col, val in g.mats:
            e0, e1 = int(rp[rb]), int(rp[re])
            mats.append(((rp[rb:re + 1] - rp[rb]).astype(np.uint32), col[e0:e1].copy(), val[e0:e1].copy()))
        res = CO.r1cs_eval_check(0, re - rb, g.n_cols, mats[0], mats[1], mats[2], w)
        first = -1 if res["first_bad_row"] < 0 else res["first_bad_row"] + rb
        t = torch.tensor([res["n_violations"], first], dtype=torch.int64)
        total, first_bad = sharding.reduce_check_result(t)
        if rank == 0:
            q.put((total, first_bad, (rb, re)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tamper,seed", [((), 20260002), ((1025 + 100,), 20260002), ((1025 + 100, 1025 + 900), 3)])
def test_sharded_check_two_ranks(tamper, seed):
    import torch.multiprocessing as mp
    import arithmetic_circuits_b200 as acg

    n, world = 1500, 2
    g, w = acg.synth_r1cs(0, n, seed)
    for t in tamper:
        w[t, 0] ^= np.uint64(1)
    full = CO.r1cs_eval_check(0, n, g.n_cols, *[(m[0], m[1], m[2]) for m in g.mats], w)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, tamper, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, first_bad, shard0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (total, first_bad) == (full["n_violations"], full["first_bad_row"])
    assert (total == 0) == (len(tamper) == 0)
    assert shard0[0] == 0 and 0 < shard0[1] < n


def test_row_shard_partitions():
    from arithmetic_circuits_b200 import sharding
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            cuts = [sharding.row_shard(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    # nnz-balanced: one huge row does not starve the others
    rp = np.concatenate([[0], np.cumsum([1] * 100 + [10000] + [1] * 100)]).astype(np.uint32)
    cuts = [sharding.row_shard_balanced([rp, rp, rp], 4, r) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 201 and all(cuts[i][1] == cuts[i + 1][0] for i in range(3))


def _gather_worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    from arithmetic_circuits_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the exchange of sharding.upload_witness_allgather on host tensors: every rank holds only its slice (and the
        # remainder) of a 32-byte-per-element vector, one in-place all-gather completes it
        full = torch.arange(n * 32, dtype=torch.int64).remainder(251).to(torch.uint8)
        mine = torch.zeros_like(full)
        s, rem0 = sharding.gather_plan(n, world)
        mine[rank * s * 32:(rank + 1) * s * 32] = full[rank * s * 32:(rank + 1) * s * 32]
        mine[rem0 * 32:] = full[rem0 * 32:]
        if s:
            dist.all_gather_into_tensor(mine[:world * s * 32], mine[rank * s * 32:(rank + 1) * s * 32].clone())
        q.put((rank, bool((mine == full).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 2, 1025 + 1500, 4097])
def test_witness_slices_one_allgather(n):
    """The slice plan of the row-sharded witness upload (equal slices + remainder) reassembles the vector with one
    all-gather, world size 2, gloo."""
    import torch.multiprocessing as mp
    from arithmetic_circuits_b200 import sharding

    for world in (1, 2, 3, 8):
        s, rem0 = sharding.gather_plan(n, world)
        assert s * world == rem0 <= n and n - rem0 < world
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [(0, True), (1, True)]


def test_rebalance_cuts():
    """Cost-balanced row blocks from one timing of equal blocks: contiguous, ordered, non-empty, and equal in cost under
    the piecewise-constant cost model."""
    from arithmetic_circuits_b200 import sharding
    n, world = 1 << 24, 8
    cuts = [sharding.row_shard(n, world, r)[0] for r in range(world)] + [n]
    times = [0.1159, 0.1174, 0.1195, 0.1225, 0.1245, 0.1273, 0.129, 0.1308]
    new = sharding.rebalance_cuts(cuts, times)
    assert new[0] == 0 and new[-1] == n and all(a < b for a, b in zip(new, new[1:]))
    dens = [t / (cuts[i + 1] - cuts[i]) for i, t in enumerate(times)]
    cost = lambda a, b: sum(dens[i] * max(0, min(b, cuts[i + 1]) - max(a, cuts[i])) for i in range(world))
    costs = [cost(new[i], new[i + 1]) for i in range(world)]
    assert max(costs) - min(costs) < 1e-6 * max(costs) * 100
    assert new[1] - new[0] > new[-1] - new[-2]      # the cheap early rows: bigger blocks
    assert sharding.rebalance_cuts([0, 10, 20], [1.0, 1.0]) == [0, 10, 20]
    assert sharding.rebalance_cuts([0, 1, 2, 3], [1.0, 5.0, 1.0]) == [0, 1, 2, 3]      # nothing to move
    assert sharding.rebalance_cuts([0, 100, 200], [3.0, 1.0]) == [0, 67, 200]
