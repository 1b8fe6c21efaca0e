"""Multi-GPU row sharding of the R1CS check: one process per GPU (torch.distributed),
rows split in contiguous blocks, witness replicated, and ONE all-reduce of the per-shard result pair
{violated-row count, first violated row}.  connect_peers() sets up the exchange over peer memory that the
check kernel performs itself (include/acg.h "multi-GPU row shards"); reduce_check_result() is the plain
collective form of the same reduction (sum, then min only when the count is non-zero) for backends without
peer access -- it is also what the CPU tests run over gloo.  No other data-path collective exists: the NTT /
QAP build stays single-GPU."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

INT64_MAX = (1 << 63) - 1


def row_shard(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rows for `rank`: sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_shard_balanced(rowptrs: Sequence[np.ndarray], world_size: int, rank: int) -> Tuple[int, int]:
    """nnz-balanced contiguous split (Split gates make rows very uneven, src/QAP.hs:443-473): cut points
    where the cumulative entry count of A+B+C crosses k/world_size of the total."""
    n_rows = len(rowptrs[0]) - 1
    cum = np.zeros(n_rows + 1, dtype=np.int64)
    for rp in rowptrs:
        cum += np.asarray(rp, dtype=np.int64)
    cum += np.arange(n_rows + 1, dtype=np.int64)  # one unit per row so empty rows still spread
    total = int(cum[-1])
    cuts = [int(np.searchsorted(cum, (total * k) // world_size, side="left")) for k in range(world_size + 1)]
    cuts[0], cuts[-1] = 0, n_rows
    for k in range(1, world_size + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts[rank], cuts[rank + 1]


def rebalance_cuts(cuts: Sequence[int], times: Sequence[float]) -> List[int]:
    """Cost-balanced contiguous row blocks from one measurement: `cuts` (world + 1 boundaries) are the current blocks
    and `times` what each block's check took.  The cost per row is taken as constant inside a measured block -- later
    rows of a circuit built gate by gate gather from a wider part of the witness and cost more -- and the new boundaries
    cut the cumulative cost into equal parts.  Rows stay contiguous and in order; every block keeps at least one row."""
    world = len(times)
    assert len(cuts) == world + 1 and all(t > 0 for t in times)
    total = float(sum(times))
    new = [cuts[0]]
    k, acc = 0, 0.0     # block being consumed, cost before it
    for j in range(1, world):
        target = total * j / world
        while k < world - 1 and acc + times[k] < target:
            acc += times[k]
            k += 1
        frac = (target - acc) / times[k]
        b = cuts[k] + int(round(frac * (cuts[k + 1] - cuts[k])))
        new.append(min(max(b, new[-1] + 1), cuts[-1] - (world - j)))
    new.append(cuts[-1])
    return new


def connect_peers(ctx, group=None):
    """PeerExchange for the ranks of a torch.distributed group (one process per GPU on one NVSwitch node): the 64-byte
    CUDA IPC handles are all-gathered through the group, then every rank maps its peers' exchange buffers."""
    import torch
    import torch.distributed as dist
    from .qap import PeerExchange

    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def all_gather_bytes(mine: bytes):
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        return [bytes(x.cpu().tolist()) for x in out]

    return PeerExchange(ctx, world, rank, all_gather_bytes)


def upload_witness_sliced(dw, w_host: np.ndarray, group=None, dw_bytes=None):
    """New witness for every rank of a row-sharded check, without every rank pulling the whole vector over PCIe:
    rank r uploads the r-th contiguous slice of `w_host` (canonical limbs, the same array on every rank) into its own
    device vector over its own PCIe link (range check + Montgomery conversion on the device), then the slices travel
    device to device -- one broadcast per rank over NVLink (NCCL).  dw_bytes: cached dw.as_torch_bytes()."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = len(dw)
    t = dw_bytes if dw_bytes is not None else dw.as_torch_bytes()
    lo, hi = row_shard(n, world, rank)
    dw.update_range(w_host[lo:hi], lo)
    for r in range(world):
        a, b = row_shard(n, world, r)
        if b > a:
            dist.broadcast(t[a * 32:b * 32], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def gather_plan(n: int, world: int):
    """How a vector of n elements is exchanged between `world` ranks with ONE all-gather: equal slices of
    n // world elements (rank r owns [r * s, (r + 1) * s)) plus a remainder of n % world elements at the end that every
    rank uploads itself.  Returns (slice length, remainder start)."""
    s = n // world
    return s, s * world


def upload_witness_allgather(dw, w_host: np.ndarray, stream, group=None, dw_bytes=None):
    """New witness for every rank of a row-sharded check, enqueue only: rank r copies ITS slice of `w_host` (canonical
    limbs, pinned, the same array on every rank) to its device vector over its own PCIe link on the library's copy
    stream (acg_witness_update_async: range check + Montgomery conversion on the device), `stream` (a torch CUDA stream)
    waits for that copy, and ONE in-place NCCL all-gather over NVLink completes the vector on every rank.  With two
    device vectors used alternately the host-to-device copy of witness i + 1 overlaps the all-gather and the check of
    witness i.  The check that follows must be enqueued on `stream`."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = len(dw)
    t = dw_bytes if dw_bytes is not None else dw.as_torch_bytes()
    s, rem0 = gather_plan(n, world)
    if s:
        dw.update_async(w_host[rank * s:(rank + 1) * s], rank * s)
    if rem0 < n:
        dw.update_async(w_host[rem0:], rem0)
    dw.stream_wait(stream.cuda_stream)
    if s:
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(t[:world * s * 32], t[rank * s * 32:(rank + 1) * s * 32], group=group)


def reduce_check_result(result, group=None):
    """result: int64 tensor [n_violations, first_bad_row] (first_bad_row = -1, i.e. UINT64_MAX, when the
    shard is clean), on the device of the process group's backend.  Returns (total violations, first bad
    row or -1).  One all-reduce on the common path."""
    import torch
    import torch.distributed as dist

    count = result[0:1]
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    total = int(count.item())
    if total == 0:
        return 0, -1
    first = result[1:2].clone()
    first[first < 0] = INT64_MAX
    dist.all_reduce(first, op=dist.ReduceOp.MIN, group=group)
    return total, int(first.item())
