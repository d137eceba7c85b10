"""Multi-GPU sort: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU; this is the partitioning BASELINE.json's north_star asks
for:

  1. every rank histograms the TOP digit of its keys (one row of the upfront digit histogram);
  2. ``all_reduce`` of the 256-bin row gives the global top-digit histogram -> every rank derives the
     same bucket->rank map (contiguous key ranges, balanced by prefix sums); an ``all_gather`` of the
     rows (2 KiB per rank) gives the all-to-all split sizes;
  3. one stable pass on the top digit groups the local keys by destination rank (lsd_sort_pass);
  4. ``all_to_all_single`` moves the buckets: each rank now owns a contiguous key range;
  5. local LSD sort of what arrived.

Rank r ends with the r-th slice of the globally sorted sequence.  The exchange is the path's one
real collective; everything else is per-rank.  Device work goes through a small ``ops`` object so
the host logic (steps 2 and 4) can be tested on CPU with gloo and a test double.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

TOP_BITS = 8
BUCKETS = 1 << TOP_BITS


def assign_buckets(global_hist: np.ndarray, nranks: int) -> np.ndarray:
    """bucket -> owning rank: contiguous, monotone, balanced on the global bucket sizes.

    Bucket b goes to rank floor(nranks * (keys before b + half of b) / total): every rank gets a
    contiguous run of buckets whose total is as close to total/nranks as whole buckets allow.  The
    balance is only as fine as one bucket (an all-equal input lands on one rank)."""
    h = np.asarray(global_hist, dtype=np.float64)
    total = h.sum()
    if total == 0:
        return np.minimum(np.arange(BUCKETS) * nranks // BUCKETS, nranks - 1).astype(np.int64)
    mid = np.cumsum(h) - h / 2.0
    owner = np.minimum((mid * nranks / total).astype(np.int64), nranks - 1)
    return np.maximum.accumulate(owner)  # monotone even with empty buckets


def split_sizes(per_rank_hist: np.ndarray, owner: np.ndarray, rank: int) -> Tuple[List[int], List[int]]:
    """(input_splits, output_splits) of this rank for all_to_all_single.

    per_rank_hist[s][b] = keys of bucket b held by rank s before the exchange."""
    nranks = per_rank_hist.shape[0]
    send = np.zeros((nranks, nranks), dtype=np.int64)  # send[s][d]
    for d in range(nranks):
        send[:, d] = per_rank_hist[:, owner == d].sum(axis=1)
    return [int(x) for x in send[rank]], [int(x) for x in send[:, rank]]


class CudaOps:
    """Device side of the distributed sort on this rank's GPU (liblsdsort through the C ABI)."""

    def __init__(self, capacity: int, r: int = 8, block: int = 0):
        from . import api

        self.api = api
        self.r, self.block = r, block
        self.sorter = api.Sorter(capacity, r=r, block=block)
        self.capacity = capacity

    def top_digit_histogram(self, keys: torch.Tensor) -> torch.Tensor:
        return self.api.digit_histograms(keys, self.r)[-1]  # [256] int64, row of the top digit

    def partition_by_top_digit(self, keys: torch.Tensor, out: torch.Tensor) -> None:
        self.api.sort_pass(keys, out, self.r, 32 // self.r - 1, self.block, workspace=self.sorter.workspace)

    def sort_(self, keys: torch.Tensor) -> None:
        self.sorter.sort_(keys)

    def empty(self, n: int) -> torch.Tensor:
        return torch.empty(n, dtype=torch.int32, device=self.sorter.device)


@dataclass
class ExchangeStats:
    n_in: int
    n_out: int
    sent_bytes: int  # bytes leaving this rank (excludes the part it keeps)
    recv_bytes: int
    owner_first_bucket: int
    owner_last_bucket: int


def distributed_sort(keys: torch.Tensor, ops, recv: torch.Tensor, staging: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, ExchangeStats]:
    """Sort the union of every rank's ``keys``; returns (this rank's sorted slice, stats).

    ``recv`` (capacity >= what this rank will own) and ``staging`` (>= len(keys)) are caller-owned
    buffers, so a timed loop does not allocate.  ``keys`` is overwritten."""
    rank, nranks = dist.get_rank(group), dist.get_world_size(group)
    n = keys.numel()
    if ops.r != TOP_BITS:
        raise ValueError("the exchange partitions on the top 8-bit digit: use r=8")

    # 1-2. top-digit histograms -> identical bucket map and split sizes on every rank
    local = ops.top_digit_histogram(keys).to(torch.int64)
    global_hist = local.clone()
    dist.all_reduce(global_hist, op=dist.ReduceOp.SUM, group=group)  # the MSD-histogram all-reduce
    gathered = [torch.empty_like(local) for _ in range(nranks)]
    dist.all_gather(gathered, local, group=group)  # per-rank rows: who sends how much to whom
    per_rank = torch.stack(gathered).cpu().numpy()
    owner = assign_buckets(global_hist.cpu().numpy(), nranks)
    in_splits, out_splits = split_sizes(per_rank, owner, rank)
    n_out = sum(out_splits)
    if n_out > recv.numel():
        raise RuntimeError(f"rank {rank}: receives {n_out} keys but recv buffer holds {recv.numel()} "
                           "(skewed top digit; raise the capacity slack)")

    # 3. stable partition by top digit == grouped by destination rank (owner is monotone in the bucket)
    ops.partition_by_top_digit(keys, staging)

    # 4. bucket exchange
    out = recv[:n_out]
    dist.all_to_all_single(out, staging[:n], out_splits, in_splits, group=group)

    # 5. local LSD sort of the owned key range
    ops.sort_(out)
    mine = np.nonzero(owner == rank)[0]
    stats = ExchangeStats(n, n_out, 4 * (n - in_splits[rank]), 4 * (n_out - out_splits[rank]),
                          int(mine[0]) if mine.size else -1, int(mine[-1]) if mine.size else -1)
    return out, stats
