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


def scatter_destinations(per_rank_hist: np.ndarray, owner: np.ndarray, rank: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Where this rank's keys go in the fused partition + exchange: (dest_rank[b], dest_offset[b], seg[b]).

    All buckets owned by one rank form one destination segment (``owner`` is monotone, so a segment is a run of
    consecutive buckets; seg[b] = first | last << 16).  The receive buffer of rank o holds one block per source rank, in
    rank order: the block of source ``rank`` starts at sum_{s < rank} send[s][o] keys (the same split all_to_all_single
    would use) and is filled tile after tile by the pass kernel.  dest_offset is equal for all buckets of a segment."""
    nranks = per_rank_hist.shape[0]
    send = np.zeros((nranks, nranks), dtype=np.int64)
    for d in range(nranks):
        send[:, d] = per_rank_hist[:, owner == d].sum(axis=1)
    before_me = send[:rank].sum(axis=0)  # [dest] keys that sources ahead of me put into dest's buffer
    dest_off = before_me[owner]
    seg = np.zeros(BUCKETS, dtype=np.int64)
    for o in range(nranks):
        idx = np.nonzero(owner == o)[0]
        if idx.size:
            seg[idx] = int(idx[0]) | (int(idx[-1]) << 16)
    return owner.astype(np.int64), dest_off.astype(np.int64), seg


class PeerExchange:
    """Receive buffers of all ranks mapped into every process (CUDA IPC over NVLink / NVSwitch peer access), so that the
    top-digit partition pass stores each bucket straight into its owner's buffer: partition and exchange are ONE kernel
    and no all-to-all runs.  Ordering between ranks is a 1-element all_reduce before and after the pass."""

    def __init__(self, recv: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
        from . import api

        self.api = api
        self.group = group
        self.rank, self.nranks = dist.get_rank(group), dist.get_world_size(group)
        self.recv = recv
        mine = api.ipc_export(recv)
        handles = [None] * self.nranks
        dist.all_gather_object(handles, mine, group=group)
        self.peer_ptr = []
        for s, (h, off) in enumerate(handles):
            self.peer_ptr.append(recv.data_ptr() if s == self.rank else api.ipc_open(h, off))
        self._token = torch.zeros(1, dtype=torch.int32, device=recv.device)
        self._dst = torch.empty(BUCKETS, dtype=torch.int64, device=recv.device)
        self._dst_host = torch.empty(BUCKETS, dtype=torch.int64).pin_memory()
        self._seg = torch.empty(BUCKETS, dtype=torch.int32, device=recv.device)
        self._seg_host = torch.empty(BUCKETS, dtype=torch.int32).pin_memory()

    def fence(self) -> None:
        """Stream-ordered barrier across ranks (no host synchronisation)."""
        dist.all_reduce(self._token, op=dist.ReduceOp.SUM, group=self.group)

    def scatter(self, keys: torch.Tensor, ops, dest_rank: np.ndarray, dest_off: np.ndarray, seg: np.ndarray) -> None:
        ptrs = np.asarray([self.peer_ptr[int(r)] for r in dest_rank], dtype=np.int64) + 4 * dest_off
        self._dst_host.copy_(torch.from_numpy(ptrs))
        self._seg_host.copy_(torch.from_numpy(seg.astype(np.int32)))
        self._dst.copy_(self._dst_host, non_blocking=True)
        self._seg.copy_(self._seg_host, non_blocking=True)
        self.fence()  # every rank is done reading what the previous exchange left in its buffer
        self.api.sort_pass_scatter(keys, self._dst, ops.r, 32 // ops.r - 1, workspace=ops.sorter.workspace,
                                   dst_seg=self._seg)
        self.fence()  # every rank's stores have landed


class CudaOps:
    """Device side of the distributed sort on this rank's GPU (liblsdsort through the C ABI)."""

    def __init__(self, capacity: int, r: int = 8, block: int = 0):
        from . import api

        self.api = api
        self.r, self.block = r, block
        self.sorter = api.Sorter(capacity, r=r, block=block)
        self.capacity = capacity

    def top_digit_histogram(self, keys: torch.Tensor) -> torch.Tensor:
        return self.api.top_digit_histogram(keys, self.r)  # [256] int64

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
    events: Optional[list] = None  # CUDA events between the stages when timing was requested

    def stage_ms(self) -> dict:
        """Device time per stage (call after a synchronize): histogram+collectives, partition, exchange, sort."""
        names = ("hist_allreduce_allgather", "partition", "all_to_all", "local_sort")
        return {k: self.events[i].elapsed_time(self.events[i + 1]) for i, k in enumerate(names)}


def distributed_sort(keys: torch.Tensor, ops, recv: torch.Tensor, staging: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None, timing: bool = False,
                     peer: Optional["PeerExchange"] = None) -> Tuple[torch.Tensor, ExchangeStats]:
    """Sort the union of every rank's ``keys``; returns (this rank's sorted slice, stats).

    ``recv`` (capacity >= what this rank will own) and ``staging`` (>= len(keys)) are caller-owned
    buffers, so a timed loop does not allocate.  ``keys`` is overwritten."""
    rank, nranks = dist.get_rank(group), dist.get_world_size(group)
    n = keys.numel()
    if ops.r != TOP_BITS:
        raise ValueError("the exchange partitions on the top 8-bit digit: use r=8")

    events = []

    def mark():
        if timing:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append(e)

    mark()
    # 1-2. top-digit histograms -> identical bucket map and split sizes on every rank
    local = ops.top_digit_histogram(keys).to(torch.int64)
    gathered = torch.empty(nranks * BUCKETS, dtype=torch.int64, device=local.device)
    # one collective: the per-rank rows (who sends how much to whom); their sum is the all-reduced MSD histogram
    dist.all_gather_into_tensor(gathered, local.contiguous().view(-1), group=group)
    per_rank = gathered.view(nranks, BUCKETS).cpu().numpy()
    owner = assign_buckets(per_rank.sum(axis=0), nranks)
    in_splits, out_splits = split_sizes(per_rank, owner, rank)
    n_out = sum(out_splits)
    if n_out > recv.numel():
        raise RuntimeError(f"rank {rank}: receives {n_out} keys but recv buffer holds {recv.numel()} "
                           "(skewed top digit; raise the capacity slack)")

    mark()
    out = recv[:n_out]
    if peer is not None:
        # 3+4 fused: the partition pass stores every bucket into its owner's receive buffer over NVLink
        dest_rank, dest_off, seg = scatter_destinations(per_rank, owner, rank)
        peer.scatter(keys, ops, dest_rank, dest_off, seg)
        mark()
        mark()
    else:
        # 3. stable partition by top digit == grouped by destination rank (owner is monotone in the bucket)
        ops.partition_by_top_digit(keys, staging)
        mark()
        # 4. bucket exchange
        dist.all_to_all_single(out, staging[:n], out_splits, in_splits, group=group)
        mark()

    # 5. local LSD sort of the owned key range
    ops.sort_(out)
    mark()
    mine = np.nonzero(owner == rank)[0]
    stats = ExchangeStats(n, n_out, 4 * (n - in_splits[rank]), 4 * (n_out - out_splits[rank]),
                          int(mine[0]) if mine.size else -1, int(mine[-1]) if mine.size else -1, events if timing else None)
    return out, stats
