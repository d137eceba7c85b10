"""Multi-GPU sort: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU; this is the partitioning BASELINE.json's north_star asks
for:

  1. every rank histograms the TOP digit of its keys (one row of the upfront digit histogram);
  2. an ``all_gather`` of the 256-bin rows (2 KiB per rank; their sum is the all-reduced MSD histogram) -> every rank
     derives the same bucket->rank map (contiguous key ranges, balanced by prefix sums) and the split sizes;
  3. one stable pass on the top digit groups the local keys by destination rank;
  4. the buckets move: each rank now owns a contiguous key range;
  5. local LSD sort of what arrived.

Rank r ends with the r-th slice of the globally sorted sequence.  Two exchanges:

* ``MultiSorter`` (default, ``distributed_sort(..., peer=...)``): a caller of the C ABI's ``lsd_sort_multi``
  (include/lsdsort.h, csrc/multi.cu).  Steps 2-4 are planned on the device and executed by ONE pass kernel that stores
  every bucket straight into its owner's receive buffer over NVLink peer memory; torch.distributed only supplies the two
  collectives (an 8 KiB all-gather, a barrier) as callbacks.  It partitions on the 8-bit window that ends at the highest
  bit that VARIES over the input (keys in any narrow range are balanced over their own 8 most significant varying bits;
  all-equal keys stay where they are).
* the NCCL path (``peer=None``): ``lsd_sort_pass`` + ``all_to_all_single`` with the plan computed on the host with numpy.
  Kept as the parity reference for the fused path and for the CPU (gloo) tests of the host logic, which replace the
  device steps by a test double.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import ctypes as C

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
    h = np.asarray(global_hist).astype(np.int64)
    total = int(h.sum())
    if total == 0:
        return np.minimum(np.arange(BUCKETS) * nranks // BUCKETS, nranks - 1).astype(np.int64)
    twice_mid = 2 * (np.cumsum(h) - h) + h  # exact integers, the same arithmetic as multi_plan_kernel (csrc/multi.cu)
    owner = np.minimum(twice_mid * nranks // (2 * total), nranks - 1).astype(np.int64)
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


class CapacityError(RuntimeError):
    """Some rank's share of the keys exceeds its receive buffer (skewed top digit).  Raised on EVERY rank together (the
    plan is identical everywhere), before any key is moved; ``needed`` is the largest share."""

    def __init__(self, needed: int, capacity: int):
        self.needed, self.capacity = needed, capacity
        super().__init__(f"a rank would own {needed} keys but the receive buffers hold {capacity}: the balance is only as "
                         "fine as one top-digit bucket; retry with more capacity")


class _RawDeviceBuffer:
    """A device pointer handed to a callback by liblsdsort, wrapped for torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class MultiSorter:
    """``lsd_sort_multi`` (C ABI) with torch.distributed supplying the two collectives.

    COLLECTIVE to construct and to call.  ``recv`` is this rank's receive buffer (CUDA int32 tensor; it is exported with
    CUDA IPC and mapped by the other ranks, so it must come from its own allocation, e.g. ``torch.empty``)."""

    def __init__(self, recv: torch.Tensor, r: int = 8, group: Optional[dist.ProcessGroup] = None,
                 max_n_local: Optional[int] = None):
        """``max_n_local``: the most keys this rank will ever bring to ``sort`` (default: ``recv.numel()``)."""
        from . import _native as N
        from . import api

        self.N, self.api, self.group = N, api, group
        self.rank, self.nranks = dist.get_rank(group), dist.get_world_size(group)
        self.recv, self.r = recv, r
        self.device = recv.device
        self._token = torch.zeros(1, dtype=torch.int32, device=recv.device)
        self._error = None
        self._timing = False

        def all_gather(_ctx, send, out, nbytes, _stream):
            try:
                src = torch.as_tensor(_RawDeviceBuffer(send, nbytes), device=self.device)
                dst = torch.as_tensor(_RawDeviceBuffer(out, nbytes * self.nranks), device=self.device)
                dist.all_gather_into_tensor(dst, src, group=self.group)
                return 0
            except Exception as e:  # noqa: BLE001 -- must not unwind through the C frame
                self._error = e
                return 1

        def barrier(_ctx, _stream):
            try:
                dist.all_reduce(self._token, op=dist.ReduceOp.SUM, group=self.group)  # stream-ordered, no host sync
                return 0
            except Exception as e:  # noqa: BLE001
                self._error = e
                return 1

        self._cb = (N.ALL_GATHER_FN(all_gather), N.BARRIER_FN(barrier))  # keep the thunks alive
        self._comm = N.MultiComm(C.sizeof(N.MultiComm), self.rank, self.nranks, self._cb[0], self._cb[1], None)
        self._ctx = C.c_void_p()
        self._scratch = torch.empty(recv.numel(), dtype=torch.int32, device=recv.device)
        self._check(N.lib().lsd_multi_ctx_create(C.byref(self._comm), recv.data_ptr(), recv.numel(),
                                                 recv.numel() if max_n_local is None else int(max_n_local), r,
                                                 C.byref(self._ctx), api._stream_ptr(recv.device)), "lsd_multi_ctx_create")

    def _check(self, status: int, where: str) -> None:
        if status == self.N.LSD_ERR_COMM and self._error is not None:
            e, self._error = self._error, None
            raise RuntimeError(f"{where}: torch.distributed callback failed") from e
        self.N.check(status, where)

    def sort(self, keys: torch.Tensor, timing: bool = False) -> Tuple[torch.Tensor, "ExchangeStats"]:
        """Sort the union of every rank's ``keys`` (not modified); returns (this rank's slice of ``recv``, stats).
        ``timing``: bracket the three stages with CUDA events (reading them synchronises the stream)."""
        if timing != self._timing:
            self.N.check(self.N.lib().lsd_multi_set_timing(self._ctx, 1 if timing else 0), "lsd_multi_set_timing")
            self._timing = timing
        if not keys.is_cuda or keys.dtype != torch.int32 or not keys.is_contiguous():
            raise TypeError("keys must be a contiguous CUDA int32 tensor")
        n_out = C.c_uint64(0)
        st = self.N.lib().lsd_sort_multi(self._ctx, keys.data_ptr(), keys.numel(), self._scratch.data_ptr(), C.byref(n_out),
                                         self.api._stream_ptr(keys.device))
        if st == self.N.LSD_ERR_CAPACITY:
            raise CapacityError(int(n_out.value), self.recv.numel())
        self._check(st, "lsd_sort_multi")
        ms = self.N.MultiStats()
        self.N.check(self.N.lib().lsd_multi_last_stats(self._ctx, C.byref(ms)), "lsd_multi_last_stats")
        owns = ms.first_bucket <= ms.last_bucket
        stats = ExchangeStats(int(ms.n_in), int(ms.n_out), int(ms.sent_bytes), 0,
                              int(ms.first_bucket) if owns else -1, int(ms.last_bucket) if owns else -1)
        stats.exchange_shift = None if ms.exchange_shift == 0xFFFFFFFF else int(ms.exchange_shift)
        if timing:
            stats.stage_ms_direct = {"hist_allgather_plan": float(ms.plan_ms), "partition": float(ms.exchange_ms),
                                     "all_to_all": 0.0, "local_sort": float(ms.sort_ms)}
        return self.recv[: int(n_out.value)], stats

    def close(self) -> None:
        if self._ctx:
            self.N.lib().lsd_multi_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


PeerExchange = MultiSorter  # round-1 name of the fused exchange object


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
    exchange_shift: Optional[int] = 24  # the exchange partitioned on bits [shift, shift + 8) (lsd_sort_multi: the window that ends at the highest varying bit; None = all keys equal)
    stage_ms_direct: Optional[dict] = None  # lsd_sort_multi reports the stage times itself

    def stage_ms(self) -> dict:
        """Device time per stage (call after a synchronize): histogram+collectives, partition, exchange, sort."""
        if self.stage_ms_direct is not None:
            return self.stage_ms_direct
        names = ("hist_allgather_plan", "partition", "all_to_all", "local_sort")
        return {k: self.events[i].elapsed_time(self.events[i + 1]) for i, k in enumerate(names)}


def distributed_sort(keys: torch.Tensor, ops, recv: torch.Tensor, staging: torch.Tensor,
                     group: Optional[dist.ProcessGroup] = None, timing: bool = False,
                     peer: Optional["MultiSorter"] = None) -> Tuple[torch.Tensor, ExchangeStats]:
    """Sort the union of every rank's ``keys``; returns (this rank's sorted slice, stats).

    ``peer`` (a MultiSorter built on ``recv``): the whole sort is one call of the C ABI's lsd_sort_multi.  Otherwise the
    NCCL path: ``recv`` (capacity >= what this rank will own) and ``staging`` (>= len(keys)) are caller-owned buffers, so
    a timed loop does not allocate, and ``keys`` is overwritten."""
    if peer is not None:
        return peer.sort(keys, timing=timing)
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
    # every rank computes every rank's share from the same table, so all of them raise together (a rank that raised
    # alone would leave the others blocked in the collective below)
    shares = [int(per_rank[:, owner == d].sum()) for d in range(nranks)]
    if max(shares) > recv.numel():
        raise CapacityError(max(shares), recv.numel())

    mark()
    out = recv[:n_out]
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
