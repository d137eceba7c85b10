"""CPU tests of the multi-GPU path's host logic: bucket->rank assignment, split sizes and
the all-to-all plumbing, run with world_size 2 and 3 on gloo.  The device steps are replaced by an
oracle-backed test double (tests may use the oracle; the product's CudaOps never does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle
from lsdradixsort_b200 import keygen, multi


def test_assign_buckets_uniform_is_even_and_monotone():
    h = np.full(256, 1000)
    for nranks in (1, 2, 4, 8):
        owner = multi.assign_buckets(h, nranks)
        assert owner[0] == 0 and owner[-1] == nranks - 1
        assert np.all(np.diff(owner) >= 0)
        assert np.all(np.bincount(owner, minlength=nranks) == 256 // nranks)


def test_assign_buckets_skewed_and_empty():
    h = np.zeros(256, dtype=np.int64)
    h[200] = 10_000  # all-equal input: one bucket, lands on exactly one rank
    owner = multi.assign_buckets(h, 4)
    assert np.all(np.diff(owner) >= 0) and 0 <= owner[200] <= 3
    h = np.zeros(256, dtype=np.int64)
    assert np.all(np.diff(multi.assign_buckets(h, 8)) >= 0)
    rng = np.random.default_rng(0)
    h = rng.integers(0, 1000, 256) * (rng.random(256) < 0.3)
    owner = multi.assign_buckets(h, 8)
    assert np.all(np.diff(owner) >= 0) and owner.max() <= 7
    load = np.bincount(owner, weights=h, minlength=8)
    assert load.max() <= h.sum() / 8 + h.max()  # within one bucket of perfect balance


def test_split_sizes_conserve_keys():
    rng = np.random.default_rng(1)
    per_rank = rng.integers(0, 500, (4, 256))
    owner = multi.assign_buckets(per_rank.sum(axis=0), 4)
    total_in, total_out = 0, 0
    mat = []
    for r in range(4):
        ins, outs = multi.split_sizes(per_rank, owner, r)
        assert sum(ins) == per_rank[r].sum()
        mat.append((ins, outs))
        total_in += sum(ins)
        total_out += sum(outs)
    assert total_in == total_out == per_rank.sum()
    for s in range(4):
        for d in range(4):
            assert mat[s][0][d] == mat[d][1][s]  # what s sends to d is what d expects from s


def test_scatter_destinations_partition_every_receive_buffer():
    """Fused partition+exchange: the blocks the sources write into one receive buffer tile it exactly (no gap, no
    overlap), in source-rank order, with the split sizes all_to_all_single would use; segments are owner runs."""
    rng = np.random.default_rng(7)
    world, n_local = 3, 4000
    keys = [rng.integers(0, 2**32, n_local, dtype=np.uint64).astype(np.uint32) for _ in range(world)]
    keys[1][:1500] = 0x7F000001  # a heavy bucket
    per_rank = np.stack([np.bincount(k >> 24, minlength=256) for k in keys]).astype(np.int64)
    owner = multi.assign_buckets(per_rank.sum(axis=0), world)
    got = [[] for _ in range(world)]
    for s in range(world):
        dest_rank, dest_off, seg = multi.scatter_destinations(per_rank, owner, s)
        ins, _ = multi.split_sizes(per_rank, owner, s)
        assert np.array_equal(dest_rank, owner)
        for o in range(world):
            idx = np.nonzero(owner == o)[0]
            if idx.size == 0:
                continue
            assert len(set(dest_off[idx].tolist())) == 1  # one block per (source, destination)
            assert all(int(v) == (int(idx[0]) | (int(idx[-1]) << 16)) for v in seg[idx])
            got[o].append((int(dest_off[idx[0]]), ins[o], s))
    for o in range(world):
        _, outs = multi.split_sizes(per_rank, owner, o)
        pos = 0
        for off, size, s in sorted(got[o]):
            assert off == pos and size == outs[s]
            pos += size
        assert pos == sum(outs)


class OracleOps:
    """Test double for multi.CudaOps: same interface, CPU tensors, oracle arithmetic."""

    r = 8

    def top_digit_histogram(self, keys):
        k = keys.numpy().view(np.uint32)
        return torch.from_numpy(_oracle.digit_histograms(k, 8)[-1].astype(np.int64))

    def partition_by_top_digit(self, keys, out):
        k = keys.numpy().view(np.uint32).copy()
        o = np.zeros_like(k)
        _oracle.oracle().lsd_oracle_sort_pass(k, o, k.size, np.zeros(256, dtype=np.uint32), 8, 3)
        out[: k.size] = torch.from_numpy(o.view(np.int32))

    def sort_(self, keys):
        k = keys.numpy().view(np.uint32)
        keys.copy_(torch.from_numpy(_oracle.sort(k, 8).view(np.int32)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, n_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys_np = keygen.make_keys(kind, n_local, seed=100 + rank)
        keys = torch.from_numpy(keys_np.view(np.int32).copy())
        recv = torch.empty(n_local * world + 16, dtype=torch.int32)
        staging = torch.empty(n_local, dtype=torch.int32)
        out, stats = multi.distributed_sort(keys, OracleOps(), recv, staging)
        q.put((rank, out.numpy().view(np.uint32).copy(), stats.n_out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "uniform"), (3, "entropy4_table"), (2, "all_equal"), (2, "sorted")])
def test_distributed_sort_gloo(world, kind):
    n_local = 5000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    whole = np.concatenate([keygen.make_keys(kind, n_local, seed=100 + r) for r in range(world)])
    got = np.concatenate([r[1] for r in results])  # rank order == global order
    assert np.array_equal(got, np.sort(whole))
    assert sum(r[2] for r in results) == whole.size


def _worker_capacity(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys = torch.from_numpy(keygen.make_keys("all_equal", 4000, seed=1).view(np.int32).copy())
        recv = torch.empty(5000, dtype=torch.int32)  # holds one rank's share of a uniform input, not all 8000 equal keys
        staging = torch.empty(4000, dtype=torch.int32)
        try:
            multi.distributed_sort(keys, OracleOps(), recv, staging)
            q.put((rank, None))
        except multi.CapacityError as e:
            q.put((rank, (e.needed, e.capacity)))
        dist.barrier()  # every rank is still alive and in step: nobody is stuck in the exchange
    finally:
        dist.destroy_process_group()


def test_distributed_sort_capacity_error_is_raised_on_every_rank():
    """An all-equal input lands on one rank.  The rank that owns nothing must raise as well, or the owner's peers would
    block in all_to_all_single (multi.py:188)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_capacity, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in results] == [(8000, 5000)] * world
