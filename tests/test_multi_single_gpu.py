"""lsd_sort_multi on ONE GPU (-m gpu): the ranks are host threads of this process that share device 0, so the whole
multi-GPU path of the C ABI -- device-side plan, peer-store exchange pass (the "peers" are buffers on the same device),
local sort, capacity status -- runs in the single-GPU test tier.  The two collectives are thread-level test doubles
(a host barrier after a stream synchronise; device-to-device copies); on real multi-GPU boxes
tests/test_multi_gpu_device.py runs the same entry over NCCL."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

from lsdradixsort_b200 import _native as N
from lsdradixsort_b200 import api, keygen, multi

pytestmark = pytest.mark.gpu


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _as_bytes(ptr, nbytes):
    return torch.as_tensor(_Raw(ptr, nbytes), device="cuda")


class ThreadComm:
    """all_gather / barrier for `world` threads that share one device."""

    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.send = [None] * world

    def callbacks(self, rank):
        def all_gather(_ctx, send, recv, nbytes, _stream):
            try:
                torch.cuda.synchronize()
                self.send[rank] = send
                self.bar.wait()
                for s in range(self.world):
                    _as_bytes(recv + s * nbytes, nbytes).copy_(_as_bytes(self.send[s], nbytes))
                torch.cuda.synchronize()
                self.bar.wait()
                return 0
            except Exception:  # noqa: BLE001
                return 1

        def barrier(_ctx, _stream):
            try:
                torch.cuda.synchronize()
                self.bar.wait()
                return 0
            except Exception:  # noqa: BLE001
                return 1

        return N.ALL_GATHER_FN(all_gather), N.BARRIER_FN(barrier)


def _rank_main(rank, comm, keys_np, capacity, results, errors):
    try:
        lib = N.lib()
        cb = comm.callbacks(rank)
        mc = N.MultiComm(C.sizeof(N.MultiComm), rank, comm.world, cb[0], cb[1], None)
        src = torch.from_numpy(keys_np.view(np.int32)).cuda()
        recv = torch.full((capacity,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        scratch = torch.empty(capacity, dtype=torch.int32, device="cuda")
        ctx = C.c_void_p()
        st = lib.lsd_multi_ctx_create(C.byref(mc), recv.data_ptr(), capacity, src.numel(), 8, C.byref(ctx), api._stream_ptr(src.device))
        assert st == N.LSD_OK, st
        n_out = C.c_uint64(0)
        st = lib.lsd_sort_multi(ctx, src.data_ptr(), src.numel(), scratch.data_ptr(), C.byref(n_out), api._stream_ptr(src.device))
        torch.cuda.synchronize()
        ms = N.MultiStats()
        lib.lsd_multi_last_stats(ctx, C.byref(ms))
        comm.bar.wait()  # nobody unmaps while a peer may still be writing
        results[rank] = (st, int(n_out.value), recv[: int(n_out.value)].cpu().numpy().view(np.uint32).copy() if st == 0 else None,
                         bool((recv == 0x5A5A5A5A).all()) if st != 0 else None,
                         (int(ms.first_bucket), int(ms.last_bucket), int(ms.n_out_max), int(ms.exchange_shift)))
        lib.lsd_multi_ctx_destroy(ctx)
    except Exception as e:  # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            comm.bar.abort()
        except Exception:  # noqa: BLE001
            pass


def _expected_shift(keys):
    """The exchange window of lsd_sort_multi: the 8 bits that end at the highest bit in which the keys differ."""
    allk = np.concatenate(keys)
    differ = int(np.bitwise_or.reduce(allk)) & ~int(np.bitwise_and.reduce(allk))
    return max(0, differ.bit_length() - 1 - 7)


def _keys(kind, n, seed):
    if kind == "small_range":     # keys below 2^12: the window is bits [4, 12) -- it straddles digits 0 and 1
        return (keygen.make_keys("uniform", n, seed) & np.uint32(0xFFF)).astype(np.uint32)
    if kind == "range17":         # keys below 2^17: the top-digit histogram is one bucket, digit 2 has two
        return (keygen.make_keys("uniform", n, seed) & np.uint32(0x1FFFF)).astype(np.uint32)
    if kind == "offset_range":    # 1000 values next to 2^31: the constant high bits are not zero
        return (np.uint32(0x80000000) + keygen.make_keys("uniform", n, seed) % np.uint32(1000)).astype(np.uint32)
    if kind == "heavy_bucket":    # 90 % of the keys share one top-digit bucket: a rank's share exceeds any sane slack
        u = keygen.make_keys("uniform", n, seed)
        return np.where(u % np.uint32(10) != 0, u & np.uint32(0x00FFFFFF), u).astype(np.uint32)
    return keygen.make_keys(kind, n, seed)


def _run(world, kind, n_local, capacity):
    comm = ThreadComm(world)
    keys = [_keys(kind, n_local, seed=900 + r) for r in range(world)]
    results, errors = [None] * world, []
    threads = [threading.Thread(target=_rank_main, args=(r, comm, keys[r], capacity, results, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    return keys, results


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["uniform", "entropy4_table", "sorted"])
def test_sort_multi_threads_on_one_gpu(world, kind):
    n_local = 200_000 + 17 * world
    keys, results = _run(world, kind, n_local, capacity=n_local * world + 64)
    assert all(r[0] == N.LSD_OK for r in results)
    whole = np.sort(np.concatenate(keys))
    assert np.array_equal(np.concatenate([r[2] for r in results]), whole)  # rank order == global order
    assert sum(r[1] for r in results) == whole.size
    # the device-side plan is the map the host-side numpy plan derives (multi.assign_buckets) from the histogram of the
    # exchange window (the top digit unless the keys are small: "sorted" at this size stays below 2^22)
    shift = _expected_shift(keys)
    per_rank = np.stack([np.bincount((k >> np.uint32(shift)) & np.uint32(255), minlength=256) for k in keys]).astype(np.int64)
    owner = multi.assign_buckets(per_rank.sum(axis=0), world)
    for r, res in enumerate(results):
        assert res[4][3] == shift
        mine = np.nonzero(owner == r)[0]
        if mine.size:
            assert res[4][0] == int(mine[0]) and res[4][1] == int(mine[-1])
        else:
            assert res[4][0] > res[4][1]
        assert res[1] == int(per_rank[:, owner == r].sum())


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind,shift", [("low_nibble", 0), ("small_range", 4), ("range17", 9), ("offset_range", 2)])
def test_sort_multi_partitions_on_the_most_significant_varying_bits(world, kind, shift):
    """BASELINE config 4 x config 5: keys that live in a narrow range are balanced over the 8-bit window that ends at the
    highest bit in which they differ instead of landing on one rank -- the ordinary 25 % slack is enough."""
    n_local = 150_000 + world
    keys, results = _run(world, kind, n_local, capacity=int(n_local * 1.25) + 64)
    assert _expected_shift(keys) == shift
    assert all(r[0] == N.LSD_OK for r in results), [r[0] for r in results]
    assert all(r[4][3] == shift for r in results)
    whole = np.sort(np.concatenate(keys))
    assert np.array_equal(np.concatenate([r[2] for r in results]), whole)
    shares = [r[1] for r in results]
    assert max(shares) <= 1.25 * n_local and min(shares) >= 0.5 * n_local, shares


@pytest.mark.parametrize("world", [1, 3])
def test_sort_multi_all_equal_keys_stay_where_they_are(world):
    n_local = 100_000 + 3 * world
    keys, results = _run(world, "all_equal", n_local, capacity=n_local + 64)
    assert all(r[0] == N.LSD_OK for r in results)
    assert all(r[4][3] == 0xFFFFFFFF and r[1] == n_local for r in results)  # nothing moved
    assert np.array_equal(np.concatenate([r[2] for r in results]), np.sort(np.concatenate(keys)))


def test_sort_multi_capacity_status_on_every_rank_and_nothing_moves():
    world, n_local = 3, 100_000
    keys, results = _run(world, "heavy_bucket", n_local, capacity=int(n_local * 1.25))  # one bucket holds 90 % of the keys
    assert [r[0] for r in results] == [N.LSD_ERR_CAPACITY] * world
    assert len({r[1] for r in results}) == 1 and results[0][1] > 2 * n_local  # *n_out reports the largest share, everywhere
    assert all(r[3] for r in results)  # receive buffers untouched
    keys, results = _run(world, "heavy_bucket", n_local, capacity=results[0][1] + 64)  # the retry the status asks for
    assert all(r[0] == N.LSD_OK for r in results)
    assert np.array_equal(np.concatenate([r[2] for r in results]), np.sort(np.concatenate(keys)))


def test_sort_multi_larger_uniform_two_ranks():
    world, n_local = 2, 1 << 23
    keys, results = _run(world, "uniform", n_local, capacity=int(n_local * 1.25))
    assert all(r[0] == N.LSD_OK for r in results)
    assert np.array_equal(np.concatenate([r[2] for r in results]), np.sort(np.concatenate(keys)))


def test_sort_multi_ragged_and_empty_ranks():
    """Ranks bring different numbers of keys, one of them none at all."""
    world = 3
    sizes = [100_003, 0, 41_000]
    comm = ThreadComm(world)
    keys = [keygen.make_keys("uniform", sizes[r], seed=70 + r) for r in range(world)]
    results, errors = [None] * world, []
    threads = [threading.Thread(target=_rank_main, args=(r, comm, keys[r], 200_000, results, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    assert all(r[0] == N.LSD_OK for r in results)
    assert np.array_equal(np.concatenate([r[2] for r in results]), np.sort(np.concatenate(keys)))
    shares = [r[1] for r in results]
    assert sum(shares) == sum(sizes) and max(shares) - min(shares) < 0.05 * sum(sizes)  # balanced over the ranks, not the sources
