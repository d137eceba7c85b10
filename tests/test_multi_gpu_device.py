"""Multi-GPU sort on real devices (-m gpu, needs >= 2 GPUs on the box; skipped otherwise): one process per GPU over NCCL,
lsd_sort_multi (the C ABI's fused peer-store exchange, planned on the device) against the NCCL all_to_all exchange and
against numpy, for uniform and skewed keys; 2 GPUs, every GPU of the box, 2^26 keys per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lsdradixsort_b200 import keygen

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, n_local, q):
    import lsdradixsort_b200 as L
    from lsdradixsort_b200 import multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    L.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        keys_np = keygen.make_keys(kind, n_local, seed=500 + rank)
        src = torch.from_numpy(keys_np.view(np.int32)).to(dev)
        cap = n_local * world + 64
        ops = multi.CudaOps(cap, r=8)
        recv = torch.empty(cap, dtype=torch.int32, device=dev)
        staging = torch.empty(n_local, dtype=torch.int32, device=dev)
        work = src.clone()
        a, _ = multi.distributed_sort(work, ops, recv, staging)  # NCCL exchange, plan on the host (numpy)
        a = a.clone()
        peer = multi.MultiSorter(recv)  # lsd_sort_multi: plan on the device, fused peer-store exchange
        recv.fill_(0x5A5A5A5A)          # nothing of the first result may survive into the second
        b, stats = peer.sort(src)
        b = b.clone()
        recv.fill_(0x3C3C3C3C)
        c, _ = peer.sort(src, timing=True)  # again, through the timed flavour
        torch.cuda.synchronize()
        assert torch.equal(src.cpu(), torch.from_numpy(keys_np.view(np.int32))), "lsd_sort_multi must not modify its input"
        # an undersized receive buffer: every rank gets the same status, nothing is moved
        small = torch.empty(max(64, n_local // 2), dtype=torch.int32, device=dev)
        tiny = multi.MultiSorter(small, max_n_local=n_local)
        small.fill_(7)
        try:
            tiny.sort(src)
            capacity_error = None
        except multi.CapacityError as e:
            capacity_error = (e.needed, e.capacity)
        torch.cuda.synchronize()
        untouched = bool((small == 7).all())
        tiny.close()
        q.put((rank, a.cpu().numpy().view(np.uint32).copy(), b.cpu().numpy().view(np.uint32).copy(),
               c.cpu().numpy().view(np.uint32).copy(), stats.n_out, capacity_error, untouched))
        dist.barrier()
        peer.close()
    finally:
        dist.destroy_process_group()


def _run(world, kind, n_local):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    whole = np.sort(np.concatenate([keygen.make_keys(kind, n_local, seed=500 + r) for r in range(world)]))
    assert np.array_equal(np.concatenate([r[1] for r in results]), whole)  # NCCL all_to_all path
    assert np.array_equal(np.concatenate([r[2] for r in results]), whole)  # lsd_sort_multi
    assert np.array_equal(np.concatenate([r[3] for r in results]), whole)  # lsd_sort_multi, timed
    assert sum(r[4] for r in results) == whole.size
    errs = {r[5] for r in results}
    assert len(errs) == 1 and None not in errs, f"every rank must report the same capacity error: {errs}"
    assert all(r[6] for r in results), "an aborted exchange must not move any key"


@pytest.mark.parametrize("kind", ["uniform", "entropy4_table", "sorted", "low_nibble", "all_equal"])
def test_distributed_sort_two_gpus_peer_and_nccl(kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    _run(2, kind, 300_000 + 13)


def test_distributed_sort_two_gpus_2pow26_keys_per_rank():
    """Bit-exact against numpy at a roofline-sized share: 2^26 keys per rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    _run(2, "uniform", 1 << 26)


@pytest.mark.parametrize("kind", ["uniform", "sorted"])
def test_distributed_sort_all_gpus(kind):
    world = torch.cuda.device_count()
    if world < 4:
        pytest.skip("needs at least four GPUs on the box")
    _run(world, kind, (1 << 21) + 5)
