"""Multi-GPU sort on real devices (-m gpu, needs >= 2 GPUs on the box; skipped otherwise): one process per GPU over NCCL,
the fused peer-scatter exchange against the NCCL all_to_all exchange and against numpy, for uniform and skewed keys."""
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
        a, _ = multi.distributed_sort(work, ops, recv, staging)  # NCCL exchange
        a = a.clone()
        peer = multi.PeerExchange(recv)
        work.copy_(src)
        b, stats = multi.distributed_sort(work, ops, recv, staging, peer=peer)  # fused peer-scatter exchange
        torch.cuda.synchronize()
        q.put((rank, a.cpu().numpy().view(np.uint32).copy(), b.cpu().numpy().view(np.uint32).copy(), stats.n_out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["uniform", "entropy4_table", "sorted"])
def test_distributed_sort_two_gpus_peer_and_nccl(kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    world, n_local = 2, 300_000 + 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    whole = np.sort(np.concatenate([keygen.make_keys(kind, n_local, seed=500 + r) for r in range(world)]))
    nccl = np.concatenate([r[1] for r in results])
    fused = np.concatenate([r[2] for r in results])
    assert np.array_equal(nccl, whole)
    assert np.array_equal(fused, whole)
    assert sum(r[3] for r in results) == whole.size
