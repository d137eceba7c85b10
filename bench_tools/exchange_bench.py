"""Multi-GPU exchange breakdown (run under torchrun, one rank per GPU): per-stage CUDA-event times of
multi.distributed_sort and the raw all_to_all_single bandwidth at the same message size."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402
from lsdradixsort_b200 import multi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n-total", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", type=str, default="")
ap.add_argument("--peer", action="store_true", help="fused partition + exchange over peer memory instead of all_to_all")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
L.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
n = (1 << args.log2n_total) // world
g = torch.Generator(device=dev).manual_seed(99 + rank)
src = torch.empty(n, dtype=torch.int32, device=dev)
for lo in range(0, n, 1 << 26):
    hi = min(n, lo + (1 << 26))
    src[lo:hi] = torch.randint(-(2**31), 2**31, (hi - lo,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
work = torch.empty_like(src)
cap = int(n * 1.25) + (1 << 16)
ops = multi.CudaOps(cap, r=8)
recv = torch.empty(cap, dtype=torch.int32, device=dev)
staging = torch.empty(n, dtype=torch.int32, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


# raw all_to_all at the same size (equal splits)
per = n // world
a2a_ms = []
for _ in range(args.reps + 1):
    dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    dist.all_to_all_single(recv[: per * world], staging[: per * world])
    e1.record()
    torch.cuda.synchronize()
    a2a_ms.append(e0.elapsed_time(e1))
a2a = min(a2a_ms[1:])
sent = 4 * per * (world - 1)

peer = multi.PeerExchange(recv) if args.peer else None
if peer is not None:
    # parity of the fused path against the NCCL path on the same keys
    work.copy_(src)
    ref_out, _ = multi.distributed_sort(work, ops, recv, staging)
    ref_copy = ref_out.clone()
    work.copy_(src)
    out, _ = multi.distributed_sort(work, ops, recv, staging, peer=peer)
    same = torch.equal(out, ref_copy)
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert bool(flag.item()), "peer-scatter exchange differs from the all_to_all exchange"
    del ref_copy
stage_best = None
for _ in range(args.reps + 1):
    work.copy_(src)
    dist.barrier()
    torch.cuda.synchronize()
    out, stats = multi.distributed_sort(work, ops, recv, staging, timing=True, peer=peer)
    torch.cuda.synchronize()
    t = stats.stage_ms()
    if stage_best is None or sum(t.values()) < sum(stage_best.values()):
        stage_best = t
if rank == 0:
    print(json.dumps({"tag": args.tag, "peer_scatter": bool(args.peer), "world": world, "keys_total_log2": args.log2n_total,
                      "raw_all_to_all_ms": round(a2a, 3), "raw_all_to_all_GBs_out_per_gpu": round(sent / a2a / 1e6, 1),
                      "stages_ms": {k: round(v, 3) for k, v in stage_best.items()},
                      "total_ms": round(sum(stage_best.values()), 3),
                      "gkeys_s": round((1 << args.log2n_total) / sum(stage_best.values()) / 1e6, 2)}))
dist.destroy_process_group()
