"""Run a few sorts of 2^log2n uniform keys with one kernel variant (profiling target for ncu)."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--r", type=int, default=8)
ap.add_argument("--scan", action="store_true", help="also run prefix_sum once")
ap.add_argument("--pairs", action="store_true", help="also run one key-value sort (lsd_sort_pairs)")
args = ap.parse_args()
n = 1 << args.log2n
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
work = torch.empty_like(src)
s = L.Sorter(n, r=args.r, variant=args.variant)
for _ in range(args.reps):
    work.copy_(src)
    s.sort_(work)
if args.pairs:
    pk, pv = src.clone(), torch.arange(n, dtype=torch.int32, device="cuda")
    L.PairSorter(n, r=args.r, variant=args.variant).sort_(pk, pv)
    torch.cuda.synchronize()
    assert bool((src[pv.long()] == pk).all()), "pairs: values do not follow their keys"
    del pk, pv
if args.scan:
    L.prefix_sum_(work, 256)
torch.cuda.synchronize()
u = work.to(torch.int64) & 0xFFFFFFFF
print("sorted" if bool((u[1:] >= u[:-1]).all()) or args.scan else "NOT SORTED")
