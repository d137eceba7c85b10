"""Key-value sort timing (GPU box): per-stage device times of lsd_sort_pairs next to lsd_sort for the same keys.
Writes JSON lines to stdout.  Algorithmic bytes: 16 B per pair per pass (key + value, read + write)."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--r", type=int, default=8)
ap.add_argument("--variants", type=str, default="0")
args = ap.parse_args()
n = 1 << args.log2n
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
keys, vals = torch.empty_like(src), torch.empty_like(src)
peak = 6551.0
for v in [int(x) for x in args.variants.split(",")]:
    ps = L.PairSorter(n, r=args.r, variant=v)
    best = None
    for _ in range(args.reps):
        keys.copy_(src)
        vals.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
        st = ps.sort_timed_(keys, vals)
        if best is None or sum(st) < sum(best):
            best = st
    u = keys.to(torch.int64) & 0xFFFFFFFF
    ok = bool((u[1:] >= u[:-1]).all()) and bool((src[vals.long()] == keys).all())
    total = sum(best)
    passes = 32 // args.r
    print(json.dumps({"what": "lsd_sort_pairs", "variant": v, "r": args.r, "log2n": args.log2n,
                      "stage_ms": [round(x, 4) for x in best], "total_ms": round(total, 4),
                      "gpairs_s": round(n / total / 1e6, 2),
                      "pass_gbs": [round(16 * n / (x * 1e6), 1) if x > 0 else 0 for x in best[1:-1]],
                      "frac_of_measured": round(16 * passes * n / (total * 1e6) / peak, 4), "verified": ok}), flush=True)
    del ps
    s = L.Sorter(n, r=args.r, variant=v)
    best = None
    for _ in range(args.reps):
        keys.copy_(src)
        st = s.sort_timed_(keys)
        if best is None or sum(st) < sum(best):
            best = st
    print(json.dumps({"what": "lsd_sort", "variant": v, "stage_ms": [round(x, 4) for x in best],
                      "total_ms": round(sum(best), 4), "gkeys_s": round(n / sum(best) / 1e6, 2)}), flush=True)
    del s
