"""lsd_sort64 (64-bit keys) timed with CUDA events against torch.sort on the same keys (library baseline, same GPU)."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lsdradixsort_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, nargs="+", default=[26, 28])
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
for lg in args.log2n:
    n = 1 << lg
    g = torch.Generator(device="cuda").manual_seed(lg)
    src = torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    s = L.Sorter64(n)
    work = torch.empty_like(src)
    best = 1e9
    for rep in range(args.reps + 2):
        work.copy_(src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.sort_(work)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            best = min(best, e0.elapsed_time(e1))
    ok = bool(torch.equal(work, torch.sort(src).values))
    del s
    tb = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = torch.sort(src).values
        e1.record()
        torch.cuda.synchronize()
        tb = min(tb, e0.elapsed_time(e1))
        del out
    print(json.dumps({"keys": n, "dtype": "int64", "lsd_sort64_ms": round(best, 4), "gkeys_s": round(n / best / 1e6, 2),
                      "algorithmic_GBs": round(8 * 16 * n / best / 1e6, 1), "bit_exact_vs_torch_sort": ok,
                      "torch_sort_ms": round(tb, 4), "torch_gkeys_s": round(n / tb / 1e6, 2)}), flush=True)
