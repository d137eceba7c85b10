"""Per-phase SM-clock timeline of the last pass (kernels that honour lsd_sort_options.debug_trace)."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--variant", type=int, default=19)
ap.add_argument("--tile", type=int, default=8352)
args = ap.parse_args()
n = 1 << args.log2n
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
tiles = (n + args.tile - 1) // args.tile
trace = torch.zeros(tiles * 16, dtype=torch.int64, device="cuda")
s = L.Sorter(n, r=8, variant=args.variant, debug_trace=trace.data_ptr())
for _ in range(2):
    work = src.clone()
    trace.zero_()
    st = s.sort_timed_(work)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(tiles, 16).astype(np.float64)
t = t[200:-200]  # steady state
names = {0: "ticket+clear", 1: "tile landed", 2: "w0 counted", 3: "count barrier", 4: "scan1+digit scan", 5: "scan2 done",
         6: "w0 ranked", 7: "w0 scattered", 8: "lookback start", 9: "lookback done", 10: "chain complete",
         11: "final barrier", 12: "w0 stores issued"}
print("pass ms", st)
for k in sorted(names):
    col = t[:, k]
    col = col[col > 0]
    print(f"{k:2d} {names[k]:18s} mean {col.mean():9.0f}  p10 {np.percentile(col,10):9.0f}  p50 {np.percentile(col,50):9.0f}  p90 {np.percentile(col,90):9.0f}  p99 {np.percentile(col,99):9.0f}")
