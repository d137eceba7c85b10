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
ap.add_argument("--r", type=int, default=8)
ap.add_argument("--family", type=str, default="lpc", choices=["lpc", "cpc", "cpcp", "wide"])
args = ap.parse_args()
n = 1 << args.log2n
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
tiles = (n + args.tile - 1) // args.tile
trace = torch.zeros(tiles * 16, dtype=torch.int64, device="cuda")
s = L.Sorter(n, r=args.r, variant=args.variant, debug_trace=trace.data_ptr())
for _ in range(2):
    work = src.clone()
    trace.zero_()
    st = s.sort_timed_(work)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(tiles, 16).astype(np.float64)
t = t[200:-200]  # steady state
names = {0: "ticket+clear", 1: "tile landed", 2: "w0 counted", 3: "count barrier", 4: "scan1+digit scan", 5: "scan2 done",
         6: "w0 ranked", 7: "w0 scattered", 8: "lookback start", 9: "lookback done", 10: "chain complete", 13: "lookback rounds", 14: "lookback hops",
         11: "final barrier", 12: "w0 stores issued", 15: "lookback load wait (sum)"}
if args.family == "wide":
    names = {1: "tile landed", 2: "w0 counted", 3: "count barrier", 4: "scan1+digit scan", 5: "scan2 done", 6: "w0 ranked+scattered",
             7: "barrier C passed", 8: "lookback start", 9: "lookback done", 10: "chain tail (last warp)", 11: "chain tail (last-1)",
             12: "w0 stores issued", 13: "lookback rounds", 14: "lookback hops", 15: "lookback load wait (sum)", 0: "lookback processing (sum)"}
if args.family == "cpc":
    names = {0: "ticket+clear", 1: "tile landed", 2: "w0 ranked", 3: "rank barrier", 4: "scan+Q done", 5: "positions done",
             6: "w0 scattered", 7: "lookback done", 8: "final barrier", 9: "w0 stores issued", 10: "ticket returned", 11: "w2 clear done", 13: "lookback rounds",
             14: "lookback hops"}
print("pass ms", st)
print("tile period per SM (cycles):", st[1] * 1e-3 * 1.965e9 * 148 / tiles)
for k in sorted(names):
    col = t[:, k]
    col = col[col > 0]
    if args.family == "cpcp" or len(col) == 0:
        continue
    print(f"{k:2d} {names[k]:18s} mean {col.mean():9.0f}  p10 {np.percentile(col,10):9.0f}  p50 {np.percentile(col,50):9.0f}  p90 {np.percentile(col,90):9.0f}  p99 {np.percentile(col,99):9.0f}")

if args.family == "cpc":
    # are tile starts phase-locked?  histogram of tile start times (ns, 250 ns bins) over the middle of the pass
    gt = trace.cpu().numpy().reshape(tiles, 16)[:, 15].astype(np.int64)
    gt = gt[gt > 0]
    gt = gt - gt.min()
    span = gt.max()
    lo, hi = int(span * 0.4), int(span * 0.4) + 20000
    sel = gt[(gt >= lo) & (gt < hi)]
    hist, _ = np.histogram(sel, bins=80, range=(lo, hi))
    print("pass span ns", span, "tile starts per 250 ns bin over 20 us in mid-pass (uniform would be", round(len(sel) / 80, 1), "):")
    print(" ".join(str(x) for x in hist))

if args.family == "cpcp":
    tt = trace.cpu().numpy().reshape(tiles, 16).astype(np.float64)[200:-200]
    ok = (tt[:, 0] > 0) & (tt[:, 8] > 0)
    tt = tt[ok]
    def d(a, b, label):
        x = tt[:, b] - tt[:, a]
        print(f"{label:34s} mean {x.mean():8.0f}  p10 {np.percentile(x,10):8.0f}  p50 {np.percentile(x,50):8.0f}  p90 {np.percentile(x,90):8.0f}")
    print("persistent pipeline, SM clocks (one CTA per SM: all roles share the clock)")
    d(0, 1, "load issue -> front starts")
    d(1, 2, "keys->regs + rank")
    d(2, 3, "scan + Q")
    d(3, 4, "positions")
    d(4, 5, "scatter")
    d(1, 5, "front total")
    d(3, 6, "hist ready -> look-back done")
    d(6, 7, "look-back done -> sorted seen")
    d(7, 8, "copy-out")
    d(0, 8, "buffer lifetime")
    print("lookback rounds", tt[:, 13].mean(), "hops", tt[:, 14].mean())
    print("tile period per SM (cycles):", st[1] * 1e-3 * 1.965e9 * 148 / tiles)
