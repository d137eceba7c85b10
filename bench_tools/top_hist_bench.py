"""lsd_top_digit_histogram / lsd_digit_histograms timed at 2^28 .. 2^31 keys (the planning step of lsd_sort_multi)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lsdradixsort_b200 as L  # noqa: E402

for lg in (28, 29, 30, 31):
    n = 1 << lg
    keys = torch.empty(n, dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(lg)
    for lo in range(0, n, 1 << 26):
        keys[lo:lo + (1 << 26)] = torch.randint(-(2**31), 2**31, (1 << 26,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    out = {}
    for name, fn in (("top", L.top_digit_histogram), ("all", L.digit_histograms)):
        best = 1e9
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h = fn(keys, 8)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 1:
                best = min(best, e0.elapsed_time(e1))
        out[name + "_ms"] = round(best, 4)
        out[name + "_GBs"] = round(4 * n / best / 1e6, 1)
    print(json.dumps({"log2n": lg, **out}), flush=True)
    del keys
