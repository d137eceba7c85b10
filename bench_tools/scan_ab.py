"""lsd_prefix_sum: L2 prefetch distance sweep (tuning aid): bit-exact check against torch.cumsum and timing."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lsdradixsort_b200 as L  # noqa: E402
from lsdradixsort_b200 import _native as N  # noqa: E402

lib = N.lib()
toggle = getattr(lib, "lsd_debug_scan_prefetch", None)
modes = tuple(m << 20 for m in (0, 6, 8, 12, 16, 20, 24)) if toggle else (None,)
for block in (256, 128, 512):
    for lg in (24, 26, 28, 30):
        if block != 256 and lg not in (28,):
            continue
        n = (1 << lg) + (17 if lg <= 24 else 0)
        g = torch.Generator(device="cuda").manual_seed(lg)
        src = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        want = None
        if lg <= 28:
            u = src.to(torch.int64) & 0xFFFFFFFF
            want = (torch.cumsum(u, 0) - u) & 0xFFFFFFFF
            want = (want - ((want >> 31) << 32)).to(torch.int32)
            del u
        words = L.GetGPUPrefixSumBlockSumsCount(n, block)
        ws = torch.empty(max(words, 64), dtype=torch.int32, device="cuda")
        work = torch.empty_like(src)
        for mode in modes:
            if toggle:
                toggle(mode)
            best = 1e9
            ok = None
            for rep in range(8):
                work.copy_(src)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.GPUPrefixSum(work, n, block, ws)
                e1.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    best = min(best, e0.elapsed_time(e1))
            if want is not None:
                ok = bool(torch.equal(work, want))
            print(json.dumps({"n": n, "block": block, "prefetch_MiB": None if mode is None else mode >> 20, "ms": round(best, 4), "GBs": round(8 * n / best / 1e6, 1),
                              "frac_of_6551": round(8 * n / best / 1e6 / 6551, 3), "bit_exact_vs_cumsum": ok}), flush=True)
        del src, work, want
