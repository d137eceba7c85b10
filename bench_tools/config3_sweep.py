"""BASELINE config 3: standalone build_histogram and exclusive prefix_sum, 2^20..2^30 uint32, at the reference's R1/R8 x
B128/B256/B512 settings (plus the one-read digit histogram).  JSON lines; GB/s = algorithmic bytes / CUDA-event time
(prefix_sum 8 B/element; build_histogram 4 B/key + G*2^r*4 B output).  Sizes <= 2^24 sit in the 126 MB L2: their GB/s is
not an HBM number."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402

ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    ts.sort()
    return ts[len(ts) // 2]


for log2n in (20, 22, 24, 26, 28, 30):
    n = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(0)
    src = torch.empty(n, dtype=torch.int32, device="cuda")
    for lo in range(0, n, 1 << 26):
        hi = min(n, lo + (1 << 26))
        src[lo:hi] = torch.randint(-(2**31), 2**31, (hi - lo,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    work = src.clone()
    for block in (128, 256, 512):
        ws = torch.empty(max(L.GetGPUPrefixSumBlockSumsCount(n, block), 64), dtype=torch.int32, device="cuda")
        t = timeit(lambda: L.GPUPrefixSum(work, n, block, ws))
        print(json.dumps({"kernel": "prefix_sum", "log2n": log2n, "block": block, "ms": round(t, 4),
                          "gbs": round(8 * n / (t * 1e6), 1), "in_l2": log2n <= 24}), flush=True)
    for r in (1, 8):
        for block in (128, 256, 512):
            out = torch.empty(((n + block - 1) // block, 1 << r), dtype=torch.int32, device="cuda")
            t = timeit(lambda: L.build_histogram(src, r, 0, block, out=out), reps=5)
            byts = 4 * n + out.numel() * 4
            print(json.dumps({"kernel": "build_histogram", "log2n": log2n, "r": r, "block": block, "ms": round(t, 4),
                              "gbs": round(byts / (t * 1e6), 1), "out_bytes": out.numel() * 4, "in_l2": log2n <= 24}), flush=True)
            del out
    t = timeit(lambda: L.digit_histograms(src, 8))
    print(json.dumps({"kernel": "digit_histograms(all 4 digits, one read)", "log2n": log2n, "r": 8, "ms": round(t, 4),
                      "gbs": round(4 * n / (t * 1e6), 1), "in_l2": log2n <= 24}), flush=True)
    del src, work
