"""Tuning sweep (GPU box): per-stage device times of the sort for every kernel variant, plus the
standalone histogram / prefix-sum kernels.  Writes JSON lines to stdout.  Not part of the product."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402


def rand_keys(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--variants", type=str, default="0,1,2,3,4,5,6,7")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--r", type=int, default=8)
    args = ap.parse_args()
    n = 1 << args.log2n
    src = rand_keys(n)
    work = torch.empty_like(src)
    peak = 6551.0
    for v in [int(x) for x in args.variants.split(",")]:
        try:
            s = L.Sorter(n, r=args.r, variant=v)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"variant": v, "error": str(e)}))
            continue
        best = None
        for _ in range(args.reps):
            work.copy_(src)
            st = s.sort_timed_(work)
            if best is None or sum(st) < sum(best):
                best = st
        total = sum(best)
        print(json.dumps({
            "variant": v, "r": args.r, "log2n": args.log2n, "stage_ms": [round(x, 4) for x in best],
            "total_ms": round(total, 4), "gkeys_s": round(n / total / 1e6, 2),
            "pass_gbs": [round(8 * n / (x * 1e6), 1) if x > 0 else 0 for x in best[1:-1]],
            "frac_of_measured": round(32 * n / (total * 1e6) / peak, 4)}), flush=True)
        del s
    # standalone kernels
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        return min(ts)

    for r in (8, 4, 1):
        t = timeit(lambda: L.digit_histograms(src, r))
        print(json.dumps({"kernel": "digit_histograms", "r": r, "ms": round(t, 4), "gbs": round(4 * n / (t * 1e6), 1)}), flush=True)
    for block in (128, 256, 512):
        ws_words = L.GetGPUPrefixSumBlockSumsCount(n, block)
        ws = torch.empty(max(ws_words, 64), dtype=torch.int32, device="cuda")
        t = timeit(lambda: L.GPUPrefixSum(work, n, block, ws))
        print(json.dumps({"kernel": "prefix_sum", "block": block, "ms": round(t, 4), "gbs": round(8 * n / (t * 1e6), 1)}), flush=True)
    for r, block in ((1, 128), (1, 512), (8, 256), (8, 512)):
        out = torch.empty(((n + block - 1) // block, 1 << r), dtype=torch.int32, device="cuda")
        t = timeit(lambda: L.build_histogram(src, r, 0, block, out=out))
        byts = 4 * n + out.numel() * 4
        print(json.dumps({"kernel": "build_histogram", "r": r, "block": block, "ms": round(t, 4), "gbs": round(byts / (t * 1e6), 1)}), flush=True)
    t = timeit(lambda: work.copy_(src))
    print(json.dumps({"kernel": "torch_copy", "ms": round(t, 4), "gbs": round(8 * n / (t * 1e6), 1)}), flush=True)


if __name__ == "__main__":
    main()
