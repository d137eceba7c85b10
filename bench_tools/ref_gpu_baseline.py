"""The reference's own CUDA kernels (oracle/_ref/libref_lsd.so: LSDRadixSort.cu rebuilt for sm_100a, unmodified)
timed on this B200 beside ours: GPULSDRadixSort (.cu:839), GPUPrefixSum (.cu:286), BuildHistogramsKernel (.cu:660).
A reported baseline, not a target.  The reference needs count % block == 0 and G*2^r < 2^31."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import _oracle  # noqa: E402
import lsdradixsort_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=26)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
ref = _oracle.ref()
assert ref is not None, "oracle/_ref/libref_lsd.so missing (built in the container by `make -C oracle ref`)"
n = 1 << args.log2n
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, prep, reps=args.reps):
    ts = []
    for _ in range(reps + 1):
        prep()
        torch.cuda.synchronize()
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    return min(ts[1:])


expected = None
for r, block in ((8, 1024), (8, 256), (4, 512)):
    grid = n // block
    h_words = 3 * grid * (1 << r)
    if grid * (1 << r) >= 2**31:
        print(json.dumps({"kernel": "ref GPULSDRadixSort", "r": r, "block": block, "skipped": "G*2^r >= 2^31"}))
        continue
    a = torch.empty_like(src)
    b = torch.empty_like(src)
    h = torch.empty(h_words, dtype=torch.int32, device="cuda")
    bs = torch.empty(ref.ref_block_sums_count(grid * (1 << r), block) + 64, dtype=torch.int32, device="cuda")

    def run():
        rc = ref.ref_gpu_sort(a.data_ptr(), b.data_ptr(), h.data_ptr(), bs.data_ptr(), n, block, r)
        assert rc == 0, rc

    t = timeit(run, lambda: a.copy_(src))
    got = a.cpu().numpy().view(np.uint32)
    if expected is None:
        w = src.clone()
        L.sort_(w, r=8)
        expected = w.cpu().numpy().view(np.uint32)
    print(json.dumps({"kernel": "ref GPULSDRadixSort", "log2n": args.log2n, "r": r, "block": block, "ms": round(t, 3),
                      "gkeys_s": round(n / t / 1e6, 3), "equals_ours": bool(np.array_equal(got, expected))}), flush=True)
    del a, b, h, bs

s = L.Sorter(n, r=8)
w = torch.empty_like(src)
t = timeit(lambda: s.sort_(w), lambda: w.copy_(src))
print(json.dumps({"kernel": "ours lsd_sort", "log2n": args.log2n, "r": 8, "ms": round(t, 3), "gkeys_s": round(n / t / 1e6, 2)}), flush=True)

for block in (128, 256, 512):
    a = torch.empty_like(src)
    bs = torch.empty(ref.ref_block_sums_count(n, block) + 64, dtype=torch.int32, device="cuda")
    t = timeit(lambda: ref.ref_gpu_prefix_sum(a.data_ptr(), n, block, bs.data_ptr()), lambda: a.copy_(src))
    ws = torch.empty(max(L.GetGPUPrefixSumBlockSumsCount(n, block), 64), dtype=torch.int32, device="cuda")
    t2 = timeit(lambda: L.GPUPrefixSum(a, n, block, ws), lambda: a.copy_(src))
    print(json.dumps({"kernel": "prefix_sum", "log2n": args.log2n, "block": block, "ref_ms": round(t, 3), "ours_ms": round(t2, 3),
                      "ref_gbs": round(8 * n / (t * 1e6), 1), "ours_gbs": round(8 * n / (t2 * 1e6), 1)}), flush=True)

for r, block in ((1, 128), (8, 256), (8, 512)):
    grid = n // block
    if grid * (1 << r) >= 2**31:
        continue
    h = torch.empty(grid * (1 << r), dtype=torch.int32, device="cuda")
    t = timeit(lambda: ref.ref_gpu_build_histograms(src.data_ptr(), h.data_ptr(), n, r, 0, block), lambda: None)
    h2 = torch.empty((grid, 1 << r), dtype=torch.int32, device="cuda")
    t2 = timeit(lambda: L.build_histogram(src, r, 0, block, out=h2), lambda: None)
    same = bool(torch.equal(h.view(grid, 1 << r), h2))
    print(json.dumps({"kernel": "build_histogram", "log2n": args.log2n, "r": r, "block": block, "ref_ms": round(t, 3),
                      "ours_ms": round(t2, 3), "equal": same}), flush=True)
