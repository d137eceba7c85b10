"""Small end-to-end run of every entry point for compute-sanitizer (memcheck / racecheck / synccheck), one tool per call."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402
from lsdradixsort_b200 import keygen  # noqa: E402

for n in (50_001, 200_000):
    for kind in ("uniform", "entropy4_table"):
        keys = keygen.make_keys(kind, n, seed=n)
        d = torch.from_numpy(keys.view(np.int32)).cuda()
        for variant in (0, 52):
            w = d.clone()
            L.sort_(w, r=8, variant=variant)
            assert np.array_equal(w.cpu().numpy().view(np.uint32), np.sort(keys)), (n, kind, variant)
        for r in (1, 4):
            w = d.clone()
            L.sort_(w, r=r)
            assert np.array_equal(w.cpu().numpy().view(np.uint32), np.sort(keys))
        kv, vv = d.clone(), torch.arange(n, dtype=torch.int32, device="cuda")
        L.sort_pairs_(kv, vv, r=8)                      # key-value pass (values staged in the dead matrix)
        assert np.array_equal(vv.cpu().numpy().view(np.uint32), np.argsort(keys, kind="stable").astype(np.uint32))
        f = torch.from_numpy(keys.view(np.float32)).cuda()
        L.sort_(f)                                      # typed flavour of the persistent kernel (f32 total order)
        L.sort_(d.clone(), r=8, key_type="i32")
        s = d.clone()
        L.prefix_sum_(s, 256)
        L.digit_histograms(d, 8)
        L.top_digit_histogram(d, 8)
        L.build_histogram(d, 8, 1, 256)
        out = torch.empty_like(d)
        L.sort_pass(d, out, 8, 3)
        starts = np.concatenate([[0], np.cumsum(np.bincount(keys >> 24, minlength=256))[:-1]]).astype(np.int64)
        o2 = torch.empty(n + 8, dtype=torch.int32, device="cuda")
        L.sort_pass_scatter(d, torch.from_numpy(o2.data_ptr() + 4 * starts).cuda(), 8, 3)
        torch.cuda.synchronize()
        assert torch.equal(o2[:n], out)
print("sanitize target ok")
