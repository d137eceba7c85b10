"""Phase clocks of scan_tma_kernel (LSD_SCAN_TRACE=1): ticket, landed, partials, look-back done, stores issued."""
import os
import sys
from pathlib import Path

os.environ["LSD_SCAN_TRACE"] = "1"
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lsdradixsort_b200 as L  # noqa: E402

n = 1 << 28
for block in (128, 256, 512):
    a = torch.randint(0, 1 << 20, (n,), dtype=torch.int32, device="cuda")
    words = L.GetGPUPrefixSumBlockSumsCount(n, block)
    ws = torch.zeros(words + 64, dtype=torch.int32, device="cuda")
    for _ in range(2):
        L.GPUPrefixSum(a, n, block, ws)
    torch.cuda.synchronize()
    tile = block * 32
    tiles = n // tile
    raw = ws.cpu().numpy().view(np.uint32)
    tr = raw[64 + 2 * tiles: 64 + 2 * tiles + 8 * tiles].reshape(tiles, 8).astype(np.float64)[500:-500]
    names = ["ticket returned", "tile landed", "partials done", "look-back done", "stores issued"]
    print(f"block {block}: tile {tile} elements, {tiles} tiles")
    for i, nm in enumerate(names):
        c = tr[:, i]
        print(f"  {nm:16s} mean {c.mean():8.0f} p10 {np.percentile(c,10):8.0f} p50 {np.percentile(c,50):8.0f} p90 {np.percentile(c,90):8.0f}")
