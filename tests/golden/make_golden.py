"""Generate tests/golden/ref_cpu_vectors.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The reference ships no golden vectors, so these fixtures are outputs of the
reference's own CPU functions -- LSDRadixSort (.cu:62), LSDRadixSortPass (.cu:25), PrefixSum (.cu:128),
BuildHistogramsCPU (.cu:643), GetGPUPrefixSumBlockSumsCount (.cu:265) -- compiled by oracle/Makefile
into oracle/_ref/libref_lsd.so and called through oracle/ref_shim.cpp.  Inputs come from
lsdradixsort_b200.keygen (portable, seeded).  tests/test_oracle.py pins oracle/lsd_oracle.c to them.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle  # noqa: E402
from lsdradixsort_b200 import keygen  # noqa: E402


def main() -> None:
    ref = _oracle.ref()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    cases = []
    # full sorts: every r the reference's CPU path accepts on sizes that stay tiny in git
    for kind, n, seed in [("uniform", 2048, 0), ("uniform", 1000, 1), ("entropy4_table", 1536, 2),
                          ("low_nibble", 777, 3), ("reverse", 1024, 0), ("all_equal", 300, 0), ("uniform", 1, 5)]:
        keys = keygen.make_keys(kind, n, seed)
        for r in (1, 2, 4, 8, 16):
            a = keys.copy()
            b = np.zeros_like(a)
            h = np.zeros(1 << r, dtype=np.uint32)
            ref.ref_cpu_sort(a, b, n, h, r)
            assert np.array_equal(a, b) and np.array_equal(b, np.sort(keys))
            tag = f"sort_{kind}_{n}_{seed}_r{r}"
            out[tag + "_out"] = b
            out[tag + "_hist"] = h  # scratch histogram as the reference leaves it (bucket starts of last pass)
            cases.append(tag)
        # one single pass on the middle digit, to pin pass-level stability
        for r, g in ((8, 1), (4, 5), (1, 17)):
            a = keys.copy()
            b = np.zeros_like(a)
            h = np.zeros(1 << r, dtype=np.uint32)
            ref.ref_cpu_sort_pass(a, b, n, h, r, g)
            out[f"pass_{kind}_{n}_{seed}_r{r}_g{g}_out"] = b
    # prefix sums (wrap-around)
    for n, seed in [(1, 0), (2, 1), (1000, 2), (4096, 3)]:
        a = keygen.uniform_u32(n, seed)
        b = a.copy()
        ref.ref_cpu_prefix_sum(b, n)
        out[f"scan_{n}_{seed}_out"] = b
    # per-tile histograms (count must be a multiple of block for the reference's CPU twin)
    for n, block, r, g, seed in [(2048, 128, 8, 0, 0), (2048, 256, 8, 3, 1), (1024, 32, 1, 31, 2),
                                 (4096, 512, 4, 7, 3), (1024, 1024, 2, 9, 4)]:
        a = keygen.uniform_u32(n, seed)
        grid = n // block
        h = np.zeros(grid * (1 << r), dtype=np.uint32)
        ref.ref_cpu_build_histograms(a, h, n, r, g, grid, block)
        out[f"hist_{n}_{block}_{r}_{g}_{seed}_out"] = h
    # scratch sizing
    sizing = []
    for count in (1, 31, 32, 33, 1024, 1 << 20, (1 << 20) + 7, 1 << 28):
        for tpb in (32, 128, 256, 1024):
            sizing.append((count, tpb, ref.ref_block_sums_count(count, tpb)))
    out["block_sums_count"] = np.array(sizing, dtype=np.int64)
    out["sort_cases"] = np.array(cases)
    dst = Path(__file__).resolve().parent / "ref_cpu_vectors.npz"
    np.savez_compressed(dst, **out)
    print(f"wrote {dst} ({dst.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()
