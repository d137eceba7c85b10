"""lsd_bench (tools/lsd_bench.cpp): the reference's Test*/Benchmark* drivers (LSDRadixSort.cu:1029-1185) over the C ABI.
The stdout blocks must carry the same field labels as the reference's BenchmarkLSDRadixSort.md / BenchmarkPrefixSum.md /
BenchmarkBuildHistogram.md so runs can be diffed field by field."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "build" / "lsd_bench"


def _build():
    res = subprocess.run(["make", "-C", str(ROOT), "tools"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert BIN.exists()


def test_harness_builds_and_rejects_bad_usage():
    _build()
    assert subprocess.run([str(BIN)], capture_output=True).returncode == 64
    assert subprocess.run([str(BIN), "nonsense"], capture_output=True).returncode == 64


def _run(*args):
    if not BIN.exists():
        _build()
    res = subprocess.run([str(BIN), *args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "CHECK FAILED" not in res.stdout
    return res.stdout


@pytest.mark.gpu
def test_harness_sort_blocks_match_reference_format():
    out = _run("sort", "--elems", "1Mi,100003", "--blocks", "128,256,1024", "--rs", "1,2,4,8")
    blocks = out.split("-- Test GPU LSD Radix Sort --\n")[1:]
    assert len(blocks) == 2 * 3 * 4
    pat = re.compile(r"Elements: \S+ GB\nHistograms: \S+ GB\nBlock Sums: \S+ GB\nBlock Size: \d+\nR: \d\n"
                     r"CPU \S+ ms\nGPU \S+ ms\nSpeedup: x\S+\n$")
    for b in blocks:
        assert pat.match(b), b


@pytest.mark.gpu
def test_harness_prefix_sum_histogram_and_pairs():
    out = _run("prefix_sum", "--elems", "1Mi,77777", "--blocks", "32,256,1024")
    assert out.count("-- Test exclusive prefix sum --") == 6 and out.count("GPU Prefix Sum: ") == 6
    out = _run("build_histogram", "--elems", "1Mi", "--blocks", "32,512", "--rs", "1,8")
    assert out.count("-- Test Build Histogram --") == 4 and out.count("Bit Group: ") == 4
    out = _run("pairs", "--elems", "300001", "--blocks", "0,256", "--rs", "4,8")
    assert out.count("(key-value)") == 4


@pytest.mark.gpu
def test_harness_wide_digits_and_64bit_keys():
    """r = 16 (the reference CPU path's other digit width, .cu:56-69) and 11 through the same driver, and the 64-bit mode;
    every block self-checks against std::sort (a failed check makes _run fail)."""
    out = _run("sort", "--elems", "1Mi,100003", "--blocks", "0", "--rs", "11,16")
    assert out.count("-- Test GPU LSD Radix Sort --") == 4 and "R: 16" in out and "R: 11" in out
    out = _run("sort64", "--elems", "1Mi,100003,7")
    assert out.count("(64-bit keys)") == 3
