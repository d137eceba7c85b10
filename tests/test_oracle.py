"""CPU tests: pin oracle/lsd_oracle.c to the reference.

Fixtures in tests/golden/ref_cpu_vectors.npz are outputs of the unmodified reference's CPU functions
(see tests/golden/make_golden.py).  When oracle/_ref/libref_lsd.so is present (build container), the
oracle is also compared live with the reference on larger inputs.  Mirrors the reference's own checks:
CheckArrays at LSDRadixSort.cu:120 (CPU sort vs std::sort), :364 (scan), :785 (histograms).
"""
from pathlib import Path

import numpy as np
import pytest

import _oracle
from lsdradixsort_b200 import keygen

GOLD = np.load(Path(__file__).parent / "golden" / "ref_cpu_vectors.npz")


def _parse_sort_case(tag):
    # sort_<kind>_<n>_<seed>_r<r>
    body, r = tag[len("sort_"):].rsplit("_r", 1)
    kind, n, seed = body.rsplit("_", 2)
    return kind, int(n), int(seed), int(r)


@pytest.mark.parametrize("tag", [str(t) for t in GOLD["sort_cases"]])
def test_oracle_sort_matches_reference_golden(tag):
    kind, n, seed, r = _parse_sort_case(tag)
    keys = keygen.make_keys(kind, n, seed)
    a = keys.copy()
    out = np.zeros_like(a)
    hist = np.zeros(1 << r, dtype=np.uint32)
    assert _oracle.oracle().lsd_oracle_sort(a, out, n, hist, r) == 0
    assert np.array_equal(out, GOLD[tag + "_out"])
    assert np.array_equal(a, out)  # reference leaves the result in both arrays (.cu:53)
    assert np.array_equal(hist, GOLD[tag + "_hist"])  # even the scratch histogram ends identical
    assert np.array_equal(out, np.sort(keys))  # the std::sort leg (.cu:97,120)


def test_oracle_single_pass_matches_reference_golden():
    seen = 0
    for name in GOLD.files:
        if not name.startswith("pass_"):
            continue
        body = name[len("pass_"):-len("_out")]
        head, g = body.rsplit("_g", 1)
        head, r = head.rsplit("_r", 1)
        kind, n, seed = head.rsplit("_", 2)
        n, seed, r, g = int(n), int(seed), int(r), int(g)
        keys = keygen.make_keys(kind, n, seed)
        a, out, hist = keys.copy(), np.zeros_like(keys), np.zeros(1 << r, dtype=np.uint32)
        _oracle.oracle().lsd_oracle_sort_pass(a, out, n, hist, r, g)
        assert np.array_equal(out, GOLD[name]), name
        # a stable pass == numpy's stable argsort on the digit
        digit = (keys >> np.uint32(g * r)) & np.uint32((1 << r) - 1)
        assert np.array_equal(out, keys[np.argsort(digit, kind="stable")]), name
        # the tile-decomposed flow of GPULSDRadixSort (.cu:839-910) is the same permutation
        tiled = np.zeros_like(keys)
        assert _oracle.oracle().lsd_oracle_tiled_pass(keys, tiled, n, r, g, 128) == 0
        assert np.array_equal(tiled, out), name
        seen += 1
    assert seen >= 10


def test_oracle_prefix_sum_matches_reference_golden():
    for n, seed in [(1, 0), (2, 1), (1000, 2), (4096, 3)]:
        a = keygen.uniform_u32(n, seed)
        got = _oracle.prefix_sum(a)
        assert np.array_equal(got, GOLD[f"scan_{n}_{seed}_out"])
        want = np.concatenate([[0], np.cumsum(a[:-1].astype(np.uint64))]).astype(np.uint64) & 0xFFFFFFFF
        assert np.array_equal(got.astype(np.uint64), want)  # exclusive, mod 2^32
    assert _oracle.prefix_sum(np.zeros(0, dtype=np.uint32)).size == 0


def test_oracle_build_histograms_matches_reference_golden():
    for n, block, r, g, seed in [(2048, 128, 8, 0, 0), (2048, 256, 8, 3, 1), (1024, 32, 1, 31, 2),
                                 (4096, 512, 4, 7, 3), (1024, 1024, 2, 9, 4)]:
        a = keygen.uniform_u32(n, seed)
        got = _oracle.build_histograms(a, r, g, block)
        assert np.array_equal(got.ravel(), GOLD[f"hist_{n}_{block}_{r}_{g}_{seed}_out"])
        assert got.sum() == n and np.all(got.sum(axis=1) == block)


def test_oracle_build_histograms_ragged_tail_counts_only_valid_keys():
    a = keygen.uniform_u32(1000, 9)
    got = _oracle.build_histograms(a, 8, 2, 256)  # 4 tiles, last one holds 232 keys (kernel guard, .cu:684)
    assert got.shape == (4, 256) and got.sum() == 1000 and got[3].sum() == 232


def test_oracle_block_sums_count_matches_reference_golden():
    for count, tpb, want in GOLD["block_sums_count"]:
        assert _oracle.oracle().lsd_oracle_block_sums_count(int(count), int(tpb)) == want


def test_oracle_digit_histograms_are_column_sums_of_tile_histograms():
    a = keygen.make_keys("entropy4_table", 5000, 4)
    for r in (1, 2, 4, 8):
        dh = _oracle.digit_histograms(a, r)
        for g in range(32 // r):
            assert np.array_equal(dh[g], _oracle.build_histograms(a, r, g, 500).sum(axis=0).astype(np.uint64))


def test_oracle_rejects_bad_radix():
    a = np.zeros(4, dtype=np.uint32)
    for r in (0, 3, 5, 32, 64):
        assert _oracle.oracle().lsd_oracle_sort(a.copy(), a.copy(), 4, np.zeros(8, dtype=np.uint32), r) == -1


@pytest.mark.skipif(_oracle.ref() is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", keygen.KINDS)
def test_oracle_equals_live_reference(kind):
    ref = _oracle.ref()
    n = 1 << 18
    keys = keygen.make_keys(kind, n, 7)
    for r in (4, 8):
        a, b, h = keys.copy(), np.zeros_like(keys), np.zeros(1 << r, dtype=np.uint32)
        ref.ref_cpu_sort(a, b, n, h, r)
        assert np.array_equal(_oracle.sort(keys, r), b)
    s = keys.copy()
    ref.ref_cpu_prefix_sum(s, n)
    assert np.array_equal(_oracle.prefix_sum(keys), s)
    hist = np.zeros((n // 256) * 256, dtype=np.uint32)
    ref.ref_cpu_build_histograms(keys.copy(), hist, n, 8, 1, n // 256, 256)
    assert np.array_equal(_oracle.build_histograms(keys, 8, 1, 256).ravel(), hist)


def test_config1_cpu_path_2pow20():
    """BASELINE config[0]: prefix_sum + build_histogram + LSD sort of 2^20 uniform keys on the CPU path."""
    n = 1 << 20
    keys = keygen.make_keys("uniform", n, 0)
    out = _oracle.sort(keys, 8)
    assert np.all(out[:-1] <= out[1:]) and np.array_equal(out, np.sort(keys))
    h = _oracle.build_histograms(keys, 8, 0, 256)
    assert h.sum() == n
    flat = _oracle.prefix_sum(h.ravel())
    assert flat[0] == 0 and flat[-1] == n - h.ravel()[-1]
