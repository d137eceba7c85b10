"""64-bit keys (lsd_sort64) and composite digit widths (r = 11, 16, ...): SURVEY 8(f)4.

The reference sorts uint32 with r in {1, 2, 4, 8} on the GPU (LSDRadixSort.cu:839, :953); its CPU path takes any factor
of 32, i.e. also r = 16 (:56-69).  CPU part: the oracle's widened loops against numpy / the compiled reference, and the
decomposition argument the CUDA path relies on (a stable pass on a w-bit digit == stable passes on its low 8 bits and
on the rest).  GPU part: the CUDA path through the C ABI against that oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import _oracle
from lsdradixsort_b200 import _native as N
from lsdradixsort_b200 import keygen


def _keys64(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 2**64, n, dtype=np.uint64)
    if kind == "low_word":      # every key below 2^32: the four high passes are skipped
        return rng.integers(0, 2**32, n, dtype=np.uint64)
    if kind == "high_word":     # only the high word varies
        return rng.integers(0, 2**32, n, dtype=np.uint64) << np.uint64(32) | np.uint64(0x12345678)
    if kind == "equal":
        return np.full(n, 0xDEADBEEFCAFEF00D, dtype=np.uint64)
    if kind == "sorted":
        return np.arange(n, dtype=np.uint64) * np.uint64(0x100000001)
    if kind == "reverse":
        return (np.arange(n, dtype=np.uint64) * np.uint64(0x100000001))[::-1].copy()
    if kind == "few":           # 16 distinct 64-bit values
        table = rng.integers(0, 2**64, 16, dtype=np.uint64)
        return table[rng.integers(0, 16, n)]
    raise ValueError(kind)


def _f64_bits(n, seed):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal(n) * 10.0 ** rng.integers(-200, 200, n)
    if n >= 12:
        f[:8] = [0.0, -0.0, np.inf, -np.inf, 1.0, -1.0, 5e-324, -5e-324]
        f[8:12] = np.array([0x7FF8000000000000, 0xFFF8000000000000, 0x7FF0000000000001, 0xFFF0000000000001],
                           dtype=np.uint64).view(np.float64)
    rng.shuffle(f)
    return f.view(np.uint64)


# ------------------------------------------------------------------ CPU: the widened oracle
@pytest.mark.parametrize("r", [8, 16])
def test_oracle_sort64_matches_numpy(r):
    for kind in ("uniform", "low_word", "high_word", "few", "reverse"):
        k = _keys64(kind, 20_003, 3)
        assert np.array_equal(_oracle.sort64(k, "u64", r), np.sort(k))
    i = _keys64("uniform", 20_003, 4)
    assert np.array_equal(_oracle.sort64(i, "i64", r).view(np.int64), np.sort(i.view(np.int64)))
    fb = _f64_bits(20_003, 5)
    got = _oracle.sort64(fb, "f64", r)
    assert np.array_equal(_oracle.to_unsigned64(got, "f64"), np.sort(_oracle.to_unsigned64(fb, "f64")))  # total order
    finite = got.view(np.float64)[~np.isnan(got.view(np.float64))]
    assert np.all(finite[:-1] <= finite[1:])


def test_oracle_r16_is_the_reference_cpu_sort():
    """r = 16 is the one digit width the reference's CPU path takes beyond 1, 2, 4, 8 (LSDRadixSort.cu:56-69)."""
    k = keygen.make_keys("uniform", 50_001, seed=16)
    want = np.sort(k)
    assert np.array_equal(_oracle.sort(k, 16), want)
    ref = _oracle.ref()
    if ref is not None:
        a, out, hist = k.copy(), np.empty_like(k), np.zeros(1 << 16, dtype=np.uint32)
        ref.ref_cpu_sort(a, out, a.size, hist, 16)
        assert np.array_equal(out, want)


@pytest.mark.parametrize("r,bit_group", [(11, 0), (11, 1), (11, 2), (16, 0), (16, 1), (13, 2), (5, 6)])
def test_wide_pass_equals_its_sub_passes(r, bit_group):
    """The decomposition lsd_sort_pass uses for composite widths, checked on the CPU: one stable pass on the w-bit
    field == a stable pass on its low 8 bits followed by a stable pass on the remaining bits; for r = 16 the field pass
    is the reference's own LSDRadixSortPass."""
    k = keygen.make_keys("uniform", 40_009, seed=r * 10 + bit_group)
    shift, width = _oracle.digit_field(r, bit_group)
    whole, starts = _oracle.sort_pass_field(k, shift, width)
    if width > 8:
        lo, _ = _oracle.sort_pass_field(k, shift, 8)
        both, _ = _oracle.sort_pass_field(lo, shift + 8, width - 8)
        assert np.array_equal(both, whole)
    hist = _oracle.field_histogram(k, shift, width)
    assert np.array_equal(starts, np.concatenate([[0], np.cumsum(hist)[:-1]]).astype(np.uint64))
    if r == 16:
        a, out, h32 = k.copy(), np.empty_like(k), np.zeros(1 << 16, dtype=np.uint32)
        _oracle.oracle().lsd_oracle_sort_pass(a, out, a.size, h32, 16, bit_group)
        assert np.array_equal(out, whole) and np.array_equal(h32.astype(np.uint64), starts)


# ------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


def _torch():
    import torch

    return torch


def _dev(a):
    torch = _torch()
    return torch.from_numpy(a.view(np.int64 if a.dtype.itemsize == 8 else np.int32).copy()).cuda()


@gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 1000, 8352 * 2 + 3, 300_011, (1 << 20) + 2])
def test_sort64_matches_oracle_sizes(n):
    import lsdradixsort_b200 as L

    k = _keys64("uniform", n, n + 1)
    d = _dev(k)
    L.sort64_(d, key_type="u64")
    assert np.array_equal(d.cpu().numpy().view(np.uint64), _oracle.sort64(k, "u64"))


@gpu
@pytest.mark.parametrize("kind", ["uniform", "low_word", "high_word", "equal", "sorted", "reverse", "few"])
def test_sort64_distributions(kind):
    import lsdradixsort_b200 as L

    for n in (100_003, 65_536):
        k = _keys64(kind, n, 7)
        d = _dev(k)
        L.sort64_(d, key_type="u64")
        assert np.array_equal(d.cpu().numpy().view(np.uint64), _oracle.sort64(k, "u64")), (kind, n)


@gpu
def test_sort64_signed_and_float_orders():
    import lsdradixsort_b200 as L

    torch = _torch()
    i = _keys64("uniform", 200_001, 11)
    d = _dev(i)
    L.sort64_(d)  # int64 tensor: signed order by default
    assert np.array_equal(d.cpu().numpy(), np.sort(i.view(np.int64)))
    assert np.array_equal(d.cpu().numpy().view(np.uint64), _oracle.sort64(i, "i64"))
    fb = _f64_bits(200_003, 12)
    f = torch.from_numpy(fb.view(np.float64).copy()).cuda()
    L.sort64_(f)
    assert np.array_equal(f.cpu().numpy().view(np.uint64), _oracle.sort64(fb, "f64"))
    clean = np.random.default_rng(13).standard_normal(100_000)
    g = torch.from_numpy(clean.copy()).cuda()
    L.sort64_(g)
    assert torch.equal(g, torch.sort(torch.from_numpy(clean).cuda()).values)


@gpu
def test_sort64_reuses_a_sorter_and_leaves_neighbours_alone():
    import lsdradixsort_b200 as L

    torch = _torch()
    s = L.Sorter64(50_000)
    for n in (50_000, 49_999, 13, 50_000):
        k = _keys64("uniform", n, n)
        buf = torch.full((n + 16,), 0x5A5A5A5A5A5A5A5A, dtype=torch.int64, device="cuda")
        buf[8:8 + n] = _dev(k)
        s.sort_(buf[8:8 + n], key_type="u64")
        out = buf.cpu().numpy().view(np.uint64)
        assert np.array_equal(out[8:8 + n], np.sort(k))
        assert np.all(out[:8] == 0x5A5A5A5A5A5A5A5A) and np.all(out[8 + n:] == 0x5A5A5A5A5A5A5A5A)


@gpu
def test_sort64_large_matches_torch():
    import lsdradixsort_b200 as L

    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(64)
    n = (1 << 25) + 3
    k = torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    want = torch.sort(k).values
    L.sort64_(k)
    assert torch.equal(k, want)


@gpu
@pytest.mark.parametrize("r", [3, 5, 6, 7, 11, 13, 16])
def test_full_sort_with_composite_digit_width(r):
    import lsdradixsort_b200 as L

    torch = _torch()
    for kind, n in (("uniform", 200_003), ("entropy4_table", 70_001), ("sorted", 8352 * 3)):
        k = keygen.make_keys(kind, n, seed=r)
        d = _dev(k)
        L.sort_(d, r=r)
        assert np.array_equal(d.cpu().numpy().view(np.uint32), _oracle.sort(k, 8)), (r, kind)
    # key-value: the stable permutation does not depend on r either
    k = keygen.make_keys("entropy4_table", 50_000, seed=r + 100)
    d = _dev(k)
    v = torch.arange(k.size, dtype=torch.int32, device="cuda")
    L.sort_pairs_(d, v, r=r)
    wk, wv = _oracle.sort_pairs(k, np.arange(k.size, dtype=np.uint32), 8)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), wk) and np.array_equal(v.cpu().numpy().view(np.uint32), wv)


@gpu
@pytest.mark.parametrize("r", [11, 16, 3, 13])
def test_sort_pass_with_composite_digit_width(r):
    """lsd_sort_pass on a composite digit: bit-exact (keys and the 2^r bucket starts) against the reference's pass on
    that bit field; chaining all digits sorts."""
    import lsdradixsort_b200 as L

    torch = _torch()
    for kind, n in (("uniform", 150_001), ("entropy4_table", 40_000), ("uniform", 5)):
        k = keygen.make_keys(kind, n, seed=r + 7)
        src = _dev(k)
        dst = torch.empty_like(src)
        digits = L.api.digit_count(r)
        cur = k
        for g in range(digits):
            offs = L.sort_pass(src, dst, r, g, want_offsets=True)
            shift, width = _oracle.digit_field(r, g)
            want, starts = _oracle.sort_pass_field(cur, shift, width)
            assert np.array_equal(dst.cpu().numpy().view(np.uint32), want), (r, g, kind)
            got_offs = offs.cpu().numpy().view(np.uint64)
            assert np.array_equal(got_offs[: 1 << width], starts)
            assert np.all(got_offs[1 << width:] == n)
            cur = want
            src, dst = dst, src
        assert np.array_equal(cur, np.sort(k))


@gpu
@pytest.mark.parametrize("r", [11, 16, 6])
def test_digit_histograms_with_composite_digit_width(r):
    import lsdradixsort_b200 as L

    k = keygen.make_keys("uniform", 300_007, seed=r)
    h = L.digit_histograms(_dev(k), r).cpu().numpy().view(np.uint64)
    assert h.shape == (L.api.digit_count(r), 1 << r)
    for g in range(h.shape[0]):
        shift, width = _oracle.digit_field(r, g)
        assert np.array_equal(h[g, : 1 << width], _oracle.field_histogram(k, shift, width))
        assert not h[g, 1 << width:].any()
