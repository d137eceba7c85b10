"""Signed and floating-point key orders (lsd_key_type): the reference sorts uint32 only
(LSDRadixSort.cu:839); i32 / f32 keys are mapped to unsigned order inside the digit histogram and the first / last
executed pass.  CPU part: the typed oracle (numpy mapping around the reference-pinned sort) against numpy's own sort.
GPU part: the CUDA path through the C ABI against that oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import _oracle
from lsdradixsort_b200 import _native as N


def _float_bits(n, seed, specials=True):
    rng = np.random.default_rng(seed)
    f = (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)).astype(np.float32)
    if specials and n >= 16:
        f[:8] = [0.0, -0.0, np.inf, -np.inf, 1.0, -1.0, np.float32(1e-45), -np.float32(1e-45)]
        f[8:12] = np.array([0x7FC00000, 0xFFC00000, 0x7F800001, 0xFF800001], dtype=np.uint32).view(np.float32)  # NaNs
    rng.shuffle(f)
    return f.view(np.uint32)


def _int_bits(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(-(2**31), 2**31, n, dtype=np.int64).astype(np.int32)
    if n >= 4:
        a[:4] = [0, -1, 2**31 - 1, -(2**31)]
    return a.view(np.uint32)


# ------------------------------------------------------------------ CPU: the typed oracle
@pytest.mark.parametrize("r", [4, 8])
def test_typed_oracle_matches_numpy_sort(r):
    ib = _int_bits(30_001, 1)
    assert np.array_equal(_oracle.sort_typed(ib, "i32", r).view(np.int32), np.sort(ib.view(np.int32)))
    fb = _float_bits(30_001, 2, specials=False)
    got = _oracle.sort_typed(fb, "f32", r).view(np.float32)
    assert np.array_equal(got, np.sort(fb.view(np.float32)))  # no NaN / signed zeros: `<` decides everything
    # with specials: total order -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN
    fb = _float_bits(5000, 3)
    got = _oracle.sort_typed(fb, "f32", r)
    gf = got.view(np.float32)
    finite = gf[~np.isnan(gf)]
    assert np.all(finite[1:] >= finite[:-1])
    assert np.isnan(gf[0]) and np.isnan(gf[1]) and np.isnan(gf[-1]) and np.isnan(gf[-2])
    z = np.flatnonzero(gf == 0.0)
    assert got[z[0]] == 0x80000000 and got[z[-1]] == 0  # -0 before +0
    assert np.array_equal(np.sort(got), np.sort(fb))     # a permutation of the input
    for kt in ("u32", "i32", "f32"):
        x = _int_bits(1000, 4)
        assert np.array_equal(_oracle.from_unsigned(_oracle.to_unsigned(x, kt), kt), x)


def test_typed_abi_validation_without_cuda():
    lib = N.lib()
    big = 1 << 30
    bad_type = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, 0, 0, 7, 0)
    assert lib.lsd_sort_ex(0x1000, 0x2000, 16, 8, 0, 0x3000, big, C.byref(bad_type), None) == N.LSD_ERR_INVALID_VALUE
    cpc_typed = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, 2, 0, N.LSD_KEY_F32, 0)  # variant 2: no typed-key mapping
    assert lib.lsd_sort_ex(0x1000, 0x2000, 16, 8, 0, 0x3000, big, C.byref(cpc_typed), None) == N.LSD_ERR_UNSUPPORTED


# ------------------------------------------------------------------ GPU
def _dev(bits, dtype):
    import torch

    return torch.from_numpy(bits.view(np.int32)).cuda().view(dtype)


def _host(t):
    import torch

    return t.view(torch.int32).cpu().numpy().view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 33, 1000, 8351, 8352, 8353, 100_003, (1 << 20) + 9])
@pytest.mark.parametrize("kt", ["i32", "f32"])
def test_typed_sort_sizes_r8(kt, n):
    import torch

    import lsdradixsort_b200 as L

    bits = _int_bits(n, n) if kt == "i32" else _float_bits(n, n)
    d = _dev(bits, torch.int32 if kt == "i32" else torch.float32)
    L.sort_(d, r=8, key_type=kt) if kt == "i32" else L.sort_(d, r=8)  # float32 tensors pick f32 by themselves
    assert np.array_equal(_host(d), _oracle.sort_typed(bits, kt, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("block", [0, 128, 256, 512, 1024])
@pytest.mark.parametrize("kt", ["i32", "f32"])
def test_typed_sort_radix_and_block_sweep(kt, r, block):
    import torch

    import lsdradixsort_b200 as L

    n = 70_000 + 13
    bits = _int_bits(n, r + block) if kt == "i32" else _float_bits(n, r + block)
    d = _dev(bits, torch.int32)
    L.sort_(d, r=r, block=block, key_type=kt)
    assert np.array_equal(_host(d), _oracle.sort_typed(bits, kt, r))


@pytest.mark.gpu
@pytest.mark.parametrize("kt", ["i32", "f32"])
def test_typed_sort_with_skipped_passes(kt):
    """Small-magnitude keys: the top digits are constant, so the first / last EXECUTED pass (where the mapping is
    applied) is not pass 0 / pass 3; and all-equal keys, where no pass runs at all."""
    import torch

    import lsdradixsort_b200 as L

    rng = np.random.default_rng(5)
    n = 200_001
    if kt == "i32":
        bits = rng.integers(0, 1 << 12, n, dtype=np.int64).astype(np.int32).view(np.uint32)  # positive, 12 bits
    else:
        bits = (1.0 + rng.integers(0, 1 << 10, n) * 2.0 ** -23).astype(np.float32).view(np.uint32)  # same exponent
    s = L.Sorter(n, r=8, key_type=kt)
    d = _dev(bits, torch.int32)
    s.sort_(d)
    assert np.array_equal(_host(d), _oracle.sort_typed(bits, kt, 8))
    assert s.info(n).skipped_mask == 0b1100
    bits = np.full(n, 0xBF800000, dtype=np.uint32)  # -1.0f everywhere
    d = _dev(bits, torch.int32)
    s.sort_(d)
    assert np.array_equal(_host(d), bits) and s.info(n).skipped_mask == 0b1111


@pytest.mark.gpu
@pytest.mark.parametrize("kt", ["i32", "f32"])
@pytest.mark.parametrize("r", [4, 8])
def test_typed_pairs_and_argsort(kt, r):
    import torch

    import lsdradixsort_b200 as L

    n = 120_007
    bits = _int_bits(n, 7) if kt == "i32" else _float_bits(n, 7)
    vals = np.arange(n, dtype=np.uint32)
    dk, dv = _dev(bits, torch.int32), _dev(vals, torch.int32)
    L.sort_pairs_(dk, dv, r=r, key_type=kt)
    wk, wv = _oracle.sort_pairs_typed(bits, vals, kt, r)
    assert np.array_equal(_host(dk), wk) and np.array_equal(_host(dv), wv)
    if kt == "f32":
        f = torch.from_numpy(bits.view(np.float32)).cuda()
        perm = L.argsort(f, r=r)
        assert np.array_equal(_host(perm), wv)


@pytest.mark.gpu
def test_typed_large_matches_torch_sort():
    import torch

    import lsdradixsort_b200 as L

    n = 1 << 26
    f = torch.randn(n, device="cuda")
    want, _ = torch.sort(f)
    L.sort_(f)
    assert bool((f == want).all())
    i = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int64, device="cuda").to(torch.int32)
    want, _ = torch.sort(i)
    L.sort_(i, key_type="i32")
    assert bool((i == want).all())
