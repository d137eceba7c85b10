"""Key-value sort (lsd_sort_pairs): the reference's stable pass with a payload written at the same slot.

CPU part pins the pairs oracle: its KEYS equal the reference-pinned lsd_oracle_sort (and the compiled reference when
present), its VALUES equal numpy's stable argsort -- the stability the reference states (LSDRadixSort.cu:25-54) made
observable.  GPU part: the CUDA path through the C ABI against that oracle, bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

import _oracle
from lsdradixsort_b200 import _native as N
from lsdradixsort_b200 import keygen


# ------------------------------------------------------------------ CPU: oracle + boundary
@pytest.mark.parametrize("r", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("kind", ["uniform", "entropy4_table", "low_nibble", "all_equal", "reverse"])
def test_pairs_oracle_is_the_stable_permutation(kind, r):
    n = 20_011
    keys = keygen.make_keys(kind, n, seed=r)
    idx = np.arange(n, dtype=np.uint32)
    k, v = _oracle.sort_pairs(keys, idx, r)
    assert np.array_equal(k, _oracle.sort(keys, r))  # same keys as the reference-pinned key-only oracle
    assert np.array_equal(v, np.argsort(keys, kind="stable").astype(np.uint32))
    assert np.array_equal(keys[v], k)
    ref = _oracle.ref()
    if ref is not None and r != 16:
        a, out, hist = keys.copy(), np.zeros_like(keys), np.zeros(1 << r, dtype=np.uint32)
        ref.ref_cpu_sort(a, out, n, hist, r)
        assert np.array_equal(k, out)


def test_pairs_oracle_arbitrary_payload_and_empty():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 16, 5000, dtype=np.uint32)  # many duplicates
    vals = rng.integers(0, 2**32, 5000, dtype=np.uint32)
    k, v = _oracle.sort_pairs(keys, vals, 8)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    k0, v0 = _oracle.sort_pairs(np.zeros(0, np.uint32), np.zeros(0, np.uint32), 8)
    assert k0.size == 0 and v0.size == 0


def test_pairs_abi_validation_without_cuda():
    lib = N.lib()
    big = 1 << 30
    assert lib.lsd_sort_pairs(None, None, None, None, 0, 8, 0, None, 0, None, None) == N.LSD_OK
    assert lib.lsd_sort_pairs(None, None, None, None, 0, 17, 0, None, 0, None, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort_pairs(0x1000, None, 0x3000, 0x4000, 16, 8, 0, 0x5000, big, None, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort_pairs(0x1000, 0x2000, 0x3000, None, 16, 8, 0, 0x5000, big, None, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort_pairs(0x1000, 0x2004, 0x3000, 0x4000, 16, 8, 0, 0x5000, big, None, None) == N.LSD_ERR_ALIGNMENT
    assert lib.lsd_sort_pairs(0x1000, 0x2000, 0x3000, 0x4000, 16, 8, 0, 0x5000, 16, None, None) == N.LSD_ERR_WORKSPACE_TOO_SMALL
    assert lib.lsd_sort_pairs(0x1000, 0x2000, 0x3000, 0x4000, 16, 0, 0, 0x5000, big, None, None) == N.LSD_ERR_INVALID_VALUE
    for r in (1, 2, 4, 8):
        assert lib.lsd_sort_pairs_workspace_bytes(1 << 20, r, 0, None) > 0
        for block in (128, 256, 512, 1024):
            assert lib.lsd_sort_pairs_workspace_bytes(1 << 20, r, block, None) > 0
    # a tuning variant that has no key-value form is refused, not silently replaced
    no_pairs = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, 2, 0, 0, 0)  # variant 2: plain keys only
    assert lib.lsd_sort_pairs_workspace_bytes(1 << 20, 8, 0, C.byref(no_pairs)) == 0
    assert lib.lsd_sort_pairs(0x1000, 0x2000, 0x3000, 0x4000, 16, 8, 0, 0x5000, big, C.byref(no_pairs), None) == N.LSD_ERR_UNSUPPORTED


# ------------------------------------------------------------------ GPU: parity through the C ABI
def _dev(a):
    import torch

    return torch.from_numpy(a.view(np.int32)).cuda()


def _host(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 8191, 8352, 8353, 16704, 100_003, (1 << 20) + 5])
def test_pairs_sizes_r8(n):
    import lsdradixsort_b200 as L

    keys = keygen.make_keys("uniform", n, seed=n)
    vals = np.arange(n, dtype=np.uint32)[::-1].copy()
    dk, dv = _dev(keys), _dev(vals)
    L.sort_pairs_(dk, dv, r=8)
    wk, wv = _oracle.sort_pairs(keys, vals, 8)
    assert np.array_equal(_host(dk), wk)
    assert np.array_equal(_host(dv), wv)


@pytest.mark.gpu
@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("block", [0, 128, 256, 512, 1024])
def test_pairs_radix_and_block_sweep(r, block):
    import lsdradixsort_b200 as L

    n = 150_000 + 41
    keys = keygen.make_keys("entropy4_table", n, seed=r + block)  # 16 distinct keys: stability is what is tested
    idx = L.argsort(_dev(keys), r=r, block=block)
    assert np.array_equal(_host(idx), np.argsort(keys, kind="stable").astype(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", keygen.KINDS)
def test_pairs_skewed_distributions_and_skipping(kind):
    import lsdradixsort_b200 as L

    n = (1 << 18) + 77
    keys = keygen.make_keys(kind, n, seed=9)
    vals = np.random.default_rng(1).integers(0, 2**32, n, dtype=np.uint32)
    dk, dv = _dev(keys), _dev(vals)
    L.sort_pairs_(dk, dv, r=8)
    wk, wv = _oracle.sort_pairs(keys, vals, 8)
    assert np.array_equal(_host(dk), wk) and np.array_equal(_host(dv), wv)


@pytest.mark.gpu
@pytest.mark.parametrize("r", [4, 8])
def test_pairs_multi_portion_and_no_skip(r):
    import lsdradixsort_b200 as L

    n = 140_001
    keys = keygen.make_keys("low_nibble", n, seed=2)
    vals = np.arange(n, dtype=np.uint32)
    dk, dv = _dev(keys), _dev(vals)
    L.sort_pairs_(dk, dv, r=r, portion_keys=16384, disable_skip=True)
    wk, wv = _oracle.sort_pairs(keys, vals, r)
    assert np.array_equal(_host(dk), wk) and np.array_equal(_host(dv), wv)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1])
def test_pairs_other_shapes_r8(variant):
    import lsdradixsort_b200 as L

    n = 90_007
    keys = keygen.make_keys("entropy4_table", n, seed=variant)
    idx = L.argsort(_dev(keys), r=8, variant=variant)
    assert np.array_equal(_host(idx), np.argsort(keys, kind="stable").astype(np.uint32))


@pytest.mark.gpu
def test_pairs_large_property_checks():
    """2^26 pairs: the oracle is too slow to run per test at full size, so check the size-independent properties --
    keys ascending, gathered keys[perm] == sorted keys, perm is a permutation, equal keys keep ascending indices."""
    import torch

    import lsdradixsort_b200 as L

    n = 1 << 26
    keys = torch.randint(0, 1 << 20, (n,), dtype=torch.int32, device="cuda")  # ~64 duplicates per key value
    perm = L.argsort(keys, r=8)
    sk = keys[perm.long()]
    assert bool((sk[1:] >= sk[:-1]).all())
    ties = sk[1:] == sk[:-1]
    assert bool((perm[1:][ties] > perm[:-1][ties]).all())  # stability
    assert int(torch.bincount(perm.long(), minlength=n).max()) == 1
    ref, _ = torch.sort(keys)
    assert bool((sk == ref).all())
