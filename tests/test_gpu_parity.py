"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs -- bit-exact, as the reference's own CheckArrays
checks are (.cu:364 scan, :785 histograms, :1018 GPU sort vs CPU sort, :120 vs std::sort)."""
import ctypes as C

import numpy as np
import pytest
import torch

import _oracle
import lsdradixsort_b200 as L
from lsdradixsort_b200 import _native as N
from lsdradixsort_b200 import keygen

pytestmark = pytest.mark.gpu


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int32)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


def test_native_library_is_loaded_and_device_is_blackwell():
    lib = N.lib()
    sm, smem, major, minor = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    N.check(lib.lsd_device_info(C.byref(sm), C.byref(smem), C.byref(major), C.byref(minor)), "lsd_device_info")
    assert sm.value > 0 and smem.value >= 128 * 1024
    assert major.value == 10, "kernels are built for sm_100a only"


# ------------------------------------------------------------------------------------------
# LSD sort
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 1000, 4096, 8191, 8192, 8193, 100_003, 1 << 20])
def test_sort_sizes_r8(n):
    keys = keygen.make_keys("uniform", n, seed=n)
    d = dev(keys)
    L.sort_(d, r=8)
    assert np.array_equal(host(d), _oracle.sort(keys, 8))


@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("block", [0, 128, 256, 512, 1024])
def test_sort_radix_and_block_sweep(r, block):
    """The reference's sweep axes: rs = {1,2,4,8} (.cu:1055-1062) x blocks (.cu:1044-1053)."""
    n = 200_000 + 37
    keys = keygen.make_keys("uniform", n, seed=r * 10 + block)
    d = dev(keys)
    L.sort_(d, r=r, block=block)
    want = _oracle.sort(keys, r)
    assert np.array_equal(host(d), want)
    assert np.array_equal(want, np.sort(keys))


@pytest.mark.parametrize("kind", keygen.KINDS)
@pytest.mark.parametrize("r", [4, 8])
def test_sort_skewed_distributions(kind, r):
    """BASELINE config 4 shapes at a size the oracle sorts in well under a second."""
    n = (1 << 19) + 123
    keys = keygen.make_keys(kind, n, seed=3)
    d = dev(keys)
    s = L.Sorter(n, r=r)
    s.sort_(d)
    assert np.array_equal(host(d), _oracle.sort(keys, r))
    info = s.info(n)
    passes = 32 // r
    if kind == "all_equal":
        assert info.skipped_mask == (1 << passes) - 1
    elif kind == "low_nibble":
        assert info.skipped_mask == ((1 << passes) - 1) & ~1
    elif kind in ("uniform", "entropy4_table"):
        assert info.skipped_mask == 0


def test_sort_skip_disabled_gives_same_result():
    n = 70_000
    keys = keygen.make_keys("low_nibble", n, 1)
    d = dev(keys)
    s = L.Sorter(n, r=8, disable_skip=True)
    s.sort_(d)
    assert s.info(n).skipped_mask == 0
    assert np.array_equal(host(d), _oracle.sort(keys, 8))


def _variant_exists(variant: int, r: int = 8) -> bool:
    opt = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, variant, 0, 0, 0)
    return N.lib().lsd_sort_workspace_bytes_ex(1 << 20, r, 0, C.byref(opt)) > 0


@pytest.mark.parametrize("kind", ["uniform", "entropy4_table", "all_equal"])
@pytest.mark.parametrize("variant", list(range(1, 96)))
def test_sort_kernel_variants_r8(variant, kind):
    """Every kernel shape of the r = 8 table.  The product library holds variants 0-2; the tuning build (make TUNING=1,
    LSDSORT_LIB=lsdradixsort_b200/liblsdsort_tuning.so) appends the measured-and-rejected families."""
    if not _variant_exists(variant):
        pytest.skip("variant not in this build of liblsdsort (tuning variants: make TUNING=1)")
    n = 300_000 + 11
    keys = keygen.make_keys(kind, n, seed=variant)
    d = dev(keys)
    L.sort_(d, r=8, variant=variant)
    assert np.array_equal(host(d), _oracle.sort(keys, 8))


@pytest.mark.parametrize("block", [128, 256, 1024])
def test_block_is_a_hint_not_a_slower_kernel(block):
    """A drop-in caller passes the reference's B (LSDRadixSort.cu:839): it must get the same kernel shape as block = 0."""
    for r in (4, 8):
        assert L.sort_workspace_bytes(1 << 24, r, block) == L.sort_workspace_bytes(1 << 24, r, 0)
    n = 1 << 22
    keys = keygen.make_keys("uniform", n, seed=block)
    a, b = dev(keys), dev(keys)
    sa, sb = L.Sorter(n, r=8, block=block), L.Sorter(n, r=8, block=0)
    sa.sort_(a)
    sb.sort_(b)
    assert torch.equal(a, b) and sa.info(n).launches == sb.info(n).launches


def test_sort_2pow32_keys_on_one_gpu():
    """n = 2^32 (the reference stops at `int count`, .cu:839): 16 GiB of keys, checked through size-independent properties
    -- ascending order, the 4 x 256 digit histograms and the 64-bit sum of the input -- plus an exact compare of a window."""
    n = 1 << 32
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs ~36 GiB of free device memory")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(32)
    step = 1 << 27
    total = 0
    for lo in range(0, n, step):
        d[lo:lo + step] = torch.randint(-(2**31), 2**31, (step,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        total += int((d[lo:lo + step].to(torch.int64) & 0xFFFFFFFF).sum().item())
    hist_before = L.digit_histograms(d, 8).cpu()
    s = L.Sorter(n, r=8)
    s.sort_(d)
    torch.cuda.synchronize()
    assert torch.equal(L.digit_histograms(d, 8).cpu(), hist_before)
    got_total, prev_last = 0, -1
    for lo in range(0, n, step):
        u = d[lo:lo + step].to(torch.int64) & 0xFFFFFFFF
        assert bool((u[1:] >= u[:-1]).all()) and int(u[0].item()) >= prev_last
        prev_last = int(u[-1].item())
        got_total += int(u.sum().item())
    assert got_total == total
    # uniform keys: key k sits near position k, so the output window [0, 2^20) is exactly the keys below its last value
    edge = int((d[(1 << 20) - 1].to(torch.int64) & 0xFFFFFFFF).item())
    assert hist_before[3][: (edge >> 24)].sum().item() <= (1 << 20)


@pytest.mark.parametrize("kind", ["entropy4_table", "uniform"])
@pytest.mark.parametrize("r", [1, 2, 4, 8])
def test_sort_multi_portion_handoff(r, kind):
    """Inputs above 2^30-1 keys run as several look-back portions; force tiny portions to cover the
    bucket-base hand-off between them (r = 4 and r = 1: the quad look-back of lookback_quad.cuh writes the bases)."""
    n = 150_001
    keys = keygen.make_keys(kind, n, seed=5)
    d = dev(keys)
    L.sort_(d, r=r, portion_keys=16384)
    assert np.array_equal(host(d), _oracle.sort(keys, r))


@pytest.mark.parametrize("r", [1, 4])
def test_sort_narrow_digits_many_tiles(r):
    """r = 4 / r = 1 at 2^24 + 77 keys (2009 tiles, ragged last tile): the quad look-back walks through deep windows."""
    n = (1 << 24) + 77
    g = torch.Generator(device="cuda").manual_seed(r)
    d = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    want = torch.sort(d.to(torch.int64) & 0xFFFFFFFF).values.to(torch.int32)
    L.sort_(d, r=r)
    assert torch.equal(d, want)


def test_sort_reference_shaped_call_leaves_result_in_a():
    n = 1 << 16
    keys = keygen.make_keys("uniform", n, 11)
    a, b = dev(keys), torch.empty(n, dtype=torch.int32, device="cuda")
    h = torch.empty(L.sort_workspace_bytes(n, 8, 256), dtype=torch.uint8, device="cuda")
    L.GPULSDRadixSort(a, b, h, n, 256, 8)
    assert np.array_equal(host(a), _oracle.sort(keys, 8))


def test_sort_is_idempotent_and_permutation_preserving():
    n = 1 << 18
    keys = keygen.make_keys("entropy4_table", n, 2)
    d = dev(keys)
    L.sort_(d)
    once = host(d).copy()
    L.sort_(d)
    assert np.array_equal(once, host(d))
    assert np.array_equal(np.bincount(once & 0xFF, minlength=256), np.bincount(keys & 0xFF, minlength=256))


def test_sort_status_codes_on_device():
    n = 4096
    d = dev(keygen.make_keys("uniform", n, 0))
    scratch = torch.empty(n, dtype=torch.int32, device="cuda")
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    st = N.lib().lsd_sort(d.data_ptr(), scratch.data_ptr(), n, 8, 0, ws.data_ptr(), 256, None)
    assert st == N.LSD_ERR_WORKSPACE_TOO_SMALL
    with pytest.raises(L.LsdError):
        L.GPULSDRadixSort(d, scratch, ws, n, 0, 8)


def test_sort_host_buffers_roundtrip():
    n = 1 << 20
    keys = keygen.make_keys("uniform", n, 21)
    hs = L.HostSorter(n, r=8)
    pinned = torch.from_numpy(keys.view(np.int32).copy()).pin_memory()
    hs.sort_(pinned)
    assert np.array_equal(pinned.numpy().view(np.uint32), _oracle.sort(keys, 8))
    pageable = keys.copy()
    hs.sort_(pageable[: n // 3])
    assert np.array_equal(pageable[: n // 3], np.sort(keys[: n // 3]))
    hs.close()


def test_sort_host_async_two_contexts():
    """lsd_sort_host_async / lsd_host_ctx_wait: two contexts used alternately (the D2H copy of one array overlaps the H2D copy
    of the next); every array is compared with the oracle after its wait."""
    n = (1 << 20) + 123
    sorters = [L.HostSorter(n, r=8), L.HostSorter(n, r=8)]
    bufs = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
    pending = [None, None]
    for step in range(6):
        slot = step & 1
        if pending[slot] is not None:
            sorters[slot].wait()
            assert np.array_equal(bufs[slot].numpy().view(np.uint32)[: pending[slot].size], _oracle.sort(pending[slot], 8))
        m = n - 1000 * step
        keys = keygen.make_keys("uniform" if step % 3 else "entropy4_table", m, 40 + step)
        bufs[slot][:m].copy_(torch.from_numpy(keys.view(np.int32)))
        sorters[slot].sort_async_(bufs[slot][:m])
        pending[slot] = keys
    for slot in range(2):
        sorters[slot].wait()
        assert np.array_equal(bufs[slot].numpy().view(np.uint32)[: pending[slot].size], _oracle.sort(pending[slot], 8))
        sorters[slot].close()


def _big_uniform(n: int, seed: int) -> np.ndarray:
    """keygen.uniform_u32 in 2^26-key pieces (same values, bounded temporaries)."""
    out = np.empty(n, dtype=np.uint32)
    step = 1 << 26
    for off in range(0, n, step):
        out[off:off + step] = keygen.uniform_u32(min(step, n - off), seed, offset=off)
    return out


def _usort(t: torch.Tensor) -> torch.Tensor:
    """torch.sort of the unsigned images (library sort, second opinion on the device)."""
    return torch.sort(t.to(torch.int64) & 0xFFFFFFFF).values.to(torch.int32)


def test_sort_full_size_2pow28_bit_exact():
    """BASELINE config 2 at full size: 2^28 uniform keys, bit-exact against the reference's own CPU sort compiled from
    /root/reference (CheckArrays at .cu:1018; the plain-C oracle when oracle/_ref is absent) and against a library sort
    on the device (the std::sort leg of .cu:120)."""
    n = 1 << 28
    keys = _big_uniform(n, seed=0)
    d = dev(keys)
    s = L.Sorter(n, r=8)
    s.sort_(d)
    torch.cuda.synchronize()
    assert s.info(n).skipped_mask == 0
    got = host(d)
    ref = _oracle.ref()
    if ref is not None:
        a, out, hist = keys.copy(), np.empty_like(keys), np.zeros(256, dtype=np.uint32)
        ref.ref_cpu_sort(a, out, n, hist, 8)  # LSDRadixSort (.cu:62-69), unmodified
        want = out
    else:
        want = _oracle.sort(keys, 8)
    assert np.array_equal(got, want)
    del want
    assert torch.equal(d, _usort(dev(keys)))
    again = d.clone()
    s.sort_(again)  # idempotence at full size
    assert torch.equal(again, d)


@pytest.mark.parametrize("kind", ["all_equal", "entropy4_table", "low_nibble", "sorted", "reverse"])
def test_sort_skewed_full_size_2pow28_bit_exact(kind):
    """BASELINE config 4 at full size, bit-exact against a library sort on the device (the oracle pins the same
    distributions at 2^19 in test_sort_skewed_distributions)."""
    n = 1 << 28
    keys = keygen.make_keys(kind, n, seed=0)
    d = dev(keys)
    want = _usort(d)
    L.Sorter(n, r=8).sort_(d)
    assert torch.equal(d, want)
    if kind in ("sorted", "all_equal"):
        assert np.array_equal(host(d[: 1 << 20]), keys[: 1 << 20])  # already in order: the sort is the identity


def test_prefix_sum_and_histograms_full_size_bit_exact():
    """BASELINE config 3 at a roofline-sized point: prefix_sum at 2^28 and 2^30 words against the oracle's wrap-around
    scan (CheckArrays at .cu:364) and build_histogram / digit histograms at 2^28 against numpy counts (.cu:785)."""
    for log2n in (28, 30):
        n = 1 << log2n
        a = _big_uniform(n, seed=log2n)
        d = dev(a)
        L.prefix_sum_(d, 256)
        want = _oracle.prefix_sum(a)
        got = host(d)
        assert np.array_equal(got, want)
        del d, want, got, a
    n = 1 << 28
    keys = _big_uniform(n, seed=7)
    d = dev(keys)
    assert np.array_equal(L.digit_histograms(d, 8).cpu().numpy().astype(np.uint64), _oracle.digit_histograms(keys, 8))
    for r, block, g in ((8, 256, 1), (8, 512, 3), (1, 128, 18)):
        got = L.build_histogram(d, r, g, block).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, _oracle.build_histograms(keys, r, g, block))  # BuildHistogramsCPU layout [G][2^r]
        del got


# ------------------------------------------------------------------------------------------
# prefix_sum
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 127, 4095, 4096, 4097, 65_537, (1 << 20) + 3])
@pytest.mark.parametrize("block", [128, 256, 512])
def test_prefix_sum_matches_oracle(n, block):
    a = keygen.uniform_u32(n, seed=n + block)  # full-range words: sums wrap mod 2^32 like the reference
    d = dev(a)
    L.prefix_sum_(d, block)
    assert np.array_equal(host(d), _oracle.prefix_sum(a))


def test_prefix_sum_reference_shaped_call():
    n = 1 << 18
    a = keygen.uniform_u32(n, 1)
    d = dev(a)
    words = L.GetGPUPrefixSumBlockSumsCount(n, 128)
    assert words < _oracle.oracle().lsd_oracle_block_sums_count(n, 128) + 128  # no more scratch than the reference
    ws = torch.empty(max(words, 64), dtype=torch.int32, device="cuda")
    L.GPUPrefixSum(d, n, 128, ws)
    assert np.array_equal(host(d), _oracle.prefix_sum(a))


def test_prefix_sum_large_small_values():
    n = (1 << 24) + 17
    a = (keygen.uniform_u32(n, 3) & np.uint32(7)).astype(np.uint32)
    d = dev(a)
    L.prefix_sum_(d)
    assert np.array_equal(host(d), _oracle.prefix_sum(a))


# ------------------------------------------------------------------------------------------
# build_histogram
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("block", [32, 128, 256, 512, 1024])
def test_build_histogram_reference_layout(r, block):
    n = 64 * 1024
    keys = keygen.make_keys("uniform", n, seed=r + block)
    for g in sorted({0, (32 // r) // 2, 32 // r - 1}):
        got = L.build_histogram(dev(keys), r, g, block).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, _oracle.build_histograms(keys, r, g, block))


@pytest.mark.parametrize("n,block", [(1, 256), (255, 256), (257, 256), (1000, 96), (5000, 1), (12345, 1000)])
def test_build_histogram_ragged_and_odd_blocks(n, block):
    keys = keygen.make_keys("entropy4_table", n, seed=n)
    got = L.build_histogram(dev(keys), 8, 2, block).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, _oracle.build_histograms(keys, 8, 2, block))


def test_build_histogram_overwrites_stale_output():
    n, r, block = 4096, 4, 128
    keys = keygen.make_keys("uniform", n, 0)
    out = torch.full((n // block, 1 << r), 12345, dtype=torch.int32, device="cuda")
    L.BuildHistograms(dev(keys), out, n, r, 3, n // block, block)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), _oracle.build_histograms(keys, r, 3, block))


@pytest.mark.parametrize("r", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", ["uniform", "all_equal", "sorted"])
def test_digit_histograms(r, kind):
    n = 1_000_003
    keys = keygen.make_keys(kind, n, seed=r)
    got = L.digit_histograms(dev(keys), r).cpu().numpy().astype(np.uint64)
    assert np.array_equal(got, _oracle.digit_histograms(keys, r))


def test_digit_histograms_tiny():
    for n in (0, 1, 3, 5):
        keys = keygen.make_keys("uniform", n, 0)
        d = dev(keys) if n else torch.empty(0, dtype=torch.int32, device="cuda")
        got = L.digit_histograms(d, 8).cpu().numpy().astype(np.uint64)
        assert np.array_equal(got, _oracle.digit_histograms(keys, 8))


# ------------------------------------------------------------------------------------------
# single pass (lsd_sort_pass): stability is observable here, unlike in the full keys-only sort
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r,bit_group", [(8, 0), (8, 3), (4, 5), (2, 9), (1, 17), (1, 31)])
@pytest.mark.parametrize("n", [1, 1000, 123_457])
def test_sort_pass_matches_reference_pass(r, bit_group, n):
    """One pass == the reference's LSDRadixSortPass (.cu:25-54): keys with equal digits keep input order."""
    keys = keygen.make_keys("uniform", n, seed=r * 100 + bit_group)
    src, dst = dev(keys), torch.zeros(n, dtype=torch.int32, device="cuda")
    offs = L.sort_pass(src, dst, r, bit_group, want_offsets=True)
    a, want, hist = keys.copy(), np.zeros_like(keys), np.zeros(1 << r, dtype=np.uint32)
    _oracle.oracle().lsd_oracle_sort_pass(a, want, n, hist, r, bit_group)
    assert np.array_equal(host(dst), want)
    assert np.array_equal(host(src), keys)  # input untouched
    assert np.array_equal(offs.cpu().numpy().astype(np.uint32), hist)  # bucket starts, as the CPU twin leaves them


def test_sort_pass_chain_equals_full_sort():
    n = 50_000
    keys = keygen.make_keys("entropy4_table", n, 8)
    a, b = dev(keys), torch.empty(n, dtype=torch.int32, device="cuda")
    for g in range(4):
        L.sort_pass(a, b, 8, g)
        a, b = b, a
    assert np.array_equal(host(a), np.sort(keys))


def test_multi_gpu_ops_single_rank_roundtrip():
    """CudaOps (the device half of multi.distributed_sort) on one GPU with a 1-rank gloo group."""
    import torch.distributed as dist

    from lsdradixsort_b200 import multi

    n = 200_000
    keys = keygen.make_keys("uniform", n, 77)
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29571", rank=0, world_size=1)
    try:
        ops = multi.CudaOps(n + 1024, r=8)
        d = dev(keys)
        recv, staging = ops.empty(n + 1024), ops.empty(n)
        # gloo moves CPU tensors only; with one rank the exchange is the identity, so emulate it on device
        hist = ops.top_digit_histogram(d).cpu().numpy()
        assert np.array_equal(hist.astype(np.uint64), _oracle.digit_histograms(keys, 8)[-1])
        ops.partition_by_top_digit(d, staging)
        part = host(staging)
        assert np.array_equal(part >> 24, np.sort(keys >> 24))  # grouped by destination bucket
        recv[:n].copy_(staging)
        ops.sort_(recv[:n])
        assert np.array_equal(host(recv[:n]), np.sort(keys))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# third parity leg: the reference's OWN CUDA kernels (oracle/_ref, rebuilt for sm_100a, unmodified)
# ------------------------------------------------------------------------------------------
def _ref_or_skip():
    ref = _oracle.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libref_lsd.so not present (built in the container from /root/reference)")
    return ref


@pytest.mark.parametrize("r,block", [(8, 1024), (8, 256), (4, 512), (1, 128)])
def test_sort_matches_reference_gpu(r, block):
    """Bit-exact against GPULSDRadixSort (.cu:839) called directly (its TestGPU... wrapper would SKIP r=8, .cu:940),
    under its preconditions: count % block == 0, powers of two."""
    ref = _ref_or_skip()
    n = 1 << 20
    keys = keygen.make_keys("uniform", n, seed=77 + r)
    grid = n // block
    a, b = dev(keys), torch.empty(n, dtype=torch.int32, device="cuda")
    h = torch.empty(3 * grid * (1 << r), dtype=torch.int32, device="cuda")
    bs = torch.empty(ref.ref_block_sums_count(grid * (1 << r), block) + 64, dtype=torch.int32, device="cuda")
    assert ref.ref_gpu_sort(a.data_ptr(), b.data_ptr(), h.data_ptr(), bs.data_ptr(), n, block, r) == 0
    assert ref.ref_device_synchronize() == 0
    ours = dev(keys)
    L.sort_(ours, r=r, block=block)
    assert np.array_equal(host(ours), host(a))
    assert np.array_equal(host(ours), np.sort(keys))


@pytest.mark.parametrize("block", [128, 256, 512])
def test_prefix_sum_matches_reference_gpu(block):
    """GPUPrefixSum (.cu:286) vs lsd_prefix_sum on the same words (the reference's own check is .cu:364)."""
    ref = _ref_or_skip()
    n = 1 << 20
    words = keygen.make_keys("uniform", n, seed=block)
    a = dev(words)
    bs = torch.empty(ref.ref_block_sums_count(n, block) + 64, dtype=torch.int32, device="cuda")
    assert ref.ref_gpu_prefix_sum(a.data_ptr(), n, block, bs.data_ptr()) == 0
    assert ref.ref_device_synchronize() == 0
    ours = dev(words)
    L.prefix_sum_(ours, block)
    assert np.array_equal(host(ours), host(a))


@pytest.mark.parametrize("r,block", [(1, 128), (8, 256), (8, 512)])
def test_build_histogram_matches_reference_gpu(r, block):
    """BuildHistogramsKernel (.cu:660) vs lsd_build_histogram, reference layout [G][2^r] (its own check is .cu:785)."""
    ref = _ref_or_skip()
    n = 1 << 20
    keys = keygen.make_keys("uniform", n, seed=r + block)
    a = dev(keys)
    grid = n // block
    h = torch.empty(grid * (1 << r), dtype=torch.int32, device="cuda")
    assert ref.ref_gpu_build_histograms(a.data_ptr(), h.data_ptr(), n, r, 1 if r == 8 else 3, block) == 0
    assert ref.ref_device_synchronize() == 0
    ours = L.build_histogram(a, r, 1 if r == 8 else 3, block)
    assert torch.equal(ours.view(-1), h)


@pytest.mark.parametrize("n", [1, 1000, 8352, 300_017])
def test_sort_pass_scatter_with_local_pointers_equals_sort_pass(n):
    """lsd_sort_pass_scatter (the fused partition + exchange step) with every bucket pointer aimed at one local buffer
    must reproduce lsd_sort_pass: bucket d at its start offset, stable inside the bucket."""
    keys = keygen.make_keys("uniform" if n != 1000 else "entropy4_table", n, seed=n)
    src = dev(keys)
    want = torch.empty_like(src)
    L.sort_pass(src, want, 8, 3)
    starts = np.concatenate([[0], np.cumsum(np.bincount(keys >> 24, minlength=256))[:-1]]).astype(np.int64)
    out = torch.full((n + 8,), -1, dtype=torch.int32, device="cuda")
    ptrs = torch.from_numpy(out.data_ptr() + 4 * starts).cuda()
    L.sort_pass_scatter(src, ptrs, 8, 3)
    torch.cuda.synchronize()
    assert np.array_equal(host(out[:n]), host(want))
    assert bool((out[n:] == -1).all())
    ref_out = np.zeros(n, dtype=np.uint32)
    _oracle.oracle().lsd_oracle_sort_pass(keys.copy(), ref_out, n, np.zeros(256, dtype=np.uint32), 8, 3)
    assert np.array_equal(host(out[:n]), ref_out)


def test_sort_pass_scatter_segments_group_buckets_per_tile():
    """Two destination segments (buckets 0..127 and 128..255): each destination receives exactly its keys, and sorting
    what arrived gives the same multiset as the oracle (the order inside a destination is (tile, bucket, position))."""
    n = 100_000
    keys = keygen.make_keys("uniform", n, seed=5)
    src = dev(keys)
    lo_cnt = int((keys >> 24 < 128).sum())
    a = torch.full((lo_cnt + 4,), -1, dtype=torch.int32, device="cuda")
    b = torch.full((n - lo_cnt + 4,), -1, dtype=torch.int32, device="cuda")
    ptrs = torch.tensor([a.data_ptr()] * 128 + [b.data_ptr()] * 128, dtype=torch.int64).cuda()
    seg = torch.tensor([0 | (127 << 16)] * 128 + [128 | (255 << 16)] * 128, dtype=torch.int32).cuda()
    L.sort_pass_scatter(src, ptrs, 8, 3, dst_seg=seg)
    torch.cuda.synchronize()
    assert bool((a[lo_cnt:] == -1).all()) and bool((b[n - lo_cnt:] == -1).all())
    ga, gb = host(a[:lo_cnt]), host(b[: n - lo_cnt])
    assert np.array_equal(np.sort(ga), np.sort(keys[keys >> 24 < 128]))
    assert np.array_equal(np.sort(gb), np.sort(keys[keys >> 24 >= 128]))


@pytest.mark.parametrize("r", [1, 4, 8])
@pytest.mark.parametrize("n", [0, 5, 100_001])
def test_top_digit_histogram_is_the_last_row_of_digit_histograms(r, n):
    keys = keygen.make_keys("uniform", n, seed=3 * r + 1)
    d = dev(keys) if n else torch.empty(0, dtype=torch.int32, device="cuda")
    top = L.top_digit_histogram(d, r).cpu().numpy().astype(np.uint64)
    assert np.array_equal(top, _oracle.digit_histograms(keys, r).reshape(32 // r, 1 << r)[-1])


def test_sort_is_cuda_graph_capturable():
    """The whole sort (memset, histogram, device-side plan, passes, copy-back) is enqueued without host synchronisation,
    so it can be captured once and replayed: same workspace, new keys each replay."""
    n = 500_000 + 3
    s = L.Sorter(n, r=8)
    buf = torch.empty(n, dtype=torch.int32, device="cuda")
    first = keygen.make_keys("uniform", n, seed=1)
    buf.copy_(dev(first))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        s.sort_(buf)  # warm-up outside capture (function attributes are set on first launch)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        buf.copy_(dev(first))
        with torch.cuda.graph(g, stream=side):
            s.sort_(buf)
    torch.cuda.synchronize()
    for seed, kind in ((2, "uniform"), (3, "low_nibble"), (4, "sorted")):
        keys = keygen.make_keys(kind, n, seed=seed)
        buf.copy_(dev(keys))
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(host(buf), _oracle.sort(keys, 8)), kind


def test_sort_writes_nothing_outside_its_buffers():
    """compute-sanitizer is closed on the GPU pool, so the bounds check is done by hand: keys, scratch and workspace are
    carved out of one allocation with guard zones in between; the guards must survive plain, typed and key-value sorts."""
    guard = 4096  # int32 words, 16 KiB
    for n in (1, 8351, 8352, 8353, 100_003, 444 * 8352 + 17):
        ws_words = (max(L.sort_workspace_bytes(n, 8, 0), 256) + 3) // 4
        ws_words = (ws_words + 63) // 64 * 64
        n_pad = (n + 63) // 64 * 64
        sizes = [n_pad, n_pad, n_pad, n_pad, ws_words]  # keys, scratch, vals, vals_scratch, workspace
        total = sum(sizes) + guard * (len(sizes) + 1)
        arena = torch.full((total,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        offs, o = [], guard
        for s in sizes:
            offs.append(o)
            o += s + guard
        keys_np = keygen.make_keys("uniform", n, seed=n)
        views = [arena[a:a + s] for a, s in zip(offs, sizes)]
        k, sc, v, vs, ws = views
        for flavour in ("plain", "f32", "pairs"):
            k[:n] = dev(keys_np)
            ws_bytes = ws.view(torch.uint8)
            if flavour == "pairs":
                v[:n] = torch.arange(n, dtype=torch.int32, device="cuda")
                st = N.lib().lsd_sort_pairs(k.data_ptr(), v.data_ptr(), sc.data_ptr(), vs.data_ptr(), n, 8, 0, ws.data_ptr(),
                                            ws_bytes.numel(), None, torch.cuda.current_stream().cuda_stream)
            else:
                o = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, 0, 0, N.LSD_KEY_F32 if flavour == "f32" else 0, 0)
                st = N.lib().lsd_sort_ex(k.data_ptr(), sc.data_ptr(), n, 8, 0, ws.data_ptr(), ws_bytes.numel(), C.byref(o),
                                         torch.cuda.current_stream().cuda_stream)
            assert st == 0
            torch.cuda.synchronize()
            want = _oracle.sort_typed(keys_np, "f32", 8) if flavour == "f32" else _oracle.sort(keys_np, 8)
            assert np.array_equal(host(k[:n]), want), (n, flavour)
            prev_end = 0
            for a, s in zip(offs + [total], sizes + [0]):
                g = arena[prev_end:a]
                assert bool((g == 0x5A5A5A5A).all()), f"guard before offset {a} was overwritten (n={n}, {flavour})"
                prev_end = a + s
            for t, s in zip((k, sc), (n, n)):  # padding words behind the n keys of keys / scratch stay untouched too
                assert bool((t[s:] == 0x5A5A5A5A).all()), (n, flavour)


def test_sort_stress_random_sizes_against_torch():
    """Many back-to-back sorts of random sizes and distributions on one Sorter (persistent kernel, ticket hand-over through
    an mbarrier, look-back across tiles): any rare ordering bug shows up as a mismatch with torch.sort."""
    g = torch.Generator(device="cuda").manual_seed(7)
    rng = np.random.default_rng(7)
    cap = 3_000_000
    s = L.Sorter(cap, r=8)
    for it in range(150):
        n = int(rng.integers(1, cap))
        bits = int(rng.choice([4, 12, 20, 32]))
        hi = (1 << bits) - 1
        k = torch.randint(0, hi + 1, (n,), dtype=torch.int64, device="cuda", generator=g)
        want, _ = torch.sort(k)
        work = (k & 0xFFFFFFFF).to(torch.int32) if bits < 32 else (k - (k >> 31 << 32)).to(torch.int32)
        s.sort_(work)
        got = work.to(torch.int64) & 0xFFFFFFFF
        assert bool((got == want).all()), (it, n, bits)
