"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the
header declares, and rejects bad arguments with status codes before touching CUDA (the reference
crashes instead: CudaUtils.h:7-8, Utils.h:6-15)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from lsdradixsort_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "lsdsort.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"LSD_API\s+[\w\s\*]+?\b(lsd_\w+)\s*\(", HEADER)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("lsd_build_histogram", "lsd_digit_histograms", "lsd_prefix_sum", "lsd_prefix_sum_workspace_bytes",
                 "lsd_sort", "lsd_sort_ex", "lsd_sort_workspace_bytes", "lsd_sort_host", "lsd_sort_timed"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), f"liblsdsort.so does not export {name}"
    # and the Python binding covers the whole header
    assert set(declared_symbols()) == set(lib._lsd_signatures)


def test_header_cites_reference_interfaces():
    for cite in ("LSDRadixSort.cu:660-702", "LSDRadixSort.cu:286-302", "LSDRadixSort.cu:839-910", ":265-276"):
        assert cite in HEADER


def test_version_and_status_strings():
    lib = N.lib()
    assert lib.lsd_version() == 100
    assert lib.lsd_status_string(0) == b"ok"
    assert lib.lsd_status_string(2) == b"workspace too small"
    assert lib.lsd_status_string(99) == b"unknown status"


def test_size_queries_need_no_gpu():
    lib = N.lib()
    # build_histogram: G * 2^r * 4 bytes, G = ceil(n / block)   (reference .cu:710-716)
    assert lib.lsd_build_histogram_bytes(1 << 20, 8, 256) == (1 << 20) // 256 * 256 * 4
    assert lib.lsd_build_histogram_bytes(1000, 1, 128) == 8 * 2 * 4
    assert lib.lsd_build_histogram_bytes(1000, 3, 128) == 0  # invalid radix
    assert lib.lsd_prefix_sum_workspace_bytes(0, 256) >= 256
    small, big = lib.lsd_prefix_sum_workspace_bytes(1 << 20, 256), lib.lsd_prefix_sum_workspace_bytes(1 << 28, 256)
    assert small < big < (1 << 28) * 4 // 16  # far below the reference's ~N/B words per level
    for r in (1, 2, 4, 8):
        a, b = lib.lsd_sort_workspace_bytes(1 << 20, r, 0), lib.lsd_sort_workspace_bytes(1 << 24, r, 0)
        assert 0 < a < b
        # never more scratch than the 3*G*H words the reference needs at B=256 (.cu:919-929) for large inputs
        assert lib.lsd_sort_workspace_bytes(1 << 28, r, 256) <= 3 * ((1 << 28) // 256) * (1 << r) * 4 + (1 << 20)
    assert lib.lsd_sort_workspace_bytes(1 << 20, 0, 0) == 0
    assert lib.lsd_sort_workspace_bytes(1 << 20, 17, 0) == 0
    # composite digit widths (every r in 3..16 other than 4 and 8): the 8-bit layout plus the temporary key
    # array and the 2^16-bin histogram of lsd_sort_pass's sub-passes
    for r in (3, 11, 16):
        assert lib.lsd_sort_workspace_bytes(1 << 20, r, 0) >= lib.lsd_sort_workspace_bytes(1 << 20, 8, 0) + (4 << 20) + (8 << 16)
    assert 0 < lib.lsd_sort64_workspace_bytes(0) <= lib.lsd_sort64_workspace_bytes(1 << 20)
    assert lib.lsd_sort_workspace_bytes(1 << 20, 8, 2048) == 0


@pytest.mark.parametrize("r,block", [(0, 256), (-8, 256), (17, 256), (32, 256), (8, -1), (8, 4096)])
def test_sort_rejects_bad_parameters_without_cuda(r, block):
    lib = N.lib()
    st = lib.lsd_sort(0x1000, 0x2000, 1024, r, block, 0x3000, 1 << 30, None)
    assert st == N.LSD_ERR_INVALID_VALUE


def test_sort_argument_validation_order():
    lib = N.lib()
    assert lib.lsd_sort(None, None, 0, 8, 0, None, 0, None) == N.LSD_OK  # n == 0 is a no-op
    assert lib.lsd_sort(None, 0x2000, 16, 8, 0, 0x3000, 1 << 30, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort(0x1000, 0x2000, 16, 8, 0, 0x3000, 16, None) == N.LSD_ERR_WORKSPACE_TOO_SMALL
    assert lib.lsd_sort(0x1004, 0x2000, 16, 8, 0, 0x3000, 1 << 30, None) == N.LSD_ERR_ALIGNMENT
    assert lib.lsd_sort(0x1000, 0x2000, 1 << 33, 8, 0, 0x3000, 1 << 40, None) == N.LSD_ERR_UNSUPPORTED
    bad = N.SortOptions(4, 0, 0, 0, 0, 0, 0)  # wrong struct_bytes
    assert lib.lsd_sort_ex(0x1000, 0x2000, 16, 8, 0, 0x3000, 1 << 30, C.byref(bad), None) == N.LSD_ERR_INVALID_VALUE
    unknown_variant = N.SortOptions(C.sizeof(N.SortOptions), 0, 0, 999, 0, 0, 0)
    assert lib.lsd_sort_ex(0x1000, 0x2000, 16, 8, 0, 0x3000, 1 << 30, C.byref(unknown_variant), None) == N.LSD_ERR_INVALID_VALUE


def test_histogram_and_scan_argument_validation():
    lib = N.lib()
    assert lib.lsd_build_histogram(0x1000, 1024, 8, 4, 256, 0x2000, None) == N.LSD_ERR_INVALID_VALUE  # bit_group == 32/r
    assert lib.lsd_build_histogram(0x1000, 1024, 8, -1, 256, 0x2000, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_build_histogram(0x1000, 1024, 8, 0, 0, 0x2000, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_build_histogram(None, 0, 8, 0, 256, None, None) == N.LSD_OK
    assert lib.lsd_digit_histograms(0x1000, 16, 17, 0x2000, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort_pass(0x1000, 0x2000, 16, 11, 3, 0, 0x3000, 1 << 30, None, None) == N.LSD_ERR_INVALID_VALUE  # r = 11 has 3 digits
    assert lib.lsd_sort_pass(0x1000, 0x2000, 16, 16, 0, 0, 0x3000, 16, None, None) == N.LSD_ERR_WORKSPACE_TOO_SMALL
    assert lib.lsd_sort64(0x1000, 0x2000, 16, 0, 0x3000, 1 << 30, None) == N.LSD_ERR_INVALID_VALUE  # a 32-bit key type
    assert lib.lsd_sort64(0x1000, 0x2008, 16, 3, 0x3000, 1 << 30, None) == N.LSD_ERR_ALIGNMENT
    assert lib.lsd_sort64(0x1000, 0x2000, 16, 3, 0x3000, 16, None) == N.LSD_ERR_WORKSPACE_TOO_SMALL
    assert lib.lsd_sort64(None, None, 0, 3, None, 0, None) == N.LSD_OK
    assert lib.lsd_prefix_sum(None, 0, 256, None, 0, None) == N.LSD_OK
    assert lib.lsd_prefix_sum(0x1000, 16, 4096, 0x2000, 1 << 20, None) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_prefix_sum(0x1004, 16, 256, 0x2000, 1 << 20, None) == N.LSD_ERR_ALIGNMENT
    assert lib.lsd_sort_host(None, None, 0) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_sort_host_async(None, None, 0) == N.LSD_ERR_INVALID_VALUE
    assert lib.lsd_host_ctx_wait(None) == N.LSD_ERR_INVALID_VALUE


def test_python_mirror_refuses_cpu_tensors():
    import torch

    import lsdradixsort_b200 as L

    with pytest.raises(TypeError):
        L.sort_(torch.zeros(8, dtype=torch.int32))
    with pytest.raises(TypeError):
        L.prefix_sum_(torch.zeros(8, dtype=torch.int32))
    with pytest.raises(TypeError):
        L.build_histogram(torch.zeros(8, dtype=torch.int32), 8, 0, 256)
