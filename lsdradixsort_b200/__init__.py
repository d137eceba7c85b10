"""lsdradixsort_b200 -- B200-native (sm_100a) LSD radix sort of uint32 keys.

A drop-in for the one hot path of emanuele-xyz/LSDRadixSort (build_histogram, exclusive
prefix_sum, LSD sort with radix bits R and block size B).  All compute is hand-written CUDA in
``csrc/`` behind the C ABI of ``include/lsdsort.h``; this package is the host-side mirror.
"""
from ._native import LsdError, LIB_PATH, build_library  # noqa: F401
from . import api  # noqa: F401
from .api import (  # noqa: F401
    BuildHistograms,
    GetGPUPrefixSumBlockSumsCount,
    GPULSDRadixSort,
    GPUPrefixSum,
    HostSorter,
    PairSorter,
    argsort,
    sort_pairs_,
    Sorter,
    Sorter64,
    sort64_,
    SortInfo,
    build_histogram,
    digit_histograms,
    prefix_sum_,
    set_device,
    sort_,
    ipc_export,
    ipc_open,
    sort_pass,
    sort_pass_scatter,
    sort_workspace_bytes,
    top_digit_histogram,
)

__version__ = "0.1.0"
