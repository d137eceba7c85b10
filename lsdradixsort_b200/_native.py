"""ctypes binding of liblsdsort.so (the C ABI in include/lsdsort.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make lib``.  There is no
fallback: if the library is missing, importing the binding raises, and every wrapper turns a
non-zero status into :class:`LsdError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = Path(os.environ["LSDSORT_LIB"]) if os.environ.get("LSDSORT_LIB") else PKG_DIR / "liblsdsort.so"  # override: tuning builds

LSD_OK = 0
LSD_ERR_INVALID_VALUE = 1
LSD_ERR_WORKSPACE_TOO_SMALL = 2
LSD_ERR_CUDA = 3
LSD_ERR_UNSUPPORTED = 4
LSD_ERR_ALIGNMENT = 5
LSD_ERR_CAPACITY = 6
LSD_ERR_COMM = 7
LSD_KEY_U32, LSD_KEY_I32, LSD_KEY_F32 = 0, 1, 2
LSD_KEY_U64, LSD_KEY_I64, LSD_KEY_F64 = 3, 4, 5


class LsdError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: lsd status {status} ({detail})")


class SortOptions(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("portion_keys", C.c_uint32),
        ("disable_skip", C.c_uint32),
        ("variant", C.c_uint32),
        ("debug_trace", C.c_uint64),
        ("key_type", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


class MultiComm(C.Structure):
    """lsd_multi_comm: rank layout + the two collectives lsd_sort_multi needs, as callbacks."""
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("rank", C.c_int),
        ("nranks", C.c_int),
        ("all_gather", ALL_GATHER_FN),
        ("barrier", BARRIER_FN),
        ("ctx", C.c_void_p),
    ]


class MultiStats(C.Structure):
    _fields_ = [
        ("n_in", C.c_uint64),
        ("n_out", C.c_uint64),
        ("n_out_max", C.c_uint64),
        ("sent_bytes", C.c_uint64),
        ("first_bucket", C.c_uint32),
        ("last_bucket", C.c_uint32),
        ("plan_ms", C.c_float),
        ("exchange_ms", C.c_float),
        ("sort_ms", C.c_float),
        ("exchange_shift", C.c_uint32),
    ]


def build_library(verbose: bool = False) -> Path:
    """Compile liblsdsort.so for sm_100a with the repo Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(REPO_ROOT), f"-j{os.cpu_count() or 4}", "lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building liblsdsort.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the native library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `make lib` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`; lsdradixsort_b200 has no CPU fallback."
        )
    l = C.CDLL(str(LIB_PATH))
    u32p, u64p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_void_p
    sig = {
        "lsd_version": (C.c_int, []),
        "lsd_status_string": (C.c_char_p, [C.c_int]),
        "lsd_last_cuda_error": (C.c_int, []),
        "lsd_set_device": (C.c_int, [C.c_int]),
        "lsd_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
        "lsd_build_histogram_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int]),
        "lsd_build_histogram": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.c_int, vp, vp]),
        "lsd_digit_histograms": (C.c_int, [vp, C.c_uint64, C.c_int, vp, vp]),
        "lsd_top_digit_histogram": (C.c_int, [vp, C.c_uint64, C.c_int, vp, vp]),
        "lsd_prefix_sum_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_int]),
        "lsd_prefix_sum": (C.c_int, [vp, C.c_uint64, C.c_int, vp, C.c_size_t, vp]),
        "lsd_sort_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int]),
        "lsd_sort_workspace_bytes_ex": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int, C.POINTER(SortOptions)]),
        "lsd_sort": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_size_t, vp]),
        "lsd_sort_ex": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(SortOptions), vp]),
        "lsd_sort_pairs_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int, C.POINTER(SortOptions)]),
        "lsd_sort_pairs": (
            C.c_int,
            [vp, vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(SortOptions), vp],
        ),
        "lsd_sort_pairs_timed": (
            C.c_int,
            [vp, vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(SortOptions), vp,
             C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)],
        ),
        "lsd_sort64_workspace_bytes": (C.c_size_t, [C.c_uint64]),
        "lsd_sort64": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, vp, C.c_size_t, vp]),
        "lsd_sort_pass": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp]),
        "lsd_sort_pass_scatter": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, vp, vp, vp, C.c_size_t, vp]),
        "lsd_ipc_export": (C.c_int, [vp, vp, u64p]),
        "lsd_ipc_open": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
        "lsd_ipc_close": (C.c_int, [vp, C.c_uint64]),
        "lsd_multi_ctx_create": (C.c_int, [C.POINTER(MultiComm), vp, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(vp), vp]),
        "lsd_multi_ctx_destroy": (C.c_int, [vp]),
        "lsd_sort_multi": (C.c_int, [vp, vp, C.c_uint64, vp, u64p, vp]),
        "lsd_multi_last_stats": (C.c_int, [vp, C.POINTER(MultiStats)]),
        "lsd_multi_set_timing": (C.c_int, [vp, C.c_int]),
        "lsd_sort_timed": (
            C.c_int,
            [vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(SortOptions), vp,
             C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)],
        ),
        "lsd_sort_read_plan": (C.c_int, [vp, C.c_uint64, C.c_int, u32p, C.POINTER(C.c_int), vp]),
        "lsd_host_ctx_create": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]),
        "lsd_host_ctx_destroy": (C.c_int, [vp]),
        "lsd_sort_host": (C.c_int, [vp, vp, C.c_uint64]),
        "lsd_sort_host_async": (C.c_int, [vp, vp, C.c_uint64]),
        "lsd_host_ctx_wait": (C.c_int, [vp]),
        "lsd_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
        "lsd_host_free": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(l, name)  # AttributeError here == the library does not export the header's symbol
        fn.restype = res
        fn.argtypes = args
    l._lsd_signatures = sig
    _lib = l
    return l


def check(status: int, where: str) -> None:
    if status == LSD_OK:
        return
    l = lib()
    detail = l.lsd_status_string(status).decode()
    if status == LSD_ERR_CUDA:
        detail += f", cudaError={l.lsd_last_cuda_error()}"
    raise LsdError(status, where, detail)
