"""ctypes doorway onto the CPU checkers in oracle/ -- TEST INFRASTRUCTURE ONLY.

``oracle()``  -> oracle/liblsd_oracle.so  (plain-C restatement; built on demand with gcc)
``ref()``     -> oracle/_ref/libref_lsd.so (the unmodified reference, compiled from /root/reference by
                 oracle/Makefile; None when it has not been built, e.g. /root/reference absent)
Nothing under lsdradixsort_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "liblsd_oracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libref_lsd.so"

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

_oracle = None
_ref = None
_ref_tried = False


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        src = ORACLE_DIR / "lsd_oracle.c"
        if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
        l = C.CDLL(str(ORACLE_SO))
        l.lsd_oracle_digit.restype = C.c_uint32
        l.lsd_oracle_digit.argtypes = [C.c_uint32, C.c_int, C.c_int]
        l.lsd_oracle_sort_pass.restype = None
        l.lsd_oracle_sort_pass.argtypes = [_u32p, _u32p, C.c_int64, _u32p, C.c_int, C.c_int]
        l.lsd_oracle_sort.restype = C.c_int
        l.lsd_oracle_sort.argtypes = [_u32p, _u32p, C.c_int64, _u32p, C.c_int]
        l.lsd_oracle_sort_pairs.restype = C.c_int
        l.lsd_oracle_sort_pairs.argtypes = [_u32p, _u32p, _u32p, _u32p, C.c_int64, _u32p, C.c_int]
        l.lsd_oracle_prefix_sum.restype = None
        l.lsd_oracle_prefix_sum.argtypes = [_u32p, C.c_int64]
        l.lsd_oracle_build_histograms.restype = None
        l.lsd_oracle_build_histograms.argtypes = [_u32p, _u32p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int]
        l.lsd_oracle_block_sums_count.restype = C.c_int64
        l.lsd_oracle_block_sums_count.argtypes = [C.c_int64, C.c_int]
        l.lsd_oracle_digit_histograms.restype = None
        l.lsd_oracle_digit_histograms.argtypes = [_u32p, C.c_int64, C.c_int, _u64p]
        l.lsd_oracle_tiled_pass.restype = C.c_int
        l.lsd_oracle_tiled_pass.argtypes = [_u32p, _u32p, C.c_int64, C.c_int, C.c_int, C.c_int]
        l.lsd_oracle_sort_pass_field.restype = None
        l.lsd_oracle_sort_pass_field.argtypes = [_u32p, _u32p, C.c_int64, _u64p, C.c_int, C.c_int]
        l.lsd_oracle_sort64.restype = C.c_int
        l.lsd_oracle_sort64.argtypes = [_u64p, _u64p, C.c_int64, _u64p, C.c_int]
        _oracle = l
    return _oracle


def ref():
    """The compiled reference, or None.  Its CPU entry points run anywhere; ref_gpu_* need a GPU."""
    global _ref, _ref_tried
    if not _ref_tried:
        _ref_tried = True
        if REF_SO.exists():
            l = C.CDLL(str(REF_SO))
            l.ref_cpu_sort.restype = None
            l.ref_cpu_sort.argtypes = [_u32p, _u32p, C.c_int, _u32p, C.c_int]
            l.ref_cpu_sort_pass.restype = None
            l.ref_cpu_sort_pass.argtypes = [_u32p, _u32p, C.c_int, _u32p, C.c_int, C.c_int]
            l.ref_cpu_prefix_sum.restype = None
            l.ref_cpu_prefix_sum.argtypes = [_u32p, C.c_int]
            l.ref_cpu_build_histograms.restype = None
            l.ref_cpu_build_histograms.argtypes = [_u32p, _u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            l.ref_block_sums_count.restype = C.c_int
            l.ref_block_sums_count.argtypes = [C.c_int, C.c_int]
            for name in ("ref_gpu_prefix_sum", "ref_gpu_build_histograms", "ref_gpu_sort"):
                getattr(l, name).restype = C.c_int
            l.ref_gpu_prefix_sum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            l.ref_gpu_build_histograms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            l.ref_gpu_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
            l.ref_device_synchronize.restype = C.c_int
            _ref = l
    return _ref


# ---- convenience wrappers (numpy in, numpy out; inputs are never modified) -----------------------
def sort(keys: np.ndarray, r: int = 8) -> np.ndarray:
    a = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    out = np.empty_like(a)
    hist = np.zeros(1 << r, dtype=np.uint32)
    rc = oracle().lsd_oracle_sort(a, out, a.size, hist, r)
    assert rc == 0, "oracle rejected r"
    assert np.array_equal(a, out), "reference post-condition: in and out both hold the result"
    return out


def sort_pairs(keys: np.ndarray, vals: np.ndarray, r: int = 8):
    """(sorted keys, values carried along): the reference's pass with a payload written at the same slot."""
    a = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    out, vout = np.empty_like(a), np.empty_like(v)
    hist = np.zeros(1 << r, dtype=np.uint32)
    rc = oracle().lsd_oracle_sort_pairs(a, v, out, vout, a.size, hist, r)
    assert rc == 0, "oracle rejected r"
    return out, vout


def prefix_sum(a: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(a, dtype=np.uint32).copy()
    oracle().lsd_oracle_prefix_sum(b, b.size)
    return b


def build_histograms(keys: np.ndarray, r: int, bit_group: int, block: int) -> np.ndarray:
    a = np.ascontiguousarray(keys, dtype=np.uint32)
    grid = (a.size + block - 1) // block
    h = np.zeros(grid * (1 << r), dtype=np.uint32)
    oracle().lsd_oracle_build_histograms(a, h, a.size, r, bit_group, grid, block)
    return h.reshape(grid, 1 << r)


def digit_histograms(keys: np.ndarray, r: int) -> np.ndarray:
    a = np.ascontiguousarray(keys, dtype=np.uint32)
    h = np.zeros((32 // r) * (1 << r), dtype=np.uint64)
    oracle().lsd_oracle_digit_histograms(a, a.size, r, h)
    return h.reshape(32 // r, 1 << r)


# ---- typed keys (lsd_key_type): the reference's sort applied to the keys' unsigned images ----------------------
def to_unsigned(bits: np.ndarray, key_type: str) -> np.ndarray:
    """Order-preserving bijection onto unsigned order: i32 flips the sign bit, f32 flips all bits of negatives and the
    sign bit of the others (IEEE total order).  Plain numpy restatement of key_to_unsigned in csrc/common.cuh."""
    b = np.ascontiguousarray(bits).view(np.uint32)
    if key_type == "u32":
        return b.copy()
    if key_type == "i32":
        return b ^ np.uint32(0x80000000)
    assert key_type == "f32"
    return np.where(b >> np.uint32(31), ~b, b ^ np.uint32(0x80000000)).astype(np.uint32)


def from_unsigned(u: np.ndarray, key_type: str) -> np.ndarray:
    if key_type == "u32":
        return u.copy()
    if key_type == "i32":
        return u ^ np.uint32(0x80000000)
    return np.where(u >> np.uint32(31), u ^ np.uint32(0x80000000), ~u).astype(np.uint32)


def sort_typed(bits: np.ndarray, key_type: str, r: int = 8) -> np.ndarray:
    """Sorted 32-bit patterns in `key_type` order: LSDRadixSort (.cu:62-69) on the unsigned images, mapped back."""
    return from_unsigned(sort(to_unsigned(bits, key_type), r), key_type)


def sort_pairs_typed(bits: np.ndarray, vals: np.ndarray, key_type: str, r: int = 8):
    k, v = sort_pairs(to_unsigned(bits, key_type), vals, r)
    return from_unsigned(k, key_type), v


# ---- composite digit widths and 64-bit keys --------------------------------------------------------
def digit_field(r: int, bit_group: int):
    """(shift, width) of digit `bit_group` at width r: bits [bit_group*r, min(32, (bit_group+1)*r))."""
    shift = bit_group * r
    return shift, min(r, 32 - shift)


def sort_pass_field(keys: np.ndarray, shift: int, width: int):
    """(out, bucket starts) of the reference's pass on the bit field [shift, shift+width)."""
    a = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty_like(a)
    hist = np.zeros(1 << width, dtype=np.uint64)
    oracle().lsd_oracle_sort_pass_field(a, out, a.size, hist, shift, width)
    return out, hist


def field_histogram(keys: np.ndarray, shift: int, width: int) -> np.ndarray:
    a = np.ascontiguousarray(keys, dtype=np.uint32)
    return np.bincount((a >> np.uint32(shift)) & np.uint32((1 << width) - 1), minlength=1 << width).astype(np.uint64)


def to_unsigned64(bits: np.ndarray, key_type: str) -> np.ndarray:
    b = np.ascontiguousarray(bits).view(np.uint64)
    if key_type == "u64":
        return b.copy()
    top = np.uint64(1 << 63)
    if key_type == "i64":
        return b ^ top
    assert key_type == "f64"
    return np.where(b >> np.uint64(63), ~b, b ^ top).astype(np.uint64)


def from_unsigned64(u: np.ndarray, key_type: str) -> np.ndarray:
    top = np.uint64(1 << 63)
    if key_type == "u64":
        return u.copy()
    if key_type == "i64":
        return u ^ top
    return np.where(u >> np.uint64(63), u ^ top, ~u).astype(np.uint64)


def sort64(bits: np.ndarray, key_type: str = "u64", r: int = 8) -> np.ndarray:
    """Sorted 64-bit patterns in `key_type` order: the reference's LSD loop on 64-bit words (lsd_oracle_sort64)."""
    a = to_unsigned64(bits, key_type)
    out = np.empty_like(a)
    hist = np.zeros(1 << r, dtype=np.uint64)
    rc = oracle().lsd_oracle_sort64(a, out, a.size, hist, r)
    assert rc == 0
    return from_unsigned64(out, key_type)
