"""Host-side mirror of the reference's three entry points, over the C ABI.

torch is used for device memory and streams only; every compute call goes through
``liblsdsort.so`` (include/lsdsort.h).  Names follow the reference
(LSDRadixSort/LSDRadixSort.cu): ``GPULSDRadixSort`` (:839), ``GPUPrefixSum`` (:286),
``GetGPUPrefixSumBlockSumsCount`` (:265), ``BuildHistograms`` (kernel at :660); the snake_case
functions are the same calls with the scratch management done for the caller.

Keys are 32-bit words: tensors may be ``torch.uint32`` or ``torch.int32`` (bit pattern is what is
sorted, as unsigned, ascending -- the reference's order).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _native as N

_KEY_DTYPES = (torch.int32, torch.uint32, torch.float32)
_KEY_TYPES = {"u32": N.LSD_KEY_U32, "i32": N.LSD_KEY_I32, "f32": N.LSD_KEY_F32}


def _typed_opts(keys: torch.Tensor, opts: dict) -> dict:
    """float32 keys are ordered as floats unless the caller says otherwise; integer words keep the reference's order."""
    if "key_type" in opts or keys.dtype != torch.float32:
        return opts
    return dict(opts, key_type="f32")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check_keys(t: torch.Tensor, what: str, allow_float: bool = False) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor (this library has no CPU path)")
    if t.dtype not in _KEY_DTYPES or (t.dtype == torch.float32 and not allow_float):
        raise TypeError(f"{what} must be int32/uint32{'/float32' if allow_float else ''}, got {t.dtype}")
    if not t.is_contiguous() or t.dim() != 1:
        raise ValueError(f"{what} must be a contiguous 1-D tensor")


def _options(portion_keys: int = 0, disable_skip: bool = False, variant: int = 0,
             debug_trace: int = 0, key_type="u32") -> Optional[N.SortOptions]:
    """``key_type``: how the 32-bit words are ordered -- "u32" (the reference's order, default), "i32" (signed) or
    "f32" (IEEE total order); see lsd_key_type in include/lsdsort.h."""
    kt = _KEY_TYPES[key_type] if isinstance(key_type, str) else int(key_type)
    if not (portion_keys or disable_skip or variant or debug_trace or kt):
        return None
    return N.SortOptions(C.sizeof(N.SortOptions), int(portion_keys), int(bool(disable_skip)), int(variant),
                         int(debug_trace), kt, 0)


def set_device(index: int) -> None:
    """Bind this process/thread to a GPU in both torch and the native library (one process per GPU)."""
    torch.cuda.set_device(index)
    N.check(N.lib().lsd_set_device(int(index)), "lsd_set_device")


# --------------------------------------------------------------------------------------------------
# build_histogram
# --------------------------------------------------------------------------------------------------
def build_histogram(keys: torch.Tensor, r: int, bit_group: int, block: int,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-tile histograms ``h[G][2^r]`` of digit ``bit_group`` (reference layout, .cu:660-702)."""
    _check_keys(keys, "keys")
    n = keys.numel()
    nbytes = N.lib().lsd_build_histogram_bytes(n, r, block)
    if nbytes == 0 and n > 0:
        raise N.LsdError(N.LSD_ERR_INVALID_VALUE, "lsd_build_histogram_bytes", "invalid r/block")
    g = (n + block - 1) // block if block > 0 else 0
    if out is None:
        out = torch.empty((g, 1 << r), dtype=torch.int32, device=keys.device)
    N.check(
        N.lib().lsd_build_histogram(keys.data_ptr(), n, r, bit_group, block, out.data_ptr(), _stream_ptr(keys.device)),
        "lsd_build_histogram",
    )
    return out


def BuildHistograms(a: torch.Tensor, h: torch.Tensor, count: int, r: int, bit_group: int, grid: int, block: int) -> None:
    """Reference-shaped call: ``BuildHistogramsKernel<<<grid, block>>>(a, h, count, r, bit_group)``."""
    if grid != (count + block - 1) // block:
        raise ValueError("grid must be ceil(count / block), as in the reference (.cu:710)")
    build_histogram(a[:count], r, bit_group, block, out=h)


def digit_count(r: int) -> int:
    """Digits of a 32-bit key at width ``r``: 32/r for 1, 2, 4, 8, 16; ceil(32/r) for the other (composite) widths,
    whose top digit is narrower (r = 11: 11 + 11 + 10 bits)."""
    return (32 + r - 1) // r


def exec_radix(r: int) -> int:
    """The digit width a FULL sort executes: composite widths (every r in 3..16 other than 4 and 8) run the 8-bit schedule
    (the result of a full sort does not depend on r)."""
    return r if r in (1, 2, 4, 8) else 8


def digit_histograms(keys: torch.Tensor, r: int = 8) -> torch.Tensor:
    """Whole-array histograms of every digit: ``[digit_count(r)][2^r]`` int64 (one read of the keys for r = 1, 2, 4, 8;
    one read per digit for the composite widths)."""
    _check_keys(keys, "keys")
    out = torch.empty((digit_count(r), 1 << r), dtype=torch.int64, device=keys.device)
    N.check(
        N.lib().lsd_digit_histograms(keys.data_ptr(), keys.numel(), r, out.data_ptr(), _stream_ptr(keys.device)),
        "lsd_digit_histograms",
    )
    return out


def top_digit_histogram(keys: torch.Tensor, r: int = 8) -> torch.Tensor:
    """Histogram of the most significant digit only: ``[2^r]`` int64 (lsd_top_digit_histogram)."""
    _check_keys(keys, "keys")
    out = torch.empty((32 // r, 1 << r), dtype=torch.int64, device=keys.device)
    N.check(
        N.lib().lsd_top_digit_histogram(keys.data_ptr(), keys.numel(), r, out.data_ptr(), _stream_ptr(keys.device)),
        "lsd_top_digit_histogram",
    )
    return out[-1]


# --------------------------------------------------------------------------------------------------
# prefix_sum
# --------------------------------------------------------------------------------------------------
def GetGPUPrefixSumBlockSumsCount(count: int, threads_per_block: int) -> int:
    """Scratch size in uint32 words for :func:`GPUPrefixSum` (reference .cu:265-276 returns words too)."""
    return (N.lib().lsd_prefix_sum_workspace_bytes(count, threads_per_block) + 3) // 4


def GPUPrefixSum(d_a: torch.Tensor, count: int, threads_per_block: int, d_block_sums: torch.Tensor) -> None:
    """In-place exclusive scan mod 2^32 of ``d_a[:count]`` (reference .cu:286-302)."""
    _check_keys(d_a, "d_a")
    if d_block_sums.data_ptr() % 256:
        raise ValueError("d_block_sums must be 256-byte aligned")
    N.check(
        N.lib().lsd_prefix_sum(d_a.data_ptr(), count, threads_per_block, d_block_sums.data_ptr(),
                               d_block_sums.numel() * d_block_sums.element_size(), _stream_ptr(d_a.device)),
        "lsd_prefix_sum",
    )


def prefix_sum_(a: torch.Tensor, block: int = 256) -> torch.Tensor:
    """In-place exclusive prefix sum (uint32 wrap-around); allocates its own scratch."""
    _check_keys(a, "a")
    words = GetGPUPrefixSumBlockSumsCount(a.numel(), block)
    ws = torch.empty(max(words, 64), dtype=torch.int32, device=a.device)
    GPUPrefixSum(a, a.numel(), block, ws)
    return a


# --------------------------------------------------------------------------------------------------
# LSD sort
# --------------------------------------------------------------------------------------------------
def sort_workspace_bytes(n: int, r: int = 8, block: int = 0, **opts) -> int:
    o = _options(**opts)
    return N.lib().lsd_sort_workspace_bytes_ex(n, r, block, C.byref(o) if o else None)


def GPULSDRadixSort(a: torch.Tensor, b: torch.Tensor, h: torch.Tensor, count: int, block: int, r: int, **opts) -> None:
    """Reference-shaped call (.cu:839): sort ``a[:count]`` ascending using ``b`` as ping-pong space and
    ``h`` as scratch; the result is left in ``a``.  ``grid``, ``h_count``, ``d``, ``block_sums`` and
    ``block_sums_count`` of the reference are derived or unused and therefore not parameters here."""
    _check_keys(a, "a", allow_float=True)
    _check_keys(b, "b", allow_float=True)
    opts = _typed_opts(a, opts)
    if b.numel() < count:
        raise ValueError("b must hold at least `count` keys")
    o = _options(**opts)
    N.check(
        N.lib().lsd_sort_ex(a.data_ptr(), b.data_ptr(), count, r, block, h.data_ptr(),
                            h.numel() * h.element_size(), C.byref(o) if o else None, _stream_ptr(a.device)),
        "lsd_sort",
    )


def sort_pass(src: torch.Tensor, dst: torch.Tensor, r: int, bit_group: int, block: int = 0,
              workspace: Optional[torch.Tensor] = None, want_offsets: bool = False):
    """One stable counting-sort pass on digit ``bit_group``: ``dst <- src`` reordered (reference
    LSDRadixSortPass, .cu:25-54, without the copy-back).  Returns the bucket start offsets if asked."""
    _check_keys(src, "src")
    _check_keys(dst, "dst")
    n = src.numel()
    if dst.numel() < n:
        raise ValueError("dst too small")
    if workspace is None:
        workspace = torch.empty(max(sort_workspace_bytes(n, r, block), 256), dtype=torch.uint8, device=src.device)
    offs = torch.empty(1 << r, dtype=torch.int64, device=src.device) if want_offsets else None
    N.check(
        N.lib().lsd_sort_pass(src.data_ptr(), dst.data_ptr(), n, r, bit_group, block, workspace.data_ptr(),
                              workspace.numel(), offs.data_ptr() if offs is not None else None,
                              _stream_ptr(src.device)),
        "lsd_sort_pass",
    )
    return offs


def sort_pass_scatter(src: torch.Tensor, dst_ptrs: torch.Tensor, r: int, bit_group: int,
                      workspace: Optional[torch.Tensor] = None, dst_seg: Optional[torch.Tensor] = None) -> None:
    """Peer-scatter pass (lsd_sort_pass_scatter): the buckets of digit ``bit_group`` go to the device pointers
    ``dst_ptrs[d]`` (int64 tensor of 2^r addresses on src's device; local or IPC-opened peer memory), grouped into the
    segments ``dst_seg[d] = first | last << 16`` (int32 tensor; None = one segment per bucket)."""
    _check_keys(src, "src")
    if dst_ptrs.dtype != torch.int64 or not dst_ptrs.is_cuda or dst_ptrs.numel() != (1 << r):
        raise TypeError("dst_ptrs must be a CUDA int64 tensor with 2^r entries")
    if dst_seg is not None and (dst_seg.dtype != torch.int32 or not dst_seg.is_cuda or dst_seg.numel() != (1 << r)):
        raise TypeError("dst_seg must be a CUDA int32 tensor with 2^r entries")
    n = src.numel()
    if workspace is None:
        workspace = torch.empty(max(sort_workspace_bytes(n, r, 0), 256), dtype=torch.uint8, device=src.device)
    N.check(
        N.lib().lsd_sort_pass_scatter(src.data_ptr(), n, r, bit_group, dst_ptrs.data_ptr(),
                                      dst_seg.data_ptr() if dst_seg is not None else None, workspace.data_ptr(),
                                      workspace.numel(), _stream_ptr(src.device)),
        "lsd_sort_pass_scatter",
    )


def ipc_export(t: torch.Tensor) -> tuple:
    """(handle bytes, offset) naming ``t``'s device memory for another process on this node (lsd_ipc_export)."""
    handle = (C.c_ubyte * 64)()
    off = C.c_uint64(0)
    N.check(N.lib().lsd_ipc_export(t.data_ptr(), handle, C.byref(off)), "lsd_ipc_export")
    return bytes(handle), int(off.value)


def ipc_open(handle: bytes, offset: int) -> int:
    """Device address in THIS process of the peer buffer exported as (handle, offset) (lsd_ipc_open)."""
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    ptr = C.c_void_p()
    N.check(N.lib().lsd_ipc_open(buf, offset, C.byref(ptr)), "lsd_ipc_open")
    return int(ptr.value)


@dataclass
class SortInfo:
    skipped_mask: int
    launches: int


class Sorter:
    """Reusable scratch + workspace for repeated sorts of up to ``max_n`` keys on one device."""

    def __init__(self, max_n: int, r: int = 8, block: int = 0, device: Optional[torch.device] = None, **opts):
        self.max_n, self.r, self.block, self.opts = int(max_n), int(r), int(block), opts
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        nbytes = sort_workspace_bytes(self.max_n, r, block, **opts)
        if nbytes == 0 and self.max_n > 0:
            raise N.LsdError(N.LSD_ERR_INVALID_VALUE, "lsd_sort_workspace_bytes", "invalid r/block/variant or n too large")
        self.scratch = torch.empty(max(self.max_n, 1), dtype=torch.int32, device=self.device)
        self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)

    def sort_(self, keys: torch.Tensor) -> torch.Tensor:
        n = keys.numel()
        if n > self.max_n:
            raise ValueError("keys larger than this Sorter's capacity")
        GPULSDRadixSort(keys, self.scratch, self.workspace, n, self.block, self.r, **self.opts)
        return keys

    def sort_timed_(self, keys: torch.Tensor) -> list:
        """Sort and return per-stage device milliseconds [hist+plan, pass0.., copy-back] (synchronises)."""
        _check_keys(keys, "keys", allow_float=True)
        stages = 32 // exec_radix(self.r) + 2
        buf = (C.c_float * stages)()
        written = C.c_int(0)
        o = _options(**_typed_opts(keys, self.opts))
        N.check(
            N.lib().lsd_sort_timed(keys.data_ptr(), self.scratch.data_ptr(), keys.numel(), self.r, self.block,
                                   self.workspace.data_ptr(), self.workspace.numel(), C.byref(o) if o else None,
                                   _stream_ptr(keys.device), buf, stages, C.byref(written)),
            "lsd_sort_timed",
        )
        return list(buf)[: written.value]

    def info(self, n: int) -> SortInfo:
        mask, launches = C.c_uint32(0), C.c_int(0)
        N.check(
            N.lib().lsd_sort_read_plan(self.workspace.data_ptr(), n, self.r, C.byref(mask), C.byref(launches),
                                       _stream_ptr(self.device)),
            "lsd_sort_read_plan",
        )
        return SortInfo(mask.value, launches.value)


class PairSorter:
    """Key-value sort (lsd_sort_pairs): reusable ping-pong buffers + workspace for up to ``max_n`` pairs."""

    def __init__(self, max_n: int, r: int = 8, block: int = 0, device: Optional[torch.device] = None, **opts):
        self.max_n, self.r, self.block, self.opts = int(max_n), int(r), int(block), opts
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        o = _options(**opts)
        nbytes = N.lib().lsd_sort_pairs_workspace_bytes(self.max_n, r, block, C.byref(o) if o else None)
        if nbytes == 0 and self.max_n > 0:
            raise N.LsdError(N.LSD_ERR_INVALID_VALUE, "lsd_sort_pairs_workspace_bytes",
                             "invalid r/block/variant (or a variant without the key-value form) or n too large")
        self.keys_scratch = torch.empty(max(self.max_n, 1), dtype=torch.int32, device=self.device)
        self.vals_scratch = torch.empty(max(self.max_n, 1), dtype=torch.int32, device=self.device)
        self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)

    def _args(self, keys: torch.Tensor, vals: torch.Tensor):
        _check_keys(keys, "keys", allow_float=True)
        _check_keys(vals, "vals", allow_float=True)
        n = keys.numel()
        if vals.numel() != n:
            raise ValueError("keys and vals must have the same length")
        if n > self.max_n:
            raise ValueError("more pairs than this PairSorter's capacity")
        o = _options(**_typed_opts(keys, self.opts))
        return (keys.data_ptr(), vals.data_ptr(), self.keys_scratch.data_ptr(), self.vals_scratch.data_ptr(), n, self.r,
                self.block, self.workspace.data_ptr(), self.workspace.numel(), C.byref(o) if o else None,
                _stream_ptr(keys.device))

    def sort_(self, keys: torch.Tensor, vals: torch.Tensor):
        """Sort ``keys`` ascending in place; ``vals`` is permuted the same way (equal keys keep their input order)."""
        N.check(N.lib().lsd_sort_pairs(*self._args(keys, vals)), "lsd_sort_pairs")
        return keys, vals

    def sort_timed_(self, keys: torch.Tensor, vals: torch.Tensor) -> list:
        stages = 32 // exec_radix(self.r) + 2
        buf = (C.c_float * stages)()
        written = C.c_int(0)
        N.check(N.lib().lsd_sort_pairs_timed(*self._args(keys, vals), buf, stages, C.byref(written)),
                "lsd_sort_pairs_timed")
        return list(buf)[: written.value]


def sort_pairs_(keys: torch.Tensor, vals: torch.Tensor, r: int = 8, block: int = 0, **opts):
    """In-place key-value sort; allocates scratch for this one call."""
    _check_keys(keys, "keys", allow_float=True)
    return PairSorter(keys.numel(), r, block, device=keys.device, **opts).sort_(keys, vals)


def argsort(keys: torch.Tensor, r: int = 8, block: int = 0, **opts) -> torch.Tensor:
    """The stable sorting permutation of ``keys`` (int32 indices; ``keys`` is left untouched): values 0..n-1 carried
    through lsd_sort_pairs.  n < 2^31."""
    _check_keys(keys, "keys", allow_float=True)
    n = keys.numel()
    idx = torch.arange(n, dtype=torch.int32, device=keys.device)
    sort_pairs_(keys.clone(), idx, r, block, **opts)
    return idx


def sort_(keys: torch.Tensor, r: int = 8, block: int = 0, **opts) -> torch.Tensor:
    """Sort ``keys`` in place; allocates scratch for this one call.  int32/uint32 tensors are ordered as UNSIGNED words
    (the reference's order) unless ``key_type="i32"``; float32 tensors as floats (IEEE total order)."""
    _check_keys(keys, "keys", allow_float=True)
    return Sorter(keys.numel(), r, block, device=keys.device, **opts).sort_(keys)


_KEY64_TYPES = {torch.int64: N.LSD_KEY_I64, torch.float64: N.LSD_KEY_F64}
if hasattr(torch, "uint64"):
    _KEY64_TYPES[torch.uint64] = N.LSD_KEY_U64
_KEY64_NAMES = {"u64": N.LSD_KEY_U64, "i64": N.LSD_KEY_I64, "f64": N.LSD_KEY_F64}


class Sorter64:
    """64-bit keys (lsd_sort64): reusable ping-pong buffer + workspace for up to ``max_n`` keys."""

    def __init__(self, max_n: int, device: Optional[torch.device] = None):
        self.max_n = int(max_n)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        nbytes = N.lib().lsd_sort64_workspace_bytes(self.max_n)
        if nbytes == 0:
            raise N.LsdError(N.LSD_ERR_UNSUPPORTED, "lsd_sort64_workspace_bytes", "n too large")
        self.scratch = torch.empty(max(self.max_n, 2), dtype=torch.int64, device=self.device)
        self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)

    def sort_(self, keys: torch.Tensor, key_type=None) -> torch.Tensor:
        """Sort in place.  ``key_type``: "u64" / "i64" / "f64"; default by dtype (int64 signed, float64 IEEE total order,
        uint64 unsigned)."""
        if not isinstance(keys, torch.Tensor) or not keys.is_cuda:
            raise TypeError("keys must be a CUDA tensor (this library has no CPU path)")
        if keys.dtype not in _KEY64_TYPES or keys.dim() != 1 or not keys.is_contiguous():
            raise TypeError("keys must be a contiguous 1-D int64/uint64/float64 tensor")
        if keys.numel() > self.max_n:
            raise ValueError("keys larger than this Sorter64's capacity")
        kt = _KEY64_TYPES[keys.dtype] if key_type is None else (_KEY64_NAMES[key_type] if isinstance(key_type, str) else int(key_type))
        N.check(
            N.lib().lsd_sort64(keys.data_ptr(), self.scratch.data_ptr(), keys.numel(), kt, self.workspace.data_ptr(),
                               self.workspace.numel(), _stream_ptr(keys.device)),
            "lsd_sort64",
        )
        return keys


def sort64_(keys: torch.Tensor, key_type=None) -> torch.Tensor:
    """Sort a 64-bit key tensor in place; allocates scratch for this one call."""
    return Sorter64(keys.numel(), device=keys.device if isinstance(keys, torch.Tensor) else None).sort_(keys, key_type)


class HostSorter:
    """Host-buffer entry: H2D copy, sort, D2H copy in one call (reference .cu:1001-1005)."""

    def __init__(self, max_n: int, r: int = 8, block: int = 0):
        self._ctx = C.c_void_p()
        N.check(N.lib().lsd_host_ctx_create(int(max_n), r, block, C.byref(self._ctx)), "lsd_host_ctx_create")
        self.max_n = int(max_n)

    @staticmethod
    def _host_ptr(host_keys):
        if isinstance(host_keys, torch.Tensor):
            if host_keys.is_cuda or host_keys.dtype not in (torch.int32, torch.uint32) or not host_keys.is_contiguous():
                raise TypeError("host_keys must be a contiguous CPU int32/uint32 tensor")
            return host_keys.data_ptr(), host_keys.numel()
        import numpy as np

        if host_keys.dtype != np.uint32 or not host_keys.flags["C_CONTIGUOUS"]:
            raise TypeError("host_keys must be a C-contiguous uint32 array")
        return host_keys.ctypes.data, host_keys.size

    def sort_(self, host_keys) -> None:
        """``host_keys``: pinned/pageable CPU torch tensor (int32/uint32) or numpy uint32 array, sorted in place."""
        ptr, n = self._host_ptr(host_keys)
        N.check(N.lib().lsd_sort_host(self._ctx, ptr, n), "lsd_sort_host")

    def sort_async_(self, host_keys) -> None:
        """Enqueue H2D + sort + D2H on the context's stream and return at once (``host_keys`` must be pinned and must stay
        alive and untouched until ``wait()``).  Two HostSorters used alternately overlap the D2H copy of one array with the
        H2D copy of the next."""
        ptr, n = self._host_ptr(host_keys)
        N.check(N.lib().lsd_sort_host_async(self._ctx, ptr, n), "lsd_sort_host_async")

    def wait(self) -> None:
        N.check(N.lib().lsd_host_ctx_wait(self._ctx), "lsd_host_ctx_wait")

    def close(self) -> None:
        if self._ctx:
            N.lib().lsd_host_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
