// sort.h -- internal interface between sort.cu and api.cu.
#pragma once
#include "common.cuh"

namespace lsd {

struct OnesweepLauncher;

struct SortLayout {
    const OnesweepLauncher* k;
    int passes;
    int H;
    uint32_t portion_keys;
    uint64_t portions;
    uint64_t total_tiles;
    size_t off_plan, off_hist, off_bases, off_tickets, off_lookback, total_bytes;
};

int make_layout(uint64_t n, int r, int block, const lsd_sort_options* opt, SortLayout* L, bool pairs = false);

// Enqueue the whole sort on `s`.  With vals / vals_scratch non-null every key carries a 32-bit value (key-value sort).  If `events` is non-null it must hold passes + 3 events; they are
// recorded before the histogram, after the plan, after every pass and after the copy-back.
int sort_enqueue(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                 const lsd_sort_options* opt, cudaStream_t s, cudaEvent_t* events, int* launches,
                 uint32_t* vals = nullptr, uint32_t* vals_scratch = nullptr);

// One stable pass on `bit_group` from `in` to `out` (no plan, never skipped).  With dst_ptrs != nullptr the pass
// runs in peer-scatter mode instead (see lsd_sort_pass_scatter in include/lsdsort.h); `out` and `hist_out` are unused.
// abort_flag (device, optional): the pass exits at once if *abort_flag != 0 when its plan kernel runs.
int pass_enqueue(const uint32_t* in, uint32_t* out, uint64_t n, int r, int bit_group, int block, void* ws,
                 size_t ws_bytes, uint64_t* hist_out, cudaStream_t s, const uint64_t* dst_ptrs = nullptr,
                 const uint32_t* dst_seg = nullptr, const uint32_t* abort_flag = nullptr, int peer_shift = -1);
// peer_shift >= 0 (peer-scatter mode only): the digit is bits [peer_shift, peer_shift + r) instead of digit `bit_group`

// Composite digit widths (r in 3..16 other than 4, 8): one pass as sub-passes of the 8-bit kernel; per-digit histograms.
size_t wide_pass_workspace_bytes(uint64_t n);
int pass_enqueue_wide(const uint32_t* in, uint32_t* out, uint64_t n, int r, int bit_group, void* ws, size_t ws_bytes,
                      uint64_t* hist_out, cudaStream_t s);
int launch_digit_histograms_wide(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s);

// 64-bit keys (keys64.cu): split into word arrays, two key-value sorts, merge.
size_t sort64_workspace_bytes(uint64_t n);
int sort64_enqueue(uint64_t* keys, uint64_t* scratch, uint64_t n, uint32_t key_type, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace lsd
