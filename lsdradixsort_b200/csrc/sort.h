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

int make_layout(uint64_t n, int r, int block, const lsd_sort_options* opt, SortLayout* L);

// Enqueue the whole sort on `s`.  If `events` is non-null it must hold passes + 3 events; they are
// recorded before the histogram, after the plan, after every pass and after the copy-back.
int sort_enqueue(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                 const lsd_sort_options* opt, cudaStream_t s, cudaEvent_t* events, int* launches);

}  // namespace lsd
