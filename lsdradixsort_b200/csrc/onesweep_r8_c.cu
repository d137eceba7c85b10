// onesweep_r8_c.cu -- 8-bit-digit kernel shapes, part C of the table assembled in onesweep_r8.cu
// (variants 55-: two-chain / cluster look-back (lpc2), store-operator and copy-out experiments, persistent prefetching kernel (lpc3)).
// The table is split over three translation units only so that they compile in parallel.
#include "onesweep_lpc32.cuh"
#include "onesweep_lpc2.cuh"
#include "onesweep_lpc3.cuh"

namespace lsd {

static const OnesweepLauncher kPart[] = {
    make_lpc2_launcher<8, 9, 29, 3, 4, 1>(),     // 55: two rank chains (packed half-word counters), no cluster
    make_lpc2_launcher<8, 9, 29, 3, 4, 2>(),     // 56: two chains + one look-back record per cluster of 2 CTAs
    make_lpc2_launcher<8, 9, 29, 3, 4, 4>(),     // 57: ... per cluster of 4
    make_lpc2_launcher<8, 9, 29, 3, 4, 8>(),     // 58: ... per cluster of 8
    make_lpc2_launcher<8, 9, 29, 3, 2, 4>(),     // 59: cluster of 4, look-back window 2
    make_lpc2_launcher<8, 9, 29, 3, 8, 4>(),     // 60: cluster of 4, look-back window 8
    make_lpc2_launcher<8, 9, 29, 3, 4, 1, false, 1>(),   // 61: two chains, no cluster, ld.global.cg polling
    make_lpc2_launcher<8, 9, 29, 3, 8, 1, false, 1>(),   // 62: ... window 8
    make_lpc2_launcher<8, 9, 29, 3, 16, 1, false, 1>(),  // 63: ... window 16
    make_lpc2_launcher<8, 9, 29, 3, 8, 1>(),             // 64: two chains, strong polling, window 8
    make_lpc32_launcher<8, 9, 29, 3, 4, 2>(),            // 65: as 0, keys stored with st.global.cg
    make_lpc32_launcher<8, 9, 29, 3, 4, 3>(),            // 66: ... st.global.cs
    make_lpc32_launcher<8, 9, 29, 3, 4, 4>(),            // 67: ... st.global.wt
    make_lpc32_launcher<8, 9, 29, 3, 4, 5>(),            // 68: ... st.global.L1::no_allocate
    make_lpc32_launcher<8, 9, 29, 3, 4, 6>(),            // 69: copy-out one bucket run per warp, lanes aligned to destination lines
    make_lpc3_launcher<8, 9, 29, 3, 4>(),                // 70: persistent LPC32, next tile prefetched into the dead counter matrix
    make_lpc3_launcher<8, 9, 29, 3, 8>(),                // 71: ... look-back window 8
    make_lpc3_launcher<8, 9, 29, 3, 4, 1>(),             // 72: ... matrix zero-filled by st.bulk
    make_lpc3_launcher<8, 11, 23, 3, 4>(),               // 73: ... 352 threads, tile 8096
    make_lpc3_launcher<8, 9, 29, 3, 2>(),                // 74: ... look-back window 2
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1>(),          // 75: as 70, ticket handed over through an mbarrier (no end-of-tile barrier)
    make_lpc3_launcher<8, 9, 29, 3, 4, 2, 0>(),          // 76: as 70, matrix zero-filled by a TMA copy of a zero page
    make_lpc3_launcher<8, 9, 29, 3, 4, 2, 1>(),          // 77: both
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1, 0, true>(),  // 78: as 75 with the per-tile phase trace compiled in (bench_tools/trace.py --variant 78)
};

const OnesweepLauncher* onesweep_r8_part_c(int* count)
{
    *count = (int)(sizeof(kPart) / sizeof(kPart[0]));
    return kPart;
}

}  // namespace lsd
