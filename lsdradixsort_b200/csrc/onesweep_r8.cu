// onesweep_r8.cu -- kernel shapes for 8-bit digits (4 passes): the headline configuration.
// Entry 0 is the default; the others are reachable through `block` (threads per CTA, the
// reference's B) and lsd_sort_options.variant (tuning sweeps from bench_tools/).
#include "onesweep_lpc32.cuh"
#include "onesweep_lpc2.cuh"
#include "onesweep_lpc3.cuh"
#include "onesweep_lpcp.cuh"
#include "onesweep_cpc.cuh"
#include "onesweep_cpcp.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1, true>(),  // 0: default -- persistent LPC32 pass, next tile prefetched into the dead counter matrix, ticket hand-over by mbarrier (= variant 75); peer-scatter / key-value / typed-key passes on onesweep_lpc32_kernel (= variant 68)
    make_launcher<8, 128, 24, kMatchBallot>(),   // 1
    make_launcher<8, 256, 24, kMatchBallot, true>(),   // 2
    make_launcher<8, 1024, 8, kMatchBallot>(),   // 3
    make_launcher<8, 256, 16, kMatchBallot>(),   // 4
    make_launcher<8, 512, 16, kMatchHw>(),       // 5: match.any instead of 8 ballots
    make_launcher<8, 256, 24, kMatchHw>(),       // 6
    make_launcher<8, 384, 20, kMatchBallot>(),   // 7
    make_lpc_launcher<8, 9, 29, 3>(),            // 8: lane-private counters, 288 threads, tile 8352, 3 CTAs/SM
    make_lpc_launcher<8, 7, 37, 3>(),            // 9: 224 threads, tile 8288
    make_lpc_launcher<8, 11, 23, 3>(),           // 10: 352 threads, tile 8096
    make_lpc_launcher<8, 9, 29, 2>(),            // 11: as 8 with 2 CTAs/SM register budget
    make_lpc_launcher<8, 13, 19, 2>(),           // 12: 416 threads, tile 7904
    make_lpc_launcher<8, 9, 15, 4>(),            // 13: 288 threads, tile 4320
    make_lpcp_launcher<8, 9, 29, 2>(),           // 14: persistent pipelined, 9 worker + 4 look-back warps, tile 8352
    make_lpcp_launcher<8, 7, 37, 2>(),           // 15: 7 worker warps, tile 8288
    make_lpcp_launcher<8, 11, 23, 2>(),          // 16: 11 worker warps, tile 8096
    make_lpcp_launcher<8, 5, 51, 2>(),           // 17: 5 worker warps, tile 8160
    make_lpcp_launcher<8, 9, 15, 3>(),           // 18: tile 4320, 3 CTAs/SM
    make_lpc32_launcher<8, 9, 29, 3>(),          // 19: 32-bit byte-offset counters, compile-time shift, tile 8352
    make_lpc32_launcher<8, 11, 23, 3>(),         // 20: 352 threads, tile 8096
    make_lpc32_launcher<8, 13, 19, 3>(),         // 21: 416 threads, tile 7904
    make_lpc32_launcher<8, 9, 29, 2>(),          // 22: as 19 with the 2-CTA register budget
    make_lpc32_launcher<8, 9, 17, 4>(),          // 23: tile 4896, 4 CTAs/SM
    make_lpc_launcher<8, 9, 29, 4>(),            // 24: packed counters, 56-register budget, 4 CTAs/SM
    make_lpc32_launcher<8, 9, 21, 3>(),          // 25: tile 6048
    make_lpc_launcher<8, 9, 21, 4>(),            // 26: packed, tile 6048, 4 CTAs/SM
    make_lpc_launcher<8, 9, 23, 4>(),            // 27: packed, tile 6624, 4 CTAs/SM
    make_launcher<8, 512, 16, kMatchBallot, true>(),   // 28: warp-multisplit (ballot) kernel, the round-1 v1 default
    make_lpc32_launcher<8, 9, 29, 3, 16>(),      // 29: look-back window 16
    make_lpc32_launcher<8, 9, 29, 3, 4>(),       // 30: look-back window 4
    make_lpc32_launcher<8, 9, 29, 3, 2>(),       // 31: look-back window 2
    make_cpc_launcher<8, 64, 3, 8>(),            // 32: column-private counters, 128 threads, tile 8192, 3 CTAs/SM
    make_cpc_launcher<8, 64, 3, 16>(),           // 33: look-back window 16
    make_cpc_launcher<8, 64, 3, 4>(),            // 34: look-back window 4
    make_cpc_launcher<8, 48, 3, 8>(),            // 35: tile 6144
    make_cpc_launcher<8, 32, 4, 8>(),            // 36: tile 4096, 4 CTAs/SM
    make_cpc_launcher<8, 64, 3, 0>(),            // 37: TIMING EXPERIMENT, no look-back (output wrong)
    make_cpc_launcher<8, 64, 3, 32>(),           // 38: look-back window 32
    make_cpc_launcher<8, 64, 3, 0, 2>(),         // 39: TIMING: no look-back, full-line stores
    make_cpc_launcher<8, 64, 3, 0, 4>(),         // 40: TIMING: no look-back, no global stores
    make_cpc_launcher<8, 64, 3, 0, 8>(),         // 41: TIMING: no look-back, conflict-free smem scatter
    make_cpc_launcher<8, 64, 3, 0, 10>(),        // 42: TIMING: no look-back, conflict-free scatter, full-line stores
    make_cpc_launcher<8, 64, 3, 0, 12>(),        // 43: TIMING: no look-back, conflict-free scatter, no stores
    make_cpcp_launcher<8, 64, 4, 8, 152>(),      // 44: persistent pipeline, 4 buffers, front groups at 152 registers
    make_cpcp_launcher<8, 64, 4, 8, 0>(),        // 45: same without register reallocation
    make_cpcp_launcher<8, 64, 4, 4, 152>(),      // 46: look-back window 4
    make_cpcp_launcher<8, 48, 5, 8, 0>(),        // 47: tile 6144, 5 buffers, no register reallocation
    make_lpc32_launcher<8, 9, 29, 3, 4, 1>(),    // 48: as 0 with the matrix zero-filled by st.bulk
    make_lpc32_launcher<8, 9, 29, 3, 4, 0, false, true>(),   // 49: single-pass matrix scan (rows kept in registers)
    make_lpc32_launcher<8, 9, 29, 3, 8, 0, false, true>(),   // 50: same, look-back window 8
    make_lpc32_launcher<8, 9, 31, 3, 4, 0, false, true>(),   // 51: single-pass scan, tile 8928
    make_cpc_launcher<8, 64, 3, 4, 16>(),        // 52: CPC, look-back window 4, 32-bit Q rows
    make_cpc_launcher<8, 64, 3, 4, 48>(),        // 53: CPC, window 4, 32-bit Q rows, skewed reorder layout
    make_cpc_launcher<8, 64, 3, 4, 32>(),        // 54: CPC, window 4, skewed reorder layout
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
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1, false, true>(),  // 78: as 75 with the per-tile phase trace compiled in (bench_tools/trace.py --variant 78)
};

const OnesweepLauncher* onesweep_table_r8(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
