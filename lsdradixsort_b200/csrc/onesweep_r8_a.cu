// onesweep_r8_a.cu -- 8-bit-digit kernel shapes, part A of the table assembled in onesweep_r8.cu.
// The product build holds entries 0-2 only: the default, the north_star's warp-multisplit pass (kept as the reference
// point of profiles/r01_microbench_rank_primitives.txt) and the non-persistent LPC32 pass.  Everything that was
// measured and rejected is compiled with -DLSD_TUNING_VARIANTS (make TUNING=1) only.
#include "onesweep_lpc32.cuh"
#include "onesweep_lpc3.cuh"
#ifdef LSD_TUNING_VARIANTS
#include "onesweep_lpc4.cuh"
#include "onesweep_lpcp.cuh"
#endif

namespace lsd {

static const OnesweepLauncher kPart[] = {
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1, 1>(),  // 0: default -- persistent LPC32 pass, next tile prefetched into the dead counter matrix, ticket hand-over by mbarrier (= variant 75); peer-scatter / key-value / typed-key passes on onesweep_lpc32_kernel (= variant 68)
    make_launcher<8, 256, 16, kMatchBallot, true>(),   // 1: warp multisplit (8 ballots + popc), every form but peer-scatter
    make_lpc32_launcher<8, 9, 29, 3>(),          // 2: non-persistent LPC32 pass, plain keys only
#ifdef LSD_TUNING_VARIANTS
    // round-2 look-back experiments on the default (profiles/r02_lookback_experiments.txt); A0..A5 = variants 3..8
    make_lpc3_launcher<8, 9, 29, 3, 4, 9, 1>(),  // A0: 0 with the pipelined look-back walk (next window in flight, whole-window fast path)
    make_lpc3_launcher<8, 9, 29, 3, 4, 9, 1, 0, true>(),  // A1: A0 with the per-tile phase trace
    make_lpc3_launcher<8, 9, 29, 3, 4, 0, 1, 0, true>(),  // A2: 0 with the per-tile phase trace
    make_lpc3_launcher<8, 9, 29, 3, 4, 8, 1>(),  // A3: 0 with the look-back records published by TMA bulk stores
    make_lpc3_launcher<8, 9, 29, 3, 4, 8, 1, 0, true>(),  // A4: A3 with the per-tile phase trace
    make_lpc4_launcher<8, 9, 29, 3, 8, 0, 1>(),  // 3: persistent pass with two rank chains, look-back window 8 (round 2: no gain, see profiles/r02_wide_pass_study.txt)
    make_lpc4_launcher<8, 9, 29, 3, 4, 0, 1>(),  // 4: two rank chains, window 4
    make_lpc4_launcher<8, 9, 29, 3, 8, 0, 1, 0, true>(),  // 5: 3 with the per-tile phase trace
    make_launcher<8, 128, 24, kMatchBallot>(),   // 3
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
    make_launcher<8, 512, 16, kMatchBallot, true>(),   // warp-multisplit (ballot) kernel, the round-1 v1 default
#endif
};

const OnesweepLauncher* onesweep_r8_part_a(int* count)
{
    *count = (int)(sizeof(kPart) / sizeof(kPart[0]));
    return kPart;
}

}  // namespace lsd
