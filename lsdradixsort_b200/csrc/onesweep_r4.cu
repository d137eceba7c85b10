// onesweep_r4.cu -- kernel shapes for 4-bit digits (8 passes).  Entry 0 is the default.
#include "onesweep_lpc3.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_lpc3_launcher<4, 9, 29, 3, 2, 10, 1, 2>(),  // 0: default for plain / typed-key sorts -- persistent LPC pass (run-time shift, dedicated prefetch buffer), quad look-back with 16 records per round (lookback_quad.cuh: 0.543 -> 0.516 ms per pass at 2^28)
    make_launcher<4, 256, 16, kMatchBallot, true>(),  // 1: warp multisplit: the key-value forms for this radix
#ifdef LSD_TUNING_VARIANTS
    make_launcher<4, 128, 16, kMatchBallot, true>(),
    make_launcher<4, 512, 16, kMatchBallot, true>(),
    make_launcher<4, 1024, 8, kMatchBallot, true>(),
    make_lpc3_launcher<4, 9, 29, 3, 4, 0, 1, 0, true>(),  // 5: the default's shape with the per-tile phase trace
    make_lpc3_launcher<4, 9, 29, 3, 8, 0, 1, 0>(),        // 6: look-back window 8
    make_lpc3_launcher<4, 9, 29, 3, 16, 0, 1, 0>(),       // 7: look-back window 16
    make_lpc3_launcher<4, 9, 23, 4, 4, 0, 1, 0>(),        // 8: 6624-key tiles, four CTAs per SM
    make_lpc3_launcher<4, 9, 23, 4, 8, 0, 1, 0>(),        // 9: same, window 8
    make_lpc3_launcher<4, 9, 29, 3, 2, 0, 1, 0>(),        // 10: window 2
    make_lpc3_launcher<4, 9, 29, 3, 1, 10, 1, 0>(),       // 11: quad look-back (lookback_quad.cuh), 8 records per round
    make_lpc3_launcher<4, 9, 29, 3, 2, 10, 1, 0>(),       // 12: 16 records per round (= the default)
    make_lpc3_launcher<4, 9, 29, 3, 4, 10, 1, 0>(),       // 13: 32 records per round
    make_lpc3_launcher<4, 9, 29, 3, 2, 10, 1, 0, true>(), // 14: 12 with the per-tile phase trace
    make_lpc3_launcher<4, 9, 29, 3, 4, 0, 1, 0>(),        // 15: the round-1 default: digit-pair look-back, window 4
    make_lpc3_launcher<4, 9, 23, 4, 2, 10, 1, 0>(),       // 16: 6624-key tiles, four CTAs per SM, quad look-back: 0.526 ms per pass against 0.530
    make_lpc3_launcher<4, 9, 21, 4, 2, 10, 1, 0>(),       // 17: 6048-key tiles, four CTAs per SM, quad look-back: 0.560
#endif
};

const OnesweepLauncher* onesweep_table_r4(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
