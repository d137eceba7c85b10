// onesweep_r1.cu -- kernel shapes for 1-bit digits (32 passes).  Entry 0 is the default.
#include "onesweep_lpc3.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_lpc3_launcher<1, 9, 29, 3, 1, 10, 1, 2>(),  // 0: default for plain / typed-key sorts -- persistent LPC pass (run-time shift, dedicated prefetch buffer), look-back with 32 records per round (lookback_quad.cuh: 0.757 -> 0.737 ms per pass at 2^28)
    make_launcher<1, 256, 16, kMatchBallot, true>(),  // 1: warp multisplit: the key-value forms for this radix
#ifdef LSD_TUNING_VARIANTS
    make_launcher<1, 128, 16, kMatchBallot, true>(),
    make_launcher<1, 512, 16, kMatchBallot, true>(),
    make_launcher<1, 1024, 8, kMatchBallot, true>(),
    make_lpc3_launcher<1, 9, 29, 3, 4, 0, 1, 0>(),        // 5: the round-1 default: digit-pair look-back, window 4
#endif
};

const OnesweepLauncher* onesweep_table_r1(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
