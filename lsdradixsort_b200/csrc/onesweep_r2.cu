// onesweep_r2.cu -- kernel shapes for 2-bit digits (16 passes).  Entry 0 is the default.
#include "onesweep_lpc3.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_lpc3_launcher<2, 9, 29, 3, 4, 0, 1, 2>(),  // 0: default for plain / typed-key sorts -- persistent LPC pass (run-time shift, dedicated prefetch buffer)
    make_launcher<2, 256, 16, kMatchBallot, true>(),  // 1: warp multisplit: the key-value forms for this radix
#ifdef LSD_TUNING_VARIANTS
    make_launcher<2, 128, 16, kMatchBallot, true>(),
    make_launcher<2, 512, 16, kMatchBallot, true>(),
    make_launcher<2, 1024, 8, kMatchBallot, true>(),
    make_lpc3_launcher<2, 9, 29, 3, 1, 10, 1, 0>(),       // 5: the default with the quad look-back (lookback_quad.cuh), 32 records per round
#endif
};

const OnesweepLauncher* onesweep_table_r2(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
