// onesweep_r4.cu -- kernel shapes for 4-bit digits (8 passes).  Entry 0 is the default.
#include "onesweep_lpc3.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_lpc3_launcher<4, 9, 29, 3, 4, 0, 1, 2>(),  // 0: default for plain / typed-key sorts -- persistent LPC pass (run-time shift, dedicated prefetch buffer)
    make_launcher<4, 256, 16, kMatchBallot, true>(),  // 1: warp multisplit: the key-value forms for this radix
#ifdef LSD_TUNING_VARIANTS
    make_launcher<4, 128, 16, kMatchBallot, true>(),
    make_launcher<4, 512, 16, kMatchBallot, true>(),
    make_launcher<4, 1024, 8, kMatchBallot, true>(),
#endif
};

const OnesweepLauncher* onesweep_table_r4(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
