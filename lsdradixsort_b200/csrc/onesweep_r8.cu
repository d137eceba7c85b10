// onesweep_r8.cu -- kernel shapes for 8-bit digits (4 passes): the headline configuration.
// Entry 0 is the default; the others are reachable through `block` (threads per CTA, the
// reference's B) and lsd_sort_options.variant (tuning sweeps from bench_tools/).
#include "onesweep_lpc.cuh"

namespace lsd {

static const OnesweepLauncher kTable[] = {
    make_launcher<8, 512, 16, kMatchBallot>(),   // 0: default
    make_launcher<8, 128, 24, kMatchBallot>(),   // 1
    make_launcher<8, 256, 24, kMatchBallot>(),   // 2
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
};

const OnesweepLauncher* onesweep_table_r8(int* count)
{
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}

}  // namespace lsd
