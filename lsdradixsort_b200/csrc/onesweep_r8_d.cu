// onesweep_r8_d.cu -- 8-bit-digit kernel shapes, part D of the table assembled in onesweep_r8.cu:
// the round-2 "wide" pass (onesweep_wide.cuh: 16 Ki-key tiles, two rank chains, dedicated look-back warps) and the
// shapes that were measured around it (profiles/r02_wide_pass_study.txt).  Tuning variants: the default stays entry 0.
#include "onesweep_wide.cuh"
#include "onesweep_lpc4.cuh"

namespace lsd {

// make_wide_launcher<RB, rank warps, keys per thread, CTAs/SM, look-back window, COPY, TRACE, look-back warps>
// COPY bit 0: position-linear copy-out instead of the bucket-walk; bit 1: scatter interleaved with the chain atomics
static const OnesweepLauncher kPart[] = {
    make_wide_launcher<8, 13, 39, 2, 8, 3, false, 2>(),   // D0: best wide shape: tile 16224, 2 look-back warps, window 8 (0.648 ms/pass)
    make_wide_launcher<8, 13, 39, 2, 8, 3, true, 2>(),    // D1: D0 with the per-tile phase trace
    make_wide_launcher<8, 13, 39, 2, 8, 2, false, 2>(),   // D2: D0 with the bucket-walk (line-aligned) copy-out (0.688)
    make_wide_launcher<8, 15, 35, 2, 4, 3>(),             // D3: 15 rank warps + one look-back warp, tile 16800 (0.723)
    make_wide_launcher<8, 15, 35, 2, 4, 1>(),             // D4: D3 with packed rank registers (0.77)
    make_wide_launcher<8, 9, 29, 3, 4, 3>(),              // D5: the round-1 tile (8352 keys, 3 CTAs/SM) in this kernel (1.05: look-back bound)
    make_wide_launcher<8, 11, 47, 2, 16, 3, false, 4>(),  // D6: 4 look-back warps, window 16 (0.79)
    // quad look-back (lookback_quad.cuh: 128-bit record accesses, window spread over lanes) on the default and on the two-chain pass
    make_lpc3_launcher<8, 9, 29, 3, 4, 10, 1>(),              // E0: default + quad look-back, 8 records per round
    make_lpc3_launcher<8, 9, 29, 3, 2, 10, 1>(),              // E1: 4 records per round with two load instructions
    make_lpc3_launcher<8, 9, 29, 3, 3, 10, 1>(),              // E2: 6 records per round
    make_lpc4_launcher<8, 9, 29, 3, 4, 10, 1>(),              // E3: two rank chains + quad look-back, 8 records per round
    make_lpc4_launcher<8, 9, 29, 3, 2, 10, 1>(),              // E4: two rank chains, 4 records per round
    make_lpc3_launcher<8, 9, 29, 3, 4, 10, 1, 0, true>(),     // E5: E0 with the per-tile phase trace
    make_lpc4_launcher<8, 9, 29, 3, 4, 10, 1, 0, true>(),     // E6: E3 with the per-tile phase trace
};

const OnesweepLauncher* onesweep_r8_part_d(int* count)
{
    *count = (int)(sizeof(kPart) / sizeof(kPart[0]));
    return kPart;
}

}  // namespace lsd
