// onesweep_r8_b.cu -- 8-bit-digit kernel shapes, part B of the table assembled in onesweep_r8.cu
// (variants 29-54: LPC32 look-back windows and scan forms, CPC and persistent CPC shapes, timing experiments).
// The table is split over three translation units only so that they compile in parallel.
#include "onesweep_lpc32.cuh"
#include "onesweep_cpc.cuh"
#include "onesweep_cpcp.cuh"

namespace lsd {

static const OnesweepLauncher kPart[] = {
    make_lpc32_launcher<8, 9, 29, 3, 16>(),      // 29: look-back window 16
    make_lpc32_launcher<8, 9, 29, 3, 4>(),       // 30: look-back window 4
    make_lpc32_launcher<8, 9, 29, 3, 2>(),       // 31: look-back window 2
    make_cpc_launcher<8, 64, 3, 8>(),            // 32: column-private counters, 128 threads, tile 8192, 3 CTAs/SM
    make_cpc_launcher<8, 64, 3, 16>(),           // 33: look-back window 16
    make_cpc_launcher<8, 64, 3, 4>(),            // 34: look-back window 4
    make_cpc_launcher<8, 48, 3, 8>(),            // 35: tile 6144
    make_cpc_launcher<8, 32, 4, 8>(),            // 36: tile 4096, 4 CTAs/SM
    make_cpc_launcher<8, 64, 3, 32>(),           // 38: look-back window 32
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
};

const OnesweepLauncher* onesweep_r8_part_b(int* count)
{
    *count = (int)(sizeof(kPart) / sizeof(kPart[0]));
    return kPart;
}

}  // namespace lsd
