// onesweep_r8.cu -- kernel shapes for 8-bit digits (4 passes): the headline configuration.
// Entry 0 is the default; the others are reachable through `block` (threads per CTA, the
// reference's B) and lsd_sort_options.variant (tuning sweeps from bench_tools/).
// The entries live in onesweep_r8_{a,b,c}.cu (three translation units that compile in parallel);
// variant numbers are positions in their concatenation.
#include <vector>

#include "onesweep.cuh"

namespace lsd {

const OnesweepLauncher* onesweep_r8_part_a(int* count);
const OnesweepLauncher* onesweep_r8_part_b(int* count);
const OnesweepLauncher* onesweep_r8_part_c(int* count);
const OnesweepLauncher* onesweep_r8_part_d(int* count);

const OnesweepLauncher* onesweep_table_r8(int* count)
{
    static const std::vector<OnesweepLauncher> table = [] {  // thread-safe one-time concatenation
        std::vector<OnesweepLauncher> t;
        for (auto part : {&onesweep_r8_part_a, &onesweep_r8_part_b, &onesweep_r8_part_c, &onesweep_r8_part_d}) {
            int n = 0;
            const OnesweepLauncher* p = part(&n);
            t.insert(t.end(), p, p + n);
        }
        return t;
    }();
    *count = (int)table.size();
    return table.data();
}

}  // namespace lsd
