// onesweep_r8.cu -- kernel shapes for 8-bit digits (4 passes): the headline configuration.
// Entry 0 is the default for every `block` (the reference's threads-per-block knob is a hint here); the others are
// reachable through lsd_sort_options.variant.  The product build holds onesweep_r8_a.cu's first entries only; with
// -DLSD_TUNING_VARIANTS (make TUNING=1) the measured-and-rejected families of onesweep_r8_{a,b,c,d}.cu are appended
// (variant numbers are positions in the concatenation).
#include <vector>

#include "onesweep.cuh"

namespace lsd {

const OnesweepLauncher* onesweep_r8_part_a(int* count);
#ifdef LSD_TUNING_VARIANTS
const OnesweepLauncher* onesweep_r8_part_b(int* count);
const OnesweepLauncher* onesweep_r8_part_c(int* count);
const OnesweepLauncher* onesweep_r8_part_d(int* count);
#endif

const OnesweepLauncher* onesweep_table_r8(int* count)
{
    static const std::vector<OnesweepLauncher> table = [] {  // thread-safe one-time concatenation
        std::vector<OnesweepLauncher> t;
#ifdef LSD_TUNING_VARIANTS
        for (auto part : {&onesweep_r8_part_a, &onesweep_r8_part_b, &onesweep_r8_part_c, &onesweep_r8_part_d}) {
#else
        for (auto part : {&onesweep_r8_part_a}) {
#endif
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
