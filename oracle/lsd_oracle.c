/*
 * lsd_oracle.c -- CPU restatement of the reference's LSD radix sort hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA product in
 * lsdradixsort_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The product
 * path never links it and has no CPU fallback.
 *
 * Parity status: PINNED.  oracle/Makefile compiles the unmodified reference
 * sources (where they lie under /root/reference) into oracle/_ref/libref_lsd.so;
 * tests/golden/make_golden.py ran that library here and committed its outputs
 * under tests/golden/, and tests/test_oracle.py checks every function below
 * against those fixtures (and live against the library when it is present).
 *
 * Each function names the reference lines whose behaviour it restates
 * (paths relative to /root/reference/LSDRadixSort/).  The code is written
 * independently: same algorithm and observable results, not the same text.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* Utils.h:22 GET_R_BITS(n, r, i): digit i (0 = least significant) of width r. */
static inline uint32_t oracle_digit(uint32_t key, int r, int bit_group)
{
    const uint32_t mask = (r >= 32) ? 0xFFFFFFFFu : ((1u << r) - 1u);
    return (key >> (bit_group * r)) & mask;
}

ORACLE_API uint32_t lsd_oracle_digit(uint32_t key, int r, int bit_group)
{
    return oracle_digit(key, r, bit_group);
}

/*
 * LSDRadixSort.cu:25-54 LSDRadixSortPass.
 * One stable counting-sort pass on digit `bit_group`: count, turn the counts
 * into bucket ends (inclusive running sum), then walk the input backwards and
 * place every key at --end[digit].  `histogram` (2^r words) is scratch and is
 * left holding the bucket START offsets, exactly like the reference (each end
 * has been decremented once per key).  The result is written to `out` and then
 * copied over `in` (reference :53), so both arrays hold it afterwards.
 */
ORACLE_API void lsd_oracle_sort_pass(uint32_t *in, uint32_t *out, int64_t count,
                                     uint32_t *histogram, int r, int bit_group)
{
    const size_t buckets = (size_t)1 << r;
    memset(histogram, 0, buckets * sizeof(uint32_t));

    for (int64_t i = 0; i < count; ++i)
        histogram[oracle_digit(in[i], r, bit_group)] += 1u;

    uint32_t running = 0;
    for (size_t b = 0; b < buckets; ++b) {
        running += histogram[b];
        histogram[b] = running; /* one past the last slot of bucket b */
    }

    for (int64_t i = count; i-- > 0;) {
        const uint32_t key = in[i];
        const uint32_t slot = --histogram[oracle_digit(key, r, bit_group)];
        out[slot] = key;
    }

    if (count > 0)
        memcpy(in, out, (size_t)count * sizeof(uint32_t));
}

/*
 * LSDRadixSort.cu:62-69 LSDRadixSort.
 * 32/r passes, least significant digit first.  Post-condition (reference
 * behaviour): `in` and `out` both hold the ascending sequence.
 * r must divide 32 and be < 32 (reference comment :60).
 */
ORACLE_API int lsd_oracle_sort(uint32_t *in, uint32_t *out, int64_t count,
                               uint32_t *histogram, int r)
{
    if (r <= 0 || r >= 32 || (32 % r) != 0)
        return -1;
    const int passes = 32 / r;
    for (int g = 0; g < passes; ++g)
        lsd_oracle_sort_pass(in, out, count, histogram, r, g);
    return 0;
}

/*
 * Key-value form of LSDRadixSortPass / LSDRadixSort (LSDRadixSort.cu:25-54, :62-69).
 * The reference moves keys only (`out[--histogram[digit]] = in[i]`, :45-50); its passes are stable by
 * construction (backwards walk over inclusive bucket ends).  Carrying a 32-bit value through the SAME
 * slot assignment makes that stability observable: this is the reference's pass with one more array
 * written at the identical index, nothing else changed.  Checker for lsd_sort_pairs; the keys it
 * produces are checked against lsd_oracle_sort / the compiled reference, the values against a stable
 * argsort (tests/test_oracle.py).
 */
ORACLE_API void lsd_oracle_sort_pairs_pass(uint32_t *in, uint32_t *vin, uint32_t *out, uint32_t *vout,
                                           int64_t count, uint32_t *histogram, int r, int bit_group)
{
    const size_t buckets = (size_t)1 << r;
    memset(histogram, 0, buckets * sizeof(uint32_t));
    for (int64_t i = 0; i < count; ++i)
        histogram[oracle_digit(in[i], r, bit_group)] += 1u;
    uint32_t running = 0;
    for (size_t b = 0; b < buckets; ++b) {
        running += histogram[b];
        histogram[b] = running;
    }
    for (int64_t i = count; i-- > 0;) {
        const uint32_t key = in[i];
        const uint32_t slot = --histogram[oracle_digit(key, r, bit_group)];
        out[slot] = key;
        vout[slot] = vin[i];
    }
    if (count > 0) {
        memcpy(in, out, (size_t)count * sizeof(uint32_t));
        memcpy(vin, vout, (size_t)count * sizeof(uint32_t));
    }
}

ORACLE_API int lsd_oracle_sort_pairs(uint32_t *in, uint32_t *vin, uint32_t *out, uint32_t *vout,
                                     int64_t count, uint32_t *histogram, int r)
{
    if (r <= 0 || r >= 32 || (32 % r) != 0)
        return -1;
    for (int g = 0; g < 32 / r; ++g)
        lsd_oracle_sort_pairs_pass(in, vin, out, vout, count, histogram, r, g);
    return 0;
}

/*
 * LSDRadixSort.cu:128-139 PrefixSum.
 * In-place EXCLUSIVE scan with uint32 wrap-around: a[i] <- sum of the original
 * a[0..i) mod 2^32, a[0] <- 0.  (The reference does an inclusive sweep then a
 * shift; a single carried sum gives the same array.)
 */
ORACLE_API void lsd_oracle_prefix_sum(uint32_t *a, int64_t count)
{
    uint32_t carry = 0;
    for (int64_t i = 0; i < count; ++i) {
        const uint32_t v = a[i];
        a[i] = carry;
        carry += v;
    }
}

/*
 * LSDRadixSort.cu:643-658 BuildHistogramsCPU (CPU twin of the kernel at
 * :660-702).  Tile-major per-tile histograms: h[g*2^r + d] += #keys of tile g
 * (keys [g*block, (g+1)*block)) whose digit `bit_group` equals d.  Like the
 * reference it ACCUMULATES into caller-zeroed memory.  The reference assumes
 * count == grid*block; here a ragged last tile simply stops at `count`, which
 * is what the GPU kernel does (`if (idx < count)`, :684).
 */
ORACLE_API void lsd_oracle_build_histograms(const uint32_t *a, uint32_t *h, int64_t count,
                                            int r, int bit_group, int64_t grid, int block)
{
    const size_t buckets = (size_t)1 << r;
    for (int64_t g = 0; g < grid; ++g) {
        uint32_t *row = h + (size_t)g * buckets;
        const int64_t lo = g * (int64_t)block;
        int64_t hi = lo + block;
        if (hi > count)
            hi = count;
        for (int64_t i = lo; i < hi; ++i)
            row[oracle_digit(a[i], r, bit_group)] += 1u;
    }
}

/*
 * LSDRadixSort.cu:265-276 GetGPUPrefixSumBlockSumsCount.
 * Scratch words the reference's recursive scan needs: one word per chunk at
 * every level (count / tpb, truncating), plus one.
 */
ORACLE_API int64_t lsd_oracle_block_sums_count(int64_t count, int threads_per_block)
{
    int64_t total = 0;
    while (count > threads_per_block) {
        const int64_t chunks = count / threads_per_block;
        total += chunks;
        count = chunks;
    }
    return total + 1;
}

/*
 * Whole-array digit histograms for every bit group: hist[g*2^r + d].  Not a
 * reference function by itself: it is the column sum of BuildHistogramsCPU's
 * output (:643-658) taken for each bit group, i.e. what LSDRadixSortPass counts
 * at :30-35 when the digit is examined on the ORIGINAL key order.  Used to
 * check the library's lsd_digit_histograms (64-bit counters).
 */
ORACLE_API void lsd_oracle_digit_histograms(const uint32_t *a, int64_t count, int r, uint64_t *hist)
{
    const int passes = 32 / r;
    const size_t buckets = (size_t)1 << r;
    memset(hist, 0, (size_t)passes * buckets * sizeof(uint64_t));
    for (int64_t i = 0; i < count; ++i)
        for (int g = 0; g < passes; ++g)
            hist[(size_t)g * buckets + oracle_digit(a[i], r, g)] += 1u;
}

/*
 * Reference-shaped GPU flow on the CPU, for documentation and cross-checks:
 * LSDRadixSort.cu:839-910 GPULSDRadixSort computes, per pass, per-tile
 * histograms (:850), per-tile exclusive scans = local offsets (:869), a
 * digit-major exclusive scan over all tiles = global offsets (:885-894), then
 * scatters key t of tile g to  t_sorted - local[g][d] + global[g][d]  after a
 * stable in-tile sort (:829-836).  The net effect of one pass is a stable
 * counting sort; this function reproduces it tile by tile so tests can check
 * that the tile decomposition is equivalent to lsd_oracle_sort_pass.
 */
ORACLE_API int lsd_oracle_tiled_pass(const uint32_t *in, uint32_t *out, int64_t count,
                                     int r, int bit_group, int block)
{
    const size_t buckets = (size_t)1 << r;
    const int64_t grid = (count + block - 1) / block;
    uint32_t *tile_hist = (uint32_t *)calloc((size_t)(grid > 0 ? grid : 1) * buckets, sizeof(uint32_t));
    uint32_t *global_off = (uint32_t *)calloc((size_t)(grid > 0 ? grid : 1) * buckets, sizeof(uint32_t));
    if (!tile_hist || !global_off) {
        free(tile_hist);
        free(global_off);
        return -1;
    }
    lsd_oracle_build_histograms(in, tile_hist, count, r, bit_group, grid, block);

    /* digit-major exclusive scan over tiles (the transposed scan of :874-895) */
    uint32_t carry = 0;
    for (size_t d = 0; d < buckets; ++d)
        for (int64_t g = 0; g < grid; ++g) {
            global_off[(size_t)g * buckets + d] = carry;
            carry += tile_hist[(size_t)g * buckets + d];
        }

    for (int64_t g = 0; g < grid; ++g) {
        uint32_t *next = global_off + (size_t)g * buckets; /* next free slot per digit */
        const int64_t lo = g * (int64_t)block;
        int64_t hi = lo + block;
        if (hi > count)
            hi = count;
        for (int64_t i = lo; i < hi; ++i) {
            const uint32_t key = in[i];
            out[next[oracle_digit(key, r, bit_group)]++] = key;
        }
    }
    free(tile_hist);
    free(global_off);
    return 0;
}

/*
 * LSDRadixSortPass (LSDRadixSort.cu:25-54) on an arbitrary bit field [shift, shift + width) of the key instead
 * of GET_R_BITS(val, r, bit_group): the digit of the composite widths (r = 11: fields 0/11, 11/11, 22/10; any
 * r: field bit_group*r / min(r, 32 - bit_group*r)).  Same three sweeps, no copy-back; `histogram` (2^width
 * words) is left holding the bucket START offsets.  Checker for lsd_sort_pass with composite r.
 */
ORACLE_API void lsd_oracle_sort_pass_field(const uint32_t *in, uint32_t *out, int64_t count, uint64_t *histogram,
                                           int shift, int width)
{
    const size_t buckets = (size_t)1 << width;
    const uint32_t mask = (uint32_t)(buckets - 1);
    memset(histogram, 0, buckets * sizeof(uint64_t));
    for (int64_t i = 0; i < count; ++i)
        histogram[(in[i] >> shift) & mask] += 1u;
    uint64_t running = 0;
    for (size_t b = 0; b < buckets; ++b) {
        running += histogram[b];
        histogram[b] = running;
    }
    for (int64_t i = count; i-- > 0;) {
        const uint32_t key = in[i];
        out[--histogram[(key >> shift) & mask]] = key;
    }
}

/*
 * LSDRadixSort (LSDRadixSort.cu:62-69) widened to 64-bit keys: 64/r passes of LSDRadixSortPass (:25-54), least
 * significant digit first, with the digit taken from a 64-bit word.  The reference sorts uint32 only; this is
 * the same loop and the same pass with the key type changed.  Checker for lsd_sort64.  `histogram`: 2^r words.
 * Returns -1 unless r divides 64 and r <= 16.
 */
ORACLE_API int lsd_oracle_sort64(uint64_t *in, uint64_t *out, int64_t count, uint64_t *histogram, int r)
{
    if (r <= 0 || r > 16 || (64 % r) != 0)
        return -1;
    const size_t buckets = (size_t)1 << r;
    const uint64_t mask = (uint64_t)buckets - 1u;
    for (int g = 0; g < 64 / r; ++g) {
        const int shift = g * r;
        memset(histogram, 0, buckets * sizeof(uint64_t));
        for (int64_t i = 0; i < count; ++i)
            histogram[(in[i] >> shift) & mask] += 1u;
        uint64_t running = 0;
        for (size_t b = 0; b < buckets; ++b) {
            running += histogram[b];
            histogram[b] = running;
        }
        for (int64_t i = count; i-- > 0;) {
            const uint64_t key = in[i];
            out[--histogram[(key >> shift) & mask]] = key;
        }
        if (count > 0)
            memcpy(in, out, (size_t)count * sizeof(uint64_t));
    }
    return 0;
}
