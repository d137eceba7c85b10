// histogram.cu -- the two histogram kernels of liblsdsort.
//
// (1) digit_hist_kernel: ONE read of the keys builds the whole-array histograms of all
//     32/r digits (north_star step 1).  Replaces the per-pass BuildHistogramsKernel launch
//     of the reference (LSDRadixSort.cu:850) on the sort path.
// (2) tile_hist_kernel: the reference-layout [G][2^r] per-tile histogram of one digit,
//     i.e. the drop-in for BuildHistogramsKernel (LSDRadixSort.cu:660-702).
#include "common.cuh"

namespace lsd {

// -------------------------------------------------------------------------------------
// (1) Whole-array digit histograms.
//
// Data layout in shared memory: cnt[row][lane], row = pass * H + digit, 32 uint32 per row.
// A thread only ever touches column `lane`, so every shared atomic of a warp instruction
// lands in 32 distinct banks whatever the key distribution is (uniform, all-equal, sorted):
// no bank conflicts and no same-address serialisation inside a warp.  R=8 needs
// 4*256*32*4 B = 128 KiB, which is why this kernel runs one 1024-thread CTA per SM
// (B200: 227 KiB opt-in shared memory per CTA).  Keys are read with 128-bit streaming
// loads, four in flight per thread.  The flush reads each row along a diagonal (column
// (j + lane) & 31) so that it is conflict-free too, and issues one 64-bit global atomic
// per non-empty (pass, digit).
// Algorithmic bytes: 4 B per key read; the 32/r * 2^r * 8 B output is negligible.
// -------------------------------------------------------------------------------------
constexpr int kHistThreads = 1024;
#ifndef LSD_HIST_UNROLL
#define LSD_HIST_UNROLL 8
#endif
constexpr int kHistUnroll = LSD_HIST_UNROLL;  // 128-bit loads in flight per thread

// TOP_ONLY: ONE digit only (run-time index `one`: the top digit for the multi-GPU planning step, any digit for lsd_sort_pass)
template <int RB, bool TOP_ONLY = false, bool TYPED = false>
__device__ __forceinline__ void hist_add_key(uint32_t* cnt_lane, uint32_t key, KeyXform xf = KeyXform{0u, 0u}, uint32_t one = 0)
{
    constexpr int NP = 32 / RB;
    constexpr int H = 1 << RB;
    if constexpr (TYPED) key = key_to_unsigned(key, xf);
    if constexpr (TOP_ONLY) {
        const uint32_t d = (key >> (one * RB)) & (H - 1);
        atomicAdd(cnt_lane + ((one * H + d) << 5), 1u);
    } else {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const uint32_t d = (key >> (p * RB)) & (H - 1);
            atomicAdd(cnt_lane + ((p * H + d) << 5), 1u);
        }
    }
}

// RF < RB (narrow digits, r = 1 / 2 / 4): the keys are counted by their 8-bit digits all the same -- four shared atomics per key
// instead of 32 / r -- and the CTA folds its 4 x 256 counts into the [32/RF][2^RF] bins before the flush (an r-bit digit is
// a bit field of one 8-bit digit, so its histogram is a sum over that digit's bins).
template <int RB, bool TOP_ONLY = false, bool TYPED = false, int RF = RB>
__global__ void __launch_bounds__(kHistThreads, 1)
digit_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, unsigned long long* __restrict__ hist, KeyXform xf,
                  uint4* __restrict__ zero_ptr, uint64_t zero_vecs, uint32_t one)
{
    constexpr int NP = 32 / RB;
    constexpr int H = 1 << RB;
    constexpr int ROWS = NP * H;
    extern __shared__ uint32_t cnt[];  // [ROWS][32]

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    for (uint32_t i = tid; i < ROWS * 32; i += kHistThreads) cnt[i] = 0;
    __syncthreads();

    uint32_t* cnt_lane = cnt + lane;
    const uint64_t nvec = n >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * kHistThreads;
    uint64_t i = (uint64_t)blockIdx.x * kHistThreads + tid;

    // The sort lets this kernel zero its look-back records on the side (sort.cu: sort_enqueue): the kernel is bound by
    // its shared atomics, not by HBM, so the ~0.5 B of stores per key ride along instead of costing a memset of their own.
    for (uint64_t z = i; z < zero_vecs; z += stride) zero_ptr[z] = make_uint4(0u, 0u, 0u, 0u);

    // main body: kHistUnroll independent 128-bit loads in flight per thread
    for (; i + (kHistUnroll - 1) * stride < nvec; i += kHistUnroll * stride) {
        uint4 v[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) v[u] = ld_stream_v4(keys + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, v[u].x, xf, one); hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, v[u].y, xf, one);
            hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, v[u].z, xf, one); hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, v[u].w, xf, one);
        }
    }
    for (; i < nvec; i += stride) {
        const uint4 a = ld_stream_v4(keys + 4 * i);
        hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, a.x, xf, one); hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, a.y, xf, one);
        hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, a.z, xf, one); hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, a.w, xf, one);
    }
    // ragged tail (n % 4 keys) -- block 0 only
    if (blockIdx.x == 0) {
        const uint64_t t = (nvec << 2) + tid;
        if (t < n) hist_add_key<RB, TOP_ONLY, TYPED>(cnt_lane, keys[t], xf, one);
    }
    __syncthreads();

    if constexpr (RF == RB) {
        for (uint32_t row = tid; row < ROWS; row += kHistThreads) {
            const uint32_t* r = cnt + (row << 5);
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += r[(j + lane) & 31];
            if (sum) atomicAdd(hist + row, (unsigned long long)sum);
        }
    } else {
        static_assert(RB == 8 && ROWS == kHistThreads && !TOP_ONLY && RB % RF == 0, "fold: one 8-bit row per thread");
        uint32_t sum = 0;
        {
            const uint32_t* r = cnt + (tid << 5);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += r[(j + lane) & 31];
        }
        __syncthreads();  // every row has been read: its first word takes the row sum
        cnt[tid << 5] = sum;
        __syncthreads();
        constexpr uint32_t HF = 1u << RF, BINS = (32u / RF) << RF;  // at most 128 bins
        if (tid < BINS) {
            const uint32_t digit = tid / HF, v = tid % HF;                // narrow digit `digit` = bits [digit*RF, digit*RF + RF)
            const uint32_t row0 = (digit * RF / 8u) << 8, off = digit * RF % 8u;
            uint32_t acc = 0;
            for (uint32_t x = 0; x < 256u; ++x)
                if (((x >> off) & (HF - 1u)) == v) acc += cnt[(row0 + x) << 5];
            if (acc) atomicAdd(hist + tid, (unsigned long long)acc);
        }
    }
}

template <int RB, bool TOP_ONLY = false, bool TYPED = false, int RF = RB>
static int launch_digit_hist_t(const uint32_t* keys, uint64_t n, uint64_t* hist, cudaStream_t s, uint32_t key_type = 0,
                               void* zero_ptr = nullptr, size_t zero_bytes = 0, uint32_t one = 32 / RB - 1)
{
    constexpr int ROWS = (32 / RB) << RB;
    constexpr int BINS = (32 / RF) << RF;
    const size_t smem = (size_t)ROWS * 32 * sizeof(uint32_t);
    LSD_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)BINS * sizeof(uint64_t), s));
    if (n == 0) return LSD_OK;
    LSD_CUDA_TRY(cudaFuncSetAttribute(digit_hist_kernel<RB, TOP_ONLY, TYPED, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one CTA per SM, but never more CTAs than there are 16 KiB slices of input
    const uint64_t slices = ((n >> 2) + kHistThreads - 1) / kHistThreads;
    int grid = sm_count();
    if ((uint64_t)grid > slices) grid = (int)(slices ? slices : 1);
    if (zero_bytes != 0 && (!aligned_to(zero_ptr, 16) || zero_bytes % 16 != 0)) return LSD_ERR_ALIGNMENT;
    digit_hist_kernel<RB, TOP_ONLY, TYPED, RF><<<grid, kHistThreads, smem, s>>>(keys, n, reinterpret_cast<unsigned long long*>(hist),
                                                                                key_xform_of(key_type), static_cast<uint4*>(zero_ptr),
                                                                                (uint64_t)(zero_bytes / 16), one);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

// One digit only (one shared atomic per key instead of 32/r): the top digit for the multi-GPU planning step, the digit of the
// pass for lsd_sort_pass.  Same [32/r][2^r] layout, the other rows are zero.
int launch_one_digit_histogram(const uint32_t* keys, uint64_t n, int r, int digit, uint64_t* hist, cudaStream_t s)
{
    if (digit < 0 || digit >= 32 / r) return LSD_ERR_INVALID_VALUE;
    switch (r) {
        case 1: return launch_digit_hist_t<1, true>(keys, n, hist, s, 0, nullptr, 0, (uint32_t)digit);
        case 2: return launch_digit_hist_t<2, true>(keys, n, hist, s, 0, nullptr, 0, (uint32_t)digit);
        case 4: return launch_digit_hist_t<4, true>(keys, n, hist, s, 0, nullptr, 0, (uint32_t)digit);
        case 8: return launch_digit_hist_t<8, true>(keys, n, hist, s, 0, nullptr, 0, (uint32_t)digit);
    }
    return LSD_ERR_INVALID_VALUE;
}

int launch_top_digit_histogram(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s)
{
    if (r != 1 && r != 2 && r != 4 && r != 8) return LSD_ERR_INVALID_VALUE;
    return launch_one_digit_histogram(keys, n, r, 32 / r - 1, hist, s);
}

int launch_digit_histograms(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s, uint32_t key_type, void* zero_ptr,
                            size_t zero_bytes)
{
    if (key_type != 0) {  // typed keys: histogram of the keys' unsigned images
        switch (r) {
            case 1: return launch_digit_hist_t<8, false, true, 1>(keys, n, hist, s, key_type, zero_ptr, zero_bytes);
            case 2: return launch_digit_hist_t<8, false, true, 2>(keys, n, hist, s, key_type, zero_ptr, zero_bytes);
            case 4: return launch_digit_hist_t<8, false, true, 4>(keys, n, hist, s, key_type, zero_ptr, zero_bytes);
            case 8: return launch_digit_hist_t<8, false, true>(keys, n, hist, s, key_type, zero_ptr, zero_bytes);
        }
        return LSD_ERR_INVALID_VALUE;
    }
    switch (r) {
        // narrow digits are counted as 8-bit digits (4 shared atomics per key instead of 32 / r) and folded before the flush
        case 1: return launch_digit_hist_t<8, false, false, 1>(keys, n, hist, s, 0, zero_ptr, zero_bytes);
        case 2: return launch_digit_hist_t<8, false, false, 2>(keys, n, hist, s, 0, zero_ptr, zero_bytes);
        case 4: return launch_digit_hist_t<8, false, false, 4>(keys, n, hist, s, 0, zero_ptr, zero_bytes);
        case 8: return launch_digit_hist_t<8>(keys, n, hist, s, 0, zero_ptr, zero_bytes);
    }
    return LSD_ERR_INVALID_VALUE;
}

// -------------------------------------------------------------------------------------
// (1b) Histogram of an arbitrary bit field (shift, bits <= 16): the composite digit widths (r = 16, which the reference's
//      CPU path accepts -- "any factor of 32", LSDRadixSort.cu:56-69 -- and the classic 11-11-10 split) and the
//      sub-passes they run as.  Up to 12 bits the counters live in shared memory (conflicts between lanes are possible:
//      this is not the hot path), wider fields count with 64-bit global atomics (512 KiB of counters: L2 resident).
// -------------------------------------------------------------------------------------
constexpr int kFieldThreads = 512;
constexpr int kFieldSmemBits = 12;

template <bool SMEM>
__global__ void __launch_bounds__(kFieldThreads)
field_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t mask, unsigned long long* __restrict__ hist)
{
    extern __shared__ uint32_t fcnt[];
    const uint32_t tid = threadIdx.x;
    if constexpr (SMEM) {
        for (uint32_t i = tid; i <= mask; i += kFieldThreads) fcnt[i] = 0;
        __syncthreads();
    }
    auto add = [&](uint32_t key) {
        const uint32_t d = (key >> shift) & mask;
        if constexpr (SMEM) atomicAdd(fcnt + d, 1u);
        else atomicAdd(hist + d, 1ull);
    };
    const uint64_t nvec = n >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * kFieldThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kFieldThreads + tid; i < nvec; i += stride) {
        const uint4 a = ld_stream_v4(keys + 4 * i);
        add(a.x); add(a.y); add(a.z); add(a.w);
    }
    if (blockIdx.x == 0) {
        const uint64_t t = (nvec << 2) + tid;
        if (t < n) add(keys[t]);
    }
    if constexpr (SMEM) {
        __syncthreads();
        for (uint32_t i = tid; i <= mask; i += kFieldThreads)
            if (fcnt[i]) atomicAdd(hist + i, (unsigned long long)fcnt[i]);
    }
}

// hist[0 .. 2^bits) (uint64) <- counts of ((key >> shift) & (2^bits - 1)); overwritten
int launch_field_histogram(const uint32_t* keys, uint64_t n, int shift, int bits, uint64_t* hist, cudaStream_t s)
{
    if (bits < 1 || bits > 16 || shift < 0 || shift + bits > 32) return LSD_ERR_INVALID_VALUE;
    LSD_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint64_t) << bits, s));
    if (n == 0) return LSD_OK;
    const uint32_t mask = (1u << bits) - 1u;
    const uint64_t slices = ((n >> 2) + kFieldThreads - 1) / kFieldThreads;
    uint64_t grid = (uint64_t)sm_count() * 4;
    if (grid > slices) grid = slices ? slices : 1;
    if (bits <= kFieldSmemBits)
        field_hist_kernel<true><<<(int)grid, kFieldThreads, sizeof(uint32_t) << bits, s>>>(keys, n, shift, mask,
                                                                                        reinterpret_cast<unsigned long long*>(hist));
    else
        field_hist_kernel<false><<<(int)grid, kFieldThreads, 0, s>>>(keys, n, shift, mask, reinterpret_cast<unsigned long long*>(hist));
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

// in-place exclusive scan of up to 2^16 uint64 counters (one CTA; every thread owns a run of consecutive entries)
__global__ void __launch_bounds__(1024)
field_scan_kernel(uint64_t* __restrict__ a, uint32_t len)
{
    __shared__ uint64_t s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (len + 1023u) / 1024u;
    const uint32_t lo = tid * per, hi = lo + per < len ? lo + per : len;
    uint64_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += a[i];
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint64_t t = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += t;
        __syncthreads();
    }
    uint64_t run = s_part[tid] - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint64_t v = a[i];
        a[i] = run;
        run += v;
    }
}

int launch_field_scan(uint64_t* a, int bits, cudaStream_t s)
{
    field_scan_kernel<<<1, 1024, 0, s>>>(a, 1u << bits);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

// -------------------------------------------------------------------------------------
// (2) Reference-layout per-tile histograms: h[g*H + d] = #{keys of tile g with digit d},
//     tile g = keys [g*block, min((g+1)*block, n)).  One warp owns a tile at a time and
//     walks tiles with a grid stride, so any `block` works (the reference ties it to the
//     CUDA block size).  H <= 4 counts with ballots (no shared memory at all: two or four
//     addresses would serialise 32 atomics); H >= 16 uses a warp-private shared histogram.
//     Output rows are written coalesced.  Algorithmic bytes: 4 B/key read + G*H*4 B written.
// -------------------------------------------------------------------------------------
constexpr int kTileHistThreads = 256;

template <int RB>
__global__ void __launch_bounds__(kTileHistThreads)
tile_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t block, uint64_t tiles,
                 uint32_t* __restrict__ hist)
{
    constexpr int H = 1 << RB;
    constexpr int WARPS = kTileHistThreads / 32;
    __shared__ uint32_t wh[WARPS][H >= 16 ? H : 1];

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t warp_stride = (uint64_t)gridDim.x * WARPS;

    for (uint64_t g = (uint64_t)blockIdx.x * WARPS + warp; g < tiles; g += warp_stride) {
        const uint64_t lo = g * block;
        const uint64_t rem = n - lo;
        const uint32_t len = rem < block ? (uint32_t)rem : block;
        const uint32_t* src = keys + lo;
        uint32_t* dst = hist + g * H;

        if constexpr (H <= 4) {
            uint32_t mine = 0;  // lane v (< H) accumulates the count of digit v
            for (uint32_t base = 0; base < len; base += 32) {
                const uint32_t i = base + lane;
                const bool ok = i < len;
                const uint32_t d = ok ? digit_of<RB>(src[i], shift) : 0u;
#pragma unroll
                for (int v = 0; v < H; ++v) {
                    const uint32_t c = __popc(__ballot_sync(kFullMask, ok && d == (uint32_t)v));
                    if (lane == (uint32_t)v) mine += c;
                }
            }
            if (lane < H) dst[lane] = mine;
        } else {
            uint32_t* my = wh[warp];
            for (uint32_t d = lane; d < H; d += 32) my[d] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < len; i += 32) atomicAdd(&my[digit_of<RB>(src[i], shift)], 1u);
            __syncwarp();
            for (uint32_t d = lane; d < H; d += 32) dst[d] = my[d];
            __syncwarp();
        }
    }
}

int launch_tile_histograms(const uint32_t* keys, uint64_t n, int r, int bit_group, int block, uint32_t* hist,
                           cudaStream_t s)
{
    if (n == 0) return LSD_OK;
    const uint64_t tiles = (n + (uint64_t)block - 1) / (uint64_t)block;
    const int shift = bit_group * r;
    const uint64_t want = (tiles + (kTileHistThreads / 32) - 1) / (kTileHistThreads / 32);
    const uint64_t cap = (uint64_t)sm_count() * 8 * 4;  // enough resident warps; the rest by grid stride
    const int grid = (int)(want < cap ? want : cap);
    switch (r) {
        case 1: tile_hist_kernel<1><<<grid, kTileHistThreads, 0, s>>>(keys, n, shift, (uint32_t)block, tiles, hist); break;
        case 2: tile_hist_kernel<2><<<grid, kTileHistThreads, 0, s>>>(keys, n, shift, (uint32_t)block, tiles, hist); break;
        case 4: tile_hist_kernel<4><<<grid, kTileHistThreads, 0, s>>>(keys, n, shift, (uint32_t)block, tiles, hist); break;
        case 8: tile_hist_kernel<8><<<grid, kTileHistThreads, 0, s>>>(keys, n, shift, (uint32_t)block, tiles, hist); break;
        default: return LSD_ERR_INVALID_VALUE;
    }
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

}  // namespace lsd
