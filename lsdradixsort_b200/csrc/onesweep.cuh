// onesweep.cuh -- one LSD digit pass as a single "onesweep" kernel (north_star step 3).
//
// Replaces, per pass, the reference's BuildHistogramsKernel + cudaMemcpy D2D + BlockPrefixSumKernel +
// 2 x TransposeSMEMKernel + recursive GPUPrefixSum + LSDRadixSortKernel (LSDRadixSort.cu:844-906):
// the keys are read once and written once (8 B per key per pass).
//
// Per tile (one CTA, THREADS x ITEMS keys):
//   1. a ticket gives the tile its place in the look-back chain (forward progress);
//   2. keys are loaded warp-striped (lane l, item i of warp w = key w*32*ITEMS + i*32 + l):
//      every warp load is one 128-byte line, and (item, lane) order == position order;
//   3. ranking: warp multisplit.  For every item the warp finds the peers holding the same
//      digit (8 ballots, or match.any), the rank inside the warp is the running warp count of
//      that digit plus popc(peers below me); the highest peer bumps the warp count.  This is
//      stable: earlier items and lower lanes (= earlier positions) always rank first;
//   4. per-digit counts are summed over warps -> the tile histogram, published to the look-back
//      chain as LOCAL; an exclusive scan over digits gives the tile-local bucket starts;
//   5. keys are scattered into the shared-memory reorder buffer at bucket_start + warp_offset +
//      rank, while the digit threads walk the chain backwards to get the tile's global prefix
//      (decoupled look-back) and publish INCLUSIVE;
//   6. the reorder buffer is streamed out in position order: consecutive threads hold
//      consecutive keys of a bucket, so global writes coalesce per digit bucket.
//
// Look-back word: flag[31:30] | value[29:0]; flag 1 = LOCAL count of this tile, 2 = INCLUSIVE
// prefix up to and including this tile, 0 = not ready.  Values are relative to the current
// "portion" (<= 2^30-1 keys); larger inputs run as consecutive portions whose last tile hands
// the advanced bucket bases to the next portion.
#pragma once
#include "common.cuh"

namespace lsd {

constexpr uint32_t kLbLocal = 1u << 30;
constexpr uint32_t kLbGlobal = 2u << 30;
constexpr uint32_t kLbValueMask = (1u << 30) - 1u;

constexpr int kMaxPasses = 32;

// Device-resident plan written by plan_kernel after the digit histogram: which passes can be
// skipped, which buffer each executed pass reads, and where the result ends up.
struct SortPlan {
    uint32_t skip[kMaxPasses];
    uint32_t src_is_scratch[kMaxPasses];
    uint32_t result_in_scratch;
    uint32_t executed_passes;
    uint32_t first_pass;  // first / last executed pass (typed keys are mapped to unsigned order on the way in and back
    uint32_t last_pass;   // on the way out); kMaxPasses when nothing runs
    uint32_t pad[60];
};

struct PassArgs {
    uint32_t* keys;        // caller's key buffer
    uint32_t* scratch;     // caller's ping-pong buffer
    const SortPlan* plan;
    const uint64_t* bases_in;  // [H] absolute exclusive bucket bases for this pass and portion
    uint64_t* bases_out;       // [H] bases for the next portion (written by the last tile) or nullptr
    uint32_t* lookback;        // [tiles][H], zero-initialised
    uint32_t* ticket;          // zero-initialised
    uint64_t portion_base;     // first key of this portion
    uint32_t portion_keys;     // keys in this portion
    uint32_t tiles;            // tiles in this portion
    int pass;
    int shift;
    unsigned long long* trace;  // optional per-tile phase clocks (tuning aid), else nullptr
    const uint64_t* dst_ptrs;   // peer-scatter mode only: [H] device pointers, one per bucket (equal inside a segment)
    const uint32_t* dst_seg;    // peer-scatter mode only: [H] first | last << 16 bucket of the bucket's segment, or nullptr
    uint32_t* vals;             // pairs mode only: the caller's value buffer (travels with `keys`)
    uint32_t* vals_scratch;     // pairs mode only: ping-pong buffer of the values (travels with `scratch`)
    uint32_t key_type;          // lsd_key_type; non-zero only with the typed-key kernels (OnesweepLauncher.launch_typed)
    uint32_t digit_mask;        // 2^r - 1; narrower for the sub-passes of composite digit widths (honoured by the run-time-shift
                                // form of onesweep_lpc3_kernel only)
};

// The key mapping a pass applies when it reads / writes keys: identity unless the keys are typed and this is the first /
// last executed pass.
__device__ __forceinline__ KeyXform pass_xform_in(const PassArgs& a)
{
    return key_xform_of((a.key_type != 0u && a.plan->first_pass == (uint32_t)a.pass) ? a.key_type : 0u);
}
__device__ __forceinline__ KeyXform pass_xform_out(const PassArgs& a)
{
    return key_xform_of((a.key_type != 0u && a.plan->last_pass == (uint32_t)a.pass) ? a.key_type : 0u);
}

constexpr int kPassPlain = 0, kPassPeer = 1, kPassPairs = 2, kPassTyped = 3, kPassPairsTyped = 4;

enum MatchMode { kMatchBallot = 0, kMatchHw = 1 };

template <int RB, int MODE>
__device__ __forceinline__ uint32_t match_peers(uint32_t d)
{
    if constexpr (MODE == kMatchHw) {
        return __match_any_sync(kFullMask, d);
    } else {
        uint32_t peers = kFullMask;
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(kFullMask, bit);
            peers &= bit ? bal : ~bal;
        }
        return peers;
    }
}

template <int RB, int THREADS, int ITEMS>
struct OnesweepShape {
    static constexpr int H = 1 << RB;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int DPT = (H + THREADS - 1) / THREADS;  // digits per digit-thread
    static constexpr int DIGIT_THREADS = H / DPT;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * ((size_t)TILE + (size_t)WARPS * H + H + 64);
    // largest multiple of TILE that keeps look-back values below 2^30
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
};

template <int RB, int THREADS, int ITEMS, int MODE, bool PAIRS = false, bool TYPED = false>
__global__ void __launch_bounds__(THREADS)
onesweep_kernel(const PassArgs a)
{
    using S = OnesweepShape<RB, THREADS, ITEMS>;
    constexpr int H = S::H, WARPS = S::WARPS, TILE = S::TILE, DPT = S::DPT, DIGIT_THREADS = S::DIGIT_THREADS;

    if (a.plan->skip[a.pass]) return;  // uniform over the grid: digit is constant, pass is the identity

    extern __shared__ uint32_t smem[];
    uint32_t* s_keys = smem;                 // [TILE] reorder buffer
    uint32_t* s_whist = s_keys + TILE;       // [WARPS][H] warp digit counts -> warp bucket offsets
    uint32_t* s_gbase = s_whist + WARPS * H; // [H] global index of bucket start minus tile-local start
    uint32_t* s_misc = s_gbase + H;          // [0..31] cross-warp scan partials, [32] tile id

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) s_misc[32] = atomicAdd(a.ticket, 1u);
    for (uint32_t i = tid; i < WARPS * H; i += THREADS) s_whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_misc[32];

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t left = a.portion_keys - tile_base;
    const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;

    // ---- 2. load, warp-striped ----
    uint32_t key[ITEMS];
    {
        const uint32_t off = warp * (32u * ITEMS) + lane;
        const uint32_t* src = in + tile_base + off;
        if (valid == (uint32_t)TILE) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = ld_stream_u32(src + i * 32);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                key[i] = (off + i * 32u < valid) ? ld_stream_u32(src + i * 32) : 0xFFFFFFFFu;  // pads sort last
        }
        if constexpr (TYPED) {  // typed keys enter unsigned order in the first executed pass (pads stay last)
            const KeyXform xin = pass_xform_in(a);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (valid == (uint32_t)TILE || off + i * 32u < valid) key[i] = key_to_unsigned(key[i], xin);
        }
    }

    // ---- 3. rank inside the warp ----
    uint32_t rank[ITEMS];
    {
        uint32_t* wh = s_whist + warp * H;
        const uint32_t lt = lanemask_lt();
        const uint32_t gt = lanemask_gt();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = digit_of<RB>(key[i], a.shift);
            const uint32_t peers = match_peers<RB, MODE>(d);
            const uint32_t r = wh[d] + __popc(peers & lt);
            __syncwarp();
            if ((peers & gt) == 0) wh[d] = r + 1;  // highest peer: count so far + whole peer group
            __syncwarp();
            rank[i] = r;
        }
    }
    __syncthreads();

    // ---- 4. tile histogram, publish LOCAL, tile-local bucket starts ----
    uint32_t count[DPT];     // tile count of my digits
    uint32_t tile_off[DPT];  // tile-local exclusive start of my digits
    uint32_t* lb_row = a.lookback + (size_t)tile * H;
    {
        uint32_t wcount[DPT][WARPS];
        uint32_t mine = 0;
        if (tid < DIGIT_THREADS) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const uint32_t d = tid * DPT + j;
                uint32_t sum = 0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    wcount[j][w] = s_whist[w * H + d];
                    sum += wcount[j][w];
                }
                if (d == H - 1) sum -= (uint32_t)TILE - valid;  // drop the pads of a ragged last tile
                count[j] = sum;
                st_relaxed_gpu(lb_row + d, (tile == 0 ? kLbGlobal : kLbLocal) | sum);
                mine += sum;
            }
        }
        // block-wide exclusive scan of `mine` over digit threads
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_misc[warp] = incl;
        __syncthreads();
        uint32_t warp_prefix = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w)
            if ((uint32_t)w < warp) warp_prefix += s_misc[w];
        uint32_t run = warp_prefix + incl - mine;
        if (tid < DIGIT_THREADS) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const uint32_t d = tid * DPT + j;
                tile_off[j] = run;
                uint32_t wrun = run;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    s_whist[w * H + d] = wrun;  // bucket start + keys of lower warps
                    wrun += wcount[j][w];
                }
                run += count[j] + ((d == H - 1) ? (uint32_t)TILE - valid : 0u);
            }
        }
    }
    __syncthreads();

    // ---- 5a. scatter into the reorder buffer ----
    {
        const uint32_t* wh = s_whist + warp * H;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = digit_of<RB>(key[i], a.shift);
            s_keys[wh[d] + rank[i]] = key[i];
        }
    }

    // ---- 5b. decoupled look-back (digit threads) ----
    if (tid < DIGIT_THREADS) {
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const uint32_t d = tid * DPT + j;
            uint32_t excl = 0;
            if (tile > 0) {
                const uint32_t* p = a.lookback + (size_t)(tile - 1) * H + d;
                while (true) {
                    const uint32_t w = ld_relaxed_gpu(p);
                    if (w == 0) continue;  // predecessor not published yet
                    excl += w & kLbValueMask;
                    if (w & kLbGlobal) break;
                    p -= H;
                }
                st_relaxed_gpu(lb_row + d, kLbGlobal | (excl + count[j]));
            }
            const uint64_t base = a.bases_in[d];
            s_gbase[d] = (uint32_t)base + excl - tile_off[j];
            if (a.bases_out != nullptr && tile == a.tiles - 1) a.bases_out[d] = base + excl + count[j];
        }
    }
    __syncthreads();

    // ---- 6. stream the reorder buffer out, coalesced per bucket ----
    const KeyXform xout = TYPED ? pass_xform_out(a) : KeyXform{0u, 0u};  // typed keys leave unsigned order in the last executed pass
    if constexpr (PAIRS) {
        // keys out, remembering the digit of every position this thread copies: the value of the key at tile
        // position p goes to the same global index.  Then the values take the keys' route through the reorder
        // buffer (same warp-striped ownership, same bucket offset + rank), so every pass moves (key, value) together.
        static_assert(!PAIRS || RB <= 8, "digits are packed four to a register");
        uint32_t dpk[(ITEMS + 3) / 4];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            uint32_t d = 0;
            if (p < valid) {
                const uint32_t k = s_keys[p];
                d = digit_of<RB>(k, a.shift);
                out[s_gbase[d] + p] = (TYPED ? key_from_unsigned(k, xout) : k);
            }
            if (i & 3) dpk[i >> 2] |= d << (8 * (i & 3)); else dpk[i >> 2] = d;
        }
        __syncthreads();  // every key has left the reorder buffer
        const uint32_t* __restrict__ vin = (src_scratch ? a.vals_scratch : a.vals) + a.portion_base;
        uint32_t* __restrict__ vout = src_scratch ? a.vals : a.vals_scratch;
        {
            const uint32_t off = warp * (32u * ITEMS) + lane;
            const uint32_t* src = vin + tile_base + off;
            const uint32_t* wh = s_whist + warp * H;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t v = (off + i * 32u < valid) ? ld_stream_u32(src + i * 32) : 0u;
                s_keys[wh[digit_of<RB>(key[i], a.shift)] + rank[i]] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t d = (dpk[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            if (p < valid) vout[s_gbase[d] + p] = s_keys[p];
        }
    } else if (valid == (uint32_t)TILE) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t k = s_keys[p];
            out[s_gbase[digit_of<RB>(k, a.shift)] + p] = (TYPED ? key_from_unsigned(k, xout) : k);
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            if (p < valid) {
                const uint32_t k = s_keys[p];
                out[s_gbase[digit_of<RB>(k, a.shift)] + p] = (TYPED ? key_from_unsigned(k, xout) : k);
            }
        }
    }
}

// One (RB, THREADS, ITEMS, MODE) shape = one launcher; sort.cu picks by (r, block, variant).
struct OnesweepLauncher {
    int radix_bits;
    int threads;
    int items;
    int mode;
    uint32_t tile;
    uint32_t portion_max;
    size_t smem_bytes;
    int (*launch)(const PassArgs& a, cudaStream_t s);
    int (*launch_peer)(const PassArgs& a, cudaStream_t s);  // bucket-pointer scatter (multi-GPU exchange), or nullptr
    int (*launch_pairs)(const PassArgs& a, cudaStream_t s);  // key-value pass, or nullptr
    int (*launch_typed)(const PassArgs& a, cudaStream_t s);        // plain pass that maps typed keys (i32 / f32), or nullptr
    int (*launch_pairs_typed)(const PassArgs& a, cudaStream_t s);  // key-value pass that maps typed keys, or nullptr
};

template <int RB, int THREADS, int ITEMS, int MODE, bool PAIRS = false, bool TYPED = false>
int onesweep_launch(const PassArgs& a, cudaStream_t s)
{
    using S = OnesweepShape<RB, THREADS, ITEMS>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<RB, THREADS, ITEMS, MODE, PAIRS, TYPED>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_BYTES));
    onesweep_kernel<RB, THREADS, ITEMS, MODE, PAIRS, TYPED><<<a.tiles, THREADS, S::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

// WITH_PAIRS: also instantiate the key-value and the typed-key (i32 / f32) forms of the kernel.
template <int RB, int THREADS, int ITEMS, int MODE, bool WITH_PAIRS = false>
constexpr OnesweepLauncher make_launcher()
{
    using S = OnesweepShape<RB, THREADS, ITEMS>;
    if constexpr (WITH_PAIRS)
        return OnesweepLauncher{RB, THREADS, ITEMS, MODE, (uint32_t)S::TILE, S::PORTION_MAX, S::SMEM_BYTES,
                                &onesweep_launch<RB, THREADS, ITEMS, MODE>, nullptr,
                                &onesweep_launch<RB, THREADS, ITEMS, MODE, true>,
                                &onesweep_launch<RB, THREADS, ITEMS, MODE, false, true>,
                                &onesweep_launch<RB, THREADS, ITEMS, MODE, true, true>};
    else
        return OnesweepLauncher{RB, THREADS, ITEMS, MODE, (uint32_t)S::TILE, S::PORTION_MAX, S::SMEM_BYTES,
                                &onesweep_launch<RB, THREADS, ITEMS, MODE>, nullptr, nullptr, nullptr, nullptr};
}

// Tables defined in onesweep_r{1,2,4,8}.cu.  Entry 0 of each table is the default shape.
const OnesweepLauncher* onesweep_table_r1(int* count);
const OnesweepLauncher* onesweep_table_r2(int* count);
const OnesweepLauncher* onesweep_table_r4(int* count);
const OnesweepLauncher* onesweep_table_r8(int* count);

}  // namespace lsd
