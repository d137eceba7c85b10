// onesweep_lpc.cuh -- onesweep digit pass, "lane-private counter" (LPC) ranking.
//
// Same contract as onesweep.cuh (one stable LSD pass, decoupled look-back, shared-memory reorder,
// coalesced per-bucket scatter), different ranking.  The ballot/match multisplit of onesweep.cuh
// costs ~48 instructions per key for an 8-bit digit (ncu: profiles/r01_v1_*).  Here ranking is two
// conflict-free shared-memory atomics per key and one scan of a small counter matrix per tile:
//
//   tile   = 32 * S keys, S = WARPS * ITEMS, S odd.  The tile is staged in shared memory by ONE TMA
//            bulk copy (cp.async.bulk + mbarrier), then read "lane-blocked": lane j of every warp owns
//            segment j = tile positions [j*S, (j+1)*S); warp w owns the sub-segment [w*ITEMS, (w+1)*ITEMS)
//            of each segment.  S odd makes that strided read bank-conflict free.
//   matrix = cnt[digit][lane]: 16-bit counters, digits 2k / 2k+1 packed in one word, row k = 32 words.
//            A thread only touches column `lane`, so every shared atomic of a warp hits 32 distinct
//            banks for ANY key distribution (uniform, all-equal, sorted ...).
//   phase 1: every key does atomicAdd(cnt[d][lane], 1)              (order irrelevant, all warps at once)
//   scan   : per row, exclusive prefix over the 32 lanes plus the tile-local bucket start; rows are
//            walked along diagonals (lane r reads column (r+k)&31) so the scan is conflict free too.
//            Row totals are the tile histogram that feeds the look-back chain.
//   phase 2: every key does rank = atomicAdd(cnt[d][lane], 1)  -- the old value IS its position in the
//            sorted tile.  Within a column, keys must take their ranks in position order: a thread's own
//            atomics are ordered by program order, and the warps of a CTA take turns (warp w waits for
//            warp w-1 on a named barrier), which is exactly sub-segment order.  Stable by construction.
//   then the keys go to the reorder buffer at `rank` and are streamed out per bucket as in onesweep.cuh.
//
// Shared-memory wavefronts per 32 keys: TMA write 1 + lane-blocked read 1 + 2 atomics + scan ~1.5 +
// reorder scatter ~3.5 (random banks) + linear read 1 + bucket-base lookup ~1.2  ~= 11, against ~18 in
// onesweep.cuh; instructions per key ~30 against ~95.
#pragma once
#include "onesweep.cuh"

namespace lsd {

// ---- PTX helpers: mbarrier + TMA bulk copy (global -> shared), named barriers ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu_v2(uint32_t* p, uint32_t a, uint32_t b)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_relaxed_gpu_v2(const uint32_t* p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

template <int RB, int WARPS, int ITEMS>
struct LpcShape {
    static constexpr int H = 1 << RB;
    static constexpr int ROWS = H / 2;                 // packed digit pairs
    static constexpr int THREADS = WARPS * 32;
    static constexpr int S = WARPS * ITEMS;            // keys per lane segment
    static constexpr int TILE = 32 * S;
    static constexpr int ROW_GROUPS = (ROWS + 31) / 32;
    static constexpr int GROUPS_PER_WARP = (ROW_GROUPS + WARPS - 1) / WARPS;
    static_assert(S % 2 == 1, "S = WARPS*ITEMS must be odd: lane-blocked shared reads stride by S words");
    static_assert(TILE < 65536, "ranks are 16-bit");
    static_assert(THREADS >= ROWS, "one digit-pair thread per matrix row");
    static_assert(WARPS <= 14, "named barriers: ids 1..WARPS-1 for the rank chain, 15 for the scan warps");
    static_assert((THREADS - ROWS) % 32 == 0 || ROWS < 32, "digit-pair threads must be whole warps (or a part of the last warp)");
    // layout (uint32 words): keys[TILE] | mat[ROWS*32] | tot[ROWS] | dp[ROWS] | gbase[H] | misc[64] ; mbarrier 8 B aligned
    static constexpr int OFF_MAT = TILE;
    static constexpr int OFF_TOT = OFF_MAT + ROWS * 32;
    static constexpr int OFF_DP = OFF_TOT + ROWS;
    static constexpr int OFF_GBASE = OFF_DP + ROWS;
    static constexpr int OFF_MISC = OFF_GBASE + H;
    static constexpr int WORDS = OFF_MISC + 64;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
};

template <int RB, int WARPS, int ITEMS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
onesweep_lpc_kernel(const PassArgs a)
{
    using S_ = LpcShape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, ROWS = S_::ROWS, THREADS = S_::THREADS, S = S_::S, TILE = S_::TILE;
    constexpr int GPW = S_::GROUPS_PER_WARP;
    constexpr int SCAN_WARPS = S_::ROW_GROUPS < WARPS ? S_::ROW_GROUPS : WARPS;  // warps 0.. own the matrix rows
    constexpr int DT0 = THREADS - ROWS;  // digit-pair threads are the LAST `ROWS` threads of the CTA: they enter the
                                         // rank chain last, so their look-back overlaps the chain of the first warps
    constexpr int LB = 4;                  // look-back window (predecessor rows fetched per round trip)
    constexpr uint32_t kScanBarrier = 15;  // named barrier: "scan pass 2 done" among the scan warps

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_keys = smem;
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_gbase = smem + S_::OFF_GBASE;
    uint32_t* s_misc = smem + S_::OFF_MISC;  // [0..15] scan partials, [32] tile id, [34..35] mbarrier
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 34);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    // ---- 0. ticket, TMA bulk load of the tile, zero the counter matrix meanwhile ----
    if (tid == 0) {
        mbar_init(s_bar, 1);
        const uint32_t t = atomicAdd(a.ticket, 1u);
        s_misc[32] = t;
        const uint32_t base = t * (uint32_t)TILE;
        if (a.portion_keys - base >= (uint32_t)TILE) {
            mbar_expect_tx(s_bar, TILE * 4);
            tma_bulk_g2s(s_keys, in + base, TILE * 4, s_bar);
        }
    }
    {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
        for (uint32_t i = tid; i < ROWS * 8; i += THREADS) m4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = s_misc[32];
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t left = a.portion_keys - tile_base;
    const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;

    if (valid == (uint32_t)TILE) {
        mbar_wait(s_bar, 0);
    } else {  // ragged last tile: guarded loads, pads sort last
        for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS) s_keys[p] = p < valid ? in[tile_base + p] : 0xFFFFFFFFu;
        __syncthreads();
    }

    // ---- 1. lane-blocked read: lane j, warp w, item i  <-  position j*S + w*ITEMS + i ----
    uint32_t key[ITEMS];
    {
        const uint32_t* src = s_keys + lane * S + warp * ITEMS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = src[i];
    }

    // ---- 2. phase 1: count.  cell(d, lane) = word (d>>1)*32 + lane, half d&1 ----
    uint32_t* col = s_mat + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = digit_of<RB>(key[i], a.shift);
        atomicAdd(col + ((d >> 1) << 5), 1u << ((d & 1u) << 4));
    }
    __syncthreads();  // counts complete; every key is in registers, so s_keys may be reused as reorder buffer

    // ---- 3. scan pass 1 (scan warps): row totals + partial sum of the columns below the diagonal start ----
    uint32_t below[GPW];  // packed sum of columns [0, lane) of my row (per row group I own)
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
        const uint32_t row = (uint32_t)(g * WARPS + warp) * 32u + lane;
        below[g] = 0;
        if (row < (uint32_t)ROWS) {
            const uint32_t* r = s_mat + row * 32u;
            uint32_t total = 0, wrapped = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const uint32_t c = (lane + k) & 31u;
                const uint32_t v = r[c];
                total += v;
                if (c < lane) wrapped += v;  // columns reached after the wrap == columns [0, lane)
            }
            below[g] = wrapped;
            s_tot[row] = total;
        }
    }
    __syncthreads();

    // ---- 4. tile histogram -> look-back (LOCAL), exclusive scan over digits -> tile-local bucket starts ----
    const bool digit_thread = tid >= (uint32_t)DT0;
    const uint32_t dt = tid - (uint32_t)DT0;  // digit pair (2*dt, 2*dt+1)
    uint32_t cnt_lo = 0, cnt_hi = 0, off_lo = 0, off_hi = 0;
    uint32_t* lb_row = a.lookback + (size_t)tile * H;
    {
        uint32_t mine = 0;
        if (digit_thread) {
            const uint32_t t = s_tot[dt];
            cnt_lo = t & 0xFFFFu;
            cnt_hi = t >> 16;
            mine = cnt_lo + cnt_hi;  // includes pads; they sit in the last digit and therefore shift nothing
            uint32_t pub_hi = cnt_hi;
            if (dt == (uint32_t)ROWS - 1) pub_hi -= (uint32_t)TILE - valid;  // pads are not keys
            const uint32_t flag = tile == 0 ? kLbGlobal : kLbLocal;
            st_relaxed_gpu_v2(lb_row + 2 * dt, flag | cnt_lo, flag | pub_hi);  // publish as early as possible
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_misc[warp] = incl;
        __syncthreads();
        uint32_t prefix = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w)
            if ((uint32_t)w < warp) prefix += s_misc[w];
        if (digit_thread) {
            off_lo = prefix + incl - mine;
            off_hi = off_lo + cnt_lo;
            s_dp[dt] = off_lo | (off_hi << 16);
            if (dt == (uint32_t)ROWS - 1) cnt_hi -= (uint32_t)TILE - valid;
        }
    }
    __syncthreads();

    // ---- 5a. look-back NOW (digit threads = tail warps), overlapping scan pass 2 and the head of the rank chain.
    //          One digit pair per thread, 64-bit words (both digits of a pair always carry the same flag). ----
    if (digit_thread) {
        uint32_t ex_lo = 0, ex_hi = 0;
        if (tile > 0) {
            // Windowed walk: LB predecessors are fetched at once (independent loads), then consumed in order
            // until an INCLUSIVE word is met; a not-yet-published predecessor restarts the window there.
            const uint32_t* p = a.lookback + (size_t)(tile - 1) * H + 2 * dt;
            uint32_t remaining = tile;  // predecessors that exist: tile-1 .. 0 (tile 0 is always INCLUSIVE)
            bool done = false;
            while (!done) {
                uint2 w[LB];
#pragma unroll
                for (int b = 0; b < LB; ++b)
                    w[b] = (uint32_t)b < remaining ? ld_relaxed_gpu_v2(p - (size_t)b * H) : make_uint2(0u, 0u);
                uint32_t consumed = 0;
#pragma unroll
                for (int b = 0; b < LB; ++b) {
                    if (!done && consumed == (uint32_t)b && w[b].x != 0) {
                        ex_lo += w[b].x & kLbValueMask;
                        ex_hi += w[b].y & kLbValueMask;
                        ++consumed;
                        if (w[b].x & kLbGlobal) done = true;
                    }
                }
                p -= (size_t)consumed * H;
                remaining -= consumed;
            }
            st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
        }
        const uint64_t b_lo = a.bases_in[2 * dt], b_hi = a.bases_in[2 * dt + 1];
        s_gbase[2 * dt] = (uint32_t)b_lo + ex_lo - off_lo;
        s_gbase[2 * dt + 1] = (uint32_t)b_hi + ex_hi - off_hi;
        if (a.bases_out != nullptr && tile == a.tiles - 1) {
            a.bases_out[2 * dt] = b_lo + ex_lo + cnt_lo;
            a.bases_out[2 * dt + 1] = b_hi + ex_hi + cnt_hi;
        }
    }

    // ---- 5b. scan pass 2 (scan warps): cell <- bucket start + keys of lower lanes ----
    if (warp < (uint32_t)SCAN_WARPS) {
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
            const uint32_t row = (uint32_t)(g * WARPS + warp) * 32u + lane;
            if (row < (uint32_t)ROWS) {
                uint32_t* r = s_mat + row * 32u;
                const uint32_t start = s_dp[row];
                uint32_t run = start + below[g];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const uint32_t c = (lane + k) & 31u;
                    if (c == 0) run = start;  // wrapped: column 0 starts at the bucket start
                    const uint32_t v = r[c];
                    r[c] = run;
                    run += v;
                }
            }
        }
        if (SCAN_WARPS > 1) named_bar_sync(kScanBarrier, SCAN_WARPS * 32);  // matrix complete before warp 0 ranks
    }

    // ---- 6. phase 2: ranks, warps in turn (sub-segment order); all atomics are issued before the baton moves on,
    //          the scatter into the reorder buffer happens after it ----
    uint32_t rk[(ITEMS + 1) / 2];  // two 16-bit ranks per register
    if (warp > 0) named_bar_sync(warp, 64);  // wait until warp-1 has issued its rank atomics
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = digit_of<RB>(key[i], a.shift);
        const uint32_t sh = (d & 1u) << 4;
        const uint32_t old = atomicAdd(col + ((d >> 1) << 5), 1u << sh);
        const uint32_t r16 = (old >> sh) & 0xFFFFu;
        if (i & 1) rk[i >> 1] |= r16 << 16; else rk[i >> 1] = r16;
    }
    if (warp + 1 < (uint32_t)WARPS) named_bar_arrive(warp + 1, 64);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) s_keys[(i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu)] = key[i];
    __syncthreads();

    // ---- 8. stream the reorder buffer out, coalesced per bucket ----
    if (valid == (uint32_t)TILE) {
#pragma unroll
        for (int i = 0; i < S / WARPS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t k = s_keys[p];
            out[s_gbase[digit_of<RB>(k, a.shift)] + p] = k;
        }
    } else {
        for (uint32_t p = tid; p < valid; p += THREADS) {
            const uint32_t k = s_keys[p];
            out[s_gbase[digit_of<RB>(k, a.shift)] + p] = k;
        }
    }
}

template <int RB, int WARPS, int ITEMS, int MINB>
int onesweep_lpc_launch(const PassArgs& a, cudaStream_t s)
{
    using S_ = LpcShape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpc_kernel<RB, WARPS, ITEMS, MINB>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<a.tiles, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

constexpr int kModeLpc = 2;

template <int RB, int WARPS, int ITEMS, int MINB = 1>
constexpr OnesweepLauncher make_lpc_launcher()
{
    using S_ = LpcShape<RB, WARPS, ITEMS>;
    return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                            &onesweep_lpc_launch<RB, WARPS, ITEMS, MINB>};
}

}  // namespace lsd
