// onesweep_lpcp.cuh -- persistent, warp-specialised, software-pipelined LPC onesweep pass.
//
// Ranking is the lane-private counter matrix of onesweep_lpc.cuh.  What changes is the schedule.
// ncu on the one-tile-per-CTA kernel (profiles/r01_lpc_*) shows ~40 % of warp stall samples waiting for
// the decoupled look-back: a tile cannot stream out before every earlier tile has published its counts,
// and while it waits it holds its shared memory and registers idle.  Here a CTA is persistent and keeps
// TWO tiles in flight, and the look-back has its own warps:
//
//   worker warps 0..W-1, iteration `it`, tile t in buffer b = it & 1:
//       wait TMA(t)  ->  keys to registers  ->  count (shared atomics)              -- barrier --
//       scan warps   : matrix scan, tile-local bucket starts; hand totals to the look-back warps
//       output warps : stream tile t-1 out of buffer b^1 (its prefix was resolved during the last iteration),
//                      then take the next ticket and issue TMA(t+1) into the drained buffer b^1
//       rank chain (warp 0, 1, ...), scatter keys into buffer b                      -- barrier --
//   look-back warps (one thread per digit pair):
//       wait totals(t) -> publish LOCAL -> windowed walk over predecessors -> publish INCLUSIVE ->
//       bucket bases of tile t for the output one iteration later.
//
// So the walk of tile t overlaps the ranking of tile t (and of t+1 if it is slow), the load of tile t+1
// overlaps the ranking of tile t, and nothing waits on global memory latency in the steady state.
// Roles synchronise through named barriers (worker-only barriers, rank chain) and mbarriers
// (TMA completion, totals ready, prefix ready).
#pragma once
#include "onesweep_lpc.cuh"

namespace lsd {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int RB, int WARPS, int ITEMS>
struct LpcpShape {
    static constexpr int H = 1 << RB;
    static constexpr int ROWS = H / 2;
    static constexpr int S = WARPS * ITEMS;
    static constexpr int TILE = 32 * S;
    static constexpr int ROW_GROUPS = (ROWS + 31) / 32;
    static constexpr int SW = ROW_GROUPS;              // scan warps (worker warps 0..SW-1)
    static constexpr int OW = WARPS - SW;              // output warps (worker warps SW..WARPS-1)
    static constexpr int LBW = ROW_GROUPS;             // look-back warps
    static constexpr int WORKER_THREADS = WARPS * 32;
    static constexpr int THREADS = (WARPS + LBW) * 32;
    static_assert(S % 2 == 1, "S = WARPS*ITEMS must be odd (conflict-free lane-blocked reads)");
    static_assert(TILE < 65536, "ranks are 16-bit");
    static_assert(OW >= 1, "need at least one output warp besides the scan warps");
    static_assert(WARPS <= 12, "named barriers 1..WARPS-1 chain, 13 workers, 14 output warps, 15 scan warps");
    // shared memory, in uint32 words
    static constexpr int OFF_BUF = 0;                          // [2][TILE]
    static constexpr int OFF_MAT = OFF_BUF + 2 * TILE;         // [ROWS][32]
    static constexpr int OFF_TOT = OFF_MAT + ROWS * 32;        // [2][ROWS]
    static constexpr int OFF_DP = OFF_TOT + 2 * ROWS;          // [2][ROWS]
    static constexpr int OFF_GBASE = OFF_DP + 2 * ROWS;        // [2][H]
    static constexpr int OFF_MISC = OFF_GBASE + 2 * H;         // [0..15] partials, [16..17] tile ids
    static constexpr int OFF_BARS = OFF_MISC + 32;             // 6 mbarriers (uint64): full[2], tot[2], prefix[2]
    static constexpr int WORDS = OFF_BARS + 12;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
};

template <int RB, int WARPS, int ITEMS, int MINB>
__global__ void __launch_bounds__((WARPS + LpcpShape<RB, WARPS, ITEMS>::LBW) * 32, MINB)
onesweep_lpcp_kernel(const PassArgs a)
{
    using S_ = LpcpShape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, ROWS = S_::ROWS, S = S_::S, TILE = S_::TILE;
    constexpr int SW = S_::SW, OW = S_::OW, LBW = S_::LBW, WT = S_::WORKER_THREADS;
    constexpr int LB = 8;  // look-back window
    constexpr uint32_t kBarWorkers = 13, kBarOutput = 14, kBarScan = 15;

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_buf = smem + S_::OFF_BUF;
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_gbase = smem + S_::OFF_GBASE;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    volatile uint32_t* s_tile = s_misc + 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_::OFF_BARS);
    uint64_t* bar_full = bars;        // [2] TMA landed
    uint64_t* bar_tot = bars + 2;     // [2] totals + bucket starts of the tile are in shared memory
    uint64_t* bar_prefix = bars + 4;  // [2] global bucket bases of the tile are in shared memory

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t tiles = a.tiles;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    // ---- prologue: barriers, first ticket, first TMA, zero matrix ----
    if (tid == 0) {
        mbar_init(&bar_full[0], 1);
        mbar_init(&bar_full[1], 1);
        mbar_init(&bar_tot[0], SW * 32);
        mbar_init(&bar_tot[1], SW * 32);
        mbar_init(&bar_prefix[0], LBW * 32);
        mbar_init(&bar_prefix[1], LBW * 32);
        const uint32_t t = atomicAdd(a.ticket, 1u);
        s_tile[0] = t;
        if (t < tiles && a.portion_keys - t * (uint32_t)TILE >= (uint32_t)TILE) {
            mbar_expect_tx(&bar_full[0], TILE * 4);
            tma_bulk_g2s(s_buf, in + (size_t)t * TILE, TILE * 4, &bar_full[0]);
        }
    }
    {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
        for (uint32_t i = tid; i < ROWS * 8; i += S_::THREADS) m4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    if (warp >= (uint32_t)WARPS) {
        // =====================================================================================
        // look-back warps
        // =====================================================================================
        const uint32_t dt = tid - (uint32_t)WT;  // digit pair (2*dt, 2*dt+1)
        for (uint32_t it = 0;; ++it) {
            const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(&bar_tot[b], ph);
            const uint32_t tile = s_tile[b];
            if (tile >= tiles) break;
            if (dt < (uint32_t)ROWS) {
                const uint32_t left = a.portion_keys - tile * (uint32_t)TILE;
                const uint32_t pads = left < (uint32_t)TILE ? (uint32_t)TILE - left : 0u;
                const uint32_t t = s_tot[b * ROWS + dt];
                const uint32_t start = s_dp[b * ROWS + dt];
                const uint32_t cnt_lo = t & 0xFFFFu;
                uint32_t cnt_hi = t >> 16;
                if (dt == (uint32_t)ROWS - 1) cnt_hi -= pads;  // pads of a ragged last tile are not keys
                uint32_t* lb_row = a.lookback + (size_t)tile * H;
                uint32_t ex_lo = 0, ex_hi = 0;
                if (tile == 0) {
                    st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | cnt_lo, kLbGlobal | cnt_hi);
                } else {
                    st_relaxed_gpu_v2(lb_row + 2 * dt, kLbLocal | cnt_lo, kLbLocal | cnt_hi);
                    const uint32_t* p = lb_row - H + 2 * dt;
                    uint32_t remaining = tile;
                    bool done = false;
                    while (!done) {
                        uint2 w[LB];
#pragma unroll
                        for (int k = 0; k < LB; ++k)
                            w[k] = (uint32_t)k < remaining ? ld_relaxed_gpu_v2(p - (size_t)k * H) : make_uint2(0u, 0u);
                        uint32_t consumed = 0;
#pragma unroll
                        for (int k = 0; k < LB; ++k) {
                            if (!done && consumed == (uint32_t)k && w[k].x != 0) {
                                ex_lo += w[k].x & kLbValueMask;
                                ex_hi += w[k].y & kLbValueMask;
                                ++consumed;
                                if (w[k].x & kLbGlobal) done = true;
                            }
                        }
                        p -= (size_t)consumed * H;
                        remaining -= consumed;
                    }
                    st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
                }
                const uint64_t b_lo = a.bases_in[2 * dt], b_hi = a.bases_in[2 * dt + 1];
                s_gbase[b * H + 2 * dt] = (uint32_t)b_lo + ex_lo - (start & 0xFFFFu);
                s_gbase[b * H + 2 * dt + 1] = (uint32_t)b_hi + ex_hi - (start >> 16);
                if (a.bases_out != nullptr && tile == tiles - 1) {
                    a.bases_out[2 * dt] = b_lo + ex_lo + cnt_lo;
                    a.bases_out[2 * dt + 1] = b_hi + ex_hi + cnt_hi;
                }
            }
            mbar_arrive(&bar_prefix[b]);
        }
        return;
    }

    // =========================================================================================
    // worker warps
    // =========================================================================================
    uint32_t* col = s_mat + lane;
    uint32_t it = 0;
    uint32_t prev_valid = 0, prev_tile = 0;
    for (;; ++it) {
        const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
        const uint32_t tile = s_tile[b];
        if (tile >= tiles) {
            if (warp < (uint32_t)SW) mbar_arrive(&bar_tot[b]);  // tell the look-back warps to stop
            break;
        }
        uint32_t* buf = s_buf + b * TILE;
        const uint32_t tile_base = tile * (uint32_t)TILE;
        const uint32_t left = a.portion_keys - tile_base;
        const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
        if (valid == (uint32_t)TILE) {
            mbar_wait(&bar_full[b], ph);
        } else {  // ragged last tile: guarded loads, pads sort last
            for (uint32_t p = tid; p < (uint32_t)TILE; p += WT) buf[p] = p < valid ? in[tile_base + p] : 0xFFFFFFFFu;
            named_bar_sync(kBarWorkers, WT);
        }

        // ---- keys to registers (lane-blocked), count ----
        uint32_t key[ITEMS];
        {
            const uint32_t* src = buf + lane * S + warp * ITEMS;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = src[i];
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = digit_of<RB>(key[i], a.shift);
            atomicAdd(col + ((d >> 1) << 5), 1u << ((d & 1u) << 4));
        }
        named_bar_sync(kBarWorkers, WT);  // counts complete; buffer b is free until the scatter

        if (warp < (uint32_t)SW) {
            // ---- scan warps: row totals, bucket starts, exclusive lane prefix ----
            const uint32_t row = warp * 32u + lane;
            uint32_t* r = s_mat + row * 32u;
            uint32_t total = 0, below = 0;
            if (row < (uint32_t)ROWS) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const uint32_t c = (lane + k) & 31u;
                    const uint32_t v = r[c];
                    total += v;
                    if (c < lane) below += v;
                }
            }
            const uint32_t lo = total & 0xFFFFu, hi = total >> 16;
            const uint32_t mine = lo + hi;
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t prefix = 0;
            if (SW > 1) {
                if (lane == 31) s_misc[warp] = incl;
                named_bar_sync(kBarScan, SW * 32);
#pragma unroll
                for (int w = 0; w < SW; ++w)
                    if ((uint32_t)w < warp) prefix += s_misc[w];
            }
            const uint32_t start_lo = prefix + incl - mine;
            const uint32_t start = start_lo | ((start_lo + lo) << 16);
            if (row < (uint32_t)ROWS) {
                s_tot[b * ROWS + row] = total;
                s_dp[b * ROWS + row] = start;
            }
            mbar_arrive(&bar_tot[b]);  // release: look-back warps may read totals/starts of this tile
            if (row < (uint32_t)ROWS) {
                uint32_t run = start + below;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const uint32_t c = (lane + k) & 31u;
                    if (c == 0) run = start;
                    const uint32_t v = r[c];
                    r[c] = run;
                    run += v;
                }
            }
            if (SW > 1) named_bar_sync(kBarScan, SW * 32);  // matrix complete before warp 0 starts the rank chain
        } else {
            // ---- output warps: stream the previous tile out, then prefetch the next one into its buffer ----
            const uint32_t otid = tid - SW * 32u;
            if (it > 0) {
                mbar_wait(&bar_prefix[b ^ 1u], ((it - 1) >> 1) & 1u);
                const uint32_t* pbuf = s_buf + (b ^ 1u) * TILE;
                const uint32_t* gb = s_gbase + (b ^ 1u) * H;
#pragma unroll 4
                for (uint32_t p = otid; p < (uint32_t)TILE; p += OW * 32) {
                    const uint32_t k = pbuf[p];
                    out[gb[digit_of<RB>(k, a.shift)] + p] = k;
                }
                if (OW > 1) named_bar_sync(kBarOutput, OW * 32);  // buffer b^1 drained
            }
            if (otid == 0) {
                const uint32_t tn = atomicAdd(a.ticket, 1u);
                s_tile[b ^ 1u] = tn;
                if (tn < tiles && a.portion_keys - tn * (uint32_t)TILE >= (uint32_t)TILE) {
                    mbar_expect_tx(&bar_full[b ^ 1u], TILE * 4);
                    tma_bulk_g2s(s_buf + (b ^ 1u) * TILE, in + (size_t)tn * TILE, TILE * 4, &bar_full[b ^ 1u]);
                }
            }
        }

        // ---- rank chain: warps in turn; all rank atomics are issued before the baton moves on ----
        uint32_t rk[(ITEMS + 1) / 2];
        if (warp > 0) named_bar_sync(warp, 64);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = digit_of<RB>(key[i], a.shift);
            const uint32_t sh = (d & 1u) << 4;
            const uint32_t old = atomicAdd(col + ((d >> 1) << 5), 1u << sh);
            const uint32_t r16 = (old >> sh) & 0xFFFFu;
            if (i & 1) rk[i >> 1] |= r16 << 16; else rk[i >> 1] = r16;
        }
        if (warp + 1 < (uint32_t)WARPS) {
            named_bar_arrive(warp + 1, 64);
        } else {
            // last warp of the chain: every rank of this tile has been taken -> clear the matrix for the next tile
            uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
            for (int i = 0; i < ROWS * 8 / 32; ++i) m4[i * 32 + lane] = make_uint4(0, 0, 0, 0);
            if ((ROWS * 8) % 32 != 0 && lane < (ROWS * 8) % 32) m4[(ROWS * 8 / 32) * 32 + lane] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) buf[(i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu)] = key[i];
        prev_valid = valid;
        prev_tile = tile;
        named_bar_sync(kBarWorkers, WT);  // scatter + matrix clear done; next tile id visible
    }

    // ---- epilogue: the last tile this CTA ranked still has to be streamed out ----
    if (it > 0) {
        const uint32_t b = (it - 1) & 1u;
        mbar_wait(&bar_prefix[b], ((it - 1) >> 1) & 1u);
        const uint32_t* pbuf = s_buf + b * TILE;
        const uint32_t* gb = s_gbase + b * H;
        for (uint32_t p = tid; p < prev_valid; p += WT) {
            const uint32_t k = pbuf[p];
            out[gb[digit_of<RB>(k, a.shift)] + p] = k;
        }
    }
    (void)prev_tile;
}

template <int RB, int WARPS, int ITEMS, int MINB>
int onesweep_lpcp_launch(const PassArgs& a, cudaStream_t s)
{
    using S_ = LpcpShape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpcp_kernel<RB, WARPS, ITEMS, MINB>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const uint32_t resident = (uint32_t)sm_count() * MINB;  // persistent: every CTA must be co-resident
    const uint32_t grid = a.tiles < resident ? a.tiles : resident;
    kern<<<grid, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

constexpr int kModeLpcp = 3;

template <int RB, int WARPS, int ITEMS, int MINB>
constexpr OnesweepLauncher make_lpcp_launcher()
{
    using S_ = LpcpShape<RB, WARPS, ITEMS>;
    return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpcp, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                            &onesweep_lpcp_launch<RB, WARPS, ITEMS, MINB>};
}

}  // namespace lsd
