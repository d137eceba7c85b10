// onesweep_lpc4.cuh -- the persistent LPC pass (onesweep_lpc3.cuh) with TWO rank chains and a wider look-back window.
//
// Why (profiles/r02_wide_pass_study.txt, bench_tools/trace.py on onesweep_lpc4_kernel): a tile of the round-1 default lives
// 18.8 K cycles, of which the rank chain is 8.6 K (9 turns x 29 returning shared atomics, one in flight per warp, ~33
// cycles each) and the look-back walk 10.7 K (28 hops, 9 rounds of 4 records); the chain's tail warps are the look-back
// warps, so the two serial stretches end together at ~15.5 K and neither alone moves the pass.  Here both shrink at once:
//   * cnt[digit][lane] keeps TWO 16-bit byte-offset counters per word -- low half: even warps, high half: odd warps; a
//     lane's segment of the tile is [even warps' keys | odd warps' keys] -- so the even and the odd warps form two
//     chains that run concurrently (5 + 4 turns);
//   * the scatter follows its atomic one step behind inside the turn instead of keeping 15 packed rank registers, which
//     frees the registers for a look-back window of 8 records per round (half the rounds).
// Same tile, workspace layout, look-back protocol, prefetch of the next tile into the dead counter matrix and key flavours
// (plain / typed) as onesweep_lpc4_kernel; 8-bit digits only (r < 8 stays on onesweep_lpc4_kernel).
#pragma once
#include "onesweep_lpc3.cuh"

namespace lsd {

// CLR: 0 = matrix cleared with 128-bit stores, 1 = st.bulk, 2 = TMA copy of a zero page (keeps the clear off the LSU pipe)
// NOB5: 1 = no CTA-wide barrier at the end of a tile: the next ticket is handed over through a second mbarrier
// TYPED: i32 / f32 keys are mapped to unsigned order when the first executed pass reads them and back when the last one writes
// TRACE: per-tile phase clocks into PassArgs.trace (bench_tools/trace.py).  Compile-time: the run-time checks and the clock
//        reads alone cost ~7 % of the pass (0.665 -> 0.715 ms), so only the tuning variant carries them.
template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CLR, int NOB5 = 0, bool TYPED = false, bool TRACE = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
onesweep_lpc4_kernel(const PassArgs a)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, THREADS = S_::THREADS, S = S_::S, TILE = S_::TILE;
    constexpr int SW = S_::SW, GPW = S_::GPW, LBT = S_::LBT, LBW = S_::LBW;
    constexpr uint32_t kBarTot = 14, kBarScan = 15;
    constexpr bool ALIAS = S_::OFF_DP - S_::OFF_MAT >= TILE;  // the dead matrix (+ tile counts) can hold the incoming tile
    constexpr int IN_OFF = (S_::WORDS + 3) & ~3;               // else: a dedicated prefetch buffer behind everything
    static_assert(LBT >= H / 2, "one digit pair per look-back thread");
    static_assert(RB == 8, "8-bit digits only");
    constexpr bool QLB = CLR == 10;  // quad look-back (lookback_quad.cuh)
    constexpr int EVEN = (WARPS + 1) / 2;  // warps of the even chain; they own the first EVEN*ITEMS keys of a lane segment
    static_assert(SW >= 2, "warps 0 and 1 (the heads of the two chains) are scan warps");
    static_assert(EVEN * ITEMS * 32 * 4 < 65536, "a row half must not carry into the other half");
    const int shift = SHIFT >= 0 ? SHIFT : a.shift;

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_keys = smem;                     // reorder buffer
    uint32_t* s_mat = smem + S_::OFF_MAT;        // counter matrix; between rank chain and next count: the incoming tile
    uint32_t* s_in = ALIAS ? s_mat : smem + IN_OFF;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_gbase = smem + S_::OFF_GBASE;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 34);
    uint64_t* s_bar2 = reinterpret_cast<uint64_t*>(s_misc + 36);  // "next ticket is in s_misc[32]" (NOB5)

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t half = warp & 1u;
    const uint32_t inc4 = half ? (4u << 16) : 4u;  // this warp's counter half, in bytes
    const uint32_t seg_off = lane * (uint32_t)S + (half ? (uint32_t)(EVEN * ITEMS) : 0u) + (warp >> 1) * (uint32_t)ITEMS;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    // ticket of the next tile + its TMA load into s_in (thread 0 only)
    auto fetch_next = [&]() {
        const uint32_t t = atomicAdd(a.ticket, 1u);
        s_misc[32] = t;
        const uint32_t base = t * (uint32_t)TILE;
        if (t < a.tiles && a.portion_keys - base >= (uint32_t)TILE) {
            mbar_expect_tx(s_bar, TILE * 4);
            tma_bulk_g2s(s_in, in + base, TILE * 4, s_bar);
        }
    };
    if (tid == 0) {
        mbar_init(s_bar, 1);
        if constexpr (NOB5) mbar_init(s_bar2, 1);
        fetch_next();
        if constexpr (NOB5) mbar_arrive_release(s_bar2);
    }
    __syncthreads();
    uint32_t phase2 = 0;

    char* mat_bytes = reinterpret_cast<char*>(s_mat);
    const uint32_t lane4 = lane << 2;
    auto cell_of = [&](uint32_t key) -> uint32_t {  // byte offset of cell (digit, lane) in the matrix
        if constexpr (SHIFT >= 0) return cell_offset<RB, SHIFT < 0 ? 0 : SHIFT>(key, lane4);
        else return (((key >> shift) & (uint32_t)(H - 1)) << 7) | lane4;
    };
    uint32_t phase = 0;
    const KeyXform xin = TYPED ? pass_xform_in(a) : KeyXform{0u, 0u};
    const bool typed_out = TYPED && a.plan->last_pass == (uint32_t)a.pass;
    const KeyXform xout = key_xform_of(typed_out ? a.key_type : 0u);

    while (true) {
        if constexpr (NOB5) {
            mbar_wait(s_bar2, phase2);
            phase2 ^= 1u;
        }
        const uint32_t tile = s_misc[32];
        if (tile >= a.tiles) break;
        const long long t_start = (TRACE && a.trace) ? clock64() : 0;  // phase clocks count from the moment the ticket is known
#define LSD_TRACE(slot)                                                                      \
    do {                                                                                     \
        if constexpr (TRACE)                                                                 \
            if (a.trace && lane == 0) a.trace[(size_t)tile * 16 + (slot)] = (unsigned long long)(clock64() - t_start); \
    } while (0)
        const uint32_t tile_base = tile * (uint32_t)TILE;
        const uint32_t left = a.portion_keys - tile_base;
        const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
        const uint32_t pads = (uint32_t)TILE - valid;

        if (valid == (uint32_t)TILE) {
            mbar_wait(s_bar, phase);
            phase ^= 1u;
        } else {
            const uint32_t pad_key = TYPED ? key_from_unsigned(0xFFFFFFFFu, xin) : 0xFFFFFFFFu;  // pads sort last
            for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS) s_in[p] = p < valid ? in[tile_base + p] : pad_key;
            __syncthreads();
        }

        if (warp == 0) LSD_TRACE(1);  // tile landed
        // ---- 1. lane-blocked read, then the matrix takes its place back ----
        uint32_t key[ITEMS];
        {
            const uint32_t* src = s_in + seg_off;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = TYPED ? key_to_unsigned(src[i], xin) : src[i];
        }
        if constexpr (CLR == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // reads before the async zero-fill
        __syncthreads();  // every key is in registers
        if constexpr (CLR == 2) {
            if (tid == 0) {
                mbar_expect_tx(s_bar, H * 128);
                tma_bulk_g2s(s_mat, g_lsd_zero_page, H * 128, s_bar);
            }
            mbar_wait(s_bar, phase);
            phase ^= 1u;
        } else if constexpr (CLR == 1) {
            // zero-fill by one st.bulk (UMEMSETS) instead of H*8 128-bit stores through the LSU
            if (tid == 0) {
                asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(smem_u32(s_mat)), "l"((uint64_t)(H * 128)) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        } else {
            uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
            for (uint32_t i = tid; i < H * 8; i += THREADS) m4[i] = make_uint4(0, 0, 0, 0);
        }
        if constexpr (CLR != 2) __syncthreads();  // matrix is zero
        if (warp == 0) LSD_TRACE(0);  // keys read, matrix cleared
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), inc4);
        if (warp == 0) LSD_TRACE(2);
        __syncthreads();  // counts complete
        if (warp == 0) LSD_TRACE(3);

        uint32_t* lb_row = a.lookback + (size_t)tile * H;
        // Pull this tile's still empty look-back record into L2 now.  The successors poll it before it is published, and
        // the lines zeroed at the start of the sort were evicted by the key stream long ago: without this their first
        // polls go to DRAM (measured: 0.663 -> 0.646 ms per pass).
        // (r = 8 only: with the 64-byte records of r = 4 it made the pass 5 % slower, 0.537 -> 0.567 ms)
        if constexpr (H >= 256)
            if (tid < (uint32_t)H / 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(lb_row + 8 * tid));  // one per 32-byte sector

        if (warp < (uint32_t)SW) {
            // ================= scan warps: totals -> bucket starts -> exclusive lane prefix =================
            const uint32_t q = lane & 7u;
            uint32_t total[GPW], below[GPW];
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                const uint32_t row = (uint32_t)(g * SW + warp) * 32u + lane;
                total[g] = 0;
                below[g] = 0;
                if (row < (uint32_t)H) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + row * 32u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t grp = (q + k) & 7u;
                        const uint4 v = r4[grp];
                        const uint32_t t = v.x + v.y + v.z + v.w;         // both halves at once (no carry, see EVEN)
                        const uint32_t s = (t & 0xFFFFu) + (t >> 16);
                        total[g] += s;
                        if (grp < q) below[g] += s;
                    }
                }
            }
            uint32_t start[GPW];
            uint32_t carry = 0;
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                uint32_t incl = total[g];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                start[g] = incl - total[g];
                if (lane == 31) s_misc[g * SW + warp] = incl;
            }
            if (SW > 1) named_bar_sync(kBarScan, SW * 32); else __syncwarp();
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                uint32_t prefix = carry;
#pragma unroll
                for (int w = 0; w < SW; ++w) {
                    const uint32_t part = s_misc[g * SW + w];
                    if ((uint32_t)w < warp) prefix += part;
                    carry += part;
                }
                start[g] += prefix;
            }
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                const uint32_t row = (uint32_t)(g * SW + warp) * 32u + lane;
                if (row < (uint32_t)H) {
                    s_tot[row] = total[g] >> 2;
                    s_dp[row] = start[g] >> 2;
                }
            }
            named_bar_arrive(kBarTot, (SW + LBW) * 32);
            if (warp == 0) LSD_TRACE(4);
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                const uint32_t row = (uint32_t)(g * SW + warp) * 32u + lane;
                if (row < (uint32_t)H) {
                    uint4* r4 = reinterpret_cast<uint4*>(s_mat + row * 32u);
                    uint32_t run = start[g] + below[g];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t grp = (q + k) & 7u;
                        if (grp == 0) run = start[g];
                        const uint4 v = r4[grp];
                        uint4 o;  // low half: where the even warps' keys of this cell start; high half: the odd warps'
                        o.x = run | ((run + (v.x & 0xFFFFu)) << 16); run += (v.x & 0xFFFFu) + (v.x >> 16);
                        o.y = run | ((run + (v.y & 0xFFFFu)) << 16); run += (v.y & 0xFFFFu) + (v.y >> 16);
                        o.z = run | ((run + (v.z & 0xFFFFu)) << 16); run += (v.z & 0xFFFFu) + (v.z >> 16);
                        o.w = run | ((run + (v.w & 0xFFFFu)) << 16); run += (v.w & 0xFFFFu) + (v.w >> 16);
                        r4[grp] = o;
                    }
                }
            }
            if (SW > 1) named_bar_sync(kBarScan, SW * 32);
            if (warp == 0) LSD_TRACE(5);
        } else if (warp >= (uint32_t)(WARPS - LBW)) {
            // ================= look-back warps (tail of the rank chain): one digit pair per thread =================
            named_bar_sync(kBarTot, (SW + LBW) * 32);
            if (warp == (uint32_t)WARPS - 1) LSD_TRACE(8);
            const uint32_t dt = tid - (uint32_t)(THREADS - LBT);
            if constexpr (QLB) {
                [[maybe_unused]] uint32_t q_rounds = 0, q_hops = 0;
                lookback_quad_tile<H, LBW, LB>(a, lb_row, tile, warp - (uint32_t)(WARPS - LBW), lane, pads, (uint32_t)(H - 1), s_tot, s_dp,
                                               s_gbase, TRACE ? &q_rounds : nullptr, TRACE ? &q_hops : nullptr);
                if constexpr (TRACE)
                    if (a.trace && warp == (uint32_t)WARPS - 1 && lane == 0) {
                        a.trace[(size_t)tile * 16 + 13] = q_rounds;
                        a.trace[(size_t)tile * 16 + 14] = q_hops;
                    }
            } else
            if (dt < (uint32_t)H / 2) {
            const uint32_t cnt_lo = s_tot[2 * dt];
            uint32_t cnt_hi = s_tot[2 * dt + 1];
            if (dt == (uint32_t)H / 2 - 1) cnt_hi -= pads;
            const uint32_t dp_lo = s_dp[2 * dt], dp_hi = s_dp[2 * dt + 1];
            uint32_t ex_lo = 0, ex_hi = 0;
            if (tile == 0) {
                st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | cnt_lo, kLbGlobal | cnt_hi);
            } else {
                st_relaxed_gpu_v2(lb_row + 2 * dt, kLbLocal | cnt_lo, kLbLocal | cnt_hi);
                const uint32_t* p = lb_row - H + 2 * dt;
                uint32_t remaining = tile;
                bool done = false;
                [[maybe_unused]] uint32_t dbg_rounds = 0, dbg_hops = 0;
                [[maybe_unused]] long long dbg_wait = 0, dbg_proc = 0;
                while (!done) {
                    if constexpr (TRACE) ++dbg_rounds;
                    [[maybe_unused]] const long long t_a = TRACE ? clock64() : 0;
                    uint2 w[LB];
#pragma unroll
                    for (int k = 0; k < LB; ++k)
                        w[k] = (uint32_t)k < remaining ? ld_relaxed_gpu_v2(p - (size_t)k * H) : make_uint2(0u, 0u);
                    [[maybe_unused]] long long t_b = 0;
                    if constexpr (TRACE) {  // all loads of the round have landed
                        uint32_t all = 0;
#pragma unroll
                        for (int k = 0; k < LB; ++k) all |= w[k].x | w[k].y;
                        asm volatile("" ::"r"(all) : "memory");
                        t_b = clock64();
                        dbg_wait += t_b - t_a;
                    }
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
                    if constexpr (TRACE) {
                        dbg_hops += consumed;
                        asm volatile("" ::"r"(consumed), "r"(ex_lo) : "memory");
                        dbg_proc += clock64() - t_b;
                    }
                }
                if constexpr (TRACE)
                    if (a.trace && warp == (uint32_t)WARPS - 1 && lane == 0) {
                        a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                        a.trace[(size_t)tile * 16 + 14] = dbg_hops;
                        a.trace[(size_t)tile * 16 + 15] = (unsigned long long)dbg_wait;  // sum over rounds: loads issued -> all landed
                        a.trace[(size_t)tile * 16 + 7] = (unsigned long long)dbg_proc;   // sum over rounds: landed -> round done
                    }
                st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
            }
            const uint64_t b_lo = a.bases_in[2 * dt], b_hi = a.bases_in[2 * dt + 1];
            s_gbase[2 * dt] = (uint32_t)b_lo + ex_lo - dp_lo;
            s_gbase[2 * dt + 1] = (uint32_t)b_hi + ex_hi - dp_hi;
            if (a.bases_out != nullptr && tile == a.tiles - 1) {
                a.bases_out[2 * dt] = b_lo + ex_lo + cnt_lo;
                a.bases_out[2 * dt + 1] = b_hi + ex_hi + cnt_hi;
            }
            }
        }

        if (warp == (uint32_t)WARPS - 1) LSD_TRACE(9);  // look-back done (last warp)
        // ---- 2. two rank chains: even warps on the low halves, odd warps on the high halves; the scatter of a key
        // ---- follows its atomic one step behind (a warp has one returning shared atomic in flight anyway) ----
        if (warp >= 2u) named_bar_sync(warp, 64);
        {
            char* kb = reinterpret_cast<char*>(s_keys);
            uint32_t prev = 0;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), inc4);
                if (i > 0) *reinterpret_cast<uint32_t*>(kb + (half ? (prev >> 16) : (prev & 0xFFFFu))) = key[i - 1];
                prev = old;
            }
            if (warp + 2u < (uint32_t)WARPS) named_bar_arrive(warp + 2u, 64);
            *reinterpret_cast<uint32_t*>(kb + (half ? (prev >> 16) : (prev & 0xFFFFu))) = key[ITEMS - 1];
        }
        if (warp == 0) LSD_TRACE(6);
        if (warp == (uint32_t)WARPS - 1) LSD_TRACE(10);
        // generic accesses to the matrix / tile counts are ordered before the async-proxy write of the next tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // reorder buffer complete; the matrix is dead
        if (warp == 0) LSD_TRACE(11);

        // ---- 3. next ticket + prefetch into the dead matrix, then stream this tile out ----
        if (tid == 0) {
            fetch_next();
            if constexpr (NOB5) mbar_arrive_release(s_bar2);
        }
        if (valid == (uint32_t)TILE && !typed_out) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t p = i * THREADS + tid;
                const uint32_t k = s_keys[p];
                st_key<5>(out + s_gbase[(k >> shift) & (H - 1)] + p, k);
            }
        } else {
            for (uint32_t p = tid; p < valid; p += THREADS) {
                const uint32_t k = s_keys[p];
                out[s_gbase[(k >> shift) & (H - 1)] + p] = TYPED ? key_from_unsigned(k, xout) : k;
            }
        }
        if (warp == 0) LSD_TRACE(12);
#undef LSD_TRACE
        // next ticket visible; the reorder buffer and the bucket bases are not written again before the next tile's
        // "every key is in registers" barrier, so with the mbarrier hand-over no barrier is needed here
        if constexpr (!NOB5) __syncthreads();
    }
}

template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CLR, int NOB5, bool TYPED, bool TRACE>
int onesweep_lpc4_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpc4_kernel<RB, WARPS, ITEMS, MINB, SHIFT, LB, CLR, NOB5, TYPED, TRACE>;
    constexpr bool ALIAS = S_::OFF_DP - S_::OFF_MAT >= S_::TILE;
    constexpr size_t SMEM = ALIAS ? S_::SMEM_BYTES : sizeof(uint32_t) * (((S_::WORDS + 3) & ~3) + S_::TILE) + 16;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const uint32_t resident = (uint32_t)sm_count() * MINB;
    const uint32_t grid = a.tiles < resident ? a.tiles : resident;  // persistent: every CTA loops over tickets
    kern<<<grid, S_::THREADS, SMEM, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CLR, int NOB5, bool TYPED = false, bool TRACE = false>
int onesweep_lpc4_launch(const PassArgs& a, cudaStream_t s)
{
    if constexpr (RB != 8) return onesweep_lpc4_launch_shift<RB, WARPS, ITEMS, MINB, -1, LB, CLR, NOB5, TYPED, TRACE>(a, s);
    switch (a.shift) {
        case 0: return onesweep_lpc4_launch_shift<RB, WARPS, ITEMS, MINB, 0, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 8: return onesweep_lpc4_launch_shift<RB, WARPS, ITEMS, MINB, 8, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 16: return onesweep_lpc4_launch_shift<RB, WARPS, ITEMS, MINB, 16, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 24: return onesweep_lpc4_launch_shift<RB, WARPS, ITEMS, MINB, 24, LB, CLR, NOB5, TYPED, TRACE>(a, s);
    }
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeLpc4 = 8;

// FORMS: 0 = plain passes only (tuning variants); 1 = the r = 8 default entry: plain and typed-key passes on the persistent
// kernel, peer-scatter and key-value passes on onesweep_lpc32_kernel (same tile size, workspace layout and look-back protocol);
// 2 = plain and typed-key passes (the r < 8 default entries; key-value sorts there use the warp-multisplit entries).
template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CLR = 0, int NOB5 = 0, int FORMS = 0, bool TRACE = false>
constexpr OnesweepLauncher make_lpc4_launcher()
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    if constexpr (FORMS == 1)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc4, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc4_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPeer, false>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPairs, false>,
                                &onesweep_lpc4_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, true>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPairsTyped, false>};
    else if constexpr (FORMS == 2)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc4, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc4_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5>, nullptr, nullptr,
                                &onesweep_lpc4_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, true>, nullptr};
    else
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc4, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc4_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, false, TRACE>, nullptr, nullptr, nullptr,
                                nullptr};
}

}  // namespace lsd
