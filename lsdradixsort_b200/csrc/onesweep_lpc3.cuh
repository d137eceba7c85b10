// onesweep_lpc3.cuh -- LPC32 pass as a PERSISTENT kernel that prefetches its next tile.
//
// In onesweep_lpc32_kernel a CTA lives for one tile: ticket, TMA load (~2 K cycles of DRAM latency under load), rank,
// look-back, copy-out, exit.  The counter matrix is dead from the end of the rank chain on, while the reorder buffer is
// still being streamed out.  Here a CTA loops over tickets and, right after the barrier that ends the ranking of tile k,
// takes the ticket of tile k+1 and lets the TMA engine load it INTO THE DEAD MATRIX (+ tile-count area; together they
// hold a tile) while tile k streams out.  Price: the matrix can only be cleared after the keys have been read out of it
// (two more CTA-wide barriers per tile, and the clear no longer hides behind the load).
// Plain and typed-key passes; same tile, look-back protocol and workspace layout as onesweep_lpc32_kernel.
// Digits narrower than 8 bits (r = 1, 2, 4: the reference's other radix settings) have a counter matrix far smaller than a
// tile, so there the prefetch goes to a dedicated second tile buffer (the shared-memory budget per CTA is the same) and the
// digit shift is a run-time value (SHIFT < 0) to keep the number of instantiations down.
#pragma once
#include "onesweep_lpc32.cuh"
#include "lookback_quad.cuh"

namespace lsd {

// 32 KiB of zeros in global memory: source of the TMA zero-fill of the counter matrix (CLR == 2)
static __device__ __align__(128) uint4 g_lsd_zero_page[2048];

__device__ __forceinline__ void mbar_arrive_release(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bulk (TMA) store shared -> global, completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// CLR: 8 = as 0, and the look-back records are PUBLISHED by TMA bulk stores from a shared-memory copy of the record (they
//      do not queue behind the copy-out stores of the three resident CTAs in the LSU pipe)
// CLR: 0 = matrix cleared with 128-bit stores, 1 = st.bulk, 2 = TMA copy of a zero page (keeps the clear off the LSU pipe)
// CLR: 9 = as 0 with the pipelined look-back walk; 10 = as 0 with the quad look-back of lookback_quad.cuh (LB load instructions
//      per round, each fetching 32 / (quads per look-back warp) records: the default of r = 4 and r = 1)
// NOB5: 1 = no CTA-wide barrier at the end of a tile: the next ticket is handed over through a second mbarrier
// TYPED: i32 / f32 keys are mapped to unsigned order when the first executed pass reads them and back when the last one writes
// TRACE: per-tile phase clocks into PassArgs.trace (bench_tools/trace.py).  Compile-time: the run-time checks and the clock
//        reads alone cost ~7 % of the pass (0.665 -> 0.715 ms), so only the tuning variant carries them.
template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CLR, int NOB5 = 0, bool TYPED = false, bool TRACE = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
onesweep_lpc3_kernel(const PassArgs a)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, THREADS = S_::THREADS, S = S_::S, TILE = S_::TILE;
    constexpr int SW = S_::SW, GPW = S_::GPW, LBT = S_::LBT, LBW = S_::LBW;
    constexpr uint32_t kBarTot = 14, kBarScan = 15;
    constexpr bool ALIAS = S_::OFF_DP - S_::OFF_MAT >= TILE;  // the dead matrix (+ tile counts) can hold the incoming tile
    constexpr int IN_OFF = (S_::WORDS + 3) & ~3;               // else: a dedicated prefetch buffer behind everything
    static_assert(LBT >= H / 2, "one digit pair per look-back thread");
    constexpr bool PUB = CLR == 8;
    constexpr bool WALK2 = CLR == 9;  // pipelined look-back walk: next window in flight, whole-window fast path
    constexpr bool QLB = CLR == 10;   // quad look-back (lookback_quad.cuh): 128-bit record accesses, the window spread over lanes
    static_assert(!PUB || (H == 256 && S_::OFF_HEADS + H + 4 - S_::OFF_DST >= 4 * H), "record copies live in the unused peer-scatter area");
    const int shift = SHIFT >= 0 ? SHIFT : a.shift;
    // run-time digits may be narrower than RB bits (sub-passes of the composite digit widths, sort.cu: pass_enqueue_wide)
    const uint32_t dmask = SHIFT >= 0 ? (uint32_t)(H - 1) : a.digit_mask;

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
    uint32_t* s_rec = smem + S_::OFF_DST;  // PUB: [tile parity][LOCAL, INCLUSIVE][H] record copies, the source of the bulk stores

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

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
        else return (((key >> shift) & dmask) << 7) | lane4;
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
            const uint32_t* src = s_in + lane * S + warp * ITEMS;
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
            atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), 4u);
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
                        const uint32_t s = v.x + v.y + v.z + v.w;
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
                        uint4 o;
                        o.x = run; run += v.x;
                        o.y = run; run += v.y;
                        o.z = run; run += v.z;
                        o.w = run; run += v.w;
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
                lookback_quad_tile<H, LBW, LB>(a, lb_row, tile, warp - (uint32_t)(WARPS - LBW), lane, pads, dmask, s_tot, s_dp, s_gbase,
                                               TRACE ? &q_rounds : nullptr, TRACE ? &q_hops : nullptr);
                if constexpr (TRACE)
                    if (a.trace && warp == (uint32_t)WARPS - 1 && lane == 0) {
                        a.trace[(size_t)tile * 16 + 13] = q_rounds;
                        a.trace[(size_t)tile * 16 + 14] = q_hops;
                    }
            } else
            if (dt < (uint32_t)H / 2) {
            const uint32_t cnt_lo = s_tot[2 * dt];
            uint32_t cnt_hi = s_tot[2 * dt + 1];
            if (2 * dt + 1 == dmask) cnt_hi -= pads;  // the pads of a ragged tile carry the largest digit in use
            const uint32_t dp_lo = s_dp[2 * dt], dp_hi = s_dp[2 * dt + 1];
            uint32_t ex_lo = 0, ex_hi = 0;
            [[maybe_unused]] uint32_t* rec = s_rec + (tile & 1u) * (2 * H);
            if constexpr (PUB) {
                // the bulk stores this thread issued for the previous tile have read their source (the same-parity copy is
                // rewritten one tile later, behind that tile's CTA-wide barriers)
                if (lane == 0) tma_bulk_wait_read_all();
                const uint32_t flag = tile == 0 ? kLbGlobal : kLbLocal;
                *reinterpret_cast<uint2*>(rec + 2 * dt) = make_uint2(flag | cnt_lo, flag | cnt_hi);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(13, H / 2);
                if (dt == 0) tma_bulk_s2g(lb_row, rec, H * 4);
            }
            if (tile == 0) {
                if constexpr (!PUB) st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | cnt_lo, kLbGlobal | cnt_hi);
            } else {
                if constexpr (!PUB) st_relaxed_gpu_v2(lb_row + 2 * dt, kLbLocal | cnt_lo, kLbLocal | cnt_hi);
                const uint32_t* p = lb_row - H + 2 * dt;
                uint32_t remaining = tile;
                bool done = false;
                [[maybe_unused]] uint32_t dbg_rounds = 0, dbg_hops = 0;
                [[maybe_unused]] long long dbg_wait = 0, dbg_proc = 0;
                if constexpr (WALK2) {
                    // A round of the plain walk costs ~1.2 K cycles, most of it the dependent consume chain of the window
                    // issued among ~27 resident warps, and the walk takes (hops / LB) rounds.  Here the NEXT window's loads
                    // are in flight while this one is consumed, and a window whose LB records are all LOCAL is consumed
                    // by two adds (LB = 4 LOCAL flags sum to 0 mod 2^32).
                    static_assert(LB == 4, "the flag bits of four LOCAL words cancel");
                    uint2 cur[LB], nxt[LB];
#pragma unroll
                    for (int k = 0; k < LB; ++k)
                        cur[k] = (uint32_t)k < remaining ? ld_relaxed_gpu_v2(p - (size_t)k * H) : make_uint2(0u, 0u);
                    while (!done) {
                        if constexpr (TRACE) ++dbg_rounds;
                        const bool more = remaining > (uint32_t)LB;
                        if (more) {
#pragma unroll
                            for (int k = 0; k < LB; ++k)
                                nxt[k] = (uint32_t)(LB + k) < remaining ? ld_relaxed_gpu_v2(p - (size_t)(LB + k) * H) : make_uint2(0u, 0u);
                        }
                        const uint32_t lo = min(min(cur[0].x, cur[1].x), min(cur[2].x, cur[3].x));
                        const uint32_t any = cur[0].x | cur[1].x | cur[2].x | cur[3].x;
                        if (lo != 0u && (any & kLbGlobal) == 0u) {  // four LOCAL records (then `more` holds: tile 0 is GLOBAL)
                            ex_lo += cur[0].x + cur[1].x + cur[2].x + cur[3].x;
                            ex_hi += cur[0].y + cur[1].y + cur[2].y + cur[3].y;
                            p -= (size_t)LB * H;
                            remaining -= (uint32_t)LB;
                            if constexpr (TRACE) dbg_hops += LB;
#pragma unroll
                            for (int k = 0; k < LB; ++k) cur[k] = nxt[k];
                            continue;
                        }
                        uint32_t consumed = 0;
#pragma unroll
                        for (int k = 0; k < LB; ++k) {
                            if (!done && consumed == (uint32_t)k && cur[k].x != 0) {
                                ex_lo += cur[k].x & kLbValueMask;
                                ex_hi += cur[k].y & kLbValueMask;
                                ++consumed;
                                if (cur[k].x & kLbGlobal) done = true;
                            }
                        }
                        if constexpr (TRACE) dbg_hops += consumed;
                        if (!done) {  // the window moved by less than LB records: fetch it again from where the walk stands
                            p -= (size_t)consumed * H;
                            remaining -= consumed;
#pragma unroll
                            for (int k = 0; k < LB; ++k)
                                cur[k] = (uint32_t)k < remaining ? ld_relaxed_gpu_v2(p - (size_t)k * H) : make_uint2(0u, 0u);
                        }
                    }
                }
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
                if constexpr (PUB) {  // one 256-byte bulk store per look-back warp (64 digits), when its slowest lane is done
                    *reinterpret_cast<uint2*>(rec + H + 2 * dt) = make_uint2(kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tma_bulk_s2g(lb_row + 2 * dt, rec + H + 2 * dt, 256);
                } else {
                    st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
                }
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
        // ---- 2. rank chain ----
        uint32_t rk[(ITEMS + 1) / 2];
        if (warp > 0) named_bar_sync(warp, 64);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), 4u);
            if (i & 1) rk[i >> 1] = __byte_perm(rk[i >> 1], old, 0x5410); else rk[i >> 1] = old;
        }
        if (warp + 1 < (uint32_t)WARPS) named_bar_arrive(warp + 1, 64);
        if (warp == 0) LSD_TRACE(6);
        if (warp == (uint32_t)WARPS - 1) LSD_TRACE(10);
        {
            char* kb = reinterpret_cast<char*>(s_keys);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t off = (i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu);
                *reinterpret_cast<uint32_t*>(kb + off) = key[i];
            }
        }
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
                st_key<5>(out + s_gbase[(k >> shift) & dmask] + p, k);
            }
        } else {
            for (uint32_t p = tid; p < valid; p += THREADS) {
                const uint32_t k = s_keys[p];
                out[s_gbase[(k >> shift) & dmask] + p] = TYPED ? key_from_unsigned(k, xout) : k;
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
int onesweep_lpc3_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpc3_kernel<RB, WARPS, ITEMS, MINB, SHIFT, LB, CLR, NOB5, TYPED, TRACE>;
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
int onesweep_lpc3_launch(const PassArgs& a, cudaStream_t s)
{
    if constexpr (RB != 8) return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, -1, LB, CLR, NOB5, TYPED, TRACE>(a, s);
    if (a.digit_mask == 0xFFu) switch (a.shift) {
        case 0: return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, 0, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 8: return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, 8, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 16: return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, 16, LB, CLR, NOB5, TYPED, TRACE>(a, s);
        case 24: return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, 24, LB, CLR, NOB5, TYPED, TRACE>(a, s);
    }
    // any other shift / a digit narrower than 8 bits: the run-time form (plain keys, no trace)
    if constexpr (!TYPED && !TRACE) return onesweep_lpc3_launch_shift<RB, WARPS, ITEMS, MINB, -1, LB, CLR, NOB5, false, false>(a, s);
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeLpc3 = 6;

// FORMS: 0 = plain passes only (tuning variants); 1 = the r = 8 default entry: plain and typed-key passes on the persistent
// kernel, peer-scatter and key-value passes on onesweep_lpc32_kernel (same tile size, workspace layout and look-back protocol);
// 2 = plain and typed-key passes (the r < 8 default entries; key-value sorts there use the warp-multisplit entries).
template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CLR = 0, int NOB5 = 0, int FORMS = 0, bool TRACE = false>
constexpr OnesweepLauncher make_lpc3_launcher()
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    if constexpr (FORMS == 1)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc3, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc3_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPeer, false>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPairs, false>,
                                &onesweep_lpc3_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, true>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, 5, kPassPairsTyped, false>};
    else if constexpr (FORMS == 2)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc3, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc3_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5>, nullptr, nullptr,
                                &onesweep_lpc3_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, true>, nullptr};
    else
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc3, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc3_launch<RB, WARPS, ITEMS, MINB, LB, CLR, NOB5, false, TRACE>, nullptr, nullptr, nullptr,
                                nullptr};
}

}  // namespace lsd
