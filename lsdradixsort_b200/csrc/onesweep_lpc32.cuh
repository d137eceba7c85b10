// onesweep_lpc32.cuh -- LPC onesweep pass, third shape: 32-bit lane-private counters that hold BYTE
// offsets, compile-time digit shift, vectorised diagonal scan, role-specific barriers.
//
// Why (ncu, profiles/r01_lpc_*): in onesweep_lpc.cuh the rank chain -- the one inherently serial part,
// one warp at a time taking its ranks from the counter matrix -- spends ~10 instructions per key on
// unpacking 16-bit counter pairs and on a run-time digit shift.  Here
//   * cnt[digit][lane] is a full 32-bit word (matrix 256 x 32 x 4 B = 32 KiB for 8-bit digits) and counts
//     in units of 4, so the value an atomicAdd returns IS the byte address of the key in the reorder
//     buffer: no unpack, no multiply;
//   * the digit shift is a template parameter: cell address = 2 instructions (shift, mask|lane);
//     => count = 3 instructions per key, rank = 4 (incl. packing two 16-bit offsets per register);
//   * the matrix scan reads rows with 128-bit shared loads along a diagonal of 16-byte groups
//     (lane r starts at group r & 7): conflict-free per quarter-warp, 4x fewer instructions;
//   * only the warps that need a hand-over synchronise: scan warps among themselves, scan -> look-back
//     warps, the rank chain; two CTA-wide barriers remain (after counting, before streaming out).
// Everything else (TMA-staged tile, lane-blocked ownership, early windowed look-back by the tail warps,
// shared-memory reorder, coalesced per-bucket scatter) is as in onesweep_lpc.cuh.
#pragma once
#include "onesweep_lpc.cuh"

namespace lsd {

template <int RB, int WARPS, int ITEMS>
struct Lpc32Shape {
    static constexpr int H = 1 << RB;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int S = WARPS * ITEMS;
    static constexpr int TILE = 32 * S;
    static constexpr int ROW_GROUPS = (H + 31) / 32;                   // 32 matrix rows (digits) per group
    static constexpr int SW = ROW_GROUPS < 4 ? ROW_GROUPS : 4;         // scan warps 0..SW-1
    static constexpr int GPW = (ROW_GROUPS + SW - 1) / SW;             // row groups per scan warp
    static constexpr int LBT = H / 2 < 32 ? 32 : H / 2;                // look-back threads (one digit pair each)
    static constexpr int LBW = LBT / 32;                               // look-back warps = the LAST warps of the CTA
    static_assert(S % 2 == 1, "S = WARPS*ITEMS must be odd (conflict-free lane-blocked reads)");
    static_assert(TILE * 4 < 65536, "reorder-buffer byte offsets are packed in 16 bits");
    static_assert(WARPS >= SW + LBW, "scan warps and look-back warps must be disjoint");
    static_assert(WARPS <= 13, "named barriers: 1..WARPS-1 rank chain, 14 totals hand-over, 15 scan warps");
    static constexpr int OFF_MAT = TILE;                 // [H][32]
    static constexpr int OFF_TOT = OFF_MAT + H * 32;     // [H] tile digit counts (keys)
    static constexpr int OFF_DP = OFF_TOT + H;           // [H] tile-local bucket starts (keys)
    static constexpr int OFF_GBASE = OFF_DP + H;         // [H]
    static constexpr int OFF_MISC = OFF_GBASE + H;       // [0..15] partials, [32] tile id, [34..35] mbarrier
    static constexpr int OFF_DST = OFF_MISC + 64;        // [H] 64-bit bucket destination pointers (peer-scatter mode only)
    static constexpr int OFF_SEG = OFF_DST + 2 * H;      // [H] first | last << 16 bucket of the bucket's destination segment
    static constexpr int OFF_HEADS = OFF_SEG + H;        // [H] compact list of segment words, [H] = number of segments
    static constexpr int WORDS = OFF_HEADS + H + 4;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
};

// byte offset of cell (digit, lane) inside the matrix: digit * 128 + lane * 4, with the digit taken at
// a compile-time shift so that it costs one shift and one LOP3.
template <int RB, int SHIFT>
__device__ __forceinline__ uint32_t cell_offset(uint32_t key, uint32_t lane4)
{
    constexpr uint32_t mask = ((1u << RB) - 1u) << 7;
    uint32_t x;
    if constexpr (SHIFT >= 7) x = key >> (SHIFT - 7);
    else x = key << (7 - SHIFT);
    return (x & mask) | lane4;
}

// key store of the plain copy-out; CLR >= 2 selects a cache operator (tuning variants, see onesweep_r8.cu)
template <int CLR>
__device__ __forceinline__ void st_key(uint32_t* p, uint32_t v)
{
    if constexpr (CLR == 2) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if constexpr (CLR == 3) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if constexpr (CLR == 4) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if constexpr (CLR == 5) asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else *p = v;
}

template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CLR, int MODE, bool SP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
onesweep_lpc32_kernel(const PassArgs a)
{
    constexpr bool PEER = MODE == kPassPeer;    // bucket-pointer scatter (multi-GPU exchange)
    constexpr bool PAIRS = MODE == kPassPairs || MODE == kPassPairsTyped;  // a 32-bit value travels with every key
    constexpr bool TYPED = MODE == kPassTyped || MODE == kPassPairsTyped;  // i32 / f32 keys: mapped to unsigned order
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, THREADS = S_::THREADS, S = S_::S, TILE = S_::TILE;
    constexpr int SW = S_::SW, GPW = S_::GPW, LBT = S_::LBT, LBW = S_::LBW;
    constexpr uint32_t kBarTot = 14, kBarScan = 15;  // LB = look-back window (rows fetched per round trip)

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_keys = smem;
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_gbase = smem + S_::OFF_GBASE;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 34);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    // the per-tile phase trace (bench_tools/trace.py) is compiled into the plain CLR == 0 tuning variants only: its run-time
    // checks and clock reads alone cost several per cent of a pass
    constexpr bool TRACE = CLR == 0 && MODE == kPassPlain;
    const long long t_start = (TRACE && a.trace) ? clock64() : 0;
#define LSD_TRACE(slot)                                                                      \
    do {                                                                                     \
        if constexpr (TRACE)                                                                 \
            if (a.trace && lane == 0) a.trace[(size_t)tile * 16 + (slot)] = (unsigned long long)(clock64() - t_start); \
    } while (0)

    // ---- 0. ticket, TMA bulk load, clear the matrix ----
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
    if constexpr (PEER) {
        // multi-GPU exchange: bucket d goes to dst_ptrs[d][rank within bucket] -- local or peer (NVLink) memory
        // Buckets that share a destination SEGMENT (seg[d] = first | last << 16) are written as one contiguous block
        // per tile -- long runs, whole lines over NVLink -- in (tile, bucket, position) order.
        uint64_t* s_dst = reinterpret_cast<uint64_t*>(smem + S_::OFF_DST);
        uint32_t* s_seg = smem + S_::OFF_SEG;
        for (uint32_t i = tid; i < (uint32_t)H; i += THREADS) {
            s_dst[i] = a.dst_ptrs[i];
            s_seg[i] = a.dst_seg ? a.dst_seg[i] : (i | (i << 16));
        }
        if (warp == 1) {  // compact list of the segments (their heads), for the line-aligned copy-out
            uint32_t* s_heads = smem + S_::OFF_HEADS;
            uint32_t count = 0;
            for (uint32_t c = 0; c < (uint32_t)H; c += 32) {
                const uint32_t d = c + lane;
                const uint32_t g = d < (uint32_t)H ? (a.dst_seg ? a.dst_seg[d] : (d | (d << 16))) : 0xFFFFFFFFu;
                const bool head = d < (uint32_t)H && (g & 0xFFFFu) == d;
                const uint32_t m = __ballot_sync(kFullMask, head);
                if (head) s_heads[count + __popc(m & ((1u << lane) - 1u))] = g;
                count += __popc(m);
            }
            if (lane == 0) s_heads[H] = count;
        }
    }
    if constexpr (CLR == 1) {
        // zero-fill by the uniform datapath (UMEMSETS): one instruction instead of H*8 128-bit stores through the LSU
        if (tid == 32) {
            asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(smem_u32(s_mat)), "l"((uint64_t)(H * 128)) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    } else {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
        for (uint32_t i = tid; i < H * 8; i += THREADS) m4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = s_misc[32];
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t left = a.portion_keys - tile_base;
    const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
    const uint32_t pads = (uint32_t)TILE - valid;

    if (warp == 0) LSD_TRACE(0);  // ticket + matrix clear done
    if (valid == (uint32_t)TILE) {
        mbar_wait(s_bar, 0);
    } else {
        // pads are the pre-image of 0xFFFFFFFF under this pass's input mapping (identity unless the keys are typed)
        const uint32_t pad_key = TYPED ? key_from_unsigned(0xFFFFFFFFu, pass_xform_in(a)) : 0xFFFFFFFFu;
        for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS) s_keys[p] = p < valid ? in[tile_base + p] : pad_key;
        __syncthreads();
    }
    if (warp == 0) LSD_TRACE(1);  // tile landed

    // ---- 1. lane-blocked read + count (cells count bytes: +4 per key) ----
    uint32_t key[ITEMS];
    {
        const uint32_t* src = s_keys + lane * S + warp * ITEMS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = src[i];
    }
    if constexpr (TYPED) {  // typed keys (i32 / f32) enter unsigned order when the first executed pass reads them
        const KeyXform xin = pass_xform_in(a);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = key_to_unsigned(key[i], xin);
    }
    char* mat_bytes = reinterpret_cast<char*>(s_mat);
    const uint32_t lane4 = lane << 2;
    // SHIFT < 0: the digit position is a run-time value (the exchange window of lsd_sort_multi: any 8 bits of the key)
    const int shift = SHIFT >= 0 ? SHIFT : a.shift;
    auto cell_of = [&](uint32_t k) -> uint32_t {
        if constexpr (SHIFT >= 0) return cell_offset<RB, SHIFT < 0 ? 0 : SHIFT>(k, lane4);
        else return (((k >> shift) & (uint32_t)(H - 1)) << 7) | lane4;
    };
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
        atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), 4u);
    if (warp == 0) LSD_TRACE(2);  // warp 0 issued its counts
    __syncthreads();  // counts complete; all keys are in registers: s_keys is now the reorder buffer
    if (warp == 0) LSD_TRACE(3);  // count barrier passed

    uint32_t* lb_row = a.lookback + (size_t)tile * H;
    // pull this tile's still empty look-back record into L2: successors poll it before it is published (see onesweep_lpc3.cuh)
    if (tid < (uint32_t)(H + 7) / 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(lb_row + 8 * tid));  // one per 32-byte sector

    if constexpr (SP) {
        // ================= single-pass scan: warps 0..7, one matrix row per thread, the row stays in registers
        // between the totals and the prefix write (one 128-bit read pass less: ~1 wavefront per 32 keys) =============
        static_assert(!SP || (H == 256 && WARPS >= 9), "single-pass scan: 256 rows over 8 warps, look-back by the last 4");
        const uint32_t q = lane & 7u;
        const uint32_t row = warp * 32u + lane;
        uint4 v[8];
        uint32_t total = 0, below = 0, start = 0;
        if (warp < 8u) {
            const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + row * 32u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                v[k] = r4[grp];
                const uint32_t sum = v[k].x + v[k].y + v[k].z + v[k].w;
                total += sum;
                if (grp < q) below += sum;
            }
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            start = incl - total;
            if (lane == 31) s_misc[warp] = incl;
            s_tot[row] = total >> 2;
        }
        __syncthreads();  // totals visible
        if (warp < 8u) {
            uint32_t prefix = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if ((uint32_t)w < warp) prefix += s_misc[w];
            start += prefix;
            s_dp[row] = start >> 2;
            uint4* r4 = reinterpret_cast<uint4*>(s_mat + row * 32u);
            uint32_t run = start + below;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                if (grp == 0) run = start;
                uint4 o;
                o.x = run; run += v[k].x;
                o.y = run; run += v[k].y;
                o.z = run; run += v[k].z;
                o.w = run; run += v[k].w;
                r4[grp] = o;
            }
        }
        __syncthreads();  // matrix, totals and bucket starts complete
        if (warp == 0) LSD_TRACE(5);
    }
    if (!SP && warp < (uint32_t)SW) {
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
                    if (grp < q) below[g] += s;  // groups reached after the wrap == columns [0, 4q)
                }
            }
        }
        // exclusive scan of the H row totals, rows ascending (group-major: g, warp, lane)
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
        if (SW > 1 || GPW > 1) {
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
        }
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
            const uint32_t row = (uint32_t)(g * SW + warp) * 32u + lane;
            if (row < (uint32_t)H) {
                s_tot[row] = total[g] >> 2;
                s_dp[row] = start[g] >> 2;
            }
        }
        named_bar_arrive(kBarTot, (SW + LBW) * 32);  // totals + starts are in shared memory: look-back warps may go
        if (warp == 0) LSD_TRACE(4);  // scan pass 1 + digit scan done
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
        if (SW > 1) named_bar_sync(kBarScan, SW * 32);  // matrix complete before warp 0 opens the rank chain
        if (warp == 0) LSD_TRACE(5);  // scan pass 2 done: rank chain opens
    } else if (warp >= (uint32_t)(WARPS - LBW)) {
        // ================= look-back warps (tail of the rank chain): one digit pair per thread =================
        if constexpr (!SP) named_bar_sync(kBarTot, (SW + LBW) * 32);
        if (warp == (uint32_t)WARPS - 1) LSD_TRACE(8);  // look-back starts
        const uint32_t dt = tid - (uint32_t)(THREADS - LBT);
        if (dt < (uint32_t)H / 2) {
            uint32_t cnt_lo = s_tot[2 * dt];
            uint32_t cnt_hi = s_tot[2 * dt + 1];
            if (dt == (uint32_t)H / 2 - 1) cnt_hi -= pads;  // pads of a ragged last tile are not keys
            uint32_t dp_lo = s_dp[2 * dt], dp_hi = s_dp[2 * dt + 1];
            if constexpr (PEER) {
                // count, look back and place per destination segment: every bucket of a segment carries the segment's
                // count, so the look-back below yields the segment's exclusive prefix for each of its buckets
                const uint32_t* s_seg = smem + S_::OFF_SEG;
                const uint32_t g_lo = s_seg[2 * dt], g_hi = s_seg[2 * dt + 1];
                const uint32_t e_lo = (g_lo >> 16) + 1u, e_hi = (g_hi >> 16) + 1u;
                dp_lo = s_dp[g_lo & 0xFFFFu];
                dp_hi = s_dp[g_hi & 0xFFFFu];
                cnt_lo = (e_lo < (uint32_t)H ? s_dp[e_lo] : valid) - dp_lo;
                cnt_hi = (e_hi < (uint32_t)H ? s_dp[e_hi] : valid) - dp_hi;
            }
            uint32_t ex_lo = 0, ex_hi = 0;
            if (tile == 0) {
                st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | cnt_lo, kLbGlobal | cnt_hi);
            } else {
                st_relaxed_gpu_v2(lb_row + 2 * dt, kLbLocal | cnt_lo, kLbLocal | cnt_hi);
                const uint32_t* p = lb_row - H + 2 * dt;
                uint32_t remaining = tile;
                bool done = false;
                uint32_t dbg_rounds = 0, dbg_hops = 0;
                while (!done) {
                    ++dbg_rounds;
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
                    dbg_hops += consumed;
                }
                if constexpr (TRACE)
                    if (a.trace && warp == (uint32_t)WARPS - 1 && lane == 0) {
                        a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                        a.trace[(size_t)tile * 16 + 14] = dbg_hops;
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

    // ---- 2. rank chain: the returned counter value is the key's byte offset in the reorder buffer ----
    uint32_t rk[(ITEMS + 1) / 2];
    if (warp > 0) named_bar_sync(warp, 64);
    // One returning shared atomic per key.  Measured alternatives on B200 (bench_tools/trace.py):
    // ld.shared + red.shared, and grouped plain ld/st with duplicates resolved in registers, both take the same
    // ~1000 cycles per warp turn (a warp has one shared atomic in flight at a time, ~33 cycles; the plain
    // version is instruction-bound instead), so the form with the fewest instructions stays.
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_of(key[i])), 4u);
        if (i & 1) rk[i >> 1] = __byte_perm(rk[i >> 1], old, 0x5410); else rk[i >> 1] = old;
    }
    if (warp + 1 < (uint32_t)WARPS) named_bar_arrive(warp + 1, 64);
    if (warp == 0) LSD_TRACE(6);                       // warp 0 took its ranks
    if (warp == (uint32_t)WARPS - 1) LSD_TRACE(10);    // last warp took its ranks: chain complete
    {
        char* kb = reinterpret_cast<char*>(s_keys);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t off = (i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu);
            *reinterpret_cast<uint32_t*>(kb + off) = key[i];
        }
    }
    if (warp == 0) LSD_TRACE(7);  // warp 0 scattered
    // key-value pass: the counter matrix (and the tile digit counts behind it) are dead after this barrier; the tile's
    // VALUES are staged there by a second TMA copy that runs while the keys stream out.  Every thread orders its generic
    // accesses to that region before the async-proxy write.
    if constexpr (PAIRS) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) LSD_TRACE(11);  // final barrier passed
    static_assert(!PAIRS || S_::OFF_DP - S_::OFF_MAT >= TILE, "the value staging area (matrix + tile counts) must hold a tile");
    if constexpr (PAIRS) {
        if (valid == (uint32_t)TILE && tid == 0) {
            const uint32_t* vsrc = (src_scratch ? a.vals_scratch : a.vals) + a.portion_base + tile_base;
            mbar_expect_tx(s_bar, TILE * 4);
            tma_bulk_g2s(s_mat, vsrc, TILE * 4, s_bar);
        }
    }

    // ---- 3. stream the reorder buffer out, coalesced per bucket ----
    // typed keys leave unsigned order when the last executed pass writes them (uniform over the grid)
    const bool typed_out = TYPED && a.plan->last_pass == (uint32_t)a.pass;
    const KeyXform xout = key_xform_of(typed_out ? a.key_type : 0u);
    if constexpr (PEER) {
        // one loop per destination segment, warps aligned to the 128-byte lines of the DESTINATION: every store of a
        // long run is one full line (NVLink packets carry whole lines, no partial sectors)
        const uint64_t* s_dst = reinterpret_cast<const uint64_t*>(smem + S_::OFF_DST);
        const uint32_t* s_heads = smem + S_::OFF_HEADS;
        const uint32_t nseg = s_heads[H];
        for (uint32_t j = 0; j < nseg; ++j) {
            const uint32_t g = s_heads[j];
            const uint32_t d0 = g & 0xFFFFu, e = (g >> 16) + 1u;
            const uint32_t lo = s_dp[d0];
            uint32_t hi = e < (uint32_t)H ? s_dp[e] : valid;
            hi = hi < valid ? hi : valid;
            uint32_t* dst = reinterpret_cast<uint32_t*>(s_dst[d0]);
            const uint32_t gb = s_gbase[d0];  // index of tile position p in the destination = gb + p
            const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(dst) >> 2) + gb + lo) & 31u;
            for (uint32_t p = lo - mis + tid; (int32_t)(p - hi) < 0; p += THREADS)
                if ((int32_t)(p - lo) >= 0) dst[gb + p] = s_keys[p];
        }
    } else if constexpr (PAIRS) {
        // keys out, remembering the digit of every position this thread copies (the values go to the same place)
        uint32_t dpk[(ITEMS + 3) / 4];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            uint32_t d = 0;
            if (p < valid) {
                const uint32_t k = s_keys[p];
                d = (k >> shift) & (H - 1);
                st_key<5>(out + s_gbase[d] + p, TYPED ? key_from_unsigned(k, xout) : k);
            }
            if (i & 3) dpk[i >> 2] |= d << (8 * (i & 3)); else dpk[i >> 2] = d;
        }
        const uint32_t* __restrict__ vin = (src_scratch ? a.vals_scratch : a.vals) + a.portion_base;
        uint32_t* __restrict__ vout = src_scratch ? a.vals : a.vals_scratch;
        uint32_t* s_vals = s_mat;  // staging area of the values: the dead counter matrix (+ tile digit counts)
        if (valid == (uint32_t)TILE) {
            mbar_wait(s_bar, 1);  // issued before the keys streamed out
        } else {
            for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS) s_vals[p] = p < valid ? vin[tile_base + p] : 0u;
            __syncthreads();
        }
        // same lane-blocked ownership as the keys, same byte offsets (rk) in the reorder buffer
        uint32_t val[ITEMS];
        {
            const uint32_t* src = s_vals + lane * S + warp * ITEMS;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) val[i] = src[i];
        }
        __syncthreads();  // every key has left the reorder buffer: the values take the keys' places
        {
            char* kb = reinterpret_cast<char*>(s_keys);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t off = (i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu);
                *reinterpret_cast<uint32_t*>(kb + off) = val[i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t d = (dpk[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            if (p < valid) st_key<5>(vout + s_gbase[d] + p, s_keys[p]);
        }
    } else if (CLR == 6 && !typed_out) {
        // tuning variant: one bucket run at a time per warp, lanes aligned to the 128-byte lines of the DESTINATION, so a
        // run of L keys costs ceil((misalignment + L) / 32) line requests instead of the ~2.9 per 32 keys of the
        // position-linear loop below (a warp there straddles two runs, each at its own alignment)
        for (uint32_t d = warp; d < (uint32_t)H; d += WARPS) {
            const uint32_t lo = s_dp[d];
            uint32_t hi = d + 1 < (uint32_t)H ? s_dp[d + 1] : (uint32_t)TILE;
            hi = hi < valid ? hi : valid;
            const uint32_t gb = s_gbase[d];
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out + gb + lo) >> 2) & 31u;
            for (uint32_t p = lo - mis + lane; (int32_t)(p - hi) < 0; p += 32)
                if ((int32_t)(p - lo) >= 0) st_key<5>(out + gb + p, s_keys[p]);
        }
    } else if (valid == (uint32_t)TILE && !typed_out) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t k = s_keys[p];
            st_key<CLR>(out + s_gbase[(k >> shift) & (H - 1)] + p, k);
        }
    } else {
        for (uint32_t p = tid; p < valid; p += THREADS) {
            const uint32_t k = s_keys[p];
            out[s_gbase[(k >> shift) & (H - 1)] + p] = TYPED ? key_from_unsigned(k, xout) : k;
        }
    }
    if (warp == 0) LSD_TRACE(12);  // warp 0 issued its stores
#undef LSD_TRACE
}

template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CLR, int MODE, bool SP>
int onesweep_lpc32_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpc32_kernel<RB, WARPS, ITEMS, MINB, SHIFT, LB, CLR, MODE, SP>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<a.tiles, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CLR, int MODE, bool SP>
int onesweep_lpc32_launch(const PassArgs& a, cudaStream_t s)
{
    static_assert(RB == 8, "shift dispatch below is written for 8-bit digits");
    switch (a.shift) {
        case 0: return onesweep_lpc32_launch_shift<RB, WARPS, ITEMS, MINB, 0, LB, CLR, MODE, SP>(a, s);
        case 8: return onesweep_lpc32_launch_shift<RB, WARPS, ITEMS, MINB, 8, LB, CLR, MODE, SP>(a, s);
        case 16: return onesweep_lpc32_launch_shift<RB, WARPS, ITEMS, MINB, 16, LB, CLR, MODE, SP>(a, s);
        case 24: return onesweep_lpc32_launch_shift<RB, WARPS, ITEMS, MINB, 24, LB, CLR, MODE, SP>(a, s);
    }
    // any other digit position: the run-time form, built for the peer-scatter pass only (exchange window of lsd_sort_multi)
    if constexpr (MODE == kPassPeer)
        if (a.shift > 0 && a.shift < 24) return onesweep_lpc32_launch_shift<RB, WARPS, ITEMS, MINB, -1, LB, CLR, MODE, SP>(a, s);
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeLpc32 = 4;

template <int RB, int WARPS, int ITEMS, int MINB, int LB = 8, int CLR = 0, bool WITH_PEER = false, bool SP = false>
constexpr OnesweepLauncher make_lpc32_launcher()
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    if constexpr (WITH_PEER)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc32, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassPlain, SP>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassPeer, SP>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassPairs, SP>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassTyped, SP>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassPairsTyped, SP>};
    else
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc32, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, LB, CLR, kPassPlain, SP>, nullptr, nullptr, nullptr,
                                nullptr};
}

}  // namespace lsd
