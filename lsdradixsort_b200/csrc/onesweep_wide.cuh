// onesweep_wide.cuh -- the r = 8 digit pass with 16 Ki-key tiles (round 2 default).
//
// Why (profiles/r01_final6_ncu_summary.csv, r01_pass_memory_skeleton.txt): onesweep_lpc3_kernel is bound
// by shared-memory (LSU) wavefronts, ~19 per 32 keys, and by the partial 32-byte sectors at both ends of every bucket
// run (256 runs of ~32 keys per 8 Ki-key tile).  Both costs are per TILE, so this kernel doubles the tile and removes
// per-key work from the copy-out:
//   * tile = 32 x S keys, S = WARPS x ITEMS (15 x 35 -> 16 800 keys), two CTAs per SM.  The keys of a tile live in
//     registers between the lane-blocked read and the scatter, so the TMA landing zone IS the reorder buffer (no second
//     tile buffer); the next tile is pulled into L2 with cp.async.bulk.prefetch.L2 while this one is ranked and its TMA
//     load is issued as soon as the reorder buffer has drained;
//   * cnt[digit][lane] holds TWO 16-bit key-index counters per word (low half: even warps, high half: odd warps), so the
//     even and the odd warps form two rank chains that run concurrently (8 + 7 turns); a lane's segment of the tile is
//     [even warps' keys | odd warps' keys].  The scan works on the packed words;
//   * one dedicated look-back warp (8 digits per lane, 128-bit relaxed loads/stores, per-word flags) walks the tile
//     records while the rank chains run: the walk is off the chain's critical path, and its record is half as many hops
//     away because tiles are twice as large;
//   * copy-out by bucket: every rank warp owns ceil(H / WARPS) buckets and writes each of their runs with lanes aligned to the
//     destination's 128-byte lines (one LDS + one single-line STG per 32-key slot, no per-key bucket-base lookup; a run
//     is only split at line boundaries, so it has exactly two partial sectors).  Tiles with a very long run (skewed
//     inputs) fall back to the position-linear copy-out, which is ideal for long runs;
//   * the counter matrix is cleared by the look-back warp after the rank chains, under the copy-out, not in front of the count.
// Ticket order, look-back protocol (flag[31:30] | value[29:0], LOCAL / INCLUSIVE), workspace layout, typed-key mapping
// and the ragged last tile (pad keys 0xFFFFFFFF, never written) are those of onesweep_lpc32.cuh.
// Replaces the reference's LSDRadixSortKernel + per-pass histogram / scan / transpose launches (LSDRadixSort.cu:795-837,
// :844-906); stable, bit-exact with LSDRadixSortPass (.cu:25-54).
#pragma once
#include "onesweep_lpc3.cuh"

namespace lsd {

template <int RB, int WARPS, int ITEMS, int NLB = 1>
struct WideShape {
    static constexpr int H = 1 << RB;
    static constexpr int RW = WARPS;                    // rank warps; warps RW .. RW+NLB-1 are the look-back warps
    static constexpr int THREADS = (WARPS + NLB) * 32;
    static constexpr int S = WARPS * ITEMS;             // keys per lane segment
    static constexpr int TILE = 32 * S;
    static constexpr int EVEN = (WARPS + 1) / 2;        // warps of chain 0 (even warps): first EVEN*ITEMS keys of a segment
    static constexpr int SW = H / 32;                   // scan warps: one matrix row per lane
    static constexpr int DPL = H / (32 * NLB);          // digits per look-back lane
    static_assert(DPL == 8 || DPL == 4 || DPL == 2, "one, two or four look-back warps");
    static_assert(RB == 8, "8-bit digits only (r < 8 stays on onesweep_lpc3_kernel)");
    static_assert(S % 2 == 1, "S = WARPS*ITEMS must be odd (conflict-free lane-blocked reads)");
    static_assert(TILE < 65536, "key indices are packed in 16 bits");
    static_assert(WARPS >= SW && WARPS <= 15, "named barriers: 1 totals, 2..WARPS-1 chains, 15 scan warps");
    static constexpr int OFF_MAT = TILE;               // [H][32] packed counters
    static constexpr int OFF_TOT = OFF_MAT + H * 32;   // [H] tile digit counts
    static constexpr int OFF_DP = OFF_TOT + H;         // [H] tile-local bucket starts
    static constexpr int OFF_G = OFF_DP + H;           // [H] global position of the bucket's first key of this tile
    static constexpr int OFF_GB = OFF_G + H;           // [H] OFF_G minus the tile-local bucket start (position-linear copy-out)
    static constexpr int OFF_MISC = OFF_GB + H;        // [0..15] scan partials, [32..33] tile ids, [34] long-run flag, [35] zero, [36..39] mbarriers
    static constexpr int WORDS = OFF_MISC + 64;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
};

__device__ __forceinline__ void st_relaxed_gpu_v4(uint32_t* p, uint4 v)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// ld_relaxed_gpu_v4: lookback_quad.cuh
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

#define LSD_TRACE(slot)                                                                      \
    do {                                                                                     \
        if constexpr (TRACE)                                                                 \
            if (a.trace && lane == 0) a.trace[(size_t)tile * 16 + (slot)] = (unsigned long long)(clock64() - t_start); \
    } while (0)

// One look-back record slice of DPL consecutive words, moved with the widest relaxed gpu-scope access.
template <int DPL>
__device__ __forceinline__ void lb_load(const uint32_t* p, uint32_t (&w)[DPL])
{
    if constexpr (DPL == 2) {
        const uint2 v = ld_relaxed_gpu_v2(p);
        w[0] = v.x; w[1] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < DPL; j += 4) {
            const uint4 v = ld_relaxed_gpu_v4(p + j);
            w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
    }
}
template <int DPL>
__device__ __forceinline__ void lb_store(uint32_t* p, const uint32_t (&w)[DPL])
{
    if constexpr (DPL == 2) {
        st_relaxed_gpu_v2(p, w[0], w[1]);
    } else {
#pragma unroll
        for (int j = 0; j < DPL; j += 4) st_relaxed_gpu_v4(p + j, make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]));
    }
}
template <int DPL>
__device__ __forceinline__ void smem_load(const uint32_t* p, uint32_t (&w)[DPL])
{
    if constexpr (DPL == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        w[0] = v.x; w[1] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < DPL; j += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(p + j);
            w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
    }
}
template <int DPL>
__device__ __forceinline__ void smem_store(uint32_t* p, const uint32_t (&w)[DPL])
{
    if constexpr (DPL == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
    } else {
#pragma unroll
        for (int j = 0; j < DPL; j += 4) *reinterpret_cast<uint4*>(p + j) = make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]);
    }
}

// ---- the look-back warps' life: next ticket + L2 prefetches, tile record walk, their share of the matrix clear.  Their own
// ---- function so that their registers (a window of LB records, DPL digits per lane) never coexist with the rank warps' keys.
template <int RB, int WARPS, int ITEMS, int NLB, int LB, bool TRACE>
__device__ __forceinline__ void wide_lookback_role(const PassArgs& a, uint32_t* smem, const uint32_t* __restrict__ in)
{
    using S_ = WideShape<RB, WARPS, ITEMS, NLB>;
    constexpr int H = S_::H, TILE = S_::TILE, SW = S_::SW, DPL = S_::DPL;
    constexpr uint32_t kBarTot = 1;
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_g = smem + S_::OFF_G;
    uint32_t* s_gb = smem + S_::OFF_GB;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lbw = (threadIdx.x >> 5) - (uint32_t)WARPS;  // look-back warp index
    const uint32_t d0 = (lbw * 32u + lane) * (uint32_t)DPL;      // first digit of this lane
    const bool boss = lbw == 0u && lane == 0u;

    for (uint32_t iter = 0;; ++iter) {
        const uint32_t tile = s_misc[32 + (iter & 1u)];
        if (tile >= a.tiles) break;
        [[maybe_unused]] const long long t_start = (TRACE && a.trace) ? clock64() : 0;
        const uint32_t left = a.portion_keys - tile * (uint32_t)TILE;
        const uint32_t pads = left < (uint32_t)TILE ? (uint32_t)TILE - left : 0u;
        if (pads) cta_sync();  // ragged last tile: the rank warps fill the landing zone by hand
        cta_sync();            // (A)
        uint32_t* lb_row = a.lookback + (size_t)tile * H;
        uint32_t nt = 0;
        if (boss) {
            nt = atomicAdd(a.ticket, 1u);  // the result is only consumed after barrier (B): its round trip hides behind the count
            s_misc[34] = 0;
        }
        // Pull this tile's still empty record into L2: the successors poll it before it is published, and the lines zeroed
        // at the start of the sort were evicted by the key stream long ago (see onesweep_lpc3.cuh).
        if (lbw == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(lb_row + 8 * lane));
        cta_sync();  // (B)
        if (boss) {
            s_misc[32 + ((iter + 1u) & 1u)] = nt;
            const uint32_t nb = nt * (uint32_t)TILE;
            if (nt < a.tiles && a.portion_keys - nb >= (uint32_t)TILE) prefetch_l2_bulk(in + nb, TILE * 4);
        }
        named_bar_sync(kBarTot, (SW + NLB) * 32);
        if (lbw == 0u) LSD_TRACE(8);
        uint32_t cnt[DPL], ex[DPL], w[DPL];
        smem_load<DPL>(s_tot + d0, cnt);
        if (d0 + DPL == (uint32_t)H) cnt[DPL - 1] -= pads;  // pads of a ragged last tile are not keys
#pragma unroll
        for (int j = 0; j < DPL; ++j) ex[j] = 0;
        uint32_t* my = lb_row + d0;
        const uint32_t flag0 = tile == 0 ? kLbGlobal : kLbLocal;
#pragma unroll
        for (int j = 0; j < DPL; ++j) w[j] = flag0 | cnt[j];
        lb_store<DPL>(my, w);
        if (tile != 0) {
            // Windowed walk, branch-free inside a round.  A record slice is READY when all its words are published and
            // either all or none of them are INCLUSIVE (a slice caught between its two states is polled again); slices are
            // consumed in order, up to the first one that is not ready and not beyond the first INCLUSIVE one.
            const uint32_t* p = my - H;
            uint32_t remaining = tile;
            bool done = false;
            [[maybe_unused]] uint32_t dbg_rounds = 0, dbg_hops = 0;
            [[maybe_unused]] long long dbg_wait = 0, dbg_proc = 0;
            while (!done) {
                if constexpr (TRACE) ++dbg_rounds;
                [[maybe_unused]] const long long t_a = TRACE ? clock64() : 0;
                uint32_t win[LB][DPL];
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    if ((uint32_t)k < remaining) {
                        lb_load<DPL>(p - (size_t)k * H, win[k]);
                    } else {
#pragma unroll
                        for (int j = 0; j < DPL; ++j) win[k][j] = 0u;
                    }
                }
                [[maybe_unused]] long long t_b = 0;
                if constexpr (TRACE) {  // all loads of the round have landed
                    uint32_t all = 0;
#pragma unroll
                    for (int k = 0; k < LB; ++k)
#pragma unroll
                        for (int j = 0; j < DPL; ++j) all |= win[k][j];
                    asm volatile("" ::"r"(all) : "memory");
                    t_b = clock64();
                    dbg_wait += t_b - t_a;
                }
                uint32_t rdy = 0, glb = 0;
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    uint32_t mn = win[k][0], mx = win[k][0];
#pragma unroll
                    for (int j = 1; j < DPL; ++j) {
                        mn = min(mn, win[k][j]);
                        mx = max(mx, win[k][j]);
                    }
                    const bool all_glb = mn >= kLbGlobal;
                    const bool ready = mn >= kLbLocal && (all_glb || mx < kLbGlobal);
                    rdy |= ready ? (1u << k) : 0u;
                    glb |= all_glb ? (1u << k) : 0u;
                }
                const uint32_t lead = (uint32_t)__ffs((int)(~rdy | (1u << LB))) - 1u;  // leading ready slices
                const uint32_t hit = glb & ((1u << lead) - 1u);                        // INCLUSIVE ones among them
                const uint32_t take = hit ? (uint32_t)__ffs((int)hit) : lead;
#pragma unroll
                for (int k = 0; k < LB; ++k)
#pragma unroll
                    for (int j = 0; j < DPL; ++j) ex[j] += (uint32_t)k < take ? (win[k][j] & kLbValueMask) : 0u;
                done = hit != 0u;
                p -= (size_t)take * H;
                remaining -= take;
                if constexpr (TRACE) {
                    dbg_hops += take;
                    asm volatile("" ::"r"(take), "r"(ex[0]) : "memory");
                    dbg_proc += clock64() - t_b;
                }
            }
            if constexpr (TRACE)
                if (a.trace && lbw == 0u && lane == 0) {
                    a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                    a.trace[(size_t)tile * 16 + 14] = dbg_hops;
                    a.trace[(size_t)tile * 16 + 15] = (unsigned long long)dbg_wait;
                    a.trace[(size_t)tile * 16 + 0] = (unsigned long long)dbg_proc;
                }
#pragma unroll
            for (int j = 0; j < DPL; ++j) w[j] = kLbGlobal | (ex[j] + cnt[j]);
            lb_store<DPL>(my, w);
        }
        uint32_t g[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            const uint64_t b = a.bases_in[d0 + j];
            g[j] = (uint32_t)b + ex[j];  // positions are < 2^32 (n <= 2^32)
            if (a.bases_out != nullptr && tile == a.tiles - 1) a.bases_out[d0 + j] = b + ex[j] + cnt[j];
        }
        smem_store<DPL>(s_g + d0, g);
        smem_load<DPL>(s_dp + d0, w);
#pragma unroll
        for (int j = 0; j < DPL; ++j) w[j] = g[j] - w[j];
        smem_store<DPL>(s_gb + d0, w);
        if (lbw == 0u) LSD_TRACE(9);
        cta_sync();  // (C) the matrix is dead: clear it (every thread its share) while the rank warps stream the tile out
        {
            uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
            for (uint32_t i = threadIdx.x; i < (uint32_t)H * 8u; i += S_::THREADS) m4[i] = make_uint4(0, 0, 0, 0);
        }
        cta_sync();  // (D)
    }
}

// LB: look-back window (records per round trip).  LONGRUN: a tile whose longest bucket run exceeds this uses the
// position-linear copy-out.  COPY bit 0: 0 = bucket-walk copy-out (default), 1 = always position-linear (tuning);
// COPY bit 1: scatter interleaved with the chain atomics instead of packed rank registers (tuning).
template <int RB, int WARPS, int ITEMS, int NLB, int MINB, int SHIFT, int LB, int COPY, bool TYPED, bool TRACE>
__global__ void __launch_bounds__((WARPS + NLB) * 32, MINB)
onesweep_wide_kernel(const PassArgs a)
{
    using S_ = WideShape<RB, WARPS, ITEMS, NLB>;
    constexpr int H = S_::H, S = S_::S, TILE = S_::TILE, RW = S_::RW;
    constexpr int RT = RW * 32;  // rank threads
    constexpr int EVEN = S_::EVEN, SW = S_::SW;
    constexpr int BPW = (H + RW - 1) / RW;  // buckets per rank warp in the copy-out
    static_assert(BPW <= 32, "bucket info is held one bucket per lane");
    constexpr uint32_t kBarTot = 1, kBarScan = 15;
    constexpr uint32_t LONGRUN = 8u * (uint32_t)(TILE / H);

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_keys = smem;  // TMA landing zone, then reorder buffer
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_tot = smem + S_::OFF_TOT;
    uint32_t* s_dp = smem + S_::OFF_DP;
    uint32_t* s_g = smem + S_::OFF_G;
    uint32_t* s_gb = smem + S_::OFF_GB;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 36);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    auto load_tile = [&](uint32_t t) {  // thread 0: TMA load of a full tile into the (drained) reorder buffer
        const uint32_t base = t * (uint32_t)TILE;
        if (t < a.tiles && a.portion_keys - base >= (uint32_t)TILE) {
            mbar_expect_tx(s_bar, TILE * 4);
            tma_bulk_g2s(s_keys, in + base, TILE * 4, s_bar);
        }
    };
    if (tid == 0) {
        mbar_init(s_bar, 1);
        const uint32_t t = atomicAdd(a.ticket, 1u);
        s_misc[32] = t;
        s_misc[35] = 0;
        load_tile(t);
    }
    {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
        for (uint32_t i = tid; i < (uint32_t)H * 8u; i += S_::THREADS) m4[i] = make_uint4(0, 0, 0, 0);
    }
    cta_sync();

    if (warp >= (uint32_t)RW) {
        wide_lookback_role<RB, WARPS, ITEMS, NLB, LB, TRACE>(a, smem, in);
        return;
    }

    // ================= rank warps =================
    const uint32_t half = warp & 1u;
    const KeyXform xin = TYPED ? pass_xform_in(a) : KeyXform{0u, 0u};
    const bool typed_out = TYPED && a.plan->last_pass == (uint32_t)a.pass;
    const KeyXform xout = key_xform_of(typed_out ? a.key_type : 0u);
    // byte offset of the key's cell in the matrix, and this warp's counter increment (low / high half)
    const uint32_t lane4 = lane << 2;
    const uint32_t inc = half ? 0x10000u : 1u;
    char* mat_bytes = reinterpret_cast<char*>(s_mat);
    // this thread's part of its lane segment: [even warps | odd warps]
    const uint32_t seg_off = lane * (uint32_t)S + (half ? (uint32_t)(EVEN * ITEMS) : 0u) + (warp >> 1) * (uint32_t)ITEMS;

    uint32_t phase = 0;
    for (uint32_t iter = 0;; ++iter) {
        const uint32_t tile = s_misc[32 + (iter & 1u)];
        if (tile >= a.tiles) break;
        [[maybe_unused]] const long long t_start = (TRACE && a.trace) ? clock64() : 0;
        const uint32_t tile_base = tile * (uint32_t)TILE;
        const uint32_t left = a.portion_keys - tile_base;
        const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
        const uint32_t pads = (uint32_t)TILE - valid;

        if (valid == (uint32_t)TILE) {
            mbar_wait(s_bar, phase);
            phase ^= 1u;
        } else {
            const uint32_t pad_key = TYPED ? key_from_unsigned(0xFFFFFFFFu, xin) : 0xFFFFFFFFu;  // pads sort last
            for (uint32_t p = tid; p < (uint32_t)TILE; p += RT) s_keys[p] = p < valid ? in[tile_base + p] : pad_key;
            cta_sync();
        }
        if (warp == 0) LSD_TRACE(1);  // tile landed

        // ---- 1. lane-blocked read: the keys of the tile move to registers ----
        uint32_t key[ITEMS];
        {
            const uint32_t* src = s_keys + seg_off;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = TYPED ? key_to_unsigned(src[i], xin) : src[i];
        }
        cta_sync();  // (A) every key is in registers: s_keys is the reorder buffer from here on

        // ---- 2. count ----
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[i], lane4)), inc);
        if (warp == 0) LSD_TRACE(2);
        cta_sync();  // (B) counts complete
        if (warp == 0) LSD_TRACE(3);

        if (warp < (uint32_t)SW) {
            // ================= scan warps: one row (digit) per lane; packed words: low = chain 0, high = chain 1 =================
            const uint32_t q = lane & 7u;
            const uint32_t row = warp * 32u + lane;
            uint4* r4 = reinterpret_cast<uint4*>(s_mat + row * 32u);
            uint32_t total = 0, below = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                const uint4 v = r4[grp];
                const uint32_t t = v.x + v.y + v.z + v.w;  // both halves at once: a row holds fewer than 2^16 keys
                const uint32_t s = (t & 0xFFFFu) + (t >> 16);
                total += s;
                if (grp < q) below += s;  // groups reached after the wrap == lanes [0, 4q)
            }
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            uint32_t start = incl - total;
            if (lane == 31) s_misc[warp] = incl;
            named_bar_sync(kBarScan, SW * 32);
#pragma unroll
            for (int w = 0; w < SW; ++w)
                if ((uint32_t)w < warp) start += s_misc[w];
            s_tot[row] = total;
            s_dp[row] = start;
            if (total > LONGRUN) s_misc[34] = 1u;
            named_bar_arrive(kBarTot, (SW + NLB) * 32);  // totals + starts are in shared memory: the look-back warps may go
            if (warp == 0) LSD_TRACE(4);
            uint32_t run = start + below;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                if (grp == 0) run = start;
                const uint4 v = r4[grp];
                uint4 o;
                o.x = run | ((run + (v.x & 0xFFFFu)) << 16); run += (v.x & 0xFFFFu) + (v.x >> 16);
                o.y = run | ((run + (v.y & 0xFFFFu)) << 16); run += (v.y & 0xFFFFu) + (v.y >> 16);
                o.z = run | ((run + (v.z & 0xFFFFu)) << 16); run += (v.z & 0xFFFFu) + (v.z >> 16);
                o.w = run | ((run + (v.w & 0xFFFFu)) << 16); run += (v.w & 0xFFFFu) + (v.w >> 16);
                r4[grp] = o;
            }
            named_bar_sync(kBarScan, SW * 32);  // matrix complete before warps 0 and 1 open the rank chains
            if (warp == 0) LSD_TRACE(5);
        }

        // ---- 3. two rank chains: the returned counter half is the key's index in the reorder buffer ----
        if (warp >= 2u) named_bar_sync(warp, 64);
        // The chain recomputes the cell offsets from key ^ z, z a zero word that only exists now (volatile load): otherwise the
        // compiler keeps the count phase's ITEMS digit terms alive next to the ITEMS keys across the scan and spills the keys.
        const uint32_t z = *reinterpret_cast<volatile uint32_t*>(s_misc + 35);
        if constexpr (COPY < 2) {
            // the turn is the ITEMS returning atomics only (a warp has one in flight at a time): the index halves are packed
            // two per register and the scatter follows after the hand-over
            uint32_t rk[(ITEMS + 1) / 2];
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[i] ^ z, lane4)), inc);
                // chain 0 wants the low half of `old`, chain 1 the high half; keep the wanted half of two atomics per register
                if (i & 1) rk[i >> 1] = __byte_perm(rk[i >> 1], old, half ? 0x7610 : 0x5410);
                else rk[i >> 1] = half ? (old >> 16) : old;
            }
            if (warp + 2u < (uint32_t)RW) named_bar_arrive(warp + 2u, 64);
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t pos = (i & 1) ? (rk[i >> 1] >> 16) : (rk[i >> 1] & 0xFFFFu);
                s_keys[pos] = key[i];
            }
        } else {
            // scatter software-pipelined by one behind the atomics: no rank registers (tuning variant)
            uint32_t prev = 0;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[i] ^ z, lane4)), inc);
                if (i > 0) s_keys[half ? (prev >> 16) : (prev & 0xFFFFu)] = key[i - 1];
                prev = old;
            }
            if (warp + 2u < (uint32_t)RW) named_bar_arrive(warp + 2u, 64);
            s_keys[half ? (prev >> 16) : (prev & 0xFFFFu)] = key[ITEMS - 1];
        }
        if (warp == 0) LSD_TRACE(6);
        if (warp == (uint32_t)RW - 1u) LSD_TRACE(10);
        if (warp == (uint32_t)RW - 2u) LSD_TRACE(11);
        cta_sync();  // (C) reorder buffer complete, bucket positions known; the matrix is dead
        if (warp == 0) LSD_TRACE(7);

        // ---- 4. clear the matrix for the next tile (every thread its share), stream the tile out ----
        {
            uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
            for (uint32_t i = tid; i < (uint32_t)H * 8u; i += S_::THREADS) m4[i] = make_uint4(0, 0, 0, 0);
        }
        if ((COPY & 1) == 0 && s_misc[34] == 0u) {
            // bucket-walk: lane j of a warp holds the run of bucket warp*BPW + j; two runs are in flight at a time
            const uint32_t b = warp * (uint32_t)BPW + lane;
            uint32_t my_dp = 0, my_n = 0, my_g = 0;
            if (lane < (uint32_t)BPW && b < (uint32_t)H) {
                my_dp = s_dp[b];
                my_n = s_tot[b] - (b == (uint32_t)H - 1u ? pads : 0u);
                my_g = s_g[b];
            }
            auto put = [&](uint32_t* p, uint32_t k) { st_key<5>(p, TYPED ? key_from_unsigned(k, xout) : k); };
#pragma unroll 1
            for (int j = 0; j < BPW; j += 2) {
                uint32_t dp[2], n[2], g[2], a0[2];
#pragma unroll
                for (int y = 0; y < 2; ++y) {
                    dp[y] = __shfl_sync(kFullMask, my_dp, j + y);
                    n[y] = (j + y < BPW) ? __shfl_sync(kFullMask, my_n, j + y) : 0u;
                    g[y] = __shfl_sync(kFullMask, my_g, j + y);
                    a0[y] = g[y] & 31u;  // position of the run's first key inside its 128-byte line
                }
                if (a0[0] + n[0] <= 128u && a0[1] + n[1] <= 128u) {
                    // both runs fit four line slots each: eight loads in flight, then eight single-line stores
                    uint32_t q[2][4], k[2][4];
#pragma unroll
                    for (int y = 0; y < 2; ++y)
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            q[y][x] = 32u * x + lane - a0[y];  // index inside the run ("negative" wraps and fails q < n)
                            k[y][x] = q[y][x] < n[y] ? s_keys[dp[y] + q[y][x]] : 0u;
                        }
#pragma unroll
                    for (int y = 0; y < 2; ++y)
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            if (q[y][x] < n[y]) put(out + (size_t)g[y] + q[y][x], k[y][x]);
                } else {
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        const uint32_t end = a0[y] + n[y];
                        for (uint32_t u = 0; u < end; u += 128u) {
                            uint32_t q[4], k[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                q[x] = u + 32u * x + lane - a0[y];
                                k[x] = q[x] < n[y] ? s_keys[dp[y] + q[x]] : 0u;
                            }
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                                if (q[x] < n[y]) put(out + (size_t)g[y] + q[x], k[x]);
                        }
                    }
                }
            }
        } else if (valid == (uint32_t)TILE) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t p = (uint32_t)i * RT + tid;
                const uint32_t k = s_keys[p];
                const uint32_t pos = s_gb[(k >> SHIFT) & (uint32_t)(H - 1)] + p;  // 32-bit wrap-around on purpose
                st_key<5>(out + (size_t)pos, TYPED ? key_from_unsigned(k, xout) : k);
            }
        } else {
            for (uint32_t p = tid; p < valid; p += RT) {
                const uint32_t k = s_keys[p];
                const uint32_t pos = s_gb[(k >> SHIFT) & (uint32_t)(H - 1)] + p;
                st_key<5>(out + (size_t)pos, TYPED ? key_from_unsigned(k, xout) : k);
            }
        }
        if (warp == 0) LSD_TRACE(12);
        // generic-proxy reads of the reorder buffer are ordered before the async-proxy write of the next tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync();  // (D) reorder buffer drained, matrix zero
        if (tid == 0) load_tile(s_misc[32 + ((iter + 1u) & 1u)]);
    }
}
#undef LSD_TRACE

template <int RB, int WARPS, int ITEMS, int NLB, int MINB, int SHIFT, int LB, int COPY, bool TYPED, bool TRACE>
int onesweep_wide_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = WideShape<RB, WARPS, ITEMS, NLB>;
    auto kern = onesweep_wide_kernel<RB, WARPS, ITEMS, NLB, MINB, SHIFT, LB, COPY, TYPED, TRACE>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const uint32_t resident = (uint32_t)sm_count() * MINB;
    const uint32_t grid = a.tiles < resident ? a.tiles : resident;  // persistent: every CTA loops over tickets
    kern<<<grid, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int WARPS, int ITEMS, int NLB, int MINB, int LB, int COPY, bool TYPED = false, bool TRACE = false>
int onesweep_wide_launch(const PassArgs& a, cudaStream_t s)
{
    switch (a.shift) {
        case 0: return onesweep_wide_launch_shift<RB, WARPS, ITEMS, NLB, MINB, 0, LB, COPY, TYPED, TRACE>(a, s);
        case 8: return onesweep_wide_launch_shift<RB, WARPS, ITEMS, NLB, MINB, 8, LB, COPY, TYPED, TRACE>(a, s);
        case 16: return onesweep_wide_launch_shift<RB, WARPS, ITEMS, NLB, MINB, 16, LB, COPY, TYPED, TRACE>(a, s);
        case 24: return onesweep_wide_launch_shift<RB, WARPS, ITEMS, NLB, MINB, 24, LB, COPY, TYPED, TRACE>(a, s);
    }
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeWide = 7;

// plain and typed-key passes; key-value and peer-scatter passes stay with the 8 Ki-key shapes (their own table entries)
template <int RB, int WARPS, int ITEMS, int MINB, int LB, int COPY = 0, bool TRACE = false, int NLB = 1>
constexpr OnesweepLauncher make_wide_launcher()
{
    using S_ = WideShape<RB, WARPS, ITEMS, NLB>;
    if constexpr (TRACE)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeWide, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_wide_launch<RB, WARPS, ITEMS, NLB, MINB, LB, COPY, false, true>, nullptr, nullptr, nullptr, nullptr};
    else
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeWide, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_wide_launch<RB, WARPS, ITEMS, NLB, MINB, LB, COPY>, nullptr, nullptr,
                                &onesweep_wide_launch<RB, WARPS, ITEMS, NLB, MINB, LB, COPY, true>, nullptr};
}

}  // namespace lsd
