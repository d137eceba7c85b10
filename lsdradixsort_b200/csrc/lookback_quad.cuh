// lookback_quad.cuh -- decoupled look-back with 128-bit record accesses and the window spread over LANES.
//
// The round-1 / round-2 walk gives every look-back thread one digit PAIR and lets it fetch LB records per round with LB
// strong 64-bit loads.  Under load every such load instruction queues behind the copy-out traffic of the three resident
// CTAs (~250 cycles apart), so a round costs ~1 K cycles whatever LB is, and the walk -- which ends when the INCLUSIVE front
// reaches it, D = c (1 + R / (W tau)) for a round of R cycles that covers W records while a new tile starts every tau --
// is the longest stretch of a tile (profiles/r02_lookback_experiments.txt).  More loads per thread do not help (R grows
// with them); WIDER loads do: here a lane owns a digit QUAD (one 128-bit load per record) and the 32 lanes of a look-back
// warp are QPW quads x G records, so ONE load instruction of the warp fetches G records of QPW quads and a round of LB
// instructions covers W = G * LB records:
//     r = 8 (256 digits, 4 look-back warps): 16 quads x 2 records per warp; LB = 4 -> 8 records per round, LB = 2 -> 4
//                                            records with half the load instructions of the pair walk
//     r = 4 (16 digits, 1 look-back warp):   4 quads x 8 records; LB = 2 -> 16 records per round
//     r = 2 (4 digits):                      1 quad x 32 records; LB = 1
// The lanes of a quad agree on how far the window can be consumed through two shuffled bit masks (record present, record
// INCLUSIVE), add the records they hold, and combine their sums once at the end.
// Record format and protocol are those of every other pass kernel (flag[31:30] | count[29:0] per digit, LOCAL then
// INCLUSIVE), so the kernels interoperate on one workspace.  A quad counts as published when all four words carry the same
// flag: a (never observed) torn 16-byte access reads as "not there yet" and is polled again.
#pragma once
#include "onesweep_lpc.cuh"

namespace lsd {

__device__ __forceinline__ void st_relaxed_gpu_v4(uint32_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int H, int LBW, int LB>
struct QuadLookback {
    static constexpr int E = H >= 4 ? 4 : 2;  // words per lane: a digit quad, or the single digit pair of r = 1 (64-bit accesses)
    static constexpr int QUADS = H / E;
    static constexpr int QPW = QUADS / LBW;  // quads per look-back warp
    static constexpr int G = 32 / QPW;       // lanes of a quad = records fetched by one load instruction
    static constexpr int W = G * LB;         // records per round
    static_assert(H >= 2 && QPW >= 1 && QPW * LBW == QUADS && G * QPW == 32, "a look-back warp is QPW quads x G records");
    static_assert(W <= 32, "the window masks are 32-bit");

    __device__ static __forceinline__ uint4 load(const uint32_t* p)
    {
        if constexpr (E == 4) return ld_relaxed_gpu_v4(p);
        const uint2 v = ld_relaxed_gpu_v2(p);
        return make_uint4(v.x, v.y, v.x, v.y);
    }
    __device__ static __forceinline__ void store(uint32_t* p, uint32_t flag, const uint4& v)
    {
        if constexpr (E == 4) st_relaxed_gpu_v4(p, flag | v.x, flag | v.y, flag | v.z, flag | v.w);
        else st_relaxed_gpu_v2(p, flag | v.x, flag | v.y);
    }

    // Exclusive prefix of the quad (digits 4*quad .. 4*quad+3) over tiles 0 .. tile-1; every lane of the quad returns it.
    // All 32 lanes of the warp must call (shuffles); `lane` = lane id, lb_row = this tile's record.
    __device__ static __forceinline__ uint4 walk(const uint32_t* lb_row, uint32_t tile, uint32_t quad, uint32_t lane,
                                                 uint32_t* rounds_out = nullptr, uint32_t* hops_out = nullptr)
    {
        const uint32_t g = lane / (uint32_t)QPW;
        uint4 ex = make_uint4(0u, 0u, 0u, 0u);
        uint32_t hops = 0;  // records consumed so far: the window starts at tile - 1 - hops
        bool done = false;
        uint32_t rounds = 0;
        while (__any_sync(kFullMask, !done)) {
            ++rounds;
            uint4 w[LB];
            uint32_t pm = 0, im = 0;  // bit m: record m of the window is published / is INCLUSIVE (own records only)
#pragma unroll
            for (int k = 0; k < LB; ++k) {
                const uint32_t m = (uint32_t)(k * G) + g;
                w[k] = make_uint4(0u, 0u, 0u, 0u);
                if (!done && hops + m < tile) w[k] = load(lb_row - (size_t)(hops + m + 1u) * H + (uint32_t)E * quad);
            }
#pragma unroll
            for (int k = 0; k < LB; ++k) {
                const uint32_t m = (uint32_t)(k * G) + g;
                const uint32_t all_and = w[k].x & w[k].y & w[k].z & w[k].w, all_or = w[k].x | w[k].y | w[k].z | w[k].w;
                const uint32_t f_and = all_and >> 30, f_or = all_or >> 30;
                if (f_and == f_or && f_and != 0u) {  // four words, one flag
                    pm |= 1u << m;
                    if (f_and == (kLbGlobal >> 30)) im |= 1u << m;
                }
            }
#pragma unroll
            for (int o = QPW; o < 32; o <<= 1) {  // OR over the G lanes of the quad
                pm |= __shfl_xor_sync(kFullMask, pm, o);
                im |= __shfl_xor_sync(kFullMask, im, o);
            }
            if (!done) {
                const uint32_t lead = (pm == 0xFFFFFFFFu) ? 32u : (uint32_t)(__ffs((int)~pm) - 1);  // leading published records
                uint32_t n = lead;
                if (im != 0u) {
                    const uint32_t fi = (uint32_t)(__ffs((int)im) - 1);
                    if (fi < lead) {
                        n = fi + 1u;
                        done = true;
                    }
                }
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    const uint32_t m = (uint32_t)(k * G) + g;
                    if (m < n) {
                        ex.x += w[k].x & kLbValueMask;
                        ex.y += w[k].y & kLbValueMask;
                        ex.z += w[k].z & kLbValueMask;
                        ex.w += w[k].w & kLbValueMask;
                    }
                }
                hops += n;
            }
        }
#pragma unroll
        for (int o = QPW; o < 32; o <<= 1) {  // sum over the G lanes of the quad
            ex.x += __shfl_xor_sync(kFullMask, ex.x, o);
            ex.y += __shfl_xor_sync(kFullMask, ex.y, o);
            ex.z += __shfl_xor_sync(kFullMask, ex.z, o);
            ex.w += __shfl_xor_sync(kFullMask, ex.w, o);
        }
        if (rounds_out) *rounds_out = rounds;
        if (hops_out) *hops_out = hops;
        return ex;
    }
};

// The look-back step of a tile for the look-back warps of the LPC kernels, in two halves so that a kernel can put work of
// its own between them: lookback_quad_publish sends out the tile's LOCAL record (GLOBAL for tile 0), lookback_quad_finish
// walks, publishes the INCLUSIVE record and leaves the buckets' global bases (minus their tile-local starts) in s_gbase.
// Nothing but the arguments is live across the walk (counts and starts are read again from shared memory behind it).
// lbw = index of this warp among the LBW look-back warps; s_tot / s_dp = tile digit counts / tile-local bucket starts.
template <int H, int LBW, int LB>
__device__ __forceinline__ uint4 lookback_quad_counts(uint32_t quad, uint32_t pads, uint32_t dmask, const uint32_t* s_tot)
{
    constexpr uint32_t E = (uint32_t)QuadLookback<H, LBW, LB>::E;
    const uint32_t d0 = E * quad;
    uint4 cnt;
    if constexpr (E == 4) cnt = *reinterpret_cast<const uint4*>(s_tot + d0);
    else cnt = make_uint4(s_tot[d0], s_tot[d0 + 1], 0u, 0u);
    if (pads != 0u && dmask / E == quad) {  // the pads of a ragged tile carry the largest digit in use
        const uint32_t j = dmask % E;
        if (j == 0u) cnt.x -= pads; else if (j == 1u) cnt.y -= pads; else if (j == 2u) cnt.z -= pads; else cnt.w -= pads;
    }
    return cnt;
}

template <int H, int LBW, int LB>
__device__ __forceinline__ void lookback_quad_publish(uint32_t* lb_row, uint32_t tile, uint32_t lbw, uint32_t lane, uint32_t pads,
                                                      uint32_t dmask, const uint32_t* s_tot)
{
    using QL = QuadLookback<H, LBW, LB>;
    const uint32_t quad = lbw * (uint32_t)QL::QPW + lane % (uint32_t)QL::QPW;
    if (lane < (uint32_t)QL::QPW)  // the first lane of every quad publishes
        QL::store(lb_row + (uint32_t)QL::E * quad, tile == 0u ? kLbGlobal : kLbLocal, lookback_quad_counts<H, LBW, LB>(quad, pads, dmask, s_tot));
}

template <int H, int LBW, int LB>
__device__ __forceinline__ void lookback_quad_finish(const PassArgs& a, uint32_t* lb_row, uint32_t tile, uint32_t lbw, uint32_t lane,
                                                     uint32_t pads, uint32_t dmask, const uint32_t* s_tot, const uint32_t* s_dp,
                                                     uint32_t* s_gbase, uint32_t* rounds_out = nullptr, uint32_t* hops_out = nullptr)
{
    using QL = QuadLookback<H, LBW, LB>;
    constexpr uint32_t E = (uint32_t)QL::E;
    const uint32_t quad = lbw * (uint32_t)QL::QPW + lane % (uint32_t)QL::QPW;
    const bool writer = lane < (uint32_t)QL::QPW;
    const uint32_t d0 = E * quad;
    uint4 ex = make_uint4(0u, 0u, 0u, 0u);
    if (tile != 0u) ex = QL::walk(lb_row, tile, quad, lane, rounds_out, hops_out);
    if (writer) {
        const uint4 cnt = lookback_quad_counts<H, LBW, LB>(quad, pads, dmask, s_tot);
        if (tile != 0u) QL::store(lb_row + d0, kLbGlobal, make_uint4(ex.x + cnt.x, ex.y + cnt.y, ex.z + cnt.z, ex.w + cnt.w));
        const uint32_t exv[4] = {ex.x, ex.y, ex.z, ex.w}, cntv[4] = {cnt.x, cnt.y, cnt.z, cnt.w};
        const bool last = a.bases_out != nullptr && tile == a.tiles - 1;
#pragma unroll
        for (uint32_t j = 0; j < E; ++j) {
            const uint64_t b = a.bases_in[d0 + j];
            s_gbase[d0 + j] = (uint32_t)b + exv[j] - s_dp[d0 + j];
            if (last) a.bases_out[d0 + j] = b + exv[j] + cntv[j];
        }
    }
}

template <int H, int LBW, int LB>
__device__ __forceinline__ void lookback_quad_tile(const PassArgs& a, uint32_t* lb_row, uint32_t tile, uint32_t lbw, uint32_t lane,
                                                   uint32_t pads, uint32_t dmask, const uint32_t* s_tot, const uint32_t* s_dp,
                                                   uint32_t* s_gbase, uint32_t* rounds_out = nullptr, uint32_t* hops_out = nullptr)
{
    lookback_quad_publish<H, LBW, LB>(lb_row, tile, lbw, lane, pads, dmask, s_tot);
    lookback_quad_finish<H, LBW, LB>(a, lb_row, tile, lbw, lane, pads, dmask, s_tot, s_dp, s_gbase, rounds_out, hops_out);
}

}  // namespace lsd
