// onesweep_cpc.cuh -- onesweep digit pass, "column-private counter" (CPC) ranking.
//
// Same contract as onesweep.cuh (one stable LSD pass on an 8-bit digit: ticket order, decoupled look-back,
// shared-memory reorder, coalesced per-bucket scatter).  What changes is how a key gets its rank inside
// the tile.  Measured on B200 (bench_tools/microbench*.cu, profiles/r01_microbench.txt):
//   * 8 ballots per key cost 24 cycles per warp instruction SM-wide, match.any 60: warp multisplit tops
//     out at ~1.3 keys/clk/SM before any other work;
//   * a conflict-free returning shared atomic costs 1 cycle per warp instruction SM-wide, but ONE warp
//     gets one every ~16 cycles: ranking must keep many warps in their atomics at the same time;
//   * the LPC kernels (onesweep_lpc*.cuh) share a counter column between the warps of a CTA, so the
//     warps take their ranks one after the other (~9 K cycles per tile, one warp active).
// Here every THREAD owns a column of the tile, so program order is position order and no hand-over
// between warps exists:
//
//   tile    = 128 columns x SPC keys (SPC = 64: 8192 keys).  Thread (warp g, lane l) of the 128-thread CTA
//             owns column c = 4*l + g = tile positions [c*SPC, (c+1)*SPC).
//   staging = 32 TMA bulk copies of 4*SPC keys each (one per lane), landing 16 bytes apart from a
//             power-of-two pitch, so the 128-bit column reads of a quarter-warp hit 8 distinct bank groups.
//   matrix  = cnt[digit][l]: one 32-bit word holds the four 8-bit counters of columns 4l..4l+3 (byte g
//             belongs to warp g).  256 x 32 words = 32 KiB.  A warp instruction touches word `lane` of 32
//             rows: 32 distinct banks whatever the key distribution.  A counter never exceeds SPC <= 255.
//   rank    : old = atomicAdd(&cnt[d][l], 1 << 8g); byte g of `old` is the number of earlier keys of the
//             column with the same digit.  One returning atomic per key, all four warps at once.
//   scan    : thread t reads rows 2t, 2t+1 (diagonal 128-bit loads, conflict-free): row totals = tile
//             histogram (published to the look-back chain straight from registers), exclusive scan over
//             digits -> bucket starts, then Q[d][l] = bucket start + keys of digit d in columns < 4l
//             (16-bit, written over the dead staging buffer).
//   position: pos = Q[d][l] + (bytes below g of cnt[d][l], one masked dp4a) + rank; keys go to the reorder
//             buffer at pos and are streamed out per bucket as in the other kernels.
//
// Shared-memory pipe cycles per 32 keys (microbench2): TMA 1 + column read 0.5 + atomic 1 + matrix clear
// 1.1 + scan 0.5 + Q 0.6 + position loads 2 + reorder scatter 3.5 + linear read 1 + bucket base 1.2 = 12.4.
#pragma once
#include "onesweep_lpc32.cuh"

namespace lsd {

template <int RB, int SPC>
struct CpcShape {
    static_assert(RB == 8, "four byte counters per word and 256 rows: written for 8-bit digits");
    static_assert(SPC % 4 == 0 && SPC <= 252, "column length: 128-bit reads, 8-bit counters");
    static constexpr int H = 1 << RB;
    static constexpr int THREADS = 128;
    static constexpr int COLS = 128;
    static constexpr int TILE = COLS * SPC;
    static constexpr int GROUP_WORDS = 4 * SPC;          // one lane's four columns, contiguous in the input
    static constexpr int PITCH = GROUP_WORDS + 4;        // +16 bytes: bank-group skew for the 128-bit reads
    static constexpr int STAGE_WORDS = 32 * PITCH;       // staging buffer, later Q (first 16 KiB), later reorder
    static constexpr int OFF_MAT = STAGE_WORDS;          // [H][32]
    static constexpr int OFF_GBASE = OFF_MAT + H * 32;   // [H]
    static constexpr int OFF_MISC = OFF_GBASE + H;       // [0..3] warp partials, [8] tile id, [10..11] mbarrier
    static constexpr int WORDS = OFF_MISC + 16;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = (uint32_t)((((1u << 30) - 1u) / TILE) * TILE);
    static_assert(TILE + TILE / 256 <= STAGE_WORDS && H * 16 <= STAGE_WORDS, "reorder buffer and Q alias the staging buffer");
    static_assert(TILE < 65536, "positions are packed in 16 bits");
};

__device__ __forceinline__ uint32_t sum_bytes4(const uint4 v, uint32_t acc)
{
    return __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, acc))));
}

template <int RB, int SPC, int MINB, int SHIFT, int LB, int DBG>
__global__ void __launch_bounds__(128, MINB)
onesweep_cpc_kernel(const PassArgs a)
{
    using S_ = CpcShape<RB, SPC>;
    constexpr int H = S_::H, THREADS = S_::THREADS, TILE = S_::TILE, PITCH = S_::PITCH;

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_stage = smem;                  // staging (padded), then Q, then reorder buffer (linear)
    uint32_t* s_mat = smem + S_::OFF_MAT;
    uint32_t* s_gbase = smem + S_::OFF_GBASE;
    uint32_t* s_misc = smem + S_::OFF_MISC;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 10);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t g = tid >> 5;  // warp = byte field

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    const long long t_start = a.trace ? clock64() : 0;
#define LSD_TRACE(slot)                                                                                          \
    do {                                                                                                         \
        if (a.trace && tid == 0) a.trace[(size_t)tile * 16 + (slot)] = (unsigned long long)(clock64() - t_start); \
    } while (0)

    // ---- 0. ticket, clear the matrix, 32 TMA bulk copies (8 per warp: issuing one costs ~100 cycles) ----
    if (tid == 0) {
        mbar_init(s_bar, 1);
        const uint32_t t = atomicAdd(a.ticket, 1u);
        s_misc[8] = t;
        if (a.trace) a.trace[(size_t)t * 16 + 10] = (unsigned long long)(clock64() - t_start);  // ticket returned
        if (a.portion_keys - t * (uint32_t)TILE >= (uint32_t)TILE) mbar_expect_tx(s_bar, TILE * 4);
    }
    {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
        for (int i = 0; i < H * 8 / THREADS; ++i) m4[i * THREADS + tid] = make_uint4(0, 0, 0, 0);
    }
    const long long t_clear = a.trace ? clock64() : 0;
    __syncthreads();
    const uint32_t tile = s_misc[8];
    if (a.trace && tid == 64) a.trace[(size_t)tile * 16 + 11] = (unsigned long long)(t_clear - t_start);  // warp 2 cleared
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t left = a.portion_keys - tile_base;
    const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
    const uint32_t pads = (uint32_t)TILE - valid;
    if (valid == (uint32_t)TILE && lane < 8) {
        const uint32_t j = g * 8u + lane;  // lane group j of the tile: 4*SPC contiguous keys
        tma_bulk_g2s(s_stage + j * PITCH, in + tile_base + j * S_::GROUP_WORDS, S_::GROUP_WORDS * 4, s_bar);
    }

    LSD_TRACE(0);  // ticket + matrix clear
    if (a.trace && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        a.trace[(size_t)tile * 16 + 15] = gt;  // wall clock (ns) when the tile id is known
    }
    if (valid == (uint32_t)TILE) {
        mbar_wait(s_bar, 0);
    } else {
        for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS)
            s_stage[p + 4u * (p / (uint32_t)S_::GROUP_WORDS)] = p < valid ? in[tile_base + p] : 0xFFFFFFFFu;
        __syncthreads();
    }
    LSD_TRACE(1);  // tile landed

    // ---- 1. my column -> registers (128-bit reads) ----
    uint32_t key[SPC];
    {
        const uint4* src = reinterpret_cast<const uint4*>(s_stage + lane * PITCH + g * SPC);
#pragma unroll
        for (int i = 0; i < SPC / 4; ++i) {
            const uint4 v = src[i];
            key[4 * i + 0] = v.x;
            key[4 * i + 1] = v.y;
            key[4 * i + 2] = v.z;
            key[4 * i + 3] = v.w;
        }
    }

    // ---- 2. rank: one returning atomic per key; byte g of the old word = earlier equal digits in my column ----
    char* mat_bytes = reinterpret_cast<char*>(s_mat);
    const uint32_t lane4 = lane << 2;
    const uint32_t inc = 1u << (8u * g);
    uint32_t rk[SPC / 4];
    {
        // PRMT selectors: byte m of the result <- byte g of `old`, the other bytes keep rk
        uint32_t sel[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) sel[m] = (0x3210u & ~(0xFu << (4 * m))) | ((4u + g) << (4 * m));
#pragma unroll
        for (int j = 0; j < SPC; ++j) {
            const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[j], lane4)), inc);
            rk[j >> 2] = __byte_perm((j & 3) ? rk[j >> 2] : 0u, old, sel[j & 3]);
        }
    }
    LSD_TRACE(2);  // warp 0 ranked
    __syncthreads();  // matrix complete; every key is in registers: the staging buffer is free
    LSD_TRACE(3);

    // ---- 3. scan: thread t owns rows (digits) 2t and 2t+1 ----
    uint32_t* lb_row = a.lookback + (size_t)tile * H;
    const uint32_t q = lane & 7u;
    uint32_t total[2], start[2];
    {
        uint32_t below[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + (2u * tid + r) * 32u);
            total[r] = 0;
            below[r] = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                const uint32_t s = sum_bytes4(r4[grp], 0u);
                total[r] += s;
                if (grp < q) below[r] += s;  // groups reached after the wrap = columns before group q
            }
        }
        // publish the tile histogram as early as possible (pads of a ragged last tile are not keys)
        {
            const uint32_t cnt_lo = total[0];
            const uint32_t cnt_hi = total[1] - (tid == (uint32_t)THREADS - 1 ? pads : 0u);
            const uint32_t flag = tile == 0 ? kLbGlobal : kLbLocal;
            st_relaxed_gpu_v2(lb_row + 2 * tid, flag | cnt_lo, flag | cnt_hi);
        }
        // exclusive scan of the 256 row totals (thread order = digit order)
        const uint32_t pair = total[0] + total[1];
        uint32_t incl = pair;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_misc[g] = incl;
        __syncthreads();
        uint32_t prefix = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w)
            if ((uint32_t)w < g) prefix += s_misc[w];
        start[0] = prefix + incl - pair;
        start[1] = start[0] + total[0];
        // Q[d][l] = bucket start + keys of digit d in columns < 4l, 16 bits each, over the dead staging buffer
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t row = 2u * tid + r;
            const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + row * 32u);
            uint2* q2 = reinterpret_cast<uint2*>(s_stage + row * 16u);
            uint4* q4 = reinterpret_cast<uint4*>(s_stage + row * 32u);  // 32-bit Q rows (DBG & 16): conflict-free reads
            uint32_t run = start[r] + below[r];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t grp = (q + k) & 7u;
                if (grp == 0) run = start[r];
                const uint4 v = r4[grp];
                const uint32_t q0 = run;
                run = __dp4a(v.x, 0x01010101u, run);
                const uint32_t q1 = run;
                run = __dp4a(v.y, 0x01010101u, run);
                const uint32_t q2v = run;
                run = __dp4a(v.z, 0x01010101u, run);
                const uint32_t q3 = run;
                run = __dp4a(v.w, 0x01010101u, run);
                if constexpr ((DBG & 16) != 0) q4[grp] = make_uint4(q0, q1, q2v, q3);
                else q2[grp] = make_uint2(q0 | (q1 << 16), q2v | (q3 << 16));
            }
        }
    }
    __syncthreads();  // Q complete
    LSD_TRACE(4);

    // ---- 4. positions: Q + counters of my word below my byte + my rank ----
    uint32_t pk[SPC / 2];
    {
        const char* q_bytes = reinterpret_cast<const char*>(s_stage);
        const uint32_t below_mask = inc - 1u;
#pragma unroll
        for (int j = 0; j < SPC; ++j) {
            const uint32_t off = cell_offset<RB, SHIFT>(key[j], lane4);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(mat_bytes + off);
            const uint32_t qv = (DBG & 16) ? *reinterpret_cast<const uint32_t*>(q_bytes + off)
                                           : (uint32_t)*reinterpret_cast<const uint16_t*>(q_bytes + (off >> 1));
            const uint32_t r = (rk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            const uint32_t pos = __dp4a(w & below_mask, 0x01010101u, qv + r);
            if (j & 1) pk[j >> 1] |= pos << 16; else pk[j >> 1] = pos;
        }
    }
    __syncthreads();  // every read of Q and of the matrix is done: the buffer becomes the reorder buffer
    LSD_TRACE(5);

    // ---- 5. scatter into the reorder buffer ----
#pragma unroll
    for (int j = 0; j < SPC; ++j) {
        uint32_t pos = (j & 1) ? (pk[j >> 1] >> 16) : (pk[j >> 1] & 0xFFFFu);
        if (DBG & 8) pos = (pos & 0x1000u) ? pos : j * THREADS + tid;  // timing experiment: conflict-free scatter
        if (DBG & 32) pos += pos >> 8;  // skewed reorder layout: one word of padding per 256 positions keeps the columns
                                        // of a single-bucket (sorted / constant-digit) tile in distinct banks
        s_stage[pos] = key[j];
    }
    LSD_TRACE(6);

    // ---- 6. decoupled look-back: thread t resolves digits 2t, 2t+1 ----
    {
        const uint32_t cnt_lo = total[0];
        const uint32_t cnt_hi = total[1] - (tid == (uint32_t)THREADS - 1 ? pads : 0u);
        uint32_t ex_lo = 0, ex_hi = 0;
        if (LB == 0) ex_lo = ex_hi = tile * (uint32_t)(TILE / H);  // timing experiment: plausible addresses for uniform keys
        if (LB > 0 && tile != 0) {  // LB == 0: timing experiment only (no look-back, output positions wrong)
            const uint32_t* p = lb_row - H + 2 * tid;
            uint32_t remaining = tile;
            bool done = false;
            uint32_t dbg_rounds = 0, dbg_hops = 0;
            while (!done) {
                ++dbg_rounds;
                uint2 w[LB > 0 ? LB : 1];
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
            if (a.trace && tid == 0) {
                a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                a.trace[(size_t)tile * 16 + 14] = dbg_hops;
            }
            st_relaxed_gpu_v2(lb_row + 2 * tid, kLbGlobal | (ex_lo + cnt_lo), kLbGlobal | (ex_hi + cnt_hi));
        }
        const uint64_t b_lo = a.bases_in[2 * tid], b_hi = a.bases_in[2 * tid + 1];
        s_gbase[2 * tid] = (uint32_t)b_lo + ex_lo - start[0];
        s_gbase[2 * tid + 1] = (uint32_t)b_hi + ex_hi - start[1];
        if (a.bases_out != nullptr && tile == a.tiles - 1) {
            a.bases_out[2 * tid] = b_lo + ex_lo + cnt_lo;
            a.bases_out[2 * tid + 1] = b_hi + ex_hi + cnt_hi;
        }
    }
    LSD_TRACE(7);  // look-back done (thread 0)
    __syncthreads();
    LSD_TRACE(8);

    // ---- 7. stream the reorder buffer out, coalesced per bucket ----
    if (valid == (uint32_t)TILE) {
#pragma unroll 16
        for (int i = 0; i < TILE / THREADS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t k = s_stage[(DBG & 32) ? p + (p >> 8) : p];
            if (DBG & 2) {  // timing experiment: full-line coalesced stores
                out[a.portion_base + tile_base + p] = k + s_gbase[(k >> SHIFT) & (H - 1)];
            } else if (DBG & 4) {  // timing experiment: no global stores
                if (k + s_gbase[(k >> SHIFT) & (H - 1)] == 0x12345u && p == 77u) out[0] = k;
            } else {
                out[s_gbase[(k >> SHIFT) & (H - 1)] + p] = k;
            }
        }
    } else {
        for (uint32_t p = tid; p < valid; p += THREADS) {
            const uint32_t k = s_stage[(DBG & 32) ? p + (p >> 8) : p];
            out[s_gbase[(k >> SHIFT) & (H - 1)] + p] = k;
        }
    }
    LSD_TRACE(9);
#undef LSD_TRACE
}

template <int RB, int SPC, int MINB, int SHIFT, int LB, int DBG>
int onesweep_cpc_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = CpcShape<RB, SPC>;
    auto kern = onesweep_cpc_kernel<RB, SPC, MINB, SHIFT, LB, DBG>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<a.tiles, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int SPC, int MINB, int LB, int DBG>
int onesweep_cpc_launch(const PassArgs& a, cudaStream_t s)
{
    switch (a.shift) {
        case 0: return onesweep_cpc_launch_shift<RB, SPC, MINB, 0, LB, DBG>(a, s);
        case 8: return onesweep_cpc_launch_shift<RB, SPC, MINB, 8, LB, DBG>(a, s);
        case 16: return onesweep_cpc_launch_shift<RB, SPC, MINB, 16, LB, DBG>(a, s);
        case 24: return onesweep_cpc_launch_shift<RB, SPC, MINB, 24, LB, DBG>(a, s);
    }
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeCpc = 5;

template <int RB, int SPC, int MINB, int LB = 8, int DBG = 0>
constexpr OnesweepLauncher make_cpc_launcher()
{
    using S_ = CpcShape<RB, SPC>;
    return OnesweepLauncher{RB, S_::THREADS, SPC, kModeCpc, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                            &onesweep_cpc_launch<RB, SPC, MINB, LB, DBG>};
}

}  // namespace lsd
