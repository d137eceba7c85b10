// onesweep_cpcp.cuh -- CPC digit pass as a persistent, warp-specialised pipeline (one CTA per SM).
//
// Why: the one-tile-per-CTA kernels (onesweep_lpc32.cuh, onesweep_cpc.cuh) run a tile's phases back to
// back -- ticket, load, rank, scan, positions, scatter, look-back, copy-out: ~19 K cycles, of which more
// than half is waiting on global round trips (profiles/r01_cpc_trace.txt) -- and only three tiles fit an
// SM, so neither the shared-memory pipe nor HBM is kept busy (ablations in profiles/r01_cpc_ablation.txt:
// the same kernel without look-back and global stores still needs 0.47 ms per pass).  Here the phases of
// DIFFERENT tiles overlap inside one CTA that stays on its SM for the whole pass:
//
//   producer warp   : takes tickets, streams tile i+NB-1.. into a ring of NB staging buffers with 16-byte
//                     cp.async copies (padded column layout, see onesweep_cpc.cuh), completion on an mbarrier;
//   2 front groups  : (4 warps each, own 32 KiB counter matrix) rank / scan / position / scatter alternate
//                     tiles exactly as the CPC kernel does, publish the tile histogram to the look-back chain
//                     as soon as it is known and leave the sorted tile in its buffer;
//   2 back groups   : (4 warps each) resolve the look-back of alternate tiles while the front groups are
//                     already on the next tiles, then stream the sorted tile out and free the buffer.
//
// Hand-overs are mbarriers in shared memory (full / hist / sorted / empty per buffer); groups synchronise
// internally on named barriers.  Tiles are still taken in ticket order and every published tile only ever
// waits on tiles with smaller tickets, which are resident and past their own histogram: no deadlock.
#pragma once
#include "onesweep_cpc.cuh"
#include "onesweep_lpcp.cuh"  // mbar_arrive

namespace lsd {

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <int RB, int SPC, int NB>
struct CpcpShape {
    using C = CpcShape<RB, SPC>;
    static constexpr int H = C::H;
    static constexpr int TILE = C::TILE;
    static constexpr int GROUP_WORDS = C::GROUP_WORDS;
    static constexpr int PITCH = C::PITCH;
    static constexpr int BUF_WORDS = C::STAGE_WORDS;
    static constexpr int GROUP_THREADS = 128;
    static constexpr int THREADS = 5 * GROUP_THREADS;  // 2 front groups, 2 back groups, producer group
    static constexpr int OFF_MAT = NB * BUF_WORDS;           // [2][H][32]
    static constexpr int OFF_TOT = OFF_MAT + 2 * H * 32;     // [NB][H] tile digit counts (pads removed)
    static constexpr int OFF_START = OFF_TOT + NB * H;       // [NB][H] tile-local bucket starts
    static constexpr int OFF_GBASE = OFF_START + NB * H;     // [NB][H]
    static constexpr int OFF_TILE = OFF_GBASE + NB * H;      // [NB] tile ids (>= tiles: stop token)
    static constexpr int OFF_PART = OFF_TILE + 8;            // [2][4] warp partials of the digit scan
    static constexpr int OFF_BAR = OFF_PART + 8;             // [4][NB] mbarriers (64-bit)
    static constexpr int WORDS = OFF_BAR + 2 * 4 * NB;
    static constexpr size_t SMEM_BYTES = sizeof(uint32_t) * WORDS + 16;
    static constexpr uint32_t PORTION_MAX = C::PORTION_MAX;
    static_assert(NB <= 8, "tile-id slots");
    static_assert(OFF_BAR % 2 == 0, "mbarriers are 8-byte aligned");
};

template <int RB, int SPC, int NB, int SHIFT, int LB, int FRONT_REGS>
__global__ void __launch_bounds__(640, 1)
onesweep_cpcp_kernel(const PassArgs a)
{
    using S_ = CpcpShape<RB, SPC, NB>;
    constexpr int H = S_::H, TILE = S_::TILE, PITCH = S_::PITCH, GT = S_::GROUP_THREADS;

    if (a.plan->skip[a.pass]) return;

    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t* s_tile = smem + S_::OFF_TILE;
    uint64_t* s_bars = reinterpret_cast<uint64_t*>(smem + S_::OFF_BAR);
    uint64_t* bar_full = s_bars;             // producer -> front: tile landed
    uint64_t* bar_hist = s_bars + NB;        // front -> back: counts and bucket starts are in shared memory
    uint64_t* bar_sorted = s_bars + 2 * NB;  // front -> back: the buffer holds the sorted tile
    uint64_t* bar_empty = s_bars + 3 * NB;   // back -> producer: the buffer may be refilled

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t role = tid >> 7;  // 0,1 front groups; 2,3 back groups; 4 producer warp
    const uint32_t gtid = tid & 127u;

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;
    const uint32_t tiles = a.tiles;

    const long long t_start = a.trace ? clock64() : 0;
#define LSD_TRACE(cond, tile_, slot)                                                                                \
    do {                                                                                                            \
        if (a.trace && (cond)) a.trace[(size_t)(tile_) * 16 + (slot)] = (unsigned long long)(clock64() - t_start);  \
    } while (0)

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            mbar_init(bar_full + b, GT);
            mbar_init(bar_hist + b, GT);
            mbar_init(bar_sorted + b, GT);
            mbar_init(bar_empty + b, GT);
        }
    }
    __syncthreads();

    if (role == 4) {
        // =========================== producer group ===========================
        // Register pool (640 threads x 96 at launch): front 2 x 128 x FRONT_REGS, back 2 x 128 x 64, producer 128 x 40.
        if constexpr (FRONT_REGS > 0) setmaxnreg_dec<40>();
        uint32_t stops = 0;
        const uint32_t pbar = 5u;
        for (uint32_t i = 0;; ++i) {
            const uint32_t b = i % NB, n = i / NB;
            if (n > 0) mbar_wait(bar_empty + b, (n - 1) & 1u);
            if (gtid == 0) s_tile[b] = atomicAdd(a.ticket, 1u);
            named_bar_sync(pbar, GT);
            const uint32_t t = s_tile[b];
            if (t >= tiles) {
                mbar_arrive(bar_full + b);  // stop token: one per front/back group pair
                if (++stops == 2) break;
                continue;
            }
            LSD_TRACE(gtid == 0, t, 0);
            uint32_t* buf = smem + b * S_::BUF_WORDS;
            const uint32_t tile_base = t * (uint32_t)TILE;
            const uint32_t left = a.portion_keys - tile_base;
            if (left >= (uint32_t)TILE) {
                const uint32_t* src = in + tile_base;
#pragma unroll
                for (int it = 0; it < TILE / 4 / GT; ++it) {
                    const uint32_t c16 = it * GT + gtid;                         // 16-byte chunk of the tile
                    const uint32_t grp = c16 / (uint32_t)(S_::GROUP_WORDS / 4);  // lane group it belongs to
                    cp_async_16(buf + c16 * 4u + grp * 4u, src + c16 * 4u);
                }
                cp_async_mbar_arrive_noinc(bar_full + b);
            } else {
                for (uint32_t p = gtid; p < (uint32_t)TILE; p += GT)
                    buf[p + 4u * (p / (uint32_t)S_::GROUP_WORDS)] = p < left ? in[tile_base + p] : 0xFFFFFFFFu;
                mbar_arrive(bar_full + b);
            }
        }
    } else if (role < 2) {
        // =========================== front groups: rank, scan, positions, scatter ===========================
        if constexpr (FRONT_REGS > 0) setmaxnreg_inc<FRONT_REGS>();
        const uint32_t f = role;
        const uint32_t g = gtid >> 5;  // warp inside the group = byte field
        uint32_t* s_mat = smem + S_::OFF_MAT + f * (H * 32);
        uint32_t* s_part = smem + S_::OFF_PART + f * 4;
        char* mat_bytes = reinterpret_cast<char*>(s_mat);
        const uint32_t lane4 = lane << 2;
        const uint32_t inc = 1u << (8u * g);
        const uint32_t bar_id = 1u + f;
        uint32_t sel[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) sel[m] = (0x3210u & ~(0xFu << (4 * m))) | ((4u + g) << (4 * m));
        const uint32_t q = lane & 7u;

        for (uint32_t i = f;; i += 2) {
            const uint32_t b = i % NB, n = i / NB;
            uint32_t* s_stage = smem + b * S_::BUF_WORDS;
            {
                uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
                for (int k = 0; k < H * 8 / GT; ++k) m4[k * GT + gtid] = make_uint4(0, 0, 0, 0);
            }
            mbar_wait(bar_full + b, n & 1u);
            const uint32_t tile = s_tile[b];
            if (tile >= tiles) {
                mbar_arrive(bar_hist + b);  // pass the stop token on
                break;
            }
            const uint32_t left = a.portion_keys - tile * (uint32_t)TILE;
            const uint32_t pads = left < (uint32_t)TILE ? (uint32_t)TILE - left : 0u;
            named_bar_sync(bar_id, GT);  // matrix cleared by everyone
            LSD_TRACE(gtid == 0, tile, 1);

            uint32_t key[SPC];
            {
                const uint4* src = reinterpret_cast<const uint4*>(s_stage + lane * PITCH + g * SPC);
#pragma unroll
                for (int k = 0; k < SPC / 4; ++k) {
                    const uint4 v = src[k];
                    key[4 * k + 0] = v.x;
                    key[4 * k + 1] = v.y;
                    key[4 * k + 2] = v.z;
                    key[4 * k + 3] = v.w;
                }
            }
            uint32_t rk[SPC / 4];
#pragma unroll
            for (int j = 0; j < SPC; ++j) {
                const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[j], lane4)), inc);
                rk[j >> 2] = __byte_perm((j & 3) ? rk[j >> 2] : 0u, old, sel[j & 3]);
            }
            named_bar_sync(bar_id, GT);  // matrix complete, every key in registers
            LSD_TRACE(gtid == 0, tile, 2);

            // ---- scan: thread t owns rows (digits) 2t, 2t+1 ----
            {
                uint32_t total[2], below[2], start[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + (2u * gtid + r) * 32u);
                    total[r] = 0;
                    below[r] = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t grp = (q + k) & 7u;
                        const uint32_t s = sum_bytes4(r4[grp], 0u);
                        total[r] += s;
                        if (grp < q) below[r] += s;
                    }
                }
                const uint32_t cnt_lo = total[0];
                const uint32_t cnt_hi = total[1] - (gtid == (uint32_t)GT - 1 ? pads : 0u);
                {
                    const uint32_t flag = tile == 0 ? kLbGlobal : kLbLocal;
                    st_relaxed_gpu_v2(a.lookback + (size_t)tile * H + 2 * gtid, flag | cnt_lo, flag | cnt_hi);
                }
                const uint32_t pair = total[0] + total[1];
                uint32_t incl = pair;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                if (lane == 31) s_part[g] = incl;
                named_bar_sync(bar_id, GT);
                uint32_t prefix = 0;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if ((uint32_t)w < g) prefix += s_part[w];
                start[0] = prefix + incl - pair;
                start[1] = start[0] + total[0];
                // hand the histogram and the bucket starts to the back group
                *reinterpret_cast<uint2*>(smem + S_::OFF_TOT + b * H + 2 * gtid) = make_uint2(cnt_lo, cnt_hi);
                *reinterpret_cast<uint2*>(smem + S_::OFF_START + b * H + 2 * gtid) = make_uint2(start[0], start[1]);
                mbar_arrive(bar_hist + b);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint32_t row = 2u * gtid + r;
                    const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + row * 32u);
                    uint2* q2 = reinterpret_cast<uint2*>(s_stage + row * 16u);
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
                        q2[grp] = make_uint2(q0 | (q1 << 16), q2v | (q3 << 16));
                    }
                }
            }
            named_bar_sync(bar_id, GT);  // Q complete
            LSD_TRACE(gtid == 0, tile, 3);

            uint32_t pk[SPC / 2];
            {
                const char* q_bytes = reinterpret_cast<const char*>(s_stage);
                const uint32_t below_mask = inc - 1u;
#pragma unroll
                for (int j = 0; j < SPC; ++j) {
                    const uint32_t off = cell_offset<RB, SHIFT>(key[j], lane4);
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(mat_bytes + off);
                    const uint32_t qv = *reinterpret_cast<const uint16_t*>(q_bytes + (off >> 1));
                    const uint32_t r = (rk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                    const uint32_t pos = __dp4a(w & below_mask, 0x01010101u, qv + r);
                    if (j & 1) pk[j >> 1] |= pos << 16; else pk[j >> 1] = pos;
                }
            }
            named_bar_sync(bar_id, GT);  // all reads of Q and the matrix done
            LSD_TRACE(gtid == 0, tile, 4);
#pragma unroll
            for (int j = 0; j < SPC; ++j) {
                const uint32_t pos = (j & 1) ? (pk[j >> 1] >> 16) : (pk[j >> 1] & 0xFFFFu);
                s_stage[pos] = key[j];
            }
            mbar_arrive(bar_sorted + b);
            LSD_TRACE(gtid == 0, tile, 5);
        }
    } else {
        // =========================== back groups: look-back, copy-out ===========================
        if constexpr (FRONT_REGS > 0) setmaxnreg_dec<64>();
        const uint32_t f = role - 2u;
        const uint32_t bar_id = 3u + f;
        for (uint32_t i = f;; i += 2) {
            const uint32_t b = i % NB, n = i / NB;
            mbar_wait(bar_hist + b, n & 1u);
            const uint32_t tile = s_tile[b];
            if (tile >= tiles) break;
            const uint32_t tile_base = tile * (uint32_t)TILE;
            const uint32_t left = a.portion_keys - tile_base;
            const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
            uint32_t* s_gbase = smem + S_::OFF_GBASE + b * H;
            {
                const uint2 cnt = *reinterpret_cast<const uint2*>(smem + S_::OFF_TOT + b * H + 2 * gtid);
                const uint2 st = *reinterpret_cast<const uint2*>(smem + S_::OFF_START + b * H + 2 * gtid);
                uint32_t* lb_row = a.lookback + (size_t)tile * H;
                uint32_t ex_lo = 0, ex_hi = 0;
                if (tile != 0) {
                    const uint32_t* p = lb_row - H + 2 * gtid;
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
                    if (a.trace && gtid == 0) {
                        a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                        a.trace[(size_t)tile * 16 + 14] = dbg_hops;
                    }
                    st_relaxed_gpu_v2(lb_row + 2 * gtid, kLbGlobal | (ex_lo + cnt.x), kLbGlobal | (ex_hi + cnt.y));
                }
                const uint64_t b_lo = a.bases_in[2 * gtid], b_hi = a.bases_in[2 * gtid + 1];
                s_gbase[2 * gtid] = (uint32_t)b_lo + ex_lo - st.x;
                s_gbase[2 * gtid + 1] = (uint32_t)b_hi + ex_hi - st.y;
                if (a.bases_out != nullptr && tile == tiles - 1) {
                    a.bases_out[2 * gtid] = b_lo + ex_lo + cnt.x;
                    a.bases_out[2 * gtid + 1] = b_hi + ex_hi + cnt.y;
                }
            }
            LSD_TRACE(gtid == 0, tile, 6);
            named_bar_sync(bar_id, GT);  // bucket bases complete
            mbar_wait(bar_sorted + b, n & 1u);
            LSD_TRACE(gtid == 0, tile, 7);
            const uint32_t* s_rb = smem + b * S_::BUF_WORDS;
            if (valid == (uint32_t)TILE) {
#pragma unroll 16
                for (int k = 0; k < TILE / GT; ++k) {
                    const uint32_t p = k * GT + gtid;
                    const uint32_t key = s_rb[p];
                    out[s_gbase[(key >> SHIFT) & (H - 1)] + p] = key;
                }
            } else {
                for (uint32_t p = gtid; p < valid; p += GT) {
                    const uint32_t key = s_rb[p];
                    out[s_gbase[(key >> SHIFT) & (H - 1)] + p] = key;
                }
            }
            mbar_arrive(bar_empty + b);
            LSD_TRACE(gtid == 0, tile, 8);
        }
    }
#undef LSD_TRACE
}

template <int RB, int SPC, int NB, int SHIFT, int LB, int FRONT_REGS>
int onesweep_cpcp_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = CpcpShape<RB, SPC, NB>;
    auto kern = onesweep_cpcp_kernel<RB, SPC, NB, SHIFT, LB, FRONT_REGS>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    const uint32_t grid = a.tiles < (uint32_t)sm_count() ? a.tiles : (uint32_t)sm_count();
    kern<<<grid, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int SPC, int NB, int LB, int FRONT_REGS>
int onesweep_cpcp_launch(const PassArgs& a, cudaStream_t s)
{
    switch (a.shift) {
        case 0: return onesweep_cpcp_launch_shift<RB, SPC, NB, 0, LB, FRONT_REGS>(a, s);
        case 8: return onesweep_cpcp_launch_shift<RB, SPC, NB, 8, LB, FRONT_REGS>(a, s);
        case 16: return onesweep_cpcp_launch_shift<RB, SPC, NB, 16, LB, FRONT_REGS>(a, s);
        case 24: return onesweep_cpcp_launch_shift<RB, SPC, NB, 24, LB, FRONT_REGS>(a, s);
    }
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeCpcp = 6;

template <int RB, int SPC, int NB, int LB = 8, int FRONT_REGS = 0>
constexpr OnesweepLauncher make_cpcp_launcher()
{
    using S_ = CpcpShape<RB, SPC, NB>;
    return OnesweepLauncher{RB, S_::THREADS, SPC, kModeCpcp, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                            &onesweep_cpcp_launch<RB, SPC, NB, LB, FRONT_REGS>};
}

}  // namespace lsd
