// onesweep_lpc2.cuh -- LPC onesweep pass, fourth shape: TWO rank chains and (optionally) one look-back record per
// thread-block CLUSTER.
//
// Why (bench_tools/trace.py on onesweep_lpc32_kernel, profiles/r01_lpc32_lookback_trace.txt): a tile lives ~18.8 K
// cycles, and two serial stretches bound it together:
//   * the rank chain -- one warp at a time takes its ranks from the counter matrix, 9 turns x ~870 cycles, because a
//     warp has one returning shared atomic in flight (~30 cycles each, 29 per turn);
//   * the look-back -- ~25 hops x ~360 cycles.  With P tiles in flight a new tile starts every tau = life / P cycles
//     (~42), and a tile meets its first INCLUSIVE predecessor after k = eps / tau hops (eps = the ~1000-cycle store ->
//     poll round trip), so the walk is long BECAUSE the tiles are small and many.
// Here
//   * cnt[digit][lane] keeps TWO 16-bit byte-offset counters per word: the low half for the even warps, the high half
//     for the odd warps.  A lane's segment of the tile is [even warps' keys | odd warps' keys], so the two halves are
//     independent columns of the position order and the even and the odd warps form two rank chains that run
//     concurrently (5 + 4 turns instead of 9).  Same 32 KiB matrix, same two atomics per key.  The scan handles the
//     packed words without unpacking (no cell exceeds 145 keys = 580 bytes, a row half never exceeds 18560: no carry
//     between the halves);
//   * with CL > 1 the kernel is launched in clusters of CL CTAs that take CL consecutive tiles under one ticket.  The
//     CTAs read each other's tile histograms through distributed shared memory (tile-exclusive offsets inside the
//     cluster), the last CTA publishes ONE look-back record for the cluster, and every CTA walks the cluster records:
//     tau grows CL-fold, the walk shrinks CL-fold, and so does the look-back traffic (1 KiB per hop per CTA).
// Everything else (TMA-staged tile, lane-blocked ownership, windowed look-back by the tail warps, byte-offset ranks,
// shared-memory reorder, coalesced per-bucket copy-out) is onesweep_lpc32.cuh.  Plain key passes only; the
// peer-scatter and key-value forms stay on onesweep_lpc32_kernel (same tile size, same workspace layout).
#pragma once
#include "onesweep_lpc32.cuh"

namespace lsd {

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
// address of `local_smem` in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_dsmem_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_dsmem_v2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

// polling load that bypasses L1 without the strong-load path: 344 vs 423 cycles unloaded, 26 vs 40 cycles per further
// load of a batch (profiles/r01_microbench_strong_loads.txt).  Flag and value share the word, so no ordering is needed.
__device__ __forceinline__ uint2 ld_cg_v2(const uint32_t* p)
{
    uint2 v;
    asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CL, int POLL>
__global__ void __launch_bounds__(WARPS * 32, MINB)
onesweep_lpc2_kernel(const PassArgs a)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    constexpr int H = S_::H, THREADS = S_::THREADS, S = S_::S, TILE = S_::TILE;
    constexpr int SW = S_::SW, GPW = S_::GPW, LBT = S_::LBT, LBW = S_::LBW;
    constexpr int EVEN = (WARPS + 1) / 2;  // warps of the even chain; they own the first EVEN*ITEMS keys of a lane segment
    constexpr uint32_t kBarTot = 14, kBarScan = 15;
    static_assert(SW >= 2, "warps 0 and 1 (the heads of the two chains) must both be scan warps");
    static_assert(LBT == H / 2, "one digit pair per look-back thread");
    static_assert(EVEN * ITEMS * 4 * 32 < 65536, "a row half must not carry into the other half");

    if (a.plan->skip[a.pass]) return;  // uniform over the grid (and so over every cluster)

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
    const uint32_t half = warp & 1u;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const bool is_scan = warp < (uint32_t)SW;
    const bool is_lb = warp >= (uint32_t)(WARPS - LBW);

    const bool src_scratch = a.plan->src_is_scratch[a.pass] != 0;
    const uint32_t* __restrict__ in = (src_scratch ? a.scratch : a.keys) + a.portion_base;
    uint32_t* __restrict__ out = src_scratch ? a.keys : a.scratch;

    const long long t_start = a.trace ? clock64() : 0;
#define LSD_TRACE(slot)                                                                      \
    do {                                                                                     \
        if (a.trace && lane == 0 && tile < a.tiles) a.trace[(size_t)tile * 16 + (slot)] = (unsigned long long)(clock64() - t_start); \
    } while (0)

    // ---- 0. ticket (one per cluster), TMA bulk load, clear the matrix ----
    if (tid == 0) {
        mbar_init(s_bar, 1);
        if constexpr (CL == 1) {
            const uint32_t t = atomicAdd(a.ticket, 1u);
            s_misc[32] = t;
            const uint32_t base = t * (uint32_t)TILE;
            if (base < a.portion_keys && a.portion_keys - base >= (uint32_t)TILE) {
                mbar_expect_tx(s_bar, TILE * 4);
                tma_bulk_g2s(s_keys, in + base, TILE * 4, s_bar);
            }
        } else if (crank == 0) {
            s_misc[33] = atomicAdd(a.ticket, 1u);
        }
    }
    {
        uint4* m4 = reinterpret_cast<uint4*>(s_mat);
#pragma unroll
        for (uint32_t i = tid; i < H * 8; i += THREADS) m4[i] = make_uint4(0, 0, 0, 0);
    }
    uint32_t tile;
    if constexpr (CL == 1) {
        __syncthreads();
        tile = s_misc[32];
    } else {
        // cluster barrier #0: every CTA of the cluster is running and rank 0's ticket is visible (it also orders this
        // CTA's own matrix clear and mbarrier init, as __syncthreads would)
        cluster_arrive();
        cluster_wait();
        tile = ld_dsmem_u32(dsmem_addr(s_misc + 33, 0)) * (uint32_t)CL + crank;
        if (tid == 0) {
            const uint32_t base = tile * (uint32_t)TILE;
            if (base < a.portion_keys && a.portion_keys - base >= (uint32_t)TILE) {
                mbar_expect_tx(s_bar, TILE * 4);
                tma_bulk_g2s(s_keys, in + base, TILE * 4, s_bar);
            }
        }
    }
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t left = tile_base < a.portion_keys ? a.portion_keys - tile_base : 0u;  // 0: a CTA past the last tile
    const uint32_t valid = left < (uint32_t)TILE ? left : (uint32_t)TILE;
    const uint32_t pads = (uint32_t)TILE - valid;

    if (warp == 0) LSD_TRACE(0);  // ticket + matrix clear done
    if (valid == (uint32_t)TILE) {
        mbar_wait(s_bar, 0);
    } else {
        for (uint32_t p = tid; p < (uint32_t)TILE; p += THREADS) s_keys[p] = p < valid ? in[tile_base + p] : 0xFFFFFFFFu;
        __syncthreads();
    }
    if (warp == 0) LSD_TRACE(1);  // tile landed

    // ---- 1. lane-blocked read + count.  Lane segment = [even warps | odd warps], sub-segment order inside each ----
    uint32_t key[ITEMS];
    {
        const uint32_t* src = s_keys + lane * S + (half ? EVEN * ITEMS : 0) + (warp >> 1) * ITEMS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = src[i];
    }
    char* mat_bytes = reinterpret_cast<char*>(s_mat);
    const uint32_t lane4 = lane << 2;
    const uint32_t addc = 4u << (16u * half);  // +4 bytes in this warp's half of the cell
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
        atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[i], lane4)), addc);
    if (warp == 0) LSD_TRACE(2);
    __syncthreads();  // counts complete; all keys are in registers: s_keys is now the reorder buffer
    if (warp == 0) LSD_TRACE(3);
    if constexpr (CL > 1) {
        if (!is_scan) cluster_arrive();  // cluster barrier #1 (the scan warps arrive once their totals are written)
    }

    if (is_scan) {
        // ================= scan warps: totals -> bucket starts -> exclusive (lane, half) prefix =================
        const uint32_t q = lane & 7u;
        uint32_t total[GPW], below[GPW];
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
            const uint32_t row = (uint32_t)(g * SW + warp) * 32u + lane;
            uint32_t tp = 0, bp = 0;  // packed sums: low half = even columns, high half = odd columns
            if (row < (uint32_t)H) {
                const uint4* r4 = reinterpret_cast<const uint4*>(s_mat + row * 32u);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t grp = (q + k) & 7u;
                    const uint4 v = r4[grp];
                    const uint32_t s = v.x + v.y + v.z + v.w;
                    tp += s;
                    if (grp < q) bp += s;
                }
            }
            total[g] = (tp & 0xFFFFu) + (tp >> 16);
            below[g] = (bp & 0xFFFFu) + (bp >> 16);
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
        named_bar_sync(kBarScan, SW * 32);
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
                s_tot[row] = (total[g] >> 2) - (row == (uint32_t)H - 1 ? pads : 0u);  // pads of a ragged tile are not keys
                s_dp[row] = start[g] >> 2;
            }
        }
        named_bar_arrive(kBarTot, (SW + LBW) * 32);  // totals + starts are in shared memory: look-back warps may go
        if constexpr (CL > 1) cluster_arrive();      // ... and, once every CTA of the cluster got here, their peers too
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
                    uint32_t lo;
                    lo = v.x & 0xFFFFu; o.x = run | ((run + lo) << 16); run += lo + (v.x >> 16);
                    lo = v.y & 0xFFFFu; o.y = run | ((run + lo) << 16); run += lo + (v.y >> 16);
                    lo = v.z & 0xFFFFu; o.z = run | ((run + lo) << 16); run += lo + (v.z >> 16);
                    lo = v.w & 0xFFFFu; o.w = run | ((run + lo) << 16); run += lo + (v.w >> 16);
                    r4[grp] = o;
                }
            }
        }
        named_bar_sync(kBarScan, SW * 32);  // matrix complete before warps 0 and 1 open the two rank chains
        if (warp == 0) LSD_TRACE(5);
    } else if (is_lb) {
        // ================= look-back warps (tails of the rank chains): one digit pair per thread =================
        named_bar_sync(kBarTot, (SW + LBW) * 32);
        if (warp == (uint32_t)WARPS - 1) LSD_TRACE(8);
        const uint32_t dt = tid - (uint32_t)(THREADS - LBT);
        const uint32_t cnt_lo = s_tot[2 * dt], cnt_hi = s_tot[2 * dt + 1];
        const uint32_t dp_lo = s_dp[2 * dt], dp_hi = s_dp[2 * dt + 1];
        uint32_t in_lo = 0, in_hi = 0;  // keys of this digit pair in the earlier tiles of the cluster
        if constexpr (CL > 1) {
            cluster_wait();  // cluster barrier #1: the tile histograms of all CTAs of the cluster are in their shared memories
#pragma unroll
            for (int r = 0; r < CL - 1; ++r)
                if ((uint32_t)r < crank) {
                    const uint2 v = ld_dsmem_v2(dsmem_addr(s_tot + 2 * dt, (uint32_t)r));
                    in_lo += v.x;
                    in_hi += v.y;
                }
            cluster_arrive();  // cluster barrier #2 (waited for at the very end): done with the peers' shared memory
        }
        const bool publisher = crank == (uint32_t)CL - 1u;  // the last CTA holds the cluster's aggregate
        const uint32_t agg_lo = in_lo + cnt_lo, agg_hi = in_hi + cnt_hi;
        const uint32_t rec = tile / (uint32_t)CL;            // one look-back record per cluster
        uint32_t* lb_row = a.lookback + (size_t)rec * H;
        uint32_t ex_lo = 0, ex_hi = 0;
        if (rec == 0) {
            if (publisher) st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | agg_lo, kLbGlobal | agg_hi);
        } else {
            if (publisher) st_relaxed_gpu_v2(lb_row + 2 * dt, kLbLocal | agg_lo, kLbLocal | agg_hi);
            const uint32_t* p = lb_row - H + 2 * dt;
            uint32_t remaining = rec;
            bool done = false;
            uint32_t dbg_rounds = 0, dbg_hops = 0;
            while (!done) {
                ++dbg_rounds;
                uint2 w[LB];
#pragma unroll
                for (int k = 0; k < LB; ++k)
                    w[k] = (uint32_t)k < remaining ? (POLL == 1 ? ld_cg_v2(p - (size_t)k * H) : ld_relaxed_gpu_v2(p - (size_t)k * H))
                                                   : make_uint2(0u, 0u);
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
            if (a.trace && warp == (uint32_t)WARPS - 1 && lane == 0 && tile < a.tiles) {
                a.trace[(size_t)tile * 16 + 13] = dbg_rounds;
                a.trace[(size_t)tile * 16 + 14] = dbg_hops;
            }
            if (publisher) st_relaxed_gpu_v2(lb_row + 2 * dt, kLbGlobal | (ex_lo + agg_lo), kLbGlobal | (ex_hi + agg_hi));
        }
        const uint64_t b_lo = a.bases_in[2 * dt], b_hi = a.bases_in[2 * dt + 1];
        s_gbase[2 * dt] = (uint32_t)b_lo + ex_lo + in_lo - dp_lo;
        s_gbase[2 * dt + 1] = (uint32_t)b_hi + ex_hi + in_hi - dp_hi;
        if (a.bases_out != nullptr && tile == a.tiles - 1) {
            a.bases_out[2 * dt] = b_lo + ex_lo + agg_lo;
            a.bases_out[2 * dt + 1] = b_hi + ex_hi + agg_hi;
        }
    }

    if (warp == (uint32_t)WARPS - 1) LSD_TRACE(9);  // look-back done (last warp)

    // ---- 2. two rank chains (even warps, odd warps): the returned half-word is the key's byte offset ----
    uint32_t rk[(ITEMS + 1) / 2];
    if (warp >= 2u) named_bar_sync(warp, 64);
    {
        const uint32_t sel_first = half ? 0x4432u : 0x4410u;   // this warp's half of `old`, zero-extended
        const uint32_t sel_second = half ? 0x7610u : 0x5410u;  // ... into the upper half of rk, keeping the lower
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(mat_bytes + cell_offset<RB, SHIFT>(key[i], lane4)), addc);
            if (i & 1) rk[i >> 1] = __byte_perm(rk[i >> 1], old, sel_second); else rk[i >> 1] = __byte_perm(old, 0u, sel_first);
        }
    }
    if (warp + 2 < (uint32_t)WARPS) named_bar_arrive(warp + 2, 64);
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
    if (warp == 0) LSD_TRACE(7);
    if constexpr (CL > 1) {
        if (!is_lb) {
            cluster_wait();    // #1 (complete long ago unless a peer CTA is far behind)
            cluster_arrive();  // #2
        }
    }
    __syncthreads();
    if (warp == 0) LSD_TRACE(11);

    // ---- 3. stream the reorder buffer out, coalesced per bucket ----
    if (valid == (uint32_t)TILE) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t p = i * THREADS + tid;
            const uint32_t k = s_keys[p];
            out[s_gbase[(k >> SHIFT) & (H - 1)] + p] = k;
        }
    } else {
        for (uint32_t p = tid; p < valid; p += THREADS) {
            const uint32_t k = s_keys[p];
            out[s_gbase[(k >> SHIFT) & (H - 1)] + p] = k;
        }
    }
    if (warp == 0) LSD_TRACE(12);
    if constexpr (CL > 1) cluster_wait();  // #2: no peer still reads this CTA's tile histogram
#undef LSD_TRACE
}

template <int RB, int WARPS, int ITEMS, int MINB, int SHIFT, int LB, int CL, int POLL>
int onesweep_lpc2_launch_shift(const PassArgs& a, cudaStream_t s)
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    auto kern = onesweep_lpc2_kernel<RB, WARPS, ITEMS, MINB, SHIFT, LB, CL, POLL>;
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_::SMEM_BYTES));
    LSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if constexpr (CL == 1) {
        kern<<<a.tiles, S_::THREADS, S_::SMEM_BYTES, s>>>(a);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((a.tiles + CL - 1) / CL * CL);  // whole clusters: CTAs past the last tile hold no keys
        cfg.blockDim = dim3(S_::THREADS);
        cfg.dynamicSmemBytes = S_::SMEM_BYTES;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        LSD_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    }
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CL, int POLL>
int onesweep_lpc2_launch(const PassArgs& a, cudaStream_t s)
{
    static_assert(RB == 8, "shift dispatch below is written for 8-bit digits");
    switch (a.shift) {
        case 0: return onesweep_lpc2_launch_shift<RB, WARPS, ITEMS, MINB, 0, LB, CL, POLL>(a, s);
        case 8: return onesweep_lpc2_launch_shift<RB, WARPS, ITEMS, MINB, 8, LB, CL, POLL>(a, s);
        case 16: return onesweep_lpc2_launch_shift<RB, WARPS, ITEMS, MINB, 16, LB, CL, POLL>(a, s);
        case 24: return onesweep_lpc2_launch_shift<RB, WARPS, ITEMS, MINB, 24, LB, CL, POLL>(a, s);
    }
    return LSD_ERR_INVALID_VALUE;
}

constexpr int kModeLpc2 = 5;

// plain passes on onesweep_lpc2_kernel; peer-scatter and key-value passes on onesweep_lpc32_kernel (same tile)
template <int RB, int WARPS, int ITEMS, int MINB, int LB, int CL, bool WITH_PEER = false, int POLL = 0>
constexpr OnesweepLauncher make_lpc2_launcher()
{
    using S_ = Lpc32Shape<RB, WARPS, ITEMS>;
    if constexpr (WITH_PEER)
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc2, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc2_launch<RB, WARPS, ITEMS, MINB, LB, CL, POLL>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, 4, 0, kPassPeer, false>,
                                &onesweep_lpc32_launch<RB, WARPS, ITEMS, MINB, 4, 0, kPassPairs, false>, nullptr, nullptr};
    else
        return OnesweepLauncher{RB, S_::THREADS, ITEMS, kModeLpc2, (uint32_t)S_::TILE, S_::PORTION_MAX, S_::SMEM_BYTES,
                                &onesweep_lpc2_launch<RB, WARPS, ITEMS, MINB, LB, CL, POLL>, nullptr, nullptr, nullptr, nullptr};
}

}  // namespace lsd
