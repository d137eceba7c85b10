// scan.cu -- single-pass exclusive prefix sum with decoupled look-back (north_star step 2).
//
// Drop-in for the reference's recursive GPUPrefixSum (LSDRadixSort.cu:286-302: BlockPrefixSumKernel
// per level + AddBlockSumsKernel per level, ~3 reads + 2 writes of the array): here every element
// is read once and written once (8 B per element), in place, uint32 wrap-around like the
// reference.
//
// Tile = THREADS x 32 elements (round 1 started with 16: ncu showed the kernel latency-bound on the
// look-back, 65 K tiles starting every ~22 cycles; larger tiles space the tile starts out so that one
// 32-wide window covers the tiles in flight).  Each thread loads eight 128-bit vectors in a vector-striped
// arrangement (vector j of thread t sits at vector index j*THREADS + t of the tile), so every
// warp load is one contiguous 512-byte run.  Tile order is handed out by an atomic ticket, so a
// tile only ever waits on tiles that are already running (forward progress without relying on
// block scheduling order).  Tile state is one 64-bit word {flag:32 | value:32}: flag 1 = tile
// aggregate, flag 2 = inclusive prefix; flag and value travel in one relaxed store, so no fence
// is needed.  Warp 0 looks back 32 predecessors at a time.
#include <cstdlib>

#include "common.cuh"

namespace lsd {

constexpr int kScanItems = 32;  // elements per thread (8 x uint4)
constexpr int kScanVecs = kScanItems / 4;

constexpr uint64_t kFlagAggregate = 1ull << 32;
constexpr uint64_t kFlagInclusive = 2ull << 32;

struct ScanWorkspace {
    uint32_t ticket;
    uint32_t pad[63];  // keep tile states on their own 256-byte line
    uint64_t state[1];  // [tiles]
};

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
scan_kernel(uint32_t* __restrict__ a, uint64_t n, ScanWorkspace* __restrict__ ws)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * kScanItems;
    constexpr int P = kScanVecs * WARPS;   // per-(vector, warp) partial sums, scanned by warp 0
    constexpr int PPL = P / 32;            // partials per lane of warp 0
    static_assert(P % 32 == 0 && PPL >= 1, "partials must fill warp 0 evenly");

    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_partial[kScanVecs * WARPS];  // [vec j][warp] inclusive sums
    __shared__ uint32_t s_tile_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * TILE;
    const uint64_t left = n - base;
    const bool full = left >= (uint64_t)TILE;

    // ---- load (vector-striped) ----
    uint4 v[kScanVecs];
    if (full) {
#pragma unroll
        for (int j = 0; j < kScanVecs; ++j)
            v[j] = __ldcs(reinterpret_cast<const uint4*>(a + base + 4ull * (j * THREADS + tid)));
    } else {
#pragma unroll
        for (int j = 0; j < kScanVecs; ++j) {
            const uint64_t e = 4ull * (j * THREADS + tid);
            v[j].x = e + 0 < left ? a[base + e + 0] : 0u;
            v[j].y = e + 1 < left ? a[base + e + 1] : 0u;
            v[j].z = e + 2 < left ? a[base + e + 2] : 0u;
            v[j].w = e + 3 < left ? a[base + e + 3] : 0u;
        }
    }

    // ---- per-vector sums, warp scans, cross-warp scan in (j, warp) order ----
    uint32_t incl[kScanVecs];
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        incl[j] = warp_inclusive_scan(v[j].x + v[j].y + v[j].z + v[j].w, lane);
        if (lane == 31) s_partial[j * WARPS + warp] = incl[j];
    }
    __syncthreads();

    if (warp == 0) {
        // scan the P partials with one warp, PPL consecutive partials per lane
        uint32_t part[PPL];
        uint32_t lane_sum = 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            part[k] = s_partial[lane * PPL + k];
            lane_sum += part[k];
        }
        const uint32_t lane_incl = warp_inclusive_scan(lane_sum, lane);
        const uint32_t tile_total = __shfl_sync(kFullMask, lane_incl, 31);
        uint32_t run = lane_incl - lane_sum;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            s_partial[lane * PPL + k] = run;  // exclusive prefix of this partial
            run += part[k];
        }

        // ---- decoupled look-back ----
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_gpu(&ws->state[0], kFlagInclusive | tile_total);
        } else {
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagAggregate | tile_total);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - (int64_t)lane;
                // virtual tiles before tile 0 read as "inclusive prefix 0"
                const uint64_t w = idx >= 0 ? ld_relaxed_gpu(&ws->state[idx]) : kFlagInclusive;
                const uint32_t ready = __ballot_sync(kFullMask, (w >> 32) != 0);
                const uint32_t incl_mask = __ballot_sync(kFullMask, (w >> 32) == 2);
                uint32_t take = (uint32_t)w;
                bool finished = false;
                if (incl_mask) {
                    // only the predecessors up to the nearest INCLUSIVE one have to be ready
                    const uint32_t first = __ffs(incl_mask) - 1;
                    const uint32_t need = (2u << first) - 1u;
                    if ((ready & need) != need) continue;  // poll again
                    if (lane > first) take = 0;
                    finished = true;
                } else if (ready != kFullMask) {
                    continue;  // window not complete yet and no inclusive word in sight: poll again
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(kFullMask, take, o);
                exclusive += take;
                if (finished) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagInclusive | (uint32_t)(exclusive + tile_total));
        }
        if (lane == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();

    // ---- exclusive results and store ----
    const uint32_t tile_prefix = s_tile_prefix;
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        const uint32_t sum = v[j].x + v[j].y + v[j].z + v[j].w;
        uint32_t run = tile_prefix + s_partial[j * WARPS + warp] + (incl[j] - sum);
        uint4 o;
        o.x = run; run += v[j].x;
        o.y = run; run += v[j].y;
        o.z = run; run += v[j].z;
        o.w = run;
        const uint64_t e = 4ull * (j * THREADS + tid);
        if (full) {
            __stcs(reinterpret_cast<uint4*>(a + base + e), o);
        } else {
            if (e + 0 < left) a[base + e + 0] = o.x;
            if (e + 1 < left) a[base + e + 1] = o.y;
            if (e + 2 < left) a[base + e + 2] = o.z;
            if (e + 3 < left) a[base + e + 3] = o.w;
        }
    }
}

// -------------------------------------------------------------------------------------
// scan_tma_kernel: the same single-pass scan, tile staged in shared memory by one TMA bulk copy.
//
// Why (ncu, profiles/r01_scan_ncu.txt): scan_kernel keeps its 32 elements per thread in registers while
// warp 0 walks the look-back chain -- 64 registers x 256 threads -> 4 CTAs per SM, 53 % of all stall samples on
// the barrier behind the look-back, 3.5 TB/s.  Here the tile lands in shared memory (16 KiB per 128-thread CTA),
// the threads only keep 8 running sums across the wait (~40 registers), and re-read their vectors from shared
// memory once the tile prefix is known: 12-13 CTAs per SM are in flight, which is what the HBM latency x bandwidth
// product of B200 needs.  Round 2: every CTA also pulls the tile that will be drawn 12 MiB of tickets later into L2
// (cp.async.bulk.prefetch.L2), so a tile's own load is an L2 round trip instead of a DRAM access under full load (the
// trace had the tile land 6.6 K cycles after the ticket): 0.505 -> 0.427 ms at 2^28 = 5.03 TB/s, 0.77 of the copy roofline.
// -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t scan_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 lds128_volatile(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(scan_smem_u32(p)));
    return v;
}

// TRACE: phase clocks per tile (LSD_SCAN_TRACE=1); compile-time so that the shipped kernel carries no run-time checks
template <int THREADS, bool TRACE = false>
__global__ void __launch_bounds__(THREADS)
scan_tma_kernel(uint32_t* __restrict__ a, uint64_t n, ScanWorkspace* __restrict__ ws, uint32_t* __restrict__ trace,
                uint32_t prefetch_tiles)
{
    const long long t_start = (TRACE && trace) ? clock64() : 0;
#define SCAN_TRACE(slot) do { if constexpr (TRACE) if (trace && tid == 0) trace[(size_t)tile * 8 + (slot)] = (uint32_t)(clock64() - t_start); } while (0)
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * kScanItems;
    constexpr int P = kScanVecs * WARPS;
    constexpr int PPL = P / 32;
    static_assert(P % 32 == 0 && PPL >= 1, "partials must fill warp 0 evenly");

    extern __shared__ __align__(128) uint32_t s_data[];  // [TILE]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_partial[kScanVecs * WARPS];
    __shared__ uint32_t s_tile_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(scan_smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t t = atomicAdd(&ws->ticket, 1u);
        s_tile = t;
        if constexpr (TRACE)
            if (trace) trace[(size_t)t * 8 + 0] = (uint32_t)(clock64() - t_start);
        const uint64_t b = (uint64_t)t * TILE;
        // pull the tile that will be drawn `prefetch_tiles` tickets from now into L2: a tile's own load then costs an L2
        // round trip instead of a DRAM access under full load (the trace had the tile land 6.6 K cycles after the ticket)
        if (prefetch_tiles != 0u) {
            const uint64_t pf = ((uint64_t)t + prefetch_tiles) * TILE;
            if (pf < n && n - pf >= (uint64_t)TILE)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a + pf), "r"(TILE * 4) : "memory");
        }
        if (n - b >= (uint64_t)TILE) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(scan_smem_u32(&s_bar)), "r"(TILE * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(scan_smem_u32(s_data)), "l"(a + b), "r"(TILE * 4), "r"(scan_smem_u32(&s_bar)) : "memory");
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * TILE;
    const uint64_t left = n - base;
    const bool full = left >= (uint64_t)TILE;
    if (full) {
        asm volatile(
            "{\n.reg .pred p;\nSCAN_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra SCAN_DONE_%=;\nbra SCAN_WAIT_%=;\nSCAN_DONE_%=:\n}\n" ::"r"(scan_smem_u32(&s_bar)) : "memory");
    } else {
        for (uint32_t i = tid; i < (uint32_t)TILE; i += THREADS) s_data[i] = i < left ? a[base + i] : 0u;
        __syncthreads();
    }

    SCAN_TRACE(1);  // tile landed
    // ---- per-vector sums, warp scans (only the 8 running sums stay in registers) ----
    uint32_t excl[kScanVecs];
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        const uint4 v = lds128_volatile(s_data + 4 * (j * THREADS + tid));
        const uint32_t sum = v.x + v.y + v.z + v.w;
        const uint32_t incl = warp_inclusive_scan(sum, lane);
        excl[j] = incl - sum;
        if (lane == 31) s_partial[j * WARPS + warp] = incl;
    }
    __syncthreads();

    SCAN_TRACE(2);  // partials done
    if (warp == 0) {
        uint32_t part[PPL];
        uint32_t lane_sum = 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            part[k] = s_partial[lane * PPL + k];
            lane_sum += part[k];
        }
        const uint32_t lane_incl = warp_inclusive_scan(lane_sum, lane);
        const uint32_t tile_total = __shfl_sync(kFullMask, lane_incl, 31);
        uint32_t run = lane_incl - lane_sum;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            s_partial[lane * PPL + k] = run;
            run += part[k];
        }
        // ---- decoupled look-back, 32 predecessors per round trip ----
        // Tried and measured slower on B200 (profiles/r01_scan_variants.txt): 64/128/256-wide windows (more strong loads
        // per round), exponential back-off on failed polls, and span records that let walkers skip what a predecessor
        // has already summed.  With ~900 tiles in flight the walk is ~18 rounds; what helped was taking the tile out of
        // the registers (this kernel) so that more tiles are in flight per SM.
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_gpu(&ws->state[0], kFlagInclusive | tile_total);
        } else {
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagAggregate | tile_total);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - (int64_t)lane;
                const uint64_t w = idx >= 0 ? ld_relaxed_gpu(&ws->state[idx]) : kFlagInclusive;  // virtual tiles: prefix 0
                const uint32_t ready = __ballot_sync(kFullMask, (w >> 32) != 0);
                const uint32_t incl_mask = __ballot_sync(kFullMask, (w >> 32) == 2);
                uint32_t take = (uint32_t)w;
                bool finished = false;
                if (incl_mask) {
                    const uint32_t first = __ffs(incl_mask) - 1;
                    const uint32_t need = (2u << first) - 1u;
                    if ((ready & need) != need) continue;
                    if (lane > first) take = 0;
                    finished = true;
                } else if (ready != kFullMask) {
                    continue;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(kFullMask, take, o);
                exclusive += take;
                if (finished) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagInclusive | (uint32_t)(exclusive + tile_total));
        }
        if (lane == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();

    SCAN_TRACE(3);  // look-back done
    // ---- re-read the vectors from shared memory, exclusive results, streaming stores ----
    const uint32_t tile_prefix = s_tile_prefix;
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        const uint4 v = lds128_volatile(s_data + 4 * (j * THREADS + tid));
        uint32_t run = tile_prefix + s_partial[j * WARPS + warp] + excl[j];
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run;
        const uint64_t e = 4ull * (j * THREADS + tid);
        if (full) {
            __stcs(reinterpret_cast<uint4*>(a + base + e), o);
        } else {
            if (e + 0 < left) a[base + e + 0] = o.x;
            if (e + 1 < left) a[base + e + 1] = o.y;
            if (e + 2 < left) a[base + e + 2] = o.z;
            if (e + 3 < left) a[base + e + 3] = o.w;
        }
    }
    SCAN_TRACE(4);  // stores issued
#undef SCAN_TRACE
}

// -------------------------------------------------------------------------------------
// scan_l2_kernel: reduce-then-scan per chunk, the second read served by the 126 MB L2.
//
// Why (profiles/r01_scan_variants.txt): scan_tma_kernel is bound by the shared memory its WAITING tiles occupy -- a tile
// lives ~26 K cycles, 13 K of them in the look-back walk, with its 32 KiB parked in shared memory, so ~33 MB are in flight
// and throughput = bytes in flight / tile life.  Here a CTA reads its chunk (THREADS x 4*VECS elements, 64 KiB) once to
// get the chunk total and its threads' running sums, publishes the total, walks the look-back chain holding NOTHING but
// VECS registers per thread, then reads the chunk again -- from L2, where the first read left it a few microseconds
// earlier -- and writes the results.  DRAM traffic stays 8 B per element as long as L2 keeps the chunks in flight
// (~6 CTAs per SM x 64 KiB = 57 MB).
// -------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_l2_v4(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int THREADS, int VECS>
__global__ void __launch_bounds__(THREADS)
scan_l2_kernel(uint32_t* __restrict__ a, uint64_t n, ScanWorkspace* __restrict__ ws)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int CHUNK = THREADS * VECS * 4;
    constexpr int P = VECS * WARPS;
    constexpr int PPL = P / 32;
    static_assert(P % 32 == 0 && PPL >= 1, "partials must fill warp 0 evenly");

    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_partial[P];
    __shared__ uint32_t s_tile_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * CHUNK;
    const uint64_t left = n - base;
    const bool full = left >= (uint64_t)CHUNK;
    const uint4* src = reinterpret_cast<const uint4*>(a + base);

    auto load_vec = [&](int j) -> uint4 {
        const uint32_t idx = (uint32_t)j * THREADS + tid;
        if (full) return ld_l2_v4(src + idx);
        const uint64_t e = 4ull * idx;
        uint4 v;
        v.x = e + 0 < left ? a[base + e + 0] : 0u;
        v.y = e + 1 < left ? a[base + e + 1] : 0u;
        v.z = e + 2 < left ? a[base + e + 2] : 0u;
        v.w = e + 3 < left ? a[base + e + 3] : 0u;
        return v;
    };

    // ---- first read: per-vector sums, warp scans ----
    uint32_t excl[VECS];
#pragma unroll
    for (int j = 0; j < VECS; ++j) {
        const uint4 v = load_vec(j);
        const uint32_t sum = v.x + v.y + v.z + v.w;
        const uint32_t incl = warp_inclusive_scan(sum, lane);
        excl[j] = incl - sum;
        if (lane == 31) s_partial[j * WARPS + warp] = incl;
    }
    __syncthreads();

    if (warp == 0) {
        uint32_t part[PPL];
        uint32_t lane_sum = 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            part[k] = s_partial[lane * PPL + k];
            lane_sum += part[k];
        }
        const uint32_t lane_incl = warp_inclusive_scan(lane_sum, lane);
        const uint32_t tile_total = __shfl_sync(kFullMask, lane_incl, 31);
        uint32_t run = lane_incl - lane_sum;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            s_partial[lane * PPL + k] = run;
            run += part[k];
        }
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_gpu(&ws->state[0], kFlagInclusive | tile_total);
        } else {
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagAggregate | tile_total);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - (int64_t)lane;
                const uint64_t w = idx >= 0 ? ld_relaxed_gpu(&ws->state[idx]) : kFlagInclusive;  // virtual chunks: prefix 0
                const uint32_t ready = __ballot_sync(kFullMask, (w >> 32) != 0);
                const uint32_t incl_mask = __ballot_sync(kFullMask, (w >> 32) == 2);
                uint32_t take = (uint32_t)w;
                bool finished = false;
                if (incl_mask) {
                    const uint32_t first = __ffs(incl_mask) - 1;
                    const uint32_t need = (2u << first) - 1u;
                    if ((ready & need) != need) continue;
                    if (lane > first) take = 0;
                    finished = true;
                } else if (ready != kFullMask) {
                    continue;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(kFullMask, take, o);
                exclusive += take;
                if (finished) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_gpu(&ws->state[tile], kFlagInclusive | (uint32_t)(exclusive + tile_total));
        }
        if (lane == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();

    // ---- second read (L2), exclusive results, streaming stores ----
    const uint32_t tile_prefix = s_tile_prefix;
#pragma unroll
    for (int j = 0; j < VECS; ++j) {
        const uint4 v = load_vec(j);
        uint32_t run = tile_prefix + s_partial[j * WARPS + warp] + excl[j];
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run;
        const uint64_t e = 4ull * ((uint32_t)j * THREADS + tid);
        if (full) {
            __stcs(reinterpret_cast<uint4*>(a + base + e), o);
        } else {
            if (e + 0 < left) a[base + e + 0] = o.x;
            if (e + 1 < left) a[base + e + 1] = o.y;
            if (e + 2 < left) a[base + e + 2] = o.z;
            if (e + 3 < left) a[base + e + 3] = o.w;
        }
    }
}

// -------------------------------------------------------------------------------------
// scan_cluster_kernel: scan_tma_kernel with thread-block clusters of 8 CTAs sharing ONE look-back record.
//
// Why (profiles/r01_scan_variants.txt): with ~900 tiles in flight every tile's look-back walks ~18 windows of 32
// predecessors at ~1 K cycles per round trip -- 18 K of the tile's 31 K-cycle life.  A cluster of 8 CTAs takes 8
// consecutive tiles; the tile totals are exchanged through distributed shared memory (st.shared::cluster, ~200 cycles),
// the cluster leader publishes one aggregate and looks back over CLUSTER records (8x fewer in flight: ~4 windows),
// and hands the cluster's exclusive prefix to its 7 peers through distributed shared memory again.  Three hardware
// cluster barriers replace ~14 global round trips.
// -------------------------------------------------------------------------------------
constexpr int kScanCluster = 8;

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    __syncwarp();  // the .aligned forms need the warp converged
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u32(const void* local_smem, uint32_t rank, uint32_t v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(scan_smem_u32(local_smem)), "r"(rank));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

template <int THREADS>
__global__ void __cluster_dims__(kScanCluster, 1, 1) __launch_bounds__(THREADS)
scan_cluster_kernel(uint32_t* __restrict__ a, uint64_t n, ScanWorkspace* __restrict__ ws, uint32_t tiles, uint32_t* __restrict__ trace)
{
    const long long t_start = trace ? clock64() : 0;
#define SCANC_TRACE(slot) do { if (trace && tid == 0 && active) trace[(size_t)tile * 8 + (slot)] = (uint32_t)(clock64() - t_start); } while (0)
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * kScanItems;
    constexpr int P = kScanVecs * WARPS;
    constexpr int PPL = P / 32;
    static_assert(P % 32 == 0 && PPL >= 1, "partials must fill warp 0 evenly");

    extern __shared__ __align__(128) uint32_t s_data[];  // [TILE]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_ticket;                 // cluster ticket, written by the leader into every CTA
    __shared__ uint32_t s_totals[kScanCluster];   // tile totals of the cluster, written by every CTA into every CTA
    __shared__ uint32_t s_cluster_prefix;         // exclusive prefix of the cluster, written by the leader into every CTA
    __shared__ uint32_t s_partial[kScanVecs * WARPS];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(scan_smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t my_ticket = 0;
    if (rank == 0 && tid == 0) my_ticket = atomicAdd(&ws->ticket, 1u);  // in flight across barrier #0
    cluster_sync_all();  // #0: every CTA of the cluster has started: its shared memory may be written remotely from here on
    if (rank == 0 && tid < (uint32_t)kScanCluster) {
        const uint32_t t = __shfl_sync((1u << kScanCluster) - 1u, my_ticket, 0);
        st_cluster_u32(&s_ticket, tid, t);
    }
    cluster_sync_all();  // #1: every CTA knows the cluster's ticket (and its mbarrier is initialised)
    const uint32_t cluster = s_ticket;
    const uint32_t tile = cluster * kScanCluster + rank;
    const bool active = tile < tiles;
    const uint64_t base = (uint64_t)tile * TILE;
    const uint64_t left = active ? n - base : 0;
    const bool full = left >= (uint64_t)TILE;
    SCANC_TRACE(0);  // ticket known (cluster barrier #1 passed)
    if (full && tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(scan_smem_u32(&s_bar)), "r"(TILE * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(scan_smem_u32(s_data)), "l"(a + base), "r"(TILE * 4), "r"(scan_smem_u32(&s_bar)) : "memory");
    }
    if (full) {
        asm volatile(
            "{\n.reg .pred p;\nSCANC_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra SCANC_DONE_%=;\nbra SCANC_WAIT_%=;\nSCANC_DONE_%=:\n}\n" ::"r"(scan_smem_u32(&s_bar)) : "memory");
    } else {
        for (uint32_t i = tid; i < (uint32_t)TILE; i += THREADS) s_data[i] = i < left ? a[base + i] : 0u;
        __syncthreads();
    }

    SCANC_TRACE(1);  // tile landed
    uint32_t excl[kScanVecs];
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        const uint4 v = lds128_volatile(s_data + 4 * (j * THREADS + tid));
        const uint32_t sum = v.x + v.y + v.z + v.w;
        const uint32_t incl = warp_inclusive_scan(sum, lane);
        excl[j] = incl - sum;
        if (lane == 31) s_partial[j * WARPS + warp] = incl;
    }
    __syncthreads();

    if (warp == 0) {
        uint32_t part[PPL];
        uint32_t lane_sum = 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            part[k] = s_partial[lane * PPL + k];
            lane_sum += part[k];
        }
        const uint32_t lane_incl = warp_inclusive_scan(lane_sum, lane);
        const uint32_t tile_total = __shfl_sync(kFullMask, lane_incl, 31);
        uint32_t run = lane_incl - lane_sum;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
            s_partial[lane * PPL + k] = run;
            run += part[k];
        }
        if (lane < (uint32_t)kScanCluster) st_cluster_u32(&s_totals[rank], lane, tile_total);  // my total into every CTA
    }
    SCANC_TRACE(2);  // partials done, total sent
    cluster_sync_all();  // #2: all tile totals of the cluster are in every CTA's shared memory

    if (rank == 0 && warp == 0) {
        uint32_t cluster_total = 0;
#pragma unroll
        for (int r = 0; r < kScanCluster; ++r) cluster_total += s_totals[r];
        uint32_t exclusive = 0;
        if (cluster == 0) {
            if (lane == 0) st_relaxed_gpu(&ws->state[0], kFlagInclusive | cluster_total);
        } else {
            if (lane == 0) st_relaxed_gpu(&ws->state[cluster], kFlagAggregate | cluster_total);
            int64_t look = (int64_t)cluster - 1;
            while (true) {
                const int64_t idx = look - (int64_t)lane;
                const uint64_t w = idx >= 0 ? ld_relaxed_gpu(&ws->state[idx]) : kFlagInclusive;  // virtual records: prefix 0
                const uint32_t ready = __ballot_sync(kFullMask, (w >> 32) != 0);
                const uint32_t incl_mask = __ballot_sync(kFullMask, (w >> 32) == 2);
                uint32_t take = (uint32_t)w;
                bool finished = false;
                if (incl_mask) {
                    const uint32_t first = __ffs(incl_mask) - 1;
                    const uint32_t need = (2u << first) - 1u;
                    if ((ready & need) != need) continue;
                    if (lane > first) take = 0;
                    finished = true;
                } else if (ready != kFullMask) {
                    continue;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(kFullMask, take, o);
                exclusive += take;
                if (finished) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_gpu(&ws->state[cluster], kFlagInclusive | (uint32_t)(exclusive + cluster_total));
        }
        if (lane < (uint32_t)kScanCluster) st_cluster_u32(&s_cluster_prefix, lane, exclusive);
    }
    cluster_sync_all();  // #3: the cluster's exclusive prefix is in every CTA; no remote access happens after this point

    SCANC_TRACE(3);  // cluster prefix known (barrier #3 passed)
    uint32_t tile_prefix = s_cluster_prefix;
    for (uint32_t r = 0; r < rank; ++r) tile_prefix += s_totals[r];
    if (!active) return;
#pragma unroll
    for (int j = 0; j < kScanVecs; ++j) {
        const uint4 v = lds128_volatile(s_data + 4 * (j * THREADS + tid));
        uint32_t run = tile_prefix + s_partial[j * WARPS + warp] + excl[j];
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run;
        const uint64_t e = 4ull * (j * THREADS + tid);
        if (full) {
            __stcs(reinterpret_cast<uint4*>(a + base + e), o);
        } else {
            if (e + 0 < left) a[base + e + 0] = o.x;
            if (e + 1 < left) a[base + e + 1] = o.y;
            if (e + 2 < left) a[base + e + 2] = o.z;
            if (e + 3 < left) a[base + e + 3] = o.w;
        }
    }
    SCANC_TRACE(4);  // stores issued
#undef SCANC_TRACE
}

// LSD_SCAN_CLUSTER=1 selects scan_cluster_kernel (one look-back record per cluster of 8 tiles).  Measured on B200 it does not
// beat scan_tma_kernel (0.528 vs 0.512 ms at 2^28): the two cluster barriers at the start (all 8 CTAs must have started before
// their shared memory may be written) and the wait for the slowest of 8 tile loads cost what the shorter look-back saves.
// The product library reads no environment: the switches exist in the tuning build (make TUNING=1) only.
#ifdef LSD_TUNING_VARIANTS
static bool scan_env(const char* name) { const char* e = getenv(name); return e && e[0] == '1'; }
#else
static constexpr bool scan_env(const char*) { return false; }
#endif
static const bool g_scan_no_cluster = !scan_env("LSD_SCAN_CLUSTER");
static const bool g_scan_register_path = scan_env("LSD_SCAN_REGISTER_PATH");

static int scan_threads_for(int block)
{
    if (block <= 0) return 256;
    if (block <= 128) return 128;
    if (block <= 256) return 256;
    return 512;
}

// LSD_SCAN_TRACE=1 (tuning aid): 8 uint32 phase clocks per tile are written after the tile states
static const bool g_scan_l2 = scan_env("LSD_SCAN_L2");
// L2 prefetch distance of scan_tma_kernel in BYTES (0 = off).  Measured at 2^28 (profiles/r02_scan_l2_prefetch.jsonl, first
// sweep in tiles, second in MiB): off 0.505-0.517 ms; 6 / 8 / 12 / 16 / 20 / 24 MiB ahead 0.437 / 0.431 / 0.427 / 0.433 / 0.445 /
// 0.456 ms; 32 MiB 0.50; 64 MiB 0.59 (the lines are evicted before their tile comes up).  12 MiB is at or next to the best
// distance for all three tile sizes (block 128 / 256 / 512) and at 2^30 (1.670 ms = 5.14 TB/s against 2.010 ms).
#ifdef LSD_TUNING_VARIANTS
static uint32_t g_scan_prefetch_bytes = 12u << 20;
extern "C" __attribute__((visibility("default"))) void lsd_debug_scan_prefetch(int bytes) { g_scan_prefetch_bytes = (uint32_t)bytes; }
#else
static constexpr uint32_t g_scan_prefetch_bytes = 12u << 20;
#endif
static const bool g_scan_trace = scan_env("LSD_SCAN_TRACE");

size_t scan_workspace_bytes(uint64_t n, int block)
{
    const uint64_t tile = (uint64_t)scan_threads_for(block) * kScanItems;
    const uint64_t tiles = (n + tile - 1) / tile;
    return sizeof(ScanWorkspace) + (size_t)(tiles ? tiles : 1) * sizeof(uint64_t) * (g_scan_trace ? 5 : 1);
}

int launch_prefix_sum(uint32_t* a, uint64_t n, int block, void* ws, size_t ws_bytes, cudaStream_t s)
{
    if (n == 0) return LSD_OK;
    const size_t need = scan_workspace_bytes(n, block);
    if (ws_bytes < need) return LSD_ERR_WORKSPACE_TOO_SMALL;
    const int threads = scan_threads_for(block);
    const uint64_t tile = (uint64_t)threads * kScanItems;
    const uint64_t tiles = (n + tile - 1) / tile;
    if (tiles > 0x7FFFFFFFull) return LSD_ERR_UNSUPPORTED;
    LSD_CUDA_TRY(cudaMemsetAsync(ws, 0, need, s));
    auto* w = static_cast<ScanWorkspace*>(ws);
    if (g_scan_l2 && aligned_to(a, 16)) {
        // 256 threads x 64 elements = 64 KiB chunks: fewer records than the workspace was sized for
        constexpr int kL2Threads = 256, kL2Vecs = 16;
        const uint64_t chunk = (uint64_t)kL2Threads * kL2Vecs * 4;
        const uint64_t chunks = (n + chunk - 1) / chunk;
        scan_l2_kernel<kL2Threads, kL2Vecs><<<(unsigned)chunks, kL2Threads, 0, s>>>(a, n, w);
        LSD_LAUNCH_CHECK();
        return LSD_OK;
    }
    const bool tma = !g_scan_register_path && aligned_to(a, 16);
    uint32_t* trace_c = g_scan_trace ? reinterpret_cast<uint32_t*>(static_cast<ScanWorkspace*>(ws)->state + tiles) : nullptr;
    if (tma && !g_scan_no_cluster) {
        const unsigned grid = (unsigned)((tiles + kScanCluster - 1) / kScanCluster * kScanCluster);
        const size_t smem_c = (size_t)threads * kScanItems * sizeof(uint32_t);
        switch (threads) {
            case 128:
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_cluster_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                scan_cluster_kernel<128><<<grid, 128, smem_c, s>>>(a, n, w, (uint32_t)tiles, trace_c);
                break;
            case 256:
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_cluster_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                scan_cluster_kernel<256><<<grid, 256, smem_c, s>>>(a, n, w, (uint32_t)tiles, trace_c);
                break;
            default:
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_cluster_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                scan_cluster_kernel<512><<<grid, 512, smem_c, s>>>(a, n, w, (uint32_t)tiles, trace_c);
                break;
        }
        LSD_LAUNCH_CHECK();
        return LSD_OK;
    }
    uint32_t* trace = g_scan_trace ? reinterpret_cast<uint32_t*>(w->state + tiles) : nullptr;
    const size_t smem = (size_t)threads * kScanItems * sizeof(uint32_t);
    if (tma) {
        switch (threads) {
            case 128:
                if (trace) {
                    LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    scan_tma_kernel<128, true><<<(unsigned)tiles, 128, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                    break;
                }
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                scan_tma_kernel<128><<<(unsigned)tiles, 128, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                break;
            case 256:
                if (trace) {
                    LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    scan_tma_kernel<256, true><<<(unsigned)tiles, 256, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                    break;
                }
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                scan_tma_kernel<256><<<(unsigned)tiles, 256, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                break;
            default:
                if (trace) {
                    LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    scan_tma_kernel<512, true><<<(unsigned)tiles, 512, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                    break;
                }
                LSD_CUDA_TRY(cudaFuncSetAttribute(scan_tma_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                scan_tma_kernel<512><<<(unsigned)tiles, 512, smem, s>>>(a, n, w, trace, g_scan_prefetch_bytes / (uint32_t)(tile * sizeof(uint32_t)));
                break;
        }
    } else {
        switch (threads) {
            case 128: scan_kernel<128><<<(unsigned)tiles, 128, 0, s>>>(a, n, w); break;
            case 256: scan_kernel<256><<<(unsigned)tiles, 256, 0, s>>>(a, n, w); break;
            default: scan_kernel<512><<<(unsigned)tiles, 512, 0, s>>>(a, n, w); break;
        }
    }
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

}  // namespace lsd
