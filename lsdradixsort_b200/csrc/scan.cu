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

static int scan_threads_for(int block)
{
    if (block <= 0) return 256;
    if (block <= 128) return 128;
    if (block <= 256) return 256;
    return 512;
}

size_t scan_workspace_bytes(uint64_t n, int block)
{
    const uint64_t tile = (uint64_t)scan_threads_for(block) * kScanItems;
    const uint64_t tiles = (n + tile - 1) / tile;
    return sizeof(ScanWorkspace) + (size_t)(tiles ? tiles : 1) * sizeof(uint64_t);
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
    switch (threads) {
        case 128: scan_kernel<128><<<(unsigned)tiles, 128, 0, s>>>(a, n, w); break;
        case 256: scan_kernel<256><<<(unsigned)tiles, 256, 0, s>>>(a, n, w); break;
        default: scan_kernel<512><<<(unsigned)tiles, 512, 0, s>>>(a, n, w); break;
    }
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

}  // namespace lsd
