// skeleton_bench.cu -- the memory skeleton of a digit pass without any ranking: every CTA TMA-loads one 8192-key tile
// and writes it back as 256 runs of 32 keys to 256 bucket streams (run = 128 bytes, optionally misaligned like real
// bucket offsets).  Question: what does the B200 memory system allow for this access pattern, independent of the SM work?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/skeleton_bench bench_tools/skeleton_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int TILE = 8192, THREADS = 256, RUN = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Ragged but ABUTTING runs (like a real pass): inside tile t bucket b starts at s(t,b) = 32 b + H(t+1,b) - H(t,b); in stream b
// the run of tile t starts at 32 t + [H(t,b+1) - H(0,b+1)] - [H(t,b) - H(0,b)], so consecutive tiles' runs touch exactly.
__device__ __forceinline__ int Hh(uint32_t t, uint32_t b)
{
    if (b == 0u || b >= 256u) return 0;
    uint32_t x = t * 0x9E3779B1u ^ b * 0x85EBCA77u;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    return (int)(x & 7u);
}
__device__ __forceinline__ uint32_t tile_start(uint32_t t, uint32_t b)
{
    if (b == 0u) return 0u;
    if (b >= 256u) return (uint32_t)TILE;
    return (uint32_t)((int)(b * RUN) + Hh(t + 1, b) - Hh(t, b));
}
__device__ __forceinline__ size_t stream_pos(uint32_t t, uint32_t b)
{
    return (size_t)((long long)t * RUN + (Hh(t, b + 1) - Hh(0, b + 1)) - (Hh(t, b) - Hh(0, b)) + 16);
}

template <int MODE>  // 0: full-line copy; 1: 256 aligned runs; 2: misaligned runs; 3: runs misaligned by whole 32 B sectors;
                     // 4: misaligned runs of ~32+-8 keys written by p-linear warps (a warp store spans two runs, like the pass kernels);
                     // 5: same runs, one warp store per (run, destination line): line-aligned, partially filled warps
__global__ void __launch_bounds__(THREADS) skeleton(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t tiles,
                                                   uint32_t stream_len)
{
    extern __shared__ __align__(128) uint32_t s[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tid = threadIdx.x, tile = blockIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(TILE * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s)), "l"(in + (size_t)tile * TILE), "r"(TILE * 4), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    if (MODE >= 6) {
        // 6: one thread per bucket: the 16-byte aligned middle of its run leaves by ONE bulk (TMA) store shared -> global (no
        //    LSU work), the <= 3 words before and after it by scalar stores; 7: bulk stores only; 8: heads / tails only.
        //    (Timing only: the source is taken at the 16-byte boundary next to the run, as if the reorder buffer had been laid
        //    out with every run at its destination's alignment.)
        const uint32_t bucket = tid;
        const uint32_t s0 = tile_start(tile, bucket), s1 = tile_start(tile, bucket + 1);
        const uint32_t mis = (bucket * 2654435761u >> 27);
        uint32_t* dst = out + (size_t)bucket * stream_len + stream_pos(tile, bucket) + mis;
        const uint32_t len = s1 - s0;
        uint32_t head = (0u - (uint32_t)((unsigned long long)dst >> 2)) & 3u;
        if (head > len) head = len;
        const uint32_t mid = (len - head) & ~3u, tail = len - head - mid;
        const uint32_t src = (s0 + head) & ~3u;
        if (MODE != 8 && mid != 0u) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head), "r"(smem_u32(s + src)), "r"(mid * 4u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (MODE != 7) {
#pragma unroll
            for (uint32_t j = 0; j < 3; ++j) {
                if (j < head) dst[j] = s[s0 + j];
                if (j < tail) dst[head + mid + j] = s[s0 + head + mid + j];
            }
        }
        if (MODE != 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        return;
    }
    if (MODE == 5) {
        // bucket-major copy-out: per-bucket info {start, length, destination} precomputed in shared memory (the pass kernel
        // has it there anyway); one warp store per (run, destination line): lanes = position inside the 128-byte line, so
        // a run is only ever split at line boundaries and has exactly two partial sectors (its two ends)
        __shared__ uint4 s_info[256];
        {
            const uint32_t bucket = tid;  // THREADS == 256
            const uint32_t s0 = tile_start(tile, bucket), s1 = tile_start(tile, bucket + 1);
            const uint32_t mis = (bucket * 2654435761u >> 27);
            const unsigned long long dst = (unsigned long long)(out + (size_t)bucket * stream_len + stream_pos(tile, bucket) + mis);
            s_info[bucket] = make_uint4(s0, s1 - s0, (uint32_t)dst, (uint32_t)(dst >> 32));
        }
        __syncthreads();
        const uint32_t warp = tid >> 5, lane = tid & 31u;
#pragma unroll 4
        for (uint32_t bucket = warp; bucket < 256u; bucket += THREADS / 32) {
            const uint4 inf = s_info[bucket];
            uint32_t* dst = reinterpret_cast<uint32_t*>((unsigned long long)inf.z | ((unsigned long long)inf.w << 32));
            const uint32_t a0 = (inf.z >> 2) & 31u;  // position of dst[0] inside its line
            const uint32_t q0 = lane - a0;           // index inside the run handled by this lane in the first store
            if (q0 < inf.y) dst[q0] = s[inf.x + q0];                      // (unsigned compare: q0 "negative" wraps to huge)
            const uint32_t q1 = q0 + 32;
            if (q1 < inf.y) dst[q1] = s[inf.x + q1];
            const uint32_t q2 = q0 + 64;
            if (q2 < inf.y) dst[q2] = s[inf.x + q2];
        }
        return;
    }
    __shared__ size_t s_base[256];
    if (MODE == 4) {
        const uint32_t bucket = tid;
        const uint32_t mis = (bucket * 2654435761u >> 27);
        s_base[bucket] = (size_t)bucket * stream_len + stream_pos(tile, bucket) + mis - tile_start(tile, bucket);
        __syncthreads();
    }
#pragma unroll 8
    for (int i = 0; i < TILE / THREADS; ++i) {
        const uint32_t p = i * THREADS + tid;
        const uint32_t k = s[p];
        if (MODE == 0) {
            out[(size_t)tile * TILE + p] = k;
        } else if (MODE == 4) {
            // the pass kernels' copy-out: key -> digit -> bucket base from shared memory -> out[base + p]
            const uint32_t bucket = k & 255u;  // the input was initialised so that position p of tile t carries its bucket
            out[s_base[bucket] + p] = k;
        } else if (MODE == 5) {
            // handled below (different loop shape)
        } else if (MODE == 3) {
            const uint32_t bucket = p / RUN, r = p % RUN;
            const uint32_t mis = (bucket * 2654435761u >> 30) * 8u;  // 0, 8, 16 or 24 keys
            out[(size_t)bucket * stream_len + (size_t)tile * RUN + r + mis] = k;
        } else {
            const uint32_t bucket = p / RUN, r = p % RUN;
            uint32_t mis = 0;
            if (MODE == 2) mis = (bucket * 2654435761u >> 27);  // fixed per bucket: every tile continues its stream
            out[(size_t)bucket * stream_len + (size_t)tile * RUN + r + mis] = k;
        }
    }
}

__global__ void init_buckets(uint32_t* in, uint32_t tiles)
{
    const uint32_t tile = blockIdx.x;
    for (uint32_t p = threadIdx.x; p < (uint32_t)TILE; p += blockDim.x) {
        uint32_t bucket = p / RUN;
        if (p < tile_start(tile, bucket)) bucket -= 1;
        else if (p >= tile_start(tile, bucket + 1)) bucket += 1;
        in[(size_t)tile * TILE + p] = bucket | (p << 8);
    }
}

template <int MODE>
float run(const uint32_t* in, uint32_t* out, uint32_t tiles, size_t smem)
{
    CK(cudaFuncSetAttribute(skeleton<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a));
        skeleton<MODE><<<tiles, THREADS, smem>>>(in, out, tiles, tiles * (RUN + 8) + 64);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t n = (size_t)1 << 28;
    const uint32_t tiles = n / TILE;
    uint32_t *in, *out;
    CK(cudaMalloc(&in, n * 4));
    CK(cudaMalloc(&out, (n + n / 4 + 256 * 128) * 4));
    init_buckets<<<tiles, 256>>>(in, tiles);
    CK(cudaDeviceSynchronize());
    printf("memory skeleton of a digit pass, 2^28 keys, tile 8192, 256 threads; ms and GB/s (8 B/key)\n");
    for (size_t smem : {(size_t)TILE * 4, (size_t)70 * 1024}) {
        const float t0 = run<0>(in, out, tiles, smem), t1 = run<1>(in, out, tiles, smem), t2 = run<2>(in, out, tiles, smem);
        const float t3 = run<3>(in, out, tiles, smem), t4 = run<4>(in, out, tiles, smem), t5 = run<5>(in, out, tiles, smem);
        printf("smem/CTA %6zu B: full-line copy %.3f ms (%.0f GB/s) | 256 aligned 128 B runs %.3f ms (%.0f GB/s) | 256 misaligned runs %.3f ms (%.0f GB/s)\n",
               smem, t0, 8.0 * n / t0 / 1e6, t1, 8.0 * n / t1 / 1e6, t2, 8.0 * n / t2 / 1e6);
        printf("                  sector-misaligned runs %.3f ms | ragged runs, p-linear warps (pass kernels today) %.3f ms | ragged runs, line-aligned warps %.3f ms\n",
               t3, t4, t5);
    }
    for (size_t smem : {(size_t)TILE * 4 + 64, (size_t)70 * 1024}) {
        const float t6 = run<6>(in, out, tiles, smem), t7 = run<7>(in, out, tiles, smem), t8 = run<8>(in, out, tiles, smem);
        printf("smem/CTA %6zu B: ragged runs by bulk S2G stores + scalar heads/tails %.3f ms (%.0f GB/s) | bulk stores only %.3f ms | heads/tails only %.3f ms\n",
               smem, t6, 8.0 * n / t6 / 1e6, t7, t8);
    }
    {   // in-place variant of the full-line copy (what an in-place scan does to DRAM: read and write streams share pages)
        const float t_in = run<0>(in, in, tiles, (size_t)TILE * 4);
        printf("in-place full-line copy (out == in): %.3f ms (%.0f GB/s)\n", t_in, 8.0 * n / t_in / 1e6);
    }
    return 0;
}
