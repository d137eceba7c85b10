// microbench2.cu -- LSU / shared-memory pipe cost table (B200, sm_100a): cycles per warp-instruction,
// SM-wide, for the access shapes a digit-pass kernel issues.  Addresses are precomputed in registers
// so that the loop body is (nearly) only the memory instruction.  Tuning tool, not part of the product.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench2 bench_tools/microbench2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int UNROLL = 16;
constexpr int ITER = 128;

enum Test {
    T_LDS32 = 0, T_LDS64, T_LDS128, T_STS32, T_STS64, T_STS128,
    T_LDS32_RANDOM, T_STS32_RANDOM, T_LDS32_BCAST, T_LDS32_RUNS,
    T_ATOMS_RET, T_ATOMS_NORET, T_ATOMS_RET_PLUS_LDS, T_ATOMS_RET_PLUS_STS_RANDOM,
    T_STG32_LINE, T_STG32_UNALIGNED, T_STG32_2RUNS, T_LDS32_PLUS_STG,
    T_LDS_U8_STS_U8_RMW,
    T_COUNT
};
static const char* kNames[T_COUNT] = {
    "lds.32 conflict-free", "lds.64 conflict-free", "lds.128 conflict-free", "sts.32 conflict-free", "sts.64 conflict-free",
    "sts.128 conflict-free", "lds.32 random word", "sts.32 random word", "lds.32 broadcast", "lds.32 runs of ~8 (gbase)",
    "atoms.ret lane-private", "atoms.noret lane-private", "atoms.ret + lds.32", "atoms.ret + sts.32 random",
    "stg.32 one aligned line", "stg.32 unaligned (2 lines)", "stg.32 two runs (3 lines)", "lds.32 + stg.32 line",
    "lds.u8+sts.u8 rmw private"};

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int TEST>
__global__ void __launch_bounds__(1024) bench_kernel(unsigned long long* cycles, uint32_t* sink, uint32_t* gbuf)
{
    extern __shared__ __align__(16) uint32_t smem[];  // 16384 words = 64 KiB
    constexpr uint32_t WORDS = 16384;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < WORDS; i += blockDim.x) smem[i] = i;
    // per-thread word offsets for the UNROLL accesses of one iteration
    uint32_t off[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint32_t h = hash32(threadIdx.x * 977u + blockIdx.x * 131071u + u * 7919u + 1u);
        const uint32_t row = h & 255u;  // 256 rows of 32 words = first 32 KiB
        if constexpr (TEST == T_LDS32 || TEST == T_STS32 || TEST == T_ATOMS_RET || TEST == T_ATOMS_NORET ||
                      TEST == T_ATOMS_RET_PLUS_LDS || TEST == T_ATOMS_RET_PLUS_STS_RANDOM || TEST == T_LDS32_PLUS_STG ||
                      TEST == T_LDS_U8_STS_U8_RMW)
            off[u] = row * 32u + lane;
        else if constexpr (TEST == T_LDS64 || TEST == T_STS64)
            off[u] = (row & 127u) * 64u + lane * 2u;
        else if constexpr (TEST == T_LDS128 || TEST == T_STS128)
            off[u] = (row & 63u) * 128u + lane * 4u;
        else if constexpr (TEST == T_LDS32_RANDOM || TEST == T_STS32_RANDOM)
            off[u] = (h >> 8) & 8191u;
        else if constexpr (TEST == T_LDS32_BCAST)
            off[u] = hash32(warp * 31u + u) & 8191u;
        else if constexpr (TEST == T_LDS32_RUNS)
            off[u] = hash32(warp * 31u + u * 5u + (lane >> 3)) & 255u;  // 4 distinct words per warp instruction
        else
            off[u] = 0;
    }
    uint32_t off2[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) off2[u] = (hash32(threadIdx.x * 131u + u * 17u + 5u) & 8191u) + 8192u;

    // global store patterns: every warp owns a 16 KiB window (L2 resident), 128-byte lines
    uint32_t* gw = gbuf + ((size_t)blockIdx.x * 32 + warp) * 4096;
    uint32_t goff[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const uint32_t line = (hash32(warp * 7u + u * 3u + blockIdx.x) & 63u) * 32u;  // 64 lines of the first 8 KiB
        if constexpr (TEST == T_STG32_LINE || TEST == T_LDS32_PLUS_STG) goff[u] = line + lane;
        else if constexpr (TEST == T_STG32_UNALIGNED) goff[u] = line + 13u + lane;
        else if constexpr (TEST == T_STG32_2RUNS) goff[u] = lane < 11 ? line + 21u + lane : (line ^ 1024u) + 7u + lane;
        else goff[u] = lane;
    }
    __syncthreads();
    uint32_t acc = 0;
    uint4 acc4 = make_uint4(0, 0, 0, 0);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if constexpr (TEST == T_LDS32 || TEST == T_LDS32_RANDOM || TEST == T_LDS32_BCAST || TEST == T_LDS32_RUNS) {
                acc ^= *(volatile uint32_t*)&smem[off[u]];
            } else if constexpr (TEST == T_LDS64) {
                uint2 v;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y)
                             : "r"((uint32_t)__cvta_generic_to_shared(&smem[off[u]])));
                acc4.x ^= v.x; acc4.y ^= v.y;
            } else if constexpr (TEST == T_LDS128) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"((uint32_t)__cvta_generic_to_shared(&smem[off[u]])));
                acc4.x ^= v.x; acc4.y ^= v.y; acc4.z ^= v.z; acc4.w ^= v.w;
            } else if constexpr (TEST == T_STS32 || TEST == T_STS32_RANDOM) {
                *(volatile uint32_t*)&smem[off[u]] = lane;
            } else if constexpr (TEST == T_STS64) {
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"((uint32_t)__cvta_generic_to_shared(&smem[off[u]])), "r"(lane), "r"(warp) : "memory");
            } else if constexpr (TEST == T_STS128) {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(&smem[off[u]])), "r"(lane), "r"(warp), "r"(lane), "r"(warp) : "memory");
            } else if constexpr (TEST == T_ATOMS_RET) {
                acc ^= atomicAdd(&smem[off[u]], 4u);
            } else if constexpr (TEST == T_ATOMS_NORET) {
                atomicAdd(&smem[off[u]], 4u);
            } else if constexpr (TEST == T_ATOMS_RET_PLUS_LDS) {
                acc ^= atomicAdd(&smem[off[u]], 4u);
                acc ^= *(volatile uint32_t*)&smem[off2[u] - lane + (lane)];
            } else if constexpr (TEST == T_ATOMS_RET_PLUS_STS_RANDOM) {
                acc ^= atomicAdd(&smem[off[u]], 4u);
                *(volatile uint32_t*)&smem[off2[u]] = lane;
            } else if constexpr (TEST == T_STG32_LINE || TEST == T_STG32_UNALIGNED || TEST == T_STG32_2RUNS) {
                __stcs(gw + goff[u], lane + it);
            } else if constexpr (TEST == T_LDS32_PLUS_STG) {
                const uint32_t v = *(volatile uint32_t*)&smem[off[u]];
                __stcs(gw + goff[u], v);
            } else if constexpr (TEST == T_LDS_U8_STS_U8_RMW) {
                volatile uint8_t* p = reinterpret_cast<volatile uint8_t*>(smem) + off[u] * 4u + (warp & 3u);
                const uint8_t v = *p;
                *p = (uint8_t)(v + 1u);
                acc ^= v;
            }
        }
    }
    const long long t1 = clock64();
    __shared__ unsigned long long s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_max, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = s_max;
    acc ^= acc4.x ^ acc4.y ^ acc4.z ^ acc4.w;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
}

template <int TEST>
void run(int sms, unsigned long long* d_cycles, uint32_t* d_sink, uint32_t* gbuf)
{
    const size_t smem = 16384 * sizeof(uint32_t);
    CK(cudaFuncSetAttribute(bench_kernel<TEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    printf("%-30s", kNames[TEST]);
    for (int nw : {1, 4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            bench_kernel<TEST><<<sms, nw * 32, smem>>>(d_cycles, d_sink, gbuf);
            CK(cudaDeviceSynchronize());
        }
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (auto c : h) mean += (double)c;
        mean /= sms;
        const double ops = (double)nw * ITER * UNROLL;
        printf(" | nw=%2d %6.2f cyc/op", nw, mean / ops);
    }
    printf("\n");
}

template <int T>
void run_all(int sms, unsigned long long* c, uint32_t* s, uint32_t* g)
{
    run<T>(sms, c, s, g);
    if constexpr (T + 1 < T_COUNT) run_all<T + 1>(sms, c, s, g);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs; one CTA per SM; cycles per warp-instruction, SM-wide (pairs count as one op)\n", p.name, sms);
    unsigned long long* d_cycles;
    uint32_t *d_sink, *gbuf;
    CK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms));
    CK(cudaMalloc(&d_sink, 4096));
    CK(cudaMalloc(&gbuf, (size_t)sms * 32 * 4096 * sizeof(uint32_t)));
    run_all<0>(sms, d_cycles, d_sink, gbuf);
    return 0;
}
