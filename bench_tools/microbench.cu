// microbench.cu -- SM-level throughput of the primitives a digit-ranking kernel can be built from
// (B200, sm_100a).  One CTA per SM, NW warps per CTA, every warp runs ITER x UNROLL independent
// operations on random 8-bit digits; the table printed is cycles per warp-instruction per SM and
// lanes (keys) per cycle per SM.  Tuning tool, not part of the product.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench bench_tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int UNROLL = 16;
constexpr int ITER = 64;
constexpr unsigned FULL = 0xFFFFFFFFu;

enum Test {
    T_ATOMS_RET = 0,      // returning atomicAdd, lane-private column (conflict-free), random row
    T_ATOMS_NORET,        // same, result unused (RED)
    T_ATOMS_RET_DEP,      // returning, each result consumed at once (serial per warp)
    T_LDS_STS,            // plain ld + st read-modify-write on the same cells
    T_LDS,                // plain ld.shared, conflict-free
    T_STS,                // plain st.shared, conflict-free
    T_MATCH,              // __match_any_sync on the 8-bit digit
    T_BALLOT8,            // 8 x ballot loop producing the same peer mask
    T_WARP_RANK_MATCH,    // full warp-private ranking step: match + popc + leader ld/st + shfl
    T_WARP_RANK_BALLOT,   // same with the ballot loop
    T_ATOMS_RAND,         // returning atomicAdd to fully random words (bank conflicts ~ birthday)
    T_ATOMS_SAME,         // returning atomicAdd, all lanes same word
    T_ATOMS_RET_16B,      // returning atomicAdd, lane-private, rows from only 16 distinct digits
    T_WARP_RANK_MATCH_ATOM, // match + popc + leader atomicAdd(returning) + shfl
    T_COUNT
};
static const char* kNames[T_COUNT] = {
    "atoms.ret lane-private", "atoms.noret lane-private", "atoms.ret dependent", "lds+sts rmw", "lds", "sts",
    "match.any 8b", "ballot x8", "warp-rank match (ld/st)", "warp-rank ballot (ld/st)", "atoms.ret random word",
    "atoms.ret same word", "atoms.ret 16 digits", "warp-rank match (atom)"};

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int TEST>
__global__ void __launch_bounds__(1024) bench_kernel(unsigned long long* cycles, uint32_t* sink)
{
    extern __shared__ uint32_t smem[];  // 256 x 32 matrix (32 KiB) + 32 x 256 warp counters (32 KiB)
    uint32_t* mat = smem;
    uint32_t* wcnt = smem + 256 * 32 + (threadIdx.x >> 5) * 256;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < 256 * 32 * 2; i += blockDim.x) smem[i] = 0;
    uint32_t key[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) key[u] = hash32(threadIdx.x * 977u + blockIdx.x * 131071u + u * 7919u + 1u);
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        uint32_t res[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t k = key[u];
            const uint32_t d = (k >> 8) & 255u;
            if constexpr (TEST == T_ATOMS_RET) {
                res[u] = atomicAdd(&mat[d * 32 + lane], 4u);
            } else if constexpr (TEST == T_ATOMS_RET_16B) {
                res[u] = atomicAdd(&mat[(d & 15u) * 32 + lane], 4u);
            } else if constexpr (TEST == T_ATOMS_NORET) {
                atomicAdd(&mat[d * 32 + lane], 4u);
                res[u] = 0;
            } else if constexpr (TEST == T_ATOMS_RET_DEP) {
                acc += atomicAdd(&mat[((d + acc) & 255u) * 32 + lane], 4u);
                res[u] = 0;
            } else if constexpr (TEST == T_LDS_STS) {
                volatile uint32_t* p = &mat[d * 32 + lane];
                const uint32_t v = *p;
                *p = v + 4u;
                res[u] = v;
            } else if constexpr (TEST == T_LDS) {
                res[u] = *(volatile uint32_t*)&mat[d * 32 + lane];
            } else if constexpr (TEST == T_STS) {
                *(volatile uint32_t*)&mat[d * 32 + lane] = k;
                res[u] = 0;
            } else if constexpr (TEST == T_MATCH) {
                res[u] = __match_any_sync(FULL, d);
            } else if constexpr (TEST == T_BALLOT8) {
                uint32_t m = FULL;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (d >> b) & 1u;
                    const uint32_t v = __ballot_sync(FULL, bit);
                    m &= bit ? v : ~v;
                }
                res[u] = m;
            } else if constexpr (TEST == T_WARP_RANK_MATCH || TEST == T_WARP_RANK_BALLOT || TEST == T_WARP_RANK_MATCH_ATOM) {
                uint32_t m;
                if constexpr (TEST == T_WARP_RANK_BALLOT) {
                    m = FULL;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const bool bit = (d >> b) & 1u;
                        const uint32_t v = __ballot_sync(FULL, bit);
                        m &= bit ? v : ~v;
                    }
                } else {
                    m = __match_any_sync(FULL, d);
                }
                const uint32_t below = __popc(m & ((1u << lane) - 1u));
                const uint32_t leader = 31u - __clz(m);  // highest peer updates the counter
                uint32_t old = 0;
                if (lane == leader) {
                    if constexpr (TEST == T_WARP_RANK_MATCH_ATOM) {
                        old = atomicAdd(&wcnt[d], below + 1u);
                    } else {
                        old = wcnt[d];
                        wcnt[d] = old + below + 1u;
                    }
                }
                old = __shfl_sync(FULL, old, leader);
                res[u] = old + below;
            } else if constexpr (TEST == T_ATOMS_RAND) {
                res[u] = atomicAdd(&mat[(k >> 3) & 8191u], 4u);
            } else if constexpr (TEST == T_ATOMS_SAME) {
                res[u] = atomicAdd(&mat[d * 32], 4u);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            acc ^= res[u];
            key[u] = key[u] * 1664525u + 1013904223u + (TEST == T_ATOMS_RET_DEP ? 0u : 0u);
        }
    }
    const long long t1 = clock64();
    __shared__ unsigned long long s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_max, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = s_max;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc + mat[threadIdx.x];
}

template <int TEST>
void run(int sms, unsigned long long* d_cycles, uint32_t* d_sink)
{
    const size_t smem = 256 * 32 * 2 * sizeof(uint32_t);
    CK(cudaFuncSetAttribute(bench_kernel<TEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    printf("%-28s", kNames[TEST]);
    for (int nw : {1, 2, 4, 8, 16, 32}) {
        bench_kernel<TEST><<<sms, nw * 32, smem>>>(d_cycles, d_sink);
        CK(cudaDeviceSynchronize());
        bench_kernel<TEST><<<sms, nw * 32, smem>>>(d_cycles, d_sink);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (auto c : h) mean += (double)c;
        mean /= sms;
        const double ops = (double)nw * ITER * UNROLL;  // warp-instructions per SM
        printf(" | nw=%2d %6.2f cyc/op %5.2f k/clk", nw, mean / ops, ops * 32.0 / mean);
    }
    printf("\n");
}

template <int T>
void run_all(int sms, unsigned long long* c, uint32_t* s)
{
    run<T>(sms, c, s);
    if constexpr (T + 1 < T_COUNT) run_all<T + 1>(sms, c, s);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs; one CTA per SM; cyc/op = cycles per warp-instruction (SM-wide), k/clk = lanes per cycle per SM\n", p.name, sms);
    unsigned long long* d_cycles;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms));
    CK(cudaMalloc(&d_sink, 4096));
    run_all<0>(sms, d_cycles, d_sink);
    return 0;
}
