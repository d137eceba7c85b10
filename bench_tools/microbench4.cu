// microbench4.cu -- shared-memory load width vs throughput on B200: does LDS.64 / LDS.128 move more than 128 B/clk/SM?
// Addresses change every iteration and every result is consumed, so nothing can be hoisted or merged.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench4 bench_tools/microbench4.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int UNROLL = 8, ITER = 512;

template <int W>  // words per lane and load: 1, 2, 4
__global__ void __launch_bounds__(1024) k(unsigned long long* cycles, uint32_t* sink)
{
    extern __shared__ __align__(16) uint32_t smem[];  // 16384 words
    for (uint32_t i = threadIdx.x; i < 16384; i += blockDim.x) smem[i] = i * 2654435761u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t off[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) off[u] = (((warp * 37u + u * 11u) * 32u * W) & 16383u) * 4u + lane * 4u * W;  // conflict-free rows
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t a = base + ((off[u] + (uint32_t)it * 128u * W) & 65535u);
            if constexpr (W == 1) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
                acc += v;
            } else if constexpr (W == 2) {
                uint32_t v0, v1;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a) : "memory");
                acc += v0 ^ v1;
            } else {
                uint32_t v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a) : "memory");
                acc += (v0 ^ v1) + (v2 ^ v3);
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
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int W>
void run(int sms, unsigned long long* d_cycles, uint32_t* d_sink, const char* name)
{
    CK(cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    printf("%-10s", name);
    for (int nw : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) { k<W><<<sms, nw * 32, 65536>>>(d_cycles, d_sink); CK(cudaDeviceSynchronize()); }
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), d_cycles, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
        double mean = 0; for (auto c : h) mean += (double)c; mean /= sms;
        const double ops = (double)nw * ITER * UNROLL;
        printf(" | nw=%2d %5.2f cyc/op %6.1f B/clk", nw, mean / ops, ops * 128.0 * W / mean);
    }
    printf("\n");
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    unsigned long long* d_cycles; uint32_t* d_sink;
    CK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms));
    CK(cudaMalloc(&d_sink, (size_t)sms * 1024 * 4));
    printf("%s: conflict-free shared loads, cycles per warp instruction SM-wide and bytes per clock per SM\n", p.name);
    run<1>(sms, d_cycles, d_sink, "lds.32");
    run<2>(sms, d_cycles, d_sink, "lds.64");
    run<4>(sms, d_cycles, d_sink, "lds.128");
    return 0;
}
