// microbench3.cu -- latency of a batch of K independent global loads issued back to back by one warp:
// strong (ld.relaxed.gpu), weak L2-only hint (ld.global.L1::no_allocate) and plain weak loads, data in L2.
// Question: does the memory system pipeline gpu-scope strong loads of one warp or serialise them?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench3 bench_tools/microbench3.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ uint2 load2(const uint32_t* p)
{
    uint2 v;
    if constexpr (MODE == 0) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else if constexpr (MODE == 1) asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else if constexpr (MODE == 2) asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

template <int MODE, int K>
__global__ void lat_kernel(const uint32_t* buf, unsigned long long* out, uint32_t* sink, int reps)
{
    const uint32_t lane = threadIdx.x & 31u;
    // every rep touches fresh lines (so plain loads cannot hit L1): rows 1 KiB apart like a look-back walk
    const uint32_t* p = buf + ((size_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * 32768 + lane * 2;
    unsigned long long total = 0;
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r) {
        uint2 w[K];
        const long long t0 = clock64();
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = load2<MODE>(p + (size_t)(r * K + k) * 256);
#pragma unroll
        for (int k = 0; k < K; ++k) acc += w[k].x + w[k].y;
        // force completion before reading the clock
        if (acc == 0x9999999u) sink[0] = acc;
        const long long t1 = clock64();
        total += (unsigned long long)(t1 - t0);
    }
    if (lane == 0) out[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = total / reps;
    if (acc == 0x12345u) sink[1] = acc;
}

template <int MODE, int K>
void run(const uint32_t* buf, unsigned long long* d_out, uint32_t* sink, int blocks, int warps)
{
    const int reps = 8;
    lat_kernel<MODE, K><<<blocks, warps * 32>>>(buf, d_out, sink, reps);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(blocks * warps);
    CK(cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
    double m = 0;
    for (auto v : h) m += (double)v;
    printf(" K=%2d %7.0f", K, m / h.size());
}

template <int MODE>
void run_mode(const char* name, const uint32_t* buf, unsigned long long* d_out, uint32_t* sink, int blocks, int warps)
{
    printf("%-34s blocks=%3d warps=%d |", name, blocks, warps);
    run<MODE, 1>(buf, d_out, sink, blocks, warps);
    run<MODE, 2>(buf, d_out, sink, blocks, warps);
    run<MODE, 4>(buf, d_out, sink, blocks, warps);
    run<MODE, 8>(buf, d_out, sink, blocks, warps);
    run<MODE, 16>(buf, d_out, sink, blocks, warps);
    printf("  cycles per batch\n");
}

int main()
{
    const int blocks = 148;
    uint32_t* buf;
    const size_t bytes = (size_t)blocks * 4 * 32768 * 4 + (1 << 20);  // one 128 KiB window per warp, ~78 MB: L2 resident
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    unsigned long long* d_out;
    uint32_t* sink;
    CK(cudaMalloc(&d_out, sizeof(unsigned long long) * blocks * 64));
    CK(cudaMalloc(&sink, 64));
    for (int warps : {1, 4}) {
        for (int b : {1, 148}) {
            run_mode<0>("ld.relaxed.gpu (STRONG.GPU)", buf, d_out, sink, b, warps);
            run_mode<1>("ld.global.L1::no_allocate (weak)", buf, d_out, sink, b, warps);
            run_mode<2>("ld.global (weak)", buf, d_out, sink, b, warps);
            run_mode<3>("ld.global.nc.L1::no_allocate", buf, d_out, sink, b, warps);
        }
    }
    return 0;
}
